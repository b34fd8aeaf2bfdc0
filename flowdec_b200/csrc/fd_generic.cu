// flowdec_b200 — shape-generic kernels for the NCSN++ configurations the tcgen05 tiles do not cover
// (the 7-level SGMSE-style backbone with bottleneck attention and a 3x3 output layer).
//
//   conv2d_direct_kernel      any Cin (multiple of 8) / Cout / H / W, 1 or 9 taps, virtual-concat sources with the
//                             GroupNorm affine (+ SiLU) applied on load; bf16 operands (the same packed K-major
//                             weights as the tensor-core kernels), fp32 accumulation on CUDA cores.  Used where the
//                             image is smaller than a 128-pixel MMA tile (24x8 ... 12x1 levels) or channel counts
//                             are not multiples of 64 — < 1 % of the network's FLOPs.
//                             Replaces nn.Conv2d / NIN there (reference layers.py:110-134, layerspp.py NIN).
//   attn_kernel               AttnBlockpp core (reference layerspp.py:72-101): softmax(q.k / sqrt(C)) v over all
//                             H*W positions of one sample.
//   gn_act_down_any_kernel    GroupNorm affine + SiLU + FIR /2 for images whose sides are not multiples of 4.
//   conv_in_any_kernel        3x3 input conv 4 -> Cout for Cout != 64.
//   output_conv3_axpy_kernel  3x3 output layer 4 -> 2 (ncsnpp.py:100 with kernel_size 3) fused with the sampler stage.
#include "fd_common.cuh"

#include <cstring>

namespace fd {

struct DirectSrc {
  const __nv_bfloat16* ptr;
  int C, c_begin, c_count, taps;
  const float* ss;      // scale/shift [B][ss_pitch][2] at this source's first consumed channel, or nullptr
  int ss_pitch;
  int kbase;            // K offset of the segment in the packed weight row
};

constexpr int kDirMaxSeg = 4;
constexpr int kDirPix = 8;      // output pixels per block
constexpr int kDirKc = 64;      // channels staged per step

struct DirectParams {
  DirectSrc src[kDirMaxSeg];
  int nseg;
  const __nv_bfloat16* w;   // [rows >= cout][ktot], K-major
  int ktot;
  const float* bias;
  void* out;
  int out_f32, cout, out_pitch;
  int H, W;
  int affine_only;          // sources with scale/shift: 0 -> SiLU(x*s+t), 1 -> x*s+t (attention's GroupNorm)
};

__global__ void __launch_bounds__(128) conv2d_direct_kernel(const DirectParams p) {
  __shared__ __align__(16) float xs[kDirPix][kDirKc];
  const int HW = p.H * p.W;
  const int b = blockIdx.y;
  const int p0 = blockIdx.x * kDirPix;
  const int co = blockIdx.z * 128 + threadIdx.x;
  const bool co_ok = co < p.cout;
  float acc[kDirPix];
#pragma unroll
  for (int i = 0; i < kDirPix; ++i) acc[i] = 0.f;
  // staging role: pixel lp, 4 channels starting at lc
  const int lp = threadIdx.x >> 4, lc = (threadIdx.x & 15) * 4;
  const int pix = p0 + lp;
  const int ph = pix / p.W, pw = pix - ph * p.W;
  for (int s = 0; s < p.nseg; ++s) {
    const DirectSrc& S = p.src[s];
    for (int tap = 0; tap < S.taps; ++tap) {
      const int dh = (S.taps == 9) ? tap / 3 - 1 : 0;
      const int dw = (S.taps == 9) ? tap % 3 - 1 : 0;
      const int hh = ph + dh, ww = pw + dw;
      const bool in_img = pix < HW && hh >= 0 && hh < p.H && ww >= 0 && ww < p.W;
      for (int c0 = 0; c0 < S.c_count; c0 += kDirKc) {
        const int kc = min(kDirKc, S.c_count - c0);
        __syncthreads();
        {
          float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
          if (in_img && lc < kc) {
            const int ch = c0 + lc;    // channel inside the segment
            const uint2 r = *reinterpret_cast<const uint2*>(
                S.ptr + ((static_cast<size_t>(b) * HW + hh * p.W + ww) * S.C + S.c_begin + ch));
            const float2 a = unpack_bf16x2(r.x), c = unpack_bf16x2(r.y);
            v = make_float4(a.x, a.y, c.x, c.y);
            if (S.ss != nullptr) {
              const float4* ss = reinterpret_cast<const float4*>(S.ss + (static_cast<size_t>(b) * S.ss_pitch + ch) * 2);
              const float4 q0 = ss[0], q1 = ss[1];
              v.x = fmaf(v.x, q0.x, q0.y);
              v.y = fmaf(v.y, q0.z, q0.w);
              v.z = fmaf(v.z, q1.x, q1.y);
              v.w = fmaf(v.w, q1.z, q1.w);
              if (!p.affine_only) {
                v.x = silu_f(v.x); v.y = silu_f(v.y); v.z = silu_f(v.z); v.w = silu_f(v.w);
              }
              // the tensor-core path rounds the transformed operand to bf16: do the same (one precision model)
              const float2 ra = unpack_bf16x2(pack_bf16x2(v.x, v.y)), rb = unpack_bf16x2(pack_bf16x2(v.z, v.w));
              v = make_float4(ra.x, ra.y, rb.x, rb.y);
            }
          }
          *reinterpret_cast<float4*>(&xs[lp][lc]) = v;
        }
        __syncthreads();
        if (co_ok) {
          const __nv_bfloat16* wr = p.w + static_cast<size_t>(co) * p.ktot + S.kbase + tap * S.c_count + c0;
          for (int k = 0; k < kc; k += 8) {
            const uint4 wq = *reinterpret_cast<const uint4*>(wr + k);
            const float2 w01 = unpack_bf16x2(wq.x), w23 = unpack_bf16x2(wq.y), w45 = unpack_bf16x2(wq.z),
                         w67 = unpack_bf16x2(wq.w);
#pragma unroll
            for (int i = 0; i < kDirPix; ++i) {
              const float4 x0 = *reinterpret_cast<const float4*>(&xs[i][k]);
              const float4 x1 = *reinterpret_cast<const float4*>(&xs[i][k + 4]);
              float a = acc[i];
              a = fmaf(x0.x, w01.x, a); a = fmaf(x0.y, w01.y, a); a = fmaf(x0.z, w23.x, a); a = fmaf(x0.w, w23.y, a);
              a = fmaf(x1.x, w45.x, a); a = fmaf(x1.y, w45.y, a); a = fmaf(x1.z, w67.x, a); a = fmaf(x1.w, w67.y, a);
              acc[i] = a;
            }
          }
        }
      }
    }
  }
  if (!co_ok) return;
  const float bv = p.bias ? p.bias[co] : 0.f;
#pragma unroll
  for (int i = 0; i < kDirPix; ++i) {
    const int q = p0 + i;
    if (q >= HW) break;
    const size_t o = (static_cast<size_t>(b) * HW + q) * p.out_pitch + co;
    if (p.out_f32) static_cast<float*>(p.out)[o] = acc[i] + bv;
    else static_cast<__nv_bfloat16*>(p.out)[o] = __float2bfloat16_rn(acc[i] + bv);
  }
}

// ------------------------------------------------------------------------------------------------
// attention core: qkv fp32 [B, T, 3C] (q | k | v per token) -> out bf16 [B, T, C]
//   out[i] = sum_j softmax_j(q_i . k_j * scale) v_j          (layerspp.py:91-96)
// one block per (query token, sample); 128 threads; scores live in shared memory
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) attn_kernel(const float* __restrict__ qkv, int T, int C, float scale,
                                                   __nv_bfloat16* __restrict__ out) {
  extern __shared__ float sm[];
  float* sq = sm;          // [C]
  float* sc = sm + C;      // [T]
  __shared__ float red[4];
  const int i = blockIdx.x, b = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float* base = qkv + static_cast<size_t>(b) * T * 3 * C;
  for (int c = threadIdx.x; c < C; c += 128) sq[c] = base[static_cast<size_t>(i) * 3 * C + c];
  __syncthreads();
  for (int j = warp; j < T; j += 4) {
    const float* kj = base + static_cast<size_t>(j) * 3 * C + C;
    float a = 0.f;
    for (int c = lane; c < C; c += 32) a = fmaf(sq[c], kj[c], a);
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) a += __shfl_xor_sync(0xffffffffu, a, off);
    if (lane == 0) sc[j] = a * scale;
  }
  __syncthreads();
  float m = -INFINITY;
  for (int j = threadIdx.x; j < T; j += 128) m = fmaxf(m, sc[j]);
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, off));
  if (lane == 0) red[warp] = m;
  __syncthreads();
  m = fmaxf(fmaxf(red[0], red[1]), fmaxf(red[2], red[3]));
  __syncthreads();
  float sum = 0.f;
  for (int j = threadIdx.x; j < T; j += 128) {
    const float e = expf(sc[j] - m);
    sc[j] = e;
    sum += e;
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, off);
  if (lane == 0) red[warp] = sum;
  __syncthreads();
  const float inv = 1.0f / ((red[0] + red[1]) + (red[2] + red[3]));
  for (int c = threadIdx.x; c < C; c += 128) {
    const float* vc = base + 2 * C + c;
    float a = 0.f;
    for (int j = 0; j < T; ++j) a = fmaf(sc[j], vc[static_cast<size_t>(j) * 3 * C], a);
    out[(static_cast<size_t>(b) * T + i) * C + c] = __float2bfloat16_rn(a * inv);
  }
}

// ------------------------------------------------------------------------------------------------
// act = FIR_down(SiLU(x*scale+shift)), raw = FIR_down(x) over the virtual concat [src1, src2]; any even H, W.
// thread = (output pixel, 4 channels)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) gn_act_down_any_kernel(const __nv_bfloat16* __restrict__ src1, int C1,
                                                              const __nv_bfloat16* __restrict__ src2, int C2,
                                                              const float* __restrict__ scale_shift,
                                                              __nv_bfloat16* __restrict__ out,
                                                              __nv_bfloat16* __restrict__ out_raw, int B, int H, int W) {
  const int C = C1 + C2, slices = C / 4;
  const int Ho = H / 2, Wo = W / 2;
  const size_t total = static_cast<size_t>(B) * Ho * Wo * slices;
  const float k[4] = {0.125f, 0.375f, 0.375f, 0.125f};
  for (size_t idx = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int c0 = static_cast<int>(idx % slices) * 4;
    const size_t pix = idx / slices;
    const int wo = static_cast<int>(pix % Wo), ho = static_cast<int>((pix / Wo) % Ho);
    const int b = static_cast<int>(pix / (static_cast<size_t>(Wo) * Ho));
    const bool first = c0 < C1;
    const int Cs = first ? C1 : C2;
    const __nv_bfloat16* img = (first ? src1 : src2) + static_cast<size_t>(b) * H * W * Cs + (first ? c0 : c0 - C1);
    float sc[4] = {1.f, 1.f, 1.f, 1.f}, sh[4] = {0.f, 0.f, 0.f, 0.f};
    if (out != nullptr) {
      const float4* ss = reinterpret_cast<const float4*>(scale_shift + (static_cast<size_t>(b) * C + c0) * 2);
      const float4 q0 = ss[0], q1 = ss[1];
      sc[0] = q0.x; sh[0] = q0.y; sc[1] = q0.z; sh[1] = q0.w; sc[2] = q1.x; sh[2] = q1.y; sc[3] = q1.z; sh[3] = q1.w;
    }
    float a[4] = {0.f, 0.f, 0.f, 0.f}, r[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int hi = 2 * ho - 1 + i;
      if (hi < 0 || hi >= H) continue;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int wi = 2 * wo - 1 + j;
        if (wi < 0 || wi >= W) continue;
        const uint2 q = *reinterpret_cast<const uint2*>(img + (static_cast<size_t>(hi) * W + wi) * Cs);
        const float2 lo = unpack_bf16x2(q.x), hi2 = unpack_bf16x2(q.y);
        const float v[4] = {lo.x, lo.y, hi2.x, hi2.y};
        const float wgt = k[i] * k[j];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          r[c] = fmaf(wgt, v[c], r[c]);
          if (out != nullptr) a[c] = fmaf(wgt, silu_f(fmaf(v[c], sc[c], sh[c])), a[c]);
        }
      }
    }
    const size_t o = pix * C + c0;
    if (out != nullptr) {
      uint2 q;
      q.x = pack_bf16x2(a[0], a[1]);
      q.y = pack_bf16x2(a[2], a[3]);
      *reinterpret_cast<uint2*>(out + o) = q;
    }
    if (out_raw != nullptr) {
      uint2 q;
      q.x = pack_bf16x2(r[0], r[1]);
      q.y = pack_bf16x2(r[2], r[3]);
      *reinterpret_cast<uint2*>(out_raw + o) = q;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// input conv 3x3, 4 -> Cout (Cout a multiple of 8, != 64), fp32 math; thread = (pixel, channel octet)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) conv_in_any_kernel(const float4* __restrict__ in, const float* __restrict__ w,
                                                          const float* __restrict__ bias,
                                                          __nv_bfloat16* __restrict__ out, int B, int H, int W, int Cout) {
  extern __shared__ float sw[];   // [36][Cout] then bias [Cout]
  float* sb = sw + 36 * Cout;
  for (int i = threadIdx.x; i < Cout * 36; i += 256) {
    const int co = i / 36, r = i % 36;      // r = ci*9 + tap (OIHW)
    sw[((r % 9) * 4 + r / 9) * Cout + co] = w[i];
  }
  for (int i = threadIdx.x; i < Cout; i += 256) sb[i] = bias[i];
  __syncthreads();
  const int oct = Cout / 8;
  const size_t total = static_cast<size_t>(B) * H * W * oct;
  for (size_t idx = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int o = static_cast<int>(idx % oct);
    const size_t pix = idx / oct;
    const int wx = static_cast<int>(pix % W), h = static_cast<int>((pix / W) % H);
    const int b = static_cast<int>(pix / (static_cast<size_t>(W) * H));
    float acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = sb[o * 8 + i];
    for (int t = 0; t < 9; ++t) {
      const int hi = h + t / 3 - 1, wi = wx + t % 3 - 1;
      if (hi < 0 || hi >= H || wi < 0 || wi >= W) continue;
      const float4 x = in[(static_cast<size_t>(b) * H + hi) * W + wi];
      const float xin[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
      for (int ci = 0; ci < 4; ++ci) {
        const float* wr = sw + (t * 4 + ci) * Cout + o * 8;
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] = fmaf(xin[ci], wr[i], acc[i]);
      }
    }
    uint4 q;
    q.x = pack_bf16x2(acc[0], acc[1]);
    q.y = pack_bf16x2(acc[2], acc[3]);
    q.z = pack_bf16x2(acc[4], acc[5]);
    q.w = pack_bf16x2(acc[6], acc[7]);
    *reinterpret_cast<uint4*>(out + pix * Cout + o * 8) = q;
  }
}

// ------------------------------------------------------------------------------------------------
// v = Conv3x3_{4->2}(pyr) (no bias); out = c1*base1 + c2*base2 + c3*base3 + coef*v   (complex as float2)
// w: device fp32 [2][4][3][3] (OIHW)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) output_conv3_axpy_kernel(const float4* __restrict__ pyr,
                                                                const float* __restrict__ w,
                                                                const float2* __restrict__ base1, float c1,
                                                                const float2* __restrict__ base2, float c2,
                                                                const float2* __restrict__ base3, float c3, float coef,
                                                                float2* __restrict__ out, float2* __restrict__ v_out,
                                                                int B, int H, int W) {
  __shared__ float sw[72];
  if (threadIdx.x < 72) sw[threadIdx.x] = w[threadIdx.x];
  __syncthreads();
  const size_t n = static_cast<size_t>(B) * H * W;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int wx = static_cast<int>(i % W), h = static_cast<int>((i / W) % H);
    const size_t img = i - (static_cast<size_t>(h) * W + wx);
    float vr = 0.f, vi = 0.f;
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      const int hi = h + t / 3 - 1, wi = wx + t % 3 - 1;
      if (hi < 0 || hi >= H || wi < 0 || wi >= W) continue;
      const float4 q = pyr[img + static_cast<size_t>(hi) * W + wi];
      vr += sw[0 * 9 + t] * q.x + sw[1 * 9 + t] * q.y + sw[2 * 9 + t] * q.z + sw[3 * 9 + t] * q.w;
      vi += sw[36 + 0 * 9 + t] * q.x + sw[36 + 1 * 9 + t] * q.y + sw[36 + 2 * 9 + t] * q.z + sw[36 + 3 * 9 + t] * q.w;
    }
    if (v_out) v_out[i] = make_float2(vr, vi);
    if (out) {
      float2 r = make_float2(coef * vr, coef * vi);
      if (base1) { const float2 a = base1[i]; r.x = fmaf(c1, a.x, r.x); r.y = fmaf(c1, a.y, r.y); }
      if (base2) { const float2 a = base2[i]; r.x = fmaf(c2, a.x, r.x); r.y = fmaf(c2, a.y, r.y); }
      if (base3) { const float2 a = base3[i]; r.x = fmaf(c3, a.x, r.x); r.y = fmaf(c3, a.y, r.y); }
      out[i] = r;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// upfirdn2d, the reference's native op (op/upfirdn2d.cpp:38-48, op/upfirdn2d_kernel.cu:118-218; semantics
// restated from op/upfirdn2d.py:182-224): zero-insertion upsampling by (up_x, up_y), padding / cropping by
// pad_{x,y}{0,1}, correlation with the FLIPPED kernel, decimation by (down_x, down_y).  fp32 NCHW, planes = N*C.
//   out[p, oy, ox] = sum_{ky,kx} kernel[kh-1-ky][kw-1-kx] * up[p, oy*down_y + ky - pad_y0, ox*down_x + kx - pad_x0]
//   up[p, y, x] = in[p, y/up_y, x/up_x] if y, x are in range and divisible, else 0
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) upfirdn2d_f32_kernel(const float* __restrict__ in, const float* __restrict__ kernel,
                                                            float* __restrict__ out, int planes, int in_h, int in_w,
                                                            int kh, int kw, int up_x, int up_y, int down_x, int down_y,
                                                            int pad_x0, int pad_y0, int out_h, int out_w) {
  extern __shared__ float sk[];
  for (int i = threadIdx.x; i < kh * kw; i += blockDim.x) sk[i] = kernel[i];
  __syncthreads();
  const size_t total = static_cast<size_t>(planes) * out_h * out_w;
  for (size_t idx = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int ox = static_cast<int>(idx % out_w), oy = static_cast<int>((idx / out_w) % out_h);
    const size_t pl = idx / (static_cast<size_t>(out_w) * out_h);
    const float* ip = in + pl * in_h * in_w;
    float acc = 0.f;
    for (int ky = 0; ky < kh; ++ky) {
      const int uy = oy * down_y + ky - pad_y0;
      if (uy < 0 || uy % up_y != 0) continue;
      const int iy = uy / up_y;
      if (iy >= in_h) continue;
      for (int kx = 0; kx < kw; ++kx) {
        const int ux = ox * down_x + kx - pad_x0;
        if (ux < 0 || ux % up_x != 0) continue;
        const int ix = ux / up_x;
        if (ix >= in_w) continue;
        acc = fmaf(sk[(kh - 1 - ky) * kw + (kw - 1 - kx)], ip[static_cast<size_t>(iy) * in_w + ix], acc);
      }
    }
    out[idx] = acc;
  }
}

static inline int gen_grid(size_t total, int block) {
  size_t g = (total + block - 1) / block;
  const size_t cap = static_cast<size_t>(148) * 16;
  return static_cast<int>(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace fd

using namespace fd;

struct fd_conv_src {   // same layout as in fd_conv_igemm.cu / include/flowdec_b200.h
  const void* ptr;
  int C, c_begin, c_count, taps;
  const float* scale_shift;
  int ss_pitch;
};

extern "C" int fd_conv2d_direct(const fd_conv_src* srcs, int nsrc, const void* wpacked, int ktot, const float* bias,
                                void* out, int out_is_f32, int cout, int out_pitch, int B, int H, int W, int flags,
                                cudaStream_t stream) {
  FD_REQUIRE(nsrc >= 1 && nsrc <= kDirMaxSeg, "fd_conv2d_direct: nsrc=%d out of range [1,%d]", nsrc, kDirMaxSeg);
  FD_REQUIRE(cout >= 1 && out_pitch >= cout, "fd_conv2d_direct: cout=%d out_pitch=%d", cout, out_pitch);
  FD_REQUIRE(B >= 1 && B <= 65535 && H >= 1 && W >= 1, "fd_conv2d_direct: bad shape B=%d H=%d W=%d", B, H, W);
  DirectParams p;
  memset(&p, 0, sizeof(p));
  int kb = 0;
  for (int s = 0; s < nsrc; ++s) {
    FD_REQUIRE(srcs[s].taps == 1 || srcs[s].taps == 9, "fd_conv2d_direct: taps must be 1 or 9");
    FD_REQUIRE(srcs[s].c_count > 0 && srcs[s].c_count % 8 == 0 && srcs[s].c_begin % 4 == 0 && srcs[s].C % 4 == 0,
               "fd_conv2d_direct: segment channels (count %d, begin %d, pitch %d) must be multiples of 8 / 4 / 4",
               srcs[s].c_count, srcs[s].c_begin, srcs[s].C);
    FD_REQUIRE(srcs[s].scale_shift == nullptr || srcs[s].ss_pitch % 2 == 0, "fd_conv2d_direct: odd ss_pitch");
    p.src[s].ptr = static_cast<const __nv_bfloat16*>(srcs[s].ptr);
    p.src[s].C = srcs[s].C;
    p.src[s].c_begin = srcs[s].c_begin;
    p.src[s].c_count = srcs[s].c_count;
    p.src[s].taps = srcs[s].taps;
    p.src[s].ss = srcs[s].scale_shift;
    p.src[s].ss_pitch = srcs[s].ss_pitch;
    p.src[s].kbase = kb;
    kb += srcs[s].c_count * srcs[s].taps;
  }
  FD_REQUIRE(kb == ktot, "fd_conv2d_direct: packed K=%d does not match segments (%d)", ktot, kb);
  p.nseg = nsrc;
  p.w = static_cast<const __nv_bfloat16*>(wpacked);
  p.ktot = ktot;
  p.bias = bias;
  p.out = out;
  p.out_f32 = out_is_f32;
  p.cout = cout;
  p.out_pitch = out_pitch;
  p.H = H;
  p.W = W;
  p.affine_only = flags & 1;
  dim3 grid((H * W + kDirPix - 1) / kDirPix, B, (cout + 127) / 128);
  conv2d_direct_kernel<<<grid, 128, 0, stream>>>(p);
  return check_launch("fd_conv2d_direct");
}

extern "C" int fd_attention(const float* qkv, int B, int T, int C, float scale, void* out, cudaStream_t stream) {
  const size_t smem = static_cast<size_t>(C + T) * sizeof(float);
  FD_REQUIRE(B >= 1 && B <= 65535 && T >= 1 && C >= 1, "fd_attention: bad shape B=%d T=%d C=%d", B, T, C);
  FD_REQUIRE(smem <= 48 * 1024, "fd_attention: %d tokens x %d channels need %zu bytes of shared memory (> 48 KB)", T, C, smem);
  attn_kernel<<<dim3(T, B), 128, smem, stream>>>(qkv, T, C, scale, static_cast<__nv_bfloat16*>(out));
  return check_launch("fd_attention");
}

extern "C" int fd_gn_act_down_any(const void* src1, int C1, const void* src2, int C2, const float* scale_shift,
                                  void* out, void* out_raw, int B, int H, int W, cudaStream_t stream) {
  FD_REQUIRE(C1 > 0 && C1 % 4 == 0 && C2 % 4 == 0 && (C1 + C2) % 4 == 0, "fd_gn_act_down_any: channels %d + %d", C1, C2);
  FD_REQUIRE(H % 2 == 0 && W % 2 == 0 && H >= 2 && W >= 2, "fd_gn_act_down_any: H=%d W=%d must be even", H, W);
  FD_REQUIRE(out != nullptr || out_raw != nullptr, "fd_gn_act_down_any: no output");
  FD_REQUIRE(out == nullptr || scale_shift != nullptr, "fd_gn_act_down_any: activated output needs scale_shift");
  const size_t total = static_cast<size_t>(B) * (H / 2) * (W / 2) * ((C1 + C2) / 4);
  gn_act_down_any_kernel<<<gen_grid(total, 256), 256, 0, stream>>>(
      static_cast<const __nv_bfloat16*>(src1), C1, static_cast<const __nv_bfloat16*>(src2), C2, scale_shift,
      static_cast<__nv_bfloat16*>(out), static_cast<__nv_bfloat16*>(out_raw), B, H, W);
  return check_launch("fd_gn_act_down_any");
}

extern "C" int fd_conv_in_any(const void* in4, const float* w, const float* bias, void* out, int B, int H, int W,
                              int cout, cudaStream_t stream) {
  FD_REQUIRE(cout % 8 == 0 && cout >= 8 && cout <= 320, "fd_conv_in_any: cout=%d", cout);
  const size_t total = static_cast<size_t>(B) * H * W * (cout / 8);
  conv_in_any_kernel<<<gen_grid(total, 256), 256, 37 * cout * sizeof(float), stream>>>(
      static_cast<const float4*>(in4), w, bias, static_cast<__nv_bfloat16*>(out), B, H, W, cout);
  return check_launch("fd_conv_in_any");
}

extern "C" int fd_output_conv3_axpy(const void* pyr4, const float* w72, const void* base1, float c1, const void* base2,
                                    float c2, const void* base3, float c3, float coef, void* out, void* v_out, int B,
                                    int H, int W, cudaStream_t stream) {
  FD_REQUIRE(out != nullptr || v_out != nullptr, "fd_output_conv3_axpy: no output");
  output_conv3_axpy_kernel<<<gen_grid(static_cast<size_t>(B) * H * W, 256), 256, 0, stream>>>(
      static_cast<const float4*>(pyr4), w72, static_cast<const float2*>(base1), c1,
      static_cast<const float2*>(base2), c2, static_cast<const float2*>(base3), c3, coef,
      static_cast<float2*>(out), static_cast<float2*>(v_out), B, H, W);
  return check_launch("fd_output_conv3_axpy");
}

extern "C" int fd_upfirdn2d_f32(const float* input, int planes, int in_h, int in_w, const float* kernel, int kh, int kw,
                                int up_x, int up_y, int down_x, int down_y, int pad_x0, int pad_x1, int pad_y0,
                                int pad_y1, float* out, cudaStream_t stream) {
  FD_REQUIRE(input != nullptr && kernel != nullptr && out != nullptr, "fd_upfirdn2d_f32: NULL pointer");
  FD_REQUIRE(planes >= 1 && in_h >= 1 && in_w >= 1 && kh >= 1 && kw >= 1 && kh * kw <= 4096,
             "fd_upfirdn2d_f32: bad shape planes=%d in=%dx%d kernel=%dx%d", planes, in_h, in_w, kh, kw);
  FD_REQUIRE(up_x >= 1 && up_y >= 1 && down_x >= 1 && down_y >= 1, "fd_upfirdn2d_f32: up/down factors must be >= 1");
  const int out_h = (in_h * up_y + pad_y0 + pad_y1 - kh) / down_y + 1;
  const int out_w = (in_w * up_x + pad_x0 + pad_x1 - kw) / down_x + 1;
  FD_REQUIRE(in_h * up_y + pad_y0 + pad_y1 >= kh && in_w * up_x + pad_x0 + pad_x1 >= kw && out_h >= 1 && out_w >= 1,
             "fd_upfirdn2d_f32: empty output (%d x %d)", out_h, out_w);
  upfirdn2d_f32_kernel<<<gen_grid(static_cast<size_t>(planes) * out_h * out_w, 256), 256, kh * kw * sizeof(float),
                         stream>>>(input, kernel, out, planes, in_h, in_w, kh, kw, up_x, up_y, down_x, down_y, pad_x0,
                                   pad_y0, out_h, out_w);
  return check_launch("fd_upfirdn2d_f32");
}
