// flowdec_b200 — upstream NDAC (DAC) decode: RVQ dequantisation + 1-D conv decoder.
//
// Replaces `dac.DAC.quantizer.from_codes` and `dac.DAC.decode` of descript-audio-codec 1.0.0
// (un-vendored dependency of the reference; call sites /root/reference/demo.ipynb:101-105):
//   from_codes : z_q = sum_i out_proj_i(codebook_i[codes[:, i, :]])
//   decode     : Conv1d k7 -> [Snake -> ConvTranspose1d(k=2s, stride s) -> 3 x ResidualUnit(k7 dilated, k1)]
//                per rate -> Snake -> Conv1d k7 -> tanh
// ~0.1 TFLOP per audio-second (1 % of the postfilter).  First implementation: fp32 on CUDA
// cores, register-tiled direct convolutions with the Snake activation fused into the operand
// load and bias / residual / tanh fused into the store.  Layout [B, C, T] fp32 as upstream.
#include "fd_common.cuh"

namespace fd {

constexpr int kCoTile = 64;   // output channels per block
constexpr int kTTile = 128;   // output time steps per block
constexpr int kCiChunk = 8;   // input channels staged per iteration

__device__ __forceinline__ float snake_f(float x, float a) {
  const float s = sinf(a * x);
  return x + s * s / (a + 1e-9f);
}

// out[b,co,t] = epi( bias[co] + sum_ci sum_k w[co,ci,k] * act(x[b,ci,t + k*dil - pad]) ) (+ res[b,co,t])
// act = Snake(alpha[ci]) if alpha != nullptr; epi = tanh if do_tanh.
// Strided form (encoder down-sampling conv): input index (t*stride + k*dil - pad).
// block: 256 threads; thread (tx = tid & 15, ty = tid >> 4) owns co = ty*4..+3, t = tx + 16*j (j < 8)
template <int kCiChunk>
__global__ void __launch_bounds__(256) dac_conv1d_kernel(const float* __restrict__ x,
                                                         const float* __restrict__ w,
                                                         const float* __restrict__ bias,
                                                         const float* __restrict__ alpha,
                                                         const float* __restrict__ res,
                                                         float* __restrict__ out, int Cin, int Cout,
                                                         int Tin, int Tout, int K, int dil, int pad,
                                                         int stride, int do_tanh) {
  extern __shared__ float smem[];
  const int span = (kTTile - 1) * stride + (K - 1) * dil + 1;   // input samples needed per channel
  float* sx = smem;                                   // [kCiChunk][span]
  float* sw = smem + kCiChunk * span;                 // [kCiChunk][K][kCoTile]
  const int t0 = blockIdx.x * kTTile, co0 = blockIdx.y * kCoTile, b = blockIdx.z;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  float acc[4][8];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
  const float* xb = x + static_cast<size_t>(b) * Cin * Tin;
  for (int c0 = 0; c0 < Cin; c0 += kCiChunk) {
    __syncthreads();
    for (int i = threadIdx.x; i < kCiChunk * span; i += 256) {
      const int ci = i / span, tt = i - ci * span;
      const int c = c0 + ci, ti = t0 * stride + tt - pad;
      float v = 0.f;
      if (c < Cin && ti >= 0 && ti < Tin) {
        v = xb[static_cast<size_t>(c) * Tin + ti];
        if (alpha) v = snake_f(v, alpha[c]);
      }
      sx[i] = v;
    }
    for (int i = threadIdx.x; i < kCiChunk * K * kCoTile; i += 256) {
      const int co = i % kCoTile, r = i / kCoTile;
      const int k = r % K, ci = r / K;
      const int c = c0 + ci, o = co0 + co;
      sw[i] = (c < Cin && o < Cout) ? w[(static_cast<size_t>(o) * Cin + c) * K + k] : 0.f;
    }
    __syncthreads();
#pragma unroll 1
    for (int ci = 0; ci < kCiChunk; ++ci) {
      for (int k = 0; k < K; ++k) {
        const float4 w4 = *reinterpret_cast<const float4*>(&sw[(ci * K + k) * kCoTile + ty * 4]);
        const float* xr = sx + ci * span + k * dil + tx * stride;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float xv = xr[16 * j * stride];
          acc[0][j] = fmaf(w4.x, xv, acc[0][j]);
          acc[1][j] = fmaf(w4.y, xv, acc[1][j]);
          acc[2][j] = fmaf(w4.z, xv, acc[2][j]);
          acc[3][j] = fmaf(w4.w, xv, acc[3][j]);
        }
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int o = co0 + ty * 4 + i;
    if (o >= Cout) continue;
    const float bv = bias ? bias[o] : 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int t = t0 + tx + 16 * j;
      if (t >= Tout) continue;
      const size_t idx = (static_cast<size_t>(b) * Cout + o) * Tout + t;
      float v = acc[i][j] + bv;
      if (res) v += res[idx];
      if (do_tanh) v = tanhf(v);
      out[idx] = v;
    }
  }
}

// ConvTranspose1d(k = 2s, stride s, padding p): with u = t + p, r = u % s, q = u / s:
//   out[co,t] = bias[co] + sum_ci ( w[ci,co,r] * act(x[ci,q]) + w[ci,co,r+s] * act(x[ci,q-1]) )
__global__ void __launch_bounds__(256) dac_convtr1d_kernel(const float* __restrict__ x,
                                                           const float* __restrict__ w,
                                                           const float* __restrict__ bias,
                                                           const float* __restrict__ alpha,
                                                           float* __restrict__ out, int Cin, int Cout,
                                                           int Tin, int Tout, int s, int pad) {
  extern __shared__ float smem[];
  const int K = 2 * s;
  const int t0 = blockIdx.x * kTTile, co0 = blockIdx.y * kCoTile, b = blockIdx.z;
  const int qbase = (t0 + pad) / s - 1;               // first input index needed
  const int span = (kTTile + s - 1) / s + 2;
  float* sx = smem;                                   // [kCiChunk][span]
  float* sw = smem + kCiChunk * span;                 // [kCiChunk][K][kCoTile]
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  int rr[8], qq[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int u = t0 + tx + 16 * j + pad;
    rr[j] = u % s;
    qq[j] = u / s - qbase;                            // >= 1
  }
  float acc[4][8];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
  const float* xb = x + static_cast<size_t>(b) * Cin * Tin;
  for (int c0 = 0; c0 < Cin; c0 += kCiChunk) {
    __syncthreads();
    for (int i = threadIdx.x; i < kCiChunk * span; i += 256) {
      const int ci = i / span, tt = i - ci * span;
      const int c = c0 + ci, ti = qbase + tt;
      float v = 0.f;
      if (c < Cin && ti >= 0 && ti < Tin) {
        v = xb[static_cast<size_t>(c) * Tin + ti];
        if (alpha) v = snake_f(v, alpha[c]);
      }
      sx[i] = v;
    }
    for (int i = threadIdx.x; i < kCiChunk * K * kCoTile; i += 256) {
      const int co = i % kCoTile, r = i / kCoTile;
      const int k = r % K, ci = r / K;
      const int c = c0 + ci, o = co0 + co;
      sw[i] = (c < Cin && o < Cout) ? w[(static_cast<size_t>(c) * Cout + o) * K + k] : 0.f;
    }
    __syncthreads();
#pragma unroll 1
    for (int ci = 0; ci < kCiChunk; ++ci) {
      const float* xr = sx + ci * span;
      const float* wr = sw + ci * K * kCoTile + ty * 4;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float x0 = xr[qq[j]], x1 = xr[qq[j] - 1];
        const float4 wa = *reinterpret_cast<const float4*>(wr + rr[j] * kCoTile);
        const float4 wb = *reinterpret_cast<const float4*>(wr + (rr[j] + s) * kCoTile);
        acc[0][j] = fmaf(wa.x, x0, fmaf(wb.x, x1, acc[0][j]));
        acc[1][j] = fmaf(wa.y, x0, fmaf(wb.y, x1, acc[1][j]));
        acc[2][j] = fmaf(wa.z, x0, fmaf(wb.z, x1, acc[2][j]));
        acc[3][j] = fmaf(wa.w, x0, fmaf(wb.w, x1, acc[3][j]));
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int o = co0 + ty * 4 + i;
    if (o >= Cout) continue;
    const float bv = bias ? bias[o] : 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int t = t0 + tx + 16 * j;
      if (t < Tout) out[(static_cast<size_t>(b) * Cout + o) * Tout + t] = acc[i][j] + bv;
    }
  }
}

// z[b,d,t] = sum_i ( bias_i[d] + sum_j W_i[d,j] * codebook_i[codes[b,i,t], j] )
__global__ void rvq_from_codes_kernel(const long long* __restrict__ codes, const float* __restrict__ cb,
                                      const float* __restrict__ w, const float* __restrict__ bias,
                                      float* __restrict__ z, int B, int nq, int T, int D, int cdim,
                                      int csize) {
  const size_t total = static_cast<size_t>(B) * D * T;
  for (size_t idx = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int t = static_cast<int>(idx % T);
    const int d = static_cast<int>((idx / T) % D);
    const int b = static_cast<int>(idx / (static_cast<size_t>(T) * D));
    float acc = 0.f;
    for (int i = 0; i < nq; ++i) {
      const long long code = codes[(static_cast<size_t>(b) * nq + i) * T + t];
      const float* e = cb + (static_cast<size_t>(i) * csize + code) * cdim;
      const float* wi = w + (static_cast<size_t>(i) * D + d) * cdim;
      float a = bias[static_cast<size_t>(i) * D + d];
      for (int j = 0; j < cdim; ++j) a = fmaf(wi[j], e[j], a);
      acc += a;
    }
    z[idx] = acc;
  }
}

// ResidualVectorQuantize.forward (eval) — one warp per time column, the running residual and the
// quantised sum of that column live in shared memory.  Per quantizer i:
//   e = in_proj_i(res) ; en = e / max(|e|, 1e-12) ; idx = argmax_c -( |en|^2 - 2 en.cn_c + |cn_c|^2 )
//   (first index wins ties, as torch.max) ; zq_i = out_proj_i(codebook_i[idx]) ; zq += zq_i ; res -= zq_i
constexpr int kRvqWarps = 4;
constexpr int kRvqMaxCdim = 16;

__global__ void __launch_bounds__(kRvqWarps * 32) rvq_encode_kernel(
    const float* __restrict__ z, const float* __restrict__ in_w, const float* __restrict__ in_b,
    const float* __restrict__ cb, const float* __restrict__ cbn, const float* __restrict__ cb2,
    const float* __restrict__ out_w, const float* __restrict__ out_b, long long* __restrict__ codes,
    float* __restrict__ zq, float* __restrict__ latents, float* __restrict__ sqerr, int B, int nq, int T,
    int D, int cdim, int csize) {
  extern __shared__ float smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int col = blockIdx.x * kRvqWarps + warp;           // b * T + t
  if (col >= B * T) return;
  const int b = col / T, t = col - b * T;
  float* res = smem + static_cast<size_t>(warp) * 2 * D;
  float* acc = res + D;
  const float* zc = z + static_cast<size_t>(b) * D * T + t;
  for (int d = lane; d < D; d += 32) {
    res[d] = zc[static_cast<size_t>(d) * T];
    acc[d] = 0.f;
  }
  __syncwarp();
  for (int i = 0; i < nq; ++i) {
    float e[kRvqMaxCdim];
#pragma unroll
    for (int j = 0; j < kRvqMaxCdim; ++j) e[j] = 0.f;
    const float* wi = in_w + static_cast<size_t>(i) * cdim * D;
    for (int d = lane; d < D; d += 32) {
      const float r = res[d];
#pragma unroll
      for (int j = 0; j < kRvqMaxCdim; ++j)
        if (j < cdim) e[j] = fmaf(wi[static_cast<size_t>(j) * D + d], r, e[j]);
    }
    float n2 = 0.f;
#pragma unroll
    for (int j = 0; j < kRvqMaxCdim; ++j) {
      if (j < cdim) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) e[j] += __shfl_xor_sync(0xffffffffu, e[j], o);
        e[j] += in_b[i * cdim + j];
        n2 = fmaf(e[j], e[j], n2);
        if (lane == j) latents[(static_cast<size_t>(b) * nq * cdim + i * cdim + j) * T + t] = e[j];
      }
    }
    const float inv = 1.f / fmaxf(sqrtf(n2), 1e-12f);
    float en[kRvqMaxCdim], l2 = 0.f;
#pragma unroll
    for (int j = 0; j < kRvqMaxCdim; ++j) {
      en[j] = (j < cdim) ? e[j] * inv : 0.f;
      l2 = fmaf(en[j], en[j], l2);
    }
    float best = -INFINITY;
    int bidx = 0x7fffffff;
    const float* cni = cbn + static_cast<size_t>(i) * csize * cdim;
    for (int c = lane; c < csize; c += 32) {
      float dot = 0.f;
#pragma unroll
      for (int j = 0; j < kRvqMaxCdim; ++j)
        if (j < cdim) dot = fmaf(en[j], cni[static_cast<size_t>(c) * cdim + j], dot);
      const float score = -((l2 - 2.f * dot) + cb2[i * csize + c]);
      if (score > best) { best = score; bidx = c; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ob = __shfl_xor_sync(0xffffffffu, best, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bidx, o);
      if (ob > best || (ob == best && oi < bidx)) { best = ob; bidx = oi; }
    }
    const float* ce = cb + (static_cast<size_t>(i) * csize + bidx) * cdim;
    float q[kRvqMaxCdim], se = 0.f;
#pragma unroll
    for (int j = 0; j < kRvqMaxCdim; ++j) {
      q[j] = (j < cdim) ? ce[j] : 0.f;
      if (j < cdim) se = fmaf(e[j] - q[j], e[j] - q[j], se);
    }
    if (lane == 0) {
      codes[(static_cast<size_t>(b) * nq + i) * T + t] = bidx;
      sqerr[(static_cast<size_t>(i) * B + b) * T + t] = se;
    }
    const float* wo = out_w + static_cast<size_t>(i) * D * cdim;
    for (int d = lane; d < D; d += 32) {
      float v = out_b[static_cast<size_t>(i) * D + d];
#pragma unroll
      for (int j = 0; j < kRvqMaxCdim; ++j)
        if (j < cdim) v = fmaf(wo[static_cast<size_t>(d) * cdim + j], q[j], v);
      acc[d] += v;
      res[d] -= v;
    }
    __syncwarp();
  }
  float* zo = zq + static_cast<size_t>(b) * D * T + t;
  for (int d = lane; d < D; d += 32) zo[static_cast<size_t>(d) * T] = acc[d];
}

// commitment / codebook loss of the eval forward: sum_i mean_b mean_{j,t} (z_e - z_q)^2 (one block, fixed order)
__global__ void __launch_bounds__(256) rvq_loss_kernel(const float* __restrict__ sqerr, float* __restrict__ loss,
                                                       int n, float scale) {
  __shared__ double part[256];
  double a = 0.0;
  for (int i = threadIdx.x; i < n; i += 256) a += static_cast<double>(sqerr[i]);
  part[threadIdx.x] = a;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) part[threadIdx.x] += part[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) loss[0] = static_cast<float>(part[0] * static_cast<double>(scale));
}

}  // namespace fd

using namespace fd;

static int launch_dac_conv1d(const float* x, const float* w, const float* bias, const float* snake_alpha,
                             const float* residual, float* out, int B, int Cin, int Cout, int Tin, int K,
                             int dilation, int pad, int stride, int do_tanh, cudaStream_t stream,
                             const char* who) {
  const int Tout = (Tin + 2 * pad - dilation * (K - 1) - 1) / stride + 1;
  FD_REQUIRE(Tin + 2 * pad - dilation * (K - 1) >= 1 && K >= 1 && K <= 16 && stride >= 1 && stride <= 16,
             "%s: bad geometry (Tin=%d, K=%d, stride=%d)", who, Tin, K, stride);
  const int span = (kTTile - 1) * stride + (K - 1) * dilation + 1;
  const int ci = stride == 1 ? kCiChunk : 4;         // strided spans are long: stage fewer channels per step
  const size_t smem = (static_cast<size_t>(ci) * span + ci * K * kCoTile) * sizeof(float);
  FD_REQUIRE(smem <= 48 * 1024, "%s: K*dilation / stride too large for the staging buffer", who);
  dim3 grid((Tout + kTTile - 1) / kTTile, (Cout + kCoTile - 1) / kCoTile, B);
  if (stride == 1)
    dac_conv1d_kernel<kCiChunk><<<grid, 256, smem, stream>>>(x, w, bias, snake_alpha, residual, out, Cin, Cout,
                                                             Tin, Tout, K, dilation, pad, 1, do_tanh);
  else
    dac_conv1d_kernel<4><<<grid, 256, smem, stream>>>(x, w, bias, snake_alpha, residual, out, Cin, Cout, Tin,
                                                      Tout, K, dilation, pad, stride, do_tanh);
  return check_launch(who);
}

extern "C" int fd_dac_conv1d(const float* x, const float* w, const float* bias, const float* snake_alpha,
                             const float* residual, float* out, int B, int Cin, int Cout, int Tin, int K,
                             int dilation, int pad, int do_tanh, cudaStream_t stream) {
  return launch_dac_conv1d(x, w, bias, snake_alpha, residual, out, B, Cin, Cout, Tin, K, dilation, pad, 1,
                           do_tanh, stream, "fd_dac_conv1d");
}

extern "C" int fd_dac_conv1d_strided(const float* x, const float* w, const float* bias,
                                     const float* snake_alpha, float* out, int B, int Cin, int Cout, int Tin,
                                     int K, int stride, int pad, cudaStream_t stream) {
  return launch_dac_conv1d(x, w, bias, snake_alpha, nullptr, out, B, Cin, Cout, Tin, K, 1, pad, stride, 0,
                           stream, "fd_dac_conv1d_strided");
}

extern "C" int fd_rvq_encode(const float* z, const float* in_proj_w, const float* in_proj_b,
                             const float* codebooks, const float* codebooks_l2n, const float* codebooks_l2n_sq,
                             const float* out_proj_w, const float* out_proj_b, long long* codes, float* zq,
                             float* latents, float* loss, float* sqerr_ws, int B, int nq, int T, int D,
                             int codebook_dim, int codebook_size, cudaStream_t stream) {
  FD_REQUIRE(codebook_dim >= 1 && codebook_dim <= kRvqMaxCdim, "fd_rvq_encode: codebook_dim %d > %d",
             codebook_dim, kRvqMaxCdim);
  FD_REQUIRE(B >= 1 && nq >= 1 && T >= 1 && codebook_size >= 1, "fd_rvq_encode: empty problem");
  const size_t smem = static_cast<size_t>(kRvqWarps) * 2 * D * sizeof(float);
  FD_REQUIRE(smem <= 48 * 1024, "fd_rvq_encode: latent dim %d too large for the column buffers", D);
  const int cols = B * T;
  rvq_encode_kernel<<<(cols + kRvqWarps - 1) / kRvqWarps, kRvqWarps * 32, smem, stream>>>(
      z, in_proj_w, in_proj_b, codebooks, codebooks_l2n, codebooks_l2n_sq, out_proj_w, out_proj_b, codes, zq,
      latents, sqerr_ws, B, nq, T, D, codebook_dim, codebook_size);
  int rc = check_launch("fd_rvq_encode");
  if (rc != 0 || loss == nullptr) return rc;
  rvq_loss_kernel<<<1, 256, 0, stream>>>(sqerr_ws, loss, nq * B * T,
                                         1.f / (static_cast<float>(codebook_dim) * T * B));
  return check_launch("fd_rvq_encode(loss)");
}

extern "C" int fd_dac_conv_transpose1d(const float* x, const float* w, const float* bias,
                                       const float* snake_alpha, float* out, int B, int Cin, int Cout,
                                       int Tin, int stride, int pad, cudaStream_t stream) {
  const int K = 2 * stride;
  const int Tout = (Tin - 1) * stride - 2 * pad + K;
  FD_REQUIRE(Tout > 0 && stride >= 1 && stride <= 16, "fd_dac_conv_transpose1d: bad geometry");
  const int span = (kTTile + stride - 1) / stride + 2;
  const size_t smem = (static_cast<size_t>(kCiChunk) * span + kCiChunk * K * kCoTile) * sizeof(float);
  FD_REQUIRE(smem <= 48 * 1024, "fd_dac_conv_transpose1d: stride too large");
  dim3 grid((Tout + kTTile - 1) / kTTile, (Cout + kCoTile - 1) / kCoTile, B);
  dac_convtr1d_kernel<<<grid, 256, smem, stream>>>(x, w, bias, snake_alpha, out, Cin, Cout, Tin, Tout,
                                                   stride, pad);
  return check_launch("fd_dac_conv_transpose1d");
}

extern "C" int fd_rvq_from_codes(const long long* codes, const float* codebooks, const float* out_proj_w,
                                 const float* out_proj_b, float* z, int B, int nq, int T, int D,
                                 int codebook_dim, int codebook_size, cudaStream_t stream) {
  const size_t total = static_cast<size_t>(B) * D * T;
  size_t g = (total + 255) / 256;
  if (g > 148 * 16) g = 148 * 16;
  rvq_from_codes_kernel<<<static_cast<int>(g), 256, 0, stream>>>(codes, codebooks, out_proj_w, out_proj_b, z,
                                                               B, nq, T, D, codebook_dim, codebook_size);
  return check_launch("fd_rvq_from_codes");
}
