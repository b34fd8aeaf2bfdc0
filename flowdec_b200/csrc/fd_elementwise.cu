// flowdec_b200 — HBM-bound kernels around the tcgen05 convolutions of the NCSN++ backbone.
//
//   GroupNorm statistics / affine+SiLU (+FIR x2 up/down)   reference layerspp.py:252-268,
//                                                          up_or_down_sampling.py:220-282,
//                                                          op/upfirdn2d_kernel.cu:118-218
//   4->64 input conv, Combine 1x1, 4-channel pyramids      reference ncsnpp.py:284,300-321,355-384
//   output 1x1 conv fused with the ODE update              reference ncsnpp.py:398 + torchdyn step
//   time embedding + per-block Dense_0 bias                reference ncsnpp.py:263-274, layerspp.py:270-272
//
// Layouts: activations bf16 NHWC [B,H,W,C]; 4-channel pyramids fp32 [B,H,W,4]; ODE state and
// spectrograms float2 [B,F,T] (bit-identical to complex64 [B,1,F,T]).
#include <type_traits>

#include "fd_common.cuh"

namespace fd {

// FIR taps of the separable [1,3,3,1] filter (normalised per axis)
// down (factor 2, pad (1,1)):  out[i]   = (x[2i-1] + 3 x[2i] + 3 x[2i+1] + x[2i+2]) / 8
// up   (factor 2, pad (2,1)):  out[2i]  = (x[i-1] + 3 x[i]) / 4 ; out[2i+1] = (3 x[i] + x[i+1]) / 4

struct bf16x8 {
  uint4 raw;
};

__device__ __forceinline__ void load8(const __nv_bfloat16* p, float (&v)[8]) {
  const uint4 r = *reinterpret_cast<const uint4*>(p);
  float2 a = unpack_bf16x2(r.x), b = unpack_bf16x2(r.y), c = unpack_bf16x2(r.z), d = unpack_bf16x2(r.w);
  v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y; v[4] = c.x; v[5] = c.y; v[6] = d.x; v[7] = d.y;
}

__device__ __forceinline__ void store8(__nv_bfloat16* p, const float (&v)[8]) {
  uint4 r;
  r.x = pack_bf16x2(v[0], v[1]);
  r.y = pack_bf16x2(v[2], v[3]);
  r.z = pack_bf16x2(v[4], v[5]);
  r.w = pack_bf16x2(v[6], v[7]);
  *reinterpret_cast<uint4*>(p) = r;
}

// fp32 activations (tf32 "precise" mode): the same helpers on float pointers
__device__ __forceinline__ void load8(const float* p, float (&v)[8]) {
  const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}

__device__ __forceinline__ void store8(float* p, const float (&v)[8]) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
}

// ------------------------------------------------------------------------------------------
// per-(sample, channel) sum / sum-of-squares, deterministic two-level reduction
//   partial[b][s][c][2]  (fp32), s = slab index; finalize reduces slabs in fp64
// ------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) chan_stats_kernel(const T* __restrict__ x, int HW,
                                                         int C, float* __restrict__ partial, int S) {
  extern __shared__ float red[];  // [256][16]
  const int b = blockIdx.y, s = blockIdx.x;
  const int oct = C >> 3;                 // threads per pixel
  const int ppi = 256 / oct;              // pixels per iteration
  const int o = threadIdx.x % oct;
  const int pl = threadIdx.x / oct;
  const int per = (HW + S - 1) / S;
  const int p0 = s * per;
  const int p1 = min(HW, p0 + per);
  float sum[8], sq[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) sum[i] = sq[i] = 0.f;
  const T* base = x + static_cast<size_t>(b) * HW * C + o * 8;
  if (pl < ppi) {
    for (int p = p0 + pl; p < p1; p += ppi) {
      float v[8];
      load8(base + static_cast<size_t>(p) * C, v);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        sum[i] += v[i];
        sq[i] = fmaf(v[i], v[i], sq[i]);
      }
    }
  }
  float* my = red + threadIdx.x * 16;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    my[i] = sum[i];
    my[8 + i] = sq[i];
  }
  __syncthreads();
  // thread t < C*2 reduces one (channel, stat) over the ppi pixel lanes, fixed order
  for (int t = threadIdx.x; t < C * 2; t += 256) {
    const int c = t >> 1, st = t & 1;
    const int oo = c >> 3, ci = c & 7;
    float acc = 0.f;
    for (int q = 0; q < ppi; ++q) acc += red[(q * oct + oo) * 16 + st * 8 + ci];
    partial[((static_cast<size_t>(b) * S + s) * C + c) * 2 + st] = acc;
  }
}

// stage 1 of the statistics reduction when a producer wrote many slabs (the conv epilogue writes
// one per warp and tile): block (chunk, b) sums `per` consecutive slabs for all channels with fully
// coalesced row reads; fp32 pairwise over <= a few hundred terms, fixed order.
__global__ void __launch_bounds__(256) slab_reduce_kernel(const float* __restrict__ in, int S, int C2x,
                                                          float* __restrict__ out, int chunks) {
  const int chunk = blockIdx.x, b = blockIdx.y;
  const int per = (S + chunks - 1) / chunks;
  const int s0 = chunk * per, s1 = min(S, s0 + per);
  for (int i = threadIdx.x; i < C2x; i += 256) {
    const float* p = in + (static_cast<size_t>(b) * S + s0) * C2x + i;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    int sl = s0;
    for (; sl + 3 < s1; sl += 4) {
      a0 += p[0];
      a1 += p[C2x];
      a2 += p[2 * static_cast<size_t>(C2x)];
      a3 += p[3 * static_cast<size_t>(C2x)];
      p += 4 * static_cast<size_t>(C2x);
    }
    for (; sl < s1; ++sl) {
      a0 += p[0];
      p += C2x;
    }
    out[(static_cast<size_t>(b) * chunks + chunk) * C2x + i] = (a0 + a1) + (a2 + a3);
  }
}

// one block per (group, sample): mean / rstd over the group's channels of the virtual concat
// [src1 (C1 channels), src2 (C2 channels)], then per-channel scale/shift.
__global__ void __launch_bounds__(128) gn_finalize_kernel(const float* __restrict__ part1, int C1, int S1,
                                                          const float* __restrict__ part2, int C2, int S2,
                                                          double count,
                                                          const float* __restrict__ gamma,
                                                          const float* __restrict__ beta, int groups,
                                                          float eps, float* __restrict__ scale_shift) {
  __shared__ double ssum[128], ssq[128];
  const int g = blockIdx.x, b = blockIdx.y;
  const int C = C1 + C2;
  const int cpg = C / groups;
  double s = 0.0, q = 0.0;
  // The group's channels are contiguous inside a source (a group may straddle the concat seam: then it has one
  // run in each source).  A thread walks whole slabs — the run's (sum, sumsq) pairs are adjacent floats, so the
  // reads are full sectors even when a producer left one slab per conv tile (S up to a few thousand) — in a
  // fixed order; the block tree below is fixed too, so the result does not depend on the batch composition.
  for (int src = 0; src < 2; ++src) {
    const int Cs = src == 0 ? C1 : C2;
    const int S = src == 0 ? S1 : S2;
    const float* part = src == 0 ? part1 : part2;
    const int g_lo = g * cpg, g_hi = g_lo + cpg;              // virtual channel range of the group
    const int lo = max(g_lo, src == 0 ? 0 : C1) - (src == 0 ? 0 : C1);
    const int hi = min(g_hi, src == 0 ? C1 : C) - (src == 0 ? 0 : C1);
    if (Cs == 0 || hi <= lo) continue;
    const float* base = part + (static_cast<size_t>(b) * S * Cs + lo) * 2;
    const int n2 = (hi - lo);                                  // float2 pairs per slab
    for (int sl = threadIdx.x; sl < S; sl += 128) {
      const float2* row = reinterpret_cast<const float2*>(base + static_cast<size_t>(sl) * Cs * 2);
      float fs = 0.f, fq = 0.f;
      for (int i = 0; i < n2; ++i) {
        const float2 v = row[i];
        fs += v.x;
        fq += v.y;
      }
      s += static_cast<double>(fs);
      q += static_cast<double>(fq);
    }
  }
  ssum[threadIdx.x] = s;
  ssq[threadIdx.x] = q;
  __syncthreads();
  for (int off = 64; off > 0; off >>= 1) {
    if (threadIdx.x < off) {
      ssum[threadIdx.x] += ssum[threadIdx.x + off];
      ssq[threadIdx.x] += ssq[threadIdx.x + off];
    }
    __syncthreads();
  }
  const double n = count * cpg;
  const double mean = ssum[0] / n;
  double var = ssq[0] / n - mean * mean;
  if (var < 0.0) var = 0.0;
  const double rstd = 1.0 / sqrt(var + static_cast<double>(eps));
  for (int i = threadIdx.x; i < cpg; i += 128) {
    const int c = g * cpg + i;
    const double sc = static_cast<double>(gamma[c]) * rstd;
    scale_shift[(static_cast<size_t>(b) * C + c) * 2 + 0] = static_cast<float>(sc);
    scale_shift[(static_cast<size_t>(b) * C + c) * 2 + 1] =
        static_cast<float>(static_cast<double>(beta[c]) - mean * sc);
  }
}

// ------------------------------------------------------------------------------------------
// act = [FIR](SiLU(x * scale + shift)) and (optionally) raw = [FIR](x) over the virtual concat
// [src1, src2] -> bf16 NHWC.  MODE 0: same resolution, 1: FIR down x2, 2: FIR up x2.
// grid.y = one input-resolution row unit, threads over (column unit, channel octet):
//   MODE 0: unit = 1 pixel;  MODE 1: unit = 2x2 output patch (6x6 inputs, each activated once);
//   MODE 2: unit = 1 input pixel -> 2x2 output quad (3x3 inputs).
// ------------------------------------------------------------------------------------------
template <int CPT, typename T = __nv_bfloat16>
struct ChanSlice {
  const T* img;              // source image base (+ channel offset) of this sample
  int Cs;                    // channel pitch of that source
  float sc[CPT], sh[CPT];
};

template <int CPT>
__device__ __forceinline__ void loadN(const __nv_bfloat16* p, float (&v)[CPT]) {
  if constexpr (CPT == 8) {
    load8(p, v);
  } else {
    const uint2 r = *reinterpret_cast<const uint2*>(p);
    const float2 a = unpack_bf16x2(r.x), b = unpack_bf16x2(r.y);
    v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
  }
}

__device__ __forceinline__ void loadN(const float* p, float (&v)[4]) {
  const float4 a = *reinterpret_cast<const float4*>(p);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
}
__device__ __forceinline__ void storeN(float* p, const float (&v)[4]) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
}

template <int CPT>
__device__ __forceinline__ void storeN(__nv_bfloat16* p, const float (&v)[CPT]) {
  if constexpr (CPT == 8) {
    store8(p, v);
  } else {
    uint2 r;
    r.x = pack_bf16x2(v[0], v[1]);
    r.y = pack_bf16x2(v[2], v[3]);
    *reinterpret_cast<uint2*>(p) = r;
  }
}

template <int CPT>
__device__ __forceinline__ void store_any(__nv_bfloat16* p, const float (&v)[CPT]) { storeN<CPT>(p, v); }
template <int CPT>
__device__ __forceinline__ void store_any(float* p, const float (&v)[CPT]) { storeN(p, v); }

template <bool ACT, int CPT, typename T>
__device__ __forceinline__ void load_act(const ChanSlice<CPT, T>& o, int H, int W, int hi, int wi,
                                         float (&a)[CPT], float (&r)[CPT], bool want_raw) {
  if (hi < 0 || hi >= H || wi < 0 || wi >= W) {
#pragma unroll
    for (int i = 0; i < CPT; ++i) a[i] = r[i] = 0.f;
    return;
  }
  float v[CPT];
  if constexpr (std::is_same<T, float>::value) loadN(o.img + (static_cast<size_t>(hi) * W + wi) * o.Cs, v);
  else loadN<CPT>(o.img + (static_cast<size_t>(hi) * W + wi) * o.Cs, v);
#pragma unroll
  for (int i = 0; i < CPT; ++i) {
    if (want_raw) r[i] = v[i];
    a[i] = ACT ? silu_fast(fmaf(v[i], o.sc[i], o.sh[i])) : v[i];
  }
}

// FIR-resampling variants.  A thread owns CPT = 4 channels (keeps the register footprint low
// enough for >= 3 blocks per SM) of one unit:
//   MODE 1: unit = 2x2 output patch (6x6 inputs, each activated once);
//   MODE 2: unit = 1 input pixel -> 2x2 output quad (3x3 inputs).
template <int MODE, bool ACT, bool RAW, typename T>
__global__ void __launch_bounds__(256) gn_act_resample_kernel(
    const T* __restrict__ src1, int C1, const T* __restrict__ src2, int C2,
    const float* __restrict__ scale_shift, T* __restrict__ out,
    T* __restrict__ out_raw, int H, int W) {
  constexpr int CPT = 4;
  const int C = C1 + C2;
  const int slices = C / CPT;
  // units along W at input resolution (MODE 1: pairs of output columns)
  const int UW = (MODE == 1) ? W / 4 : W;
  const int UH = (MODE == 1) ? H / 4 : H;
  const int row_unit = blockIdx.y % UH;
  const int b = blockIdx.y / UH;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= UW * slices) return;
  const int uw = t / slices;
  const int c0 = (t - uw * slices) * CPT;
  ChanSlice<CPT, T> o;
  {
    const bool first = c0 < C1;
    o.Cs = first ? C1 : C2;
    o.img = (first ? src1 : src2) + static_cast<size_t>(b) * H * W * o.Cs + (first ? c0 : c0 - C1);
    if (ACT) {
      const float4* ss = reinterpret_cast<const float4*>(scale_shift + (static_cast<size_t>(b) * C + c0) * 2);
#pragma unroll
      for (int i = 0; i < CPT / 2; ++i) {
        const float4 q = ss[i];
        o.sc[2 * i] = q.x; o.sh[2 * i] = q.y; o.sc[2 * i + 1] = q.z; o.sh[2 * i + 1] = q.w;
      }
    }
  }
  if (MODE == 1) {
    // outputs (2*row_unit + {0,1}, 2*uw + {0,1}); inputs rows 4*row_unit-1 .. +4, cols 4*uw-1 .. +4
    const int Ho = H / 2, Wo = W / 2;
    float acc[2][2][CPT], racc[2][2][CPT];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int j = 0; j < 2; ++j)
#pragma unroll
        for (int c = 0; c < CPT; ++c) acc[i][j][c] = racc[i][j][c] = 0.f;
    const float k[4] = {0.125f, 0.375f, 0.375f, 0.125f};
#pragma unroll
    for (int ri = 0; ri < 6; ++ri) {
      const int hi = 4 * row_unit - 1 + ri;
      float h0[CPT], h1[CPT], rh0[CPT], rh1[CPT];  // horizontal FIR of this input row, two output columns
#pragma unroll
      for (int c = 0; c < CPT; ++c) h0[c] = h1[c] = rh0[c] = rh1[c] = 0.f;
#pragma unroll
      for (int ci = 0; ci < 6; ++ci) {
        float a[CPT], r[CPT];
        load_act<ACT, CPT, T>(o, H, W, hi, 4 * uw - 1 + ci, a, r, RAW);
#pragma unroll
        for (int c = 0; c < CPT; ++c) {
          if (ci < 4) { h0[c] = fmaf(k[ci], a[c], h0[c]); if (RAW) rh0[c] = fmaf(k[ci], r[c], rh0[c]); }
          if (ci >= 2) { h1[c] = fmaf(k[ci - 2], a[c], h1[c]); if (RAW) rh1[c] = fmaf(k[ci - 2], r[c], rh1[c]); }
        }
      }
#pragma unroll
      for (int c = 0; c < CPT; ++c) {
        if (ri < 4) {
          acc[0][0][c] = fmaf(k[ri], h0[c], acc[0][0][c]);
          acc[0][1][c] = fmaf(k[ri], h1[c], acc[0][1][c]);
          if (RAW) { racc[0][0][c] = fmaf(k[ri], rh0[c], racc[0][0][c]); racc[0][1][c] = fmaf(k[ri], rh1[c], racc[0][1][c]); }
        }
        if (ri >= 2) {
          acc[1][0][c] = fmaf(k[ri - 2], h0[c], acc[1][0][c]);
          acc[1][1][c] = fmaf(k[ri - 2], h1[c], acc[1][1][c]);
          if (RAW) { racc[1][0][c] = fmaf(k[ri - 2], rh0[c], racc[1][0][c]); racc[1][1][c] = fmaf(k[ri - 2], rh1[c], racc[1][1][c]); }
        }
      }
    }
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const size_t off = ((static_cast<size_t>(b) * Ho + 2 * row_unit + i) * Wo + 2 * uw + j) * C + c0;
        store_any<CPT>(out + off, acc[i][j]);
        if (RAW) store_any<CPT>(out_raw + off, racc[i][j]);
      }
  } else {
    // input pixel (row_unit, uw) -> outputs (2*row_unit + {0,1}, 2*uw + {0,1})
    const int Ho = H * 2, Wo = W * 2;
    float top[3][CPT], bot[3][CPT], rtop[3][CPT], rbot[3][CPT];
#pragma unroll
    for (int cj = 0; cj < 3; ++cj) {
      float a0[CPT], a1[CPT], a2[CPT], r0[CPT], r1[CPT], r2[CPT];
      load_act<ACT, CPT, T>(o, H, W, row_unit - 1, uw - 1 + cj, a0, r0, RAW);
      load_act<ACT, CPT, T>(o, H, W, row_unit, uw - 1 + cj, a1, r1, RAW);
      load_act<ACT, CPT, T>(o, H, W, row_unit + 1, uw - 1 + cj, a2, r2, RAW);
#pragma unroll
      for (int c = 0; c < CPT; ++c) {
        top[cj][c] = 0.25f * a0[c] + 0.75f * a1[c];
        bot[cj][c] = 0.75f * a1[c] + 0.25f * a2[c];
        if (RAW) {
          rtop[cj][c] = 0.25f * r0[c] + 0.75f * r1[c];
          rbot[cj][c] = 0.75f * r1[c] + 0.25f * r2[c];
        }
      }
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      float e[CPT], f[CPT], re[CPT], rf[CPT];
#pragma unroll
      for (int c = 0; c < CPT; ++c) {
        const float* l = i ? bot[0] : top[0];
        const float* m = i ? bot[1] : top[1];
        const float* r = i ? bot[2] : top[2];
        e[c] = 0.25f * l[c] + 0.75f * m[c];
        f[c] = 0.75f * m[c] + 0.25f * r[c];
        if (RAW) {
          const float* rl = i ? rbot[0] : rtop[0];
          const float* rm = i ? rbot[1] : rtop[1];
          const float* rr = i ? rbot[2] : rtop[2];
          re[c] = 0.25f * rl[c] + 0.75f * rm[c];
          rf[c] = 0.75f * rm[c] + 0.25f * rr[c];
        }
      }
      const size_t off = ((static_cast<size_t>(b) * Ho + 2 * row_unit + i) * Wo + 2 * uw) * C + c0;
      store_any<CPT>(out + off, e);
      store_any<CPT>(out + off + C, f);
      if (RAW) {
        store_any<CPT>(out_raw + off, re);
        store_any<CPT>(out_raw + off + C, rf);
      }
    }
  }
}

// MODE 0 specialisation: a = SiLU(x*scale+shift), same resolution.  A thread owns one channel
// octet (scale/shift in registers) and walks kPix pixels spaced one block-row apart, so the 64 B
// of per-channel parameters are loaded once per 8 x 16 B of activations.
template <int kPix, typename T>
__global__ void __launch_bounds__(256) gn_act_kernel(const T* __restrict__ src1, int C1,
                                                     const T* __restrict__ src2, int C2,
                                                     const float* __restrict__ scale_shift,
                                                     T* __restrict__ out, int HW) {
  const int C = C1 + C2;
  const int oct = C >> 3;
  const int ppb = 256 / oct;               // pixels covered by one block pass
  const int o8 = threadIdx.x % oct;
  const int pl = threadIdx.x / oct;
  if (pl >= ppb) return;
  const int b = blockIdx.y;
  const int c0 = o8 * 8;
  const bool first = c0 < C1;
  const int Cs = first ? C1 : C2;
  const T* img = (first ? src1 : src2) + static_cast<size_t>(b) * HW * Cs + (first ? c0 : c0 - C1);
  float sc[8], sh[8];
  const float4* ss = reinterpret_cast<const float4*>(scale_shift + (static_cast<size_t>(b) * C + c0) * 2);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float4 q = ss[i];
    sc[2 * i] = q.x; sh[2 * i] = q.y; sc[2 * i + 1] = q.z; sh[2 * i + 1] = q.w;
  }
  const int p0 = blockIdx.x * (ppb * kPix) + pl;
  T* ob = out + static_cast<size_t>(b) * HW * C + c0;
  if constexpr (std::is_same<T, float>::value) {
#pragma unroll
    for (int j = 0; j < kPix; ++j) {
      const int p = p0 + j * ppb;
      if (p >= HW) continue;
      float v[8];
      load8(img + static_cast<size_t>(p) * Cs, v);
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = silu_f(fmaf(v[i], sc[i], sh[i]));
      store8(ob + static_cast<size_t>(p) * C, v);
    }
  } else {
    uint4 raw[kPix];
#pragma unroll
    for (int j = 0; j < kPix; ++j) {
      const int p = p0 + j * ppb;
      if (p < HW) raw[j] = *reinterpret_cast<const uint4*>(img + static_cast<size_t>(p) * Cs);
    }
#pragma unroll
    for (int j = 0; j < kPix; ++j) {
      const int p = p0 + j * ppb;
      if (p >= HW) continue;
      const float2 a = unpack_bf16x2(raw[j].x), bq = unpack_bf16x2(raw[j].y), c = unpack_bf16x2(raw[j].z),
                   d = unpack_bf16x2(raw[j].w);
      float v[8] = {a.x, a.y, bq.x, bq.y, c.x, c.y, d.x, d.y};
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = silu_f(fmaf(v[i], sc[i], sh[i]));
      store8(ob + static_cast<size_t>(p) * C, v);
    }
  }
}

// ------------------------------------------------------------------------------------------
// 4-channel fp32 helpers
// ------------------------------------------------------------------------------------------
__global__ void pack4_kernel(const float2* __restrict__ x, const float2* __restrict__ y,
                             float4* __restrict__ out, size_t n) {
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const float2 a = x[i], b = y[i];
    out[i] = make_float4(a.x, a.y, b.x, b.y);
  }
}

__global__ void fir_down4_kernel(const float4* __restrict__ in, float4* __restrict__ out, int B, int H,
                                 int W) {
  const int Ho = H / 2, Wo = W / 2;
  const size_t total = static_cast<size_t>(B) * Ho * Wo;
  const float k[4] = {0.125f, 0.375f, 0.375f, 0.125f};
  for (size_t idx = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int wo = static_cast<int>(idx % Wo);
    const int ho = static_cast<int>((idx / Wo) % Ho);
    const int b = static_cast<int>(idx / (static_cast<size_t>(Wo) * Ho));
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      const int hi = 2 * ho - 1 + a;
      if (hi < 0 || hi >= H) continue;
#pragma unroll
      for (int bb = 0; bb < 4; ++bb) {
        const int wi = 2 * wo - 1 + bb;
        if (wi < 0 || wi >= W) continue;
        const float4 v = in[(static_cast<size_t>(b) * H + hi) * W + wi];
        const float wgt = k[a] * k[bb];
        acc.x = fmaf(wgt, v.x, acc.x);
        acc.y = fmaf(wgt, v.y, acc.y);
        acc.z = fmaf(wgt, v.z, acc.z);
        acc.w = fmaf(wgt, v.w, acc.w);
      }
    }
    out[idx] = acc;
  }
}

// out[2H,2W] = FIR_up(lo[H,W]) + add[2H,2W]   (out may alias add)
__global__ void pyramid_up_add_kernel(const float4* __restrict__ lo, const float4* add,
                                      float4* out, int B, int H, int W) {
  const int Ho = 2 * H, Wo = 2 * W;
  const size_t total = static_cast<size_t>(B) * Ho * Wo;
  for (size_t idx = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int wo = static_cast<int>(idx % Wo);
    const int ho = static_cast<int>((idx / Wo) % Ho);
    const int b = static_cast<int>(idx / (static_cast<size_t>(Wo) * Ho));
    const int hi = ho >> 1, wi = wo >> 1;
    const int hn = (ho & 1) ? hi + 1 : hi - 1;
    const int wn = (wo & 1) ? wi + 1 : wi - 1;
    float4 acc = add[idx];
    auto tap = [&](int h, int w, float wgt) {
      if (h < 0 || h >= H || w < 0 || w >= W) return;
      const float4 v = lo[(static_cast<size_t>(b) * H + h) * W + w];
      acc.x = fmaf(wgt, v.x, acc.x);
      acc.y = fmaf(wgt, v.y, acc.y);
      acc.z = fmaf(wgt, v.z, acc.z);
      acc.w = fmaf(wgt, v.w, acc.w);
    };
    tap(hi, wi, 0.5625f);
    tap(hi, wn, 0.1875f);
    tap(hn, wi, 0.1875f);
    tap(hn, wn, 0.0625f);
    out[idx] = acc;
  }
}

// 3x3 conv to 4 channels evaluated as "GEMM first, shift after": the tensor-core kernel produces
// per-pixel partial products part[b,h,w,tap*4+co] = W_tap[co,:] . a[b,h,w,:] (one pass over a),
// this kernel sums the 9 shifted partials (+ bias, + FIR-upsampled coarser pyramid):
//   out[p,co] = bias[co] + sum_tap part[p + delta_tap][tap*4+co] (+ FIR_up(lo)[p,co])
__global__ void pyramid_gather_kernel(const float* __restrict__ part, int pc, const float* __restrict__ bias,
                                      const float4* __restrict__ lo, float4* __restrict__ out, int B, int H,
                                      int W) {
  const size_t total = static_cast<size_t>(B) * H * W;
  const float4 bv = make_float4(bias[0], bias[1], bias[2], bias[3]);
  for (size_t idx = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int w = static_cast<int>(idx % W);
    const int h = static_cast<int>((idx / W) % H);
    const int b = static_cast<int>(idx / (static_cast<size_t>(W) * H));
    float4 acc = bv;
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      const int hh = h + t / 3 - 1, ww = w + t % 3 - 1;
      if (hh < 0 || hh >= H || ww < 0 || ww >= W) continue;
      const float4 v = *reinterpret_cast<const float4*>(
          part + ((static_cast<size_t>(b) * H + hh) * W + ww) * pc + t * 4);
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    if (lo != nullptr) {
      const int Hl = H / 2, Wl = W / 2;
      const int hi = h >> 1, wi = w >> 1;
      const int hn = (h & 1) ? hi + 1 : hi - 1;
      const int wn = (w & 1) ? wi + 1 : wi - 1;
      auto tap = [&](int y, int x, float wgt) {
        if (y < 0 || y >= Hl || x < 0 || x >= Wl) return;
        const float4 v = lo[(static_cast<size_t>(b) * Hl + y) * Wl + x];
        acc.x = fmaf(wgt, v.x, acc.x);
        acc.y = fmaf(wgt, v.y, acc.y);
        acc.z = fmaf(wgt, v.z, acc.z);
        acc.w = fmaf(wgt, v.w, acc.w);
      };
      tap(hi, wi, 0.5625f);
      tap(hi, wn, 0.1875f);
      tap(hn, wi, 0.1875f);
      tap(hn, wn, 0.0625f);
    }
    out[idx] = acc;
  }
}

// ------------------------------------------------------------------------------------------
// input conv 3x3, 4 -> 64 channels (fp32 math on CUDA cores: K = 36 is too thin for UMMA and the
// ODE state stays fp32 on the way in).  Block = 8 warps = 8 channel octets; each lane computes
// two adjacent pixels x 8 channels so every LDS.128 of weights feeds 8 FMAs.
// smem weights: [tap][ci][64]
// ------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) conv_in_kernel(const float4* __restrict__ in,
                                                      const float* __restrict__ w,
                                                      const float* __restrict__ bias,
                                                      T* __restrict__ out, int B, int H,
                                                      int W) {
  __shared__ __align__(16) float sw[36][64];
  __shared__ float sb[64];
  for (int i = threadIdx.x; i < 64 * 36; i += 256) {
    const int co = i / 36, r = i % 36;      // r = ci*9 + kh*3 + kw   (OIHW)
    const int ci = r / 9, t = r % 9;
    sw[t * 4 + ci][co] = w[i];
  }
  if (threadIdx.x < 64) sb[threadIdx.x] = bias[threadIdx.x];
  __syncthreads();
  const int lane = threadIdx.x & 31, o = threadIdx.x >> 5;
  const int wtiles = (W + 63) / 64;
  const int ntiles = B * H * wtiles;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int wt = tile % wtiles;
    const int h = (tile / wtiles) % H;
    const int b = tile / (wtiles * H);
    const int wx = wt * 64 + lane * 2;
    if (wx >= W) continue;
    // packed fp32x2 accumulators (FFMA2): channel pairs (2i, 2i+1) of the octet for the two pixels
    float2 a0[4], a1[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) a0[i] = a1[i] = make_float2(sb[o * 8 + 2 * i], sb[o * 8 + 2 * i + 1]);
#pragma unroll
    for (int kh = 0; kh < 3; ++kh) {
      const int hi = h + kh - 1;
      if (hi < 0 || hi >= H) continue;
      const float4* rowp = in + (static_cast<size_t>(b) * H + hi) * W;
      float4 v[4];   // input columns wx-1 .. wx+2
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int wi = wx - 1 + j;
        v[j] = (wi >= 0 && wi < W) ? rowp[wi] : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int kw = 0; kw < 3; ++kw) {
        const int t = kh * 3 + kw;
        const float4 x0 = v[kw], x1 = v[kw + 1];
        const float xin0[4] = {x0.x, x0.y, x0.z, x0.w};
        const float xin1[4] = {x1.x, x1.y, x1.z, x1.w};
#pragma unroll
        for (int ci = 0; ci < 4; ++ci) {
          const float4 wa = *reinterpret_cast<const float4*>(&sw[t * 4 + ci][o * 8]);
          const float4 wb = *reinterpret_cast<const float4*>(&sw[t * 4 + ci][o * 8 + 4]);
          const float2 w2[4] = {make_float2(wa.x, wa.y), make_float2(wa.z, wa.w), make_float2(wb.x, wb.y),
                                make_float2(wb.z, wb.w)};
          const float2 p0 = make_float2(xin0[ci], xin0[ci]), p1 = make_float2(xin1[ci], xin1[ci]);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            a0[i] = ffma2(p0, w2[i], a0[i]);
            a1[i] = ffma2(p1, w2[i], a1[i]);
          }
        }
      }
    }
    const float acc0[8] = {a0[0].x, a0[0].y, a0[1].x, a0[1].y, a0[2].x, a0[2].y, a0[3].x, a0[3].y};
    const float acc1[8] = {a1[0].x, a1[0].y, a1[1].x, a1[1].y, a1[2].x, a1[2].y, a1[3].x, a1[3].y};
    T* op = out + ((static_cast<size_t>(b) * H + h) * W + wx) * 64 + o * 8;
    store8(op, acc0);
    if (wx + 1 < W) store8(op + 64, acc1);
  }
}

// Combine(method='sum'): out = h + Conv1x1_{4->C}(pyr) + bias   (layerspp.py:62-69)
// thread = one channel octet (weights held in registers) walking pixels with the block stride
template <typename T>
__global__ void __launch_bounds__(256) combine_kernel(const float4* __restrict__ pyr,
                                                      const float* __restrict__ w,   // [C][4]
                                                      const float* __restrict__ bias,
                                                      const T* __restrict__ h,
                                                      T* __restrict__ out, int npix,
                                                      int C) {
  const int oct = C >> 3;
  const int o = threadIdx.x % oct;
  const int pl = threadIdx.x / oct;
  const int ppb = 256 / oct;          // pixels per block iteration
  if (pl >= ppb) return;
  float4 wr[8];
  float br[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    wr[i] = *reinterpret_cast<const float4*>(w + (o * 8 + i) * 4);
    br[i] = bias[o * 8 + i];
  }
  for (int pix = blockIdx.x * ppb + pl; pix < npix; pix += gridDim.x * ppb) {
    const float4 p = pyr[pix];
    float v[8];
    load8(h + static_cast<size_t>(pix) * C + o * 8, v);
#pragma unroll
    for (int i = 0; i < 8; ++i)
      v[i] += br[i] + wr[i].x * p.x + wr[i].y * p.y + wr[i].z * p.z + wr[i].w * p.w;
    store8(out + static_cast<size_t>(pix) * C + o * 8, v);
  }
}

// v = W_out[2x4] * pyr ; out = c1*base1 + c2*base2 + c3*base3 + coef*v   (complex as float2).
// One fused stage of every sampler: Euler / midpoint / Heun (base1 = x, base2 = predictor state),
// reverse-diffusion predictor (x, y, noise) and annealed-Langevin corrector (x, noise).
__global__ void output_axpy_kernel(const float4* __restrict__ pyr, float w00, float w01, float w02,
                                   float w03, float w10, float w11, float w12, float w13,
                                   const float2* __restrict__ base1, float c1,
                                   const float2* __restrict__ base2, float c2,
                                   const float2* __restrict__ base3, float c3, float coef,
                                   float2* __restrict__ out, float2* __restrict__ v_out, size_t n) {
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const float4 p = pyr[i];
    const float vr = w00 * p.x + w01 * p.y + w02 * p.z + w03 * p.w;
    const float vi = w10 * p.x + w11 * p.y + w12 * p.z + w13 * p.w;
    if (v_out) v_out[i] = make_float2(vr, vi);
    if (out) {
      float2 r = make_float2(coef * vr, coef * vi);
      if (base1) {
        const float2 a = base1[i];
        r.x = fmaf(c1, a.x, r.x);
        r.y = fmaf(c1, a.y, r.y);
      }
      if (base2) {
        const float2 a = base2[i];
        r.x = fmaf(c2, a.x, r.x);
        r.y = fmaf(c2, a.y, r.y);
      }
      if (base3) {
        const float2 a = base3[i];
        r.x = fmaf(c3, a.x, r.x);
        r.y = fmaf(c3, a.y, r.y);
      }
      out[i] = r;
    }
  }
}

// x0 = Y + fac * (float)(sigma[f] (f64) * eps)      (model.py:512,530-536)
__global__ void x0_kernel(const float2* __restrict__ Y, const double* __restrict__ sigma,
                          const float2* __restrict__ eps, float fac, float2* __restrict__ out, int Fq,
                          int T, size_t n) {
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int f = static_cast<int>((i / T) % Fq);
    const double s = sigma[f];
    const float2 e = eps[i], y = Y[i];
    const float nr = static_cast<float>(s * static_cast<double>(e.x));
    const float ni = static_cast<float>(s * static_cast<double>(e.y));
    out[i] = make_float2(y.x + fac * nr, y.y + fac * ni);
  }
}

// ------------------------------------------------------------------------------------------
// time embedding (tiny; fp64 trig because 2*pi*t*W reaches hundreds of radians)
// ------------------------------------------------------------------------------------------
__global__ void fourier_embed_kernel(float t, const float* __restrict__ Wf, int nf,
                                     float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < nf) {
    // reference computes x[:,None] * W[None,:] * 2 * np.pi in fp32 (layerspp.py:50)
    const float proj = t * Wf[i] * 2.0f * 3.14159265358979323846f;
    out[i] = static_cast<float>(sin(static_cast<double>(proj)));
    out[nf + i] = static_cast<float>(cos(static_cast<double>(proj)));
  }
}

// out[m] = (add ? add[m] : 0) + out_scale * (b[m] + sum_k W[m][k] * act(in[k]));  one warp per row
__global__ void matvec_kernel(const float* __restrict__ in, int K, int silu_in,
                              const float* __restrict__ Wm, const float* __restrict__ b,
                              const float* __restrict__ add, float out_scale,
                              float* __restrict__ out, int M) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= M) return;
  float acc = 0.f;
  for (int k = lane; k < K; k += 32) {
    float v = in[k];
    if (silu_in) v = v / (1.0f + expf(-v));
    acc = fmaf(Wm[static_cast<size_t>(row) * K + k], v, acc);
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
  if (lane == 0) out[row] = (add ? add[row] : 0.f) + out_scale * (acc + (b ? b[row] : 0.f));
}

static inline int grid_for(size_t total, int block) {
  size_t g = (total + block - 1) / block;
  const size_t cap = static_cast<size_t>(148) * 16;
  return static_cast<int>(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace fd

using namespace fd;

namespace fd {   // fd_fir_tiles.cu
bool fir_tiles_eligible(int C1, int C2, int H, int W, int mode);
int fir_tiles_launch(const void* src1, int C1, const void* src2, int C2, const float* scale_shift, void* out,
                     void* out_raw, int B, int H, int W, int mode, cudaStream_t stream);
int fir_tiles_set(int on);
}  // namespace fd

extern "C" int fd_fir_tiles_enable(int on) { return fir_tiles_set(on); }
typedef __nv_bfloat16 bf16;

template <typename T>
static int chan_stats_launch(const void* x, int B, int HW, int C, float* partial, int S, cudaStream_t stream) {
  FD_REQUIRE(C % 8 == 0 && C >= 8 && C <= 2048, "fd_chan_stats: unsupported C=%d", C);
  FD_REQUIRE(S >= 1, "fd_chan_stats: S must be >= 1");
  chan_stats_kernel<T><<<dim3(S, B), 256, 256 * 16 * sizeof(float), stream>>>(
      static_cast<const T*>(x), HW, C, partial, S);
  return check_launch("fd_chan_stats");
}

extern "C" int fd_chan_stats(const void* x, int B, int HW, int C, float* partial, int S,
                             cudaStream_t stream) {
  return chan_stats_launch<bf16>(x, B, HW, C, partial, S, stream);
}

// the `_f32` entry points take fp32 NHWC activations (tf32 "precise" mode of the backbone); same contracts
extern "C" int fd_chan_stats_f32(const void* x, int B, int HW, int C, float* partial, int S,
                                 cudaStream_t stream) {
  return chan_stats_launch<float>(x, B, HW, C, partial, S, stream);
}

extern "C" int fd_slab_reduce(const float* in, int B, int S, int C, float* out, int chunks,
                              cudaStream_t stream) {
  FD_REQUIRE(chunks >= 1 && chunks <= S, "fd_slab_reduce: chunks=%d out of range for S=%d", chunks, S);
  slab_reduce_kernel<<<dim3(chunks, B), 256, 0, stream>>>(in, S, C * 2, out, chunks);
  return check_launch("fd_slab_reduce");
}

extern "C" int fd_gn_finalize(const float* part1, int C1, int S1, const float* part2, int C2, int S2,
                              int B, double count, const float* gamma, const float* beta, int groups,
                              float eps, float* scale_shift, cudaStream_t stream) {
  FD_REQUIRE((C1 + C2) % groups == 0, "fd_gn_finalize: C=%d not divisible by groups=%d", C1 + C2, groups);
  gn_finalize_kernel<<<dim3(groups, B), 128, 0, stream>>>(part1, C1, S1, part2, C2, S2, count, gamma,
                                                          beta, groups, eps, scale_shift);
  return check_launch("fd_gn_finalize");
}

template <typename T>
static int gn_act_resample_launch(const void* src1, int C1, const void* src2, int C2,
                                  const float* scale_shift, void* out, void* out_raw, int B, int H,
                                  int W, int mode, cudaStream_t stream) {
  FD_REQUIRE(C1 % 8 == 0 && C2 % 8 == 0 && C1 > 0, "fd_gn_act_resample: channels must be multiples of 8");
  FD_REQUIRE(mode >= 0 && mode <= 2, "fd_gn_act_resample: mode %d", mode);
  FD_REQUIRE(mode != 1 || (H % 4 == 0 && W % 4 == 0), "fd_gn_act_resample: down-sampling needs H, W % 4 == 0");
  FD_REQUIRE(out != nullptr || out_raw != nullptr, "fd_gn_act_resample: no output");
  FD_REQUIRE(out == nullptr || scale_shift != nullptr, "fd_gn_act_resample: activated output needs scale_shift");
  FD_REQUIRE(mode != 0 || out != nullptr, "fd_gn_act_resample: mode 0 without activation is a copy");
  if (std::is_same<T, bf16>::value && out != nullptr && out_raw != nullptr && fir_tiles_eligible(C1, C2, H, W, mode))
    return fir_tiles_launch(src1, C1, src2, C2, scale_shift, out, out_raw, B, H, W, mode, stream);
  const int oct = (C1 + C2) / 8;
  const int slices = (C1 + C2) / 4;   // FIR variants: 4 channels per thread
  const int UW = mode == 1 ? W / 4 : W;
  const int UH = mode == 1 ? H / 4 : H;
  FD_REQUIRE(static_cast<long long>(B) * UH <= 65535 && oct <= 256,
             "fd_gn_act_resample: B*rows=%lld exceeds grid.y (or C > 2048)", static_cast<long long>(B) * UH);
  dim3 grid((UW * slices + 255) / 256, B * UH);
  const T* s1 = static_cast<const T*>(src1);
  const T* s2 = static_cast<const T*>(src2);
  T* o = static_cast<T*>(out);
  T* r = static_cast<T*>(out_raw);
#define FD_LAUNCH_GN(M, A, R) \
  gn_act_resample_kernel<M, A, R, T><<<grid, 256, 0, stream>>>(s1, C1, s2, C2, scale_shift, A ? o : r, r, H, W)
  if (o != nullptr && r != nullptr) {
    if (mode == 1) FD_LAUNCH_GN(1, true, true);
    else if (mode == 2) FD_LAUNCH_GN(2, true, true);
    else FD_REQUIRE(false, "fd_gn_act_resample: raw output only with resampling modes");
  } else if (o != nullptr) {
    if (mode == 0) {
      const int ppb = 256 / oct;
      constexpr int kPix = 8;
      dim3 g0((H * W + ppb * kPix - 1) / (ppb * kPix), B);
      gn_act_kernel<kPix, T><<<g0, 256, 0, stream>>>(s1, C1, s2, C2, scale_shift, o, H * W);
    } else if (mode == 1) FD_LAUNCH_GN(1, true, false);
    else FD_LAUNCH_GN(2, true, false);
  } else {
    if (mode == 1) FD_LAUNCH_GN(1, false, false);
    else FD_LAUNCH_GN(2, false, false);
  }
#undef FD_LAUNCH_GN
  return check_launch("fd_gn_act_resample");
}

extern "C" int fd_gn_act_resample(const void* src1, int C1, const void* src2, int C2,
                                  const float* scale_shift, void* out, void* out_raw, int B, int H,
                                  int W, int mode, cudaStream_t stream) {
  return gn_act_resample_launch<bf16>(src1, C1, src2, C2, scale_shift, out, out_raw, B, H, W, mode, stream);
}

extern "C" int fd_gn_act_resample_f32(const void* src1, int C1, const void* src2, int C2,
                                      const float* scale_shift, void* out, void* out_raw, int B, int H,
                                      int W, int mode, cudaStream_t stream) {
  return gn_act_resample_launch<float>(src1, C1, src2, C2, scale_shift, out, out_raw, B, H, W, mode, stream);
}

extern "C" int fd_pack4(const void* x, const void* y, void* out, size_t npix, cudaStream_t stream) {
  pack4_kernel<<<grid_for(npix, 256), 256, 0, stream>>>(static_cast<const float2*>(x),
                                                       static_cast<const float2*>(y),
                                                       static_cast<float4*>(out), npix);
  return check_launch("fd_pack4");
}

extern "C" int fd_fir_down4(const void* in, void* out, int B, int H, int W, cudaStream_t stream) {
  FD_REQUIRE(H % 2 == 0 && W % 2 == 0, "fd_fir_down4: odd size");
  fir_down4_kernel<<<grid_for(static_cast<size_t>(B) * (H / 2) * (W / 2), 256), 256, 0, stream>>>(
      static_cast<const float4*>(in), static_cast<float4*>(out), B, H, W);
  return check_launch("fd_fir_down4");
}

extern "C" int fd_pyramid_up_add(const void* lo, const void* add, void* out, int B, int H, int W,
                                 cudaStream_t stream) {
  pyramid_up_add_kernel<<<grid_for(static_cast<size_t>(B) * H * W * 4, 256), 256, 0, stream>>>(
      static_cast<const float4*>(lo), static_cast<const float4*>(add), static_cast<float4*>(out), B, H,
      W);
  return check_launch("fd_pyramid_up_add");
}

extern "C" int fd_pyramid_gather(const float* part, int part_channels, const float* bias, const void* lo4,
                                 void* out4, int B, int H, int W, cudaStream_t stream) {
  FD_REQUIRE(part_channels >= 36 && part_channels % 4 == 0, "fd_pyramid_gather: part_channels=%d", part_channels);
  pyramid_gather_kernel<<<grid_for(static_cast<size_t>(B) * H * W, 256), 256, 0, stream>>>(
      part, part_channels, bias, static_cast<const float4*>(lo4), static_cast<float4*>(out4), B, H, W);
  return check_launch("fd_pyramid_gather");
}

extern "C" int fd_conv_in(const void* in4, const float* w, const float* bias, void* out, int B, int H,
                          int W, cudaStream_t stream) {
  const size_t ntiles = static_cast<size_t>(B) * H * ((W + 63) / 64);
  FD_REQUIRE(ntiles < (1u << 31), "fd_conv_in: too many tiles");
  const int grid = static_cast<int>(ntiles < 148 * 8 ? ntiles : 148 * 8);
  conv_in_kernel<bf16><<<grid, 256, 0, stream>>>(static_cast<const float4*>(in4), w, bias,
                                                 static_cast<bf16*>(out), B, H, W);
  return check_launch("fd_conv_in");
}

extern "C" int fd_conv_in_f32(const void* in4, const float* w, const float* bias, void* out, int B, int H,
                              int W, cudaStream_t stream) {
  const size_t ntiles = static_cast<size_t>(B) * H * ((W + 63) / 64);
  FD_REQUIRE(ntiles < (1u << 31), "fd_conv_in_f32: too many tiles");
  const int grid = static_cast<int>(ntiles < 148 * 8 ? ntiles : 148 * 8);
  conv_in_kernel<float><<<grid, 256, 0, stream>>>(static_cast<const float4*>(in4), w, bias,
                                                  static_cast<float*>(out), B, H, W);
  return check_launch("fd_conv_in_f32");
}

extern "C" int fd_combine(const void* pyr4, const float* w, const float* bias, const void* h, void* out,
                          size_t npix, int C, cudaStream_t stream) {
  FD_REQUIRE(C % 8 == 0, "fd_combine: C=%d", C);
  FD_REQUIRE(C <= 2048 && npix < (1u << 31), "fd_combine: C=%d / npix out of range", C);
  const int ppb = 256 / (C / 8);
  combine_kernel<bf16><<<grid_for((npix + ppb - 1) / ppb, 1), 256, 0, stream>>>(
      static_cast<const float4*>(pyr4), w, bias, static_cast<const bf16*>(h), static_cast<bf16*>(out),
      static_cast<int>(npix), C);
  return check_launch("fd_combine");
}

extern "C" int fd_combine_f32(const void* pyr4, const float* w, const float* bias, const void* h, void* out,
                              size_t npix, int C, cudaStream_t stream) {
  FD_REQUIRE(C % 8 == 0 && C <= 2048 && npix < (1u << 31), "fd_combine_f32: C=%d / npix out of range", C);
  const int ppb = 256 / (C / 8);
  combine_kernel<float><<<grid_for((npix + ppb - 1) / ppb, 1), 256, 0, stream>>>(
      static_cast<const float4*>(pyr4), w, bias, static_cast<const float*>(h), static_cast<float*>(out),
      static_cast<int>(npix), C);
  return check_launch("fd_combine_f32");
}

extern "C" int fd_output_axpy(const void* pyr4, const float* w_out_host8, const void* base1, float c1,
                              const void* base2, float c2, const void* base3, float c3, float coef,
                              void* out, void* v_out, size_t npix, cudaStream_t stream) {
  const float* w = w_out_host8;  // host pointer: 2x4 weights, passed by value into the launch
  output_axpy_kernel<<<grid_for(npix, 256), 256, 0, stream>>>(
      static_cast<const float4*>(pyr4), w[0], w[1], w[2], w[3], w[4], w[5], w[6], w[7],
      static_cast<const float2*>(base1), c1, static_cast<const float2*>(base2), c2,
      static_cast<const float2*>(base3), c3, coef, static_cast<float2*>(out), static_cast<float2*>(v_out),
      npix);
  return check_launch("fd_output_axpy");
}

extern "C" int fd_x0(const void* Y, const double* sigma, const void* eps, float fac, void* out, int B,
                     int Fq, int T, cudaStream_t stream) {
  const size_t n = static_cast<size_t>(B) * Fq * T;
  x0_kernel<<<grid_for(n, 256), 256, 0, stream>>>(static_cast<const float2*>(Y), sigma,
                                                  static_cast<const float2*>(eps), fac,
                                                  static_cast<float2*>(out), Fq, T, n);
  return check_launch("fd_x0");
}

extern "C" int fd_fourier_embed(float t, const float* Wf, int nf, float* out, cudaStream_t stream) {
  fourier_embed_kernel<<<(nf + 127) / 128, 128, 0, stream>>>(t, Wf, nf, out);
  return check_launch("fd_fourier_embed");
}

extern "C" int fd_matvec(const float* in, int K, int silu_in, const float* Wm, const float* b,
                         const float* add, float out_scale, float* out, int M, cudaStream_t stream) {
  matvec_kernel<<<(M + 7) / 8, 256, 0, stream>>>(in, K, silu_in, Wm, b, add, out_scale, out, M);
  return check_launch("fd_matvec");
}
