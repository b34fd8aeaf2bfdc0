// flowdec_b200 — upstream NDAC (DAC) decoder on tcgen05 tensor cores.
//
// Replaces the 1-D convolutions of `dac.DAC.decode` (descript-audio-codec 1.0.0, call site
// /root/reference/demo.ipynb:105): WNConv1d k=7 dilated / k=1, WNConvTranspose1d(k = 2s, stride s) and the
// Snake1d activations between them.  The fp32 CUDA-core kernels of fd_dac.cu ran the decoder at ~9 TFLOP/s
// (683 ms for 32 x 2 s, 57 % of a codes -> waveform -> enhance pipeline); here every layer is one implicit GEMM
//
//   D[t, n] = sum_tap sum_c  X[t + off_tap, c] * Wp[n, tap*Cin + c]          (kind::tf32, fp32 accumulation)
//
// on activations stored time-major, fp32 [B, T, C]:
//   * the A operand of ALL taps of a k-slice is ONE TMA box of 128 + (max_off - min_off) time steps x 32 channels;
//     tap j is the same box addressed from row off_j - min_off (the 128B swizzle is address-based, so any
//     128-byte-aligned start row works: tools/umma_probe.py).  Out-of-range time steps arrive as TMA zero fill =
//     the conv's zero padding.
//   * ConvTranspose1d(k = 2s, stride s, pad p) is the 2-tap conv  Y[q, r*Cout + co] = W[:,co,r] . x[q] +
//     W[:,co,r+s] . x[q-1]  with s*Cout output columns; row-major [q, r, co] IS the up-sampled signal
//     [q*s + r, co], so the result is written once and consumed through a strided view (offset p).
//   * Snake1d of the NEXT layer (x + sin^2(a x)/(a + 1e-9)) is applied in the epilogue, which writes the raw
//     tensor (when a later residual needs it) and/or the activated tensor; the residual add of a ResidualUnit
//     is an fp32 add in the epilogue (exact, not a K segment, so the residual stream is never rounded).
// warps: 0 A producer, 1 MMA issuer, 2-5 epilogue, 6 weight producer.  One persistent CTA per SM.
#include "fd_common.cuh"

#include <algorithm>
#include <cstring>

namespace fd {

constexpr int kDtTile = 128;                 // time steps per tile (UMMA M)
constexpr int kDtMaxRows = 184;              // 128 + 6*9 (k = 7, dilation 9) rounded up to a multiple of 8
constexpr int kDtAStage = kDtMaxRows * 128;  // 23552 bytes, multiple of 1024
constexpr int kDtMaxN = 192;                 // output columns per tile
constexpr int kDtBStage = kDtMaxN * 128;     // 24576
constexpr int kDtSA = 3, kDtSB = 4;
constexpr int kDtOutStage = kDtTile * 128;   // one 128 x 32-float staging tile
constexpr int kDtSmem = 1024 + kDtSA * kDtAStage + kDtSB * kDtBStage + 2 * kDtOutStage + 3 * kDtMaxN * 4 + 256;
constexpr int kDtMaxTaps = 8;

struct DacTcParams {
  CUtensorMap a_map, b_map, raw_map, act_map;
  int has_raw, has_act;
  int ntaps, tap_row[kDtMaxTaps];
  int box_rows, box_t0;
  int kslices, Cin;
  int N, n_tiles, t_tiles, B;
  int T_out, Ntot;
  const float* bias;
  const float* alpha;
  int alpha_mod;
  const float* residual;
  long long res_bstride;
};

__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}

__device__ __forceinline__ void tma_store_3d(const CUtensorMap* m, const void* smem_src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}

// Snake1d: x + sin^2(a x) / (a + 1e-9) = x + (1 - cos(2 a x)) * 0.5 / (a + 1e-9).  cos through one MUFU op after
// an explicit reduction of the argument to [-pi, pi] (2 FMAs; cos.approx alone loses accuracy for large arguments):
// absolute error ~1e-6 of the sin^2 term, three orders below the tf32 rounding of the stored activation.  sinf()
// in the epilogue cost ~40 instructions per element and made the 48 kHz layers epilogue-bound (profiles/r2_ndac_*).
__device__ __forceinline__ float snake_fast(float x, float two_alpha, float half_inv) {
  const float z = two_alpha * x;
  const float k = rintf(z * 0.15915494309189535f);                 // z / 2 pi
  float r = fmaf(k, -6.2831854820251465f, z);                      // z - k * fl(2 pi)
  r = fmaf(k, 1.7484555e-7f, r);                                   // ... - k * (2 pi - fl(2 pi))
  return fmaf(1.0f - __cosf(r), half_inv, x);
}

__global__ void __launch_bounds__(224, 1) dac_conv_tc_kernel(const __grid_constant__ DacTcParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* sA = smem;
  uint8_t* sB = sA + kDtSA * kDtAStage;
  uint8_t* sOut = sB + kDtSB * kDtBStage;            // [0]: raw staging, [1]: activated staging
  float* sBias = reinterpret_cast<float*>(sOut + 2 * kDtOutStage);
  float* sAlpha = sBias + kDtMaxN;
  float* sInv = sAlpha + kDtMaxN;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sInv + kDtMaxN);
  uint64_t* fullA = bars;
  uint64_t* emptyA = bars + kDtSA;
  uint64_t* fullB = bars + 2 * kDtSA;
  uint64_t* emptyB = fullB + kDtSB;
  uint64_t* tfull = emptyB + kDtSB;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < kDtSA; ++s) { mbar_init(&fullA[s], 1); mbar_init(&emptyA[s], 1); }
    for (int s = 0; s < kDtSB; ++s) { mbar_init(&fullB[s], 1); mbar_init(&emptyB[s], 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(&tfull[a], 1); mbar_init(&tempty[a], 128); }
    fence_mbar_init();
    tma_prefetch_desc(&p.a_map);
    tma_prefetch_desc(&p.b_map);
    if (p.has_raw) tma_prefetch_desc(&p.raw_map);
    if (p.has_act) tma_prefetch_desc(&p.act_map);
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  const int num_tiles = p.B * p.t_tiles * p.n_tiles;
  const int a_bytes = p.box_rows * 128, b_bytes = p.N * 128;

  if (warp == 0) {
    if (lane == 0) {
      int sa = 0;
      uint32_t pa = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int tt = (tile / p.n_tiles) % p.t_tiles, b = tile / (p.n_tiles * p.t_tiles);
        for (int ks = 0; ks < p.kslices; ++ks) {
          mbar_wait(&emptyA[sa], pa ^ 1u);
          mbar_expect_tx(&fullA[sa], a_bytes);
          tma_load_3d(sA + sa * kDtAStage, &p.a_map, &fullA[sa], ks * 32, tt * kDtTile + p.box_t0, b);
          if (++sa == kDtSA) { sa = 0; pa ^= 1u; }
        }
      }
    }
  } else if (warp == 6) {
    if (lane == 0) {
      int sb = 0;
      uint32_t pb = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int nt = tile % p.n_tiles;
        for (int ks = 0; ks < p.kslices; ++ks) {
          for (int tap = 0; tap < p.ntaps; ++tap) {
            mbar_wait(&emptyB[sb], pb ^ 1u);
            mbar_expect_tx(&fullB[sb], b_bytes);
            tma_load_2d(sB + sb * kDtBStage, &p.b_map, &fullB[sb], tap * p.Cin + ks * 32, nt * p.N);
            if (++sb == kDtSB) { sb = 0; pb ^= 1u; }
          }
        }
      }
    }
  } else if (warp == 1) {
    const uint32_t idesc = umma_idesc_tf32(kDtTile, p.N);
    int sa = 0, sb = 0, acc = 0;
    uint32_t pa = 0, pb = 0, acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      mbar_wait(&tempty[acc], acc_phase ^ 1u);
      tc_fence_after_sync();
      const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc * 256);
      uint32_t first = 1;
      for (int ks = 0; ks < p.kslices; ++ks) {
        mbar_wait(&fullA[sa], pa);
        tc_fence_after_sync();
        const uint32_t a_base = smem_u32(sA + sa * kDtAStage);
        for (int tap = 0; tap < p.ntaps; ++tap) {
          mbar_wait(&fullB[sb], pb);
          tc_fence_after_sync();
          const uint64_t da = umma_desc_k_sw128(a_base + static_cast<uint32_t>(p.tap_row[tap]) * 128u);
          const uint64_t db = umma_desc_k_sw128(smem_u32(sB + sb * kDtBStage));
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_tf32(d_tmem, da + static_cast<uint64_t>(k * 2), db + static_cast<uint64_t>(k * 2), idesc,
                        (first && k == 0) ? 0u : 1u);
            umma_commit(&emptyB[sb]);
            if (tap == p.ntaps - 1) {
              umma_commit(&emptyA[sa]);
              if (ks == p.kslices - 1) umma_commit(&tfull[acc]);
            }
          }
          __syncwarp();
          first = 0;
          if (++sb == kDtSB) { sb = 0; pb ^= 1u; }
        }
        if (++sa == kDtSA) { sa = 0; pa ^= 1u; }
      }
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1u;
    }
  } else if (warp < 6) {
    const int ew = warp & 3;
    const int row = ew * 32 + lane;
    const int et = threadIdx.x - 64;             // 0..127 among the epilogue threads
    const bool leader = (et == 0);
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int nt = tile % p.n_tiles, tt = (tile / p.n_tiles) % p.t_tiles, b = tile / (p.n_tiles * p.t_tiles);
      const int n0 = nt * p.N, t = tt * kDtTile + row;
      // per-tile column parameters (the previous tile's readers are past their last use: same threads, in order)
      named_bar_sync(3, 128);
      for (int c = et; c < p.N; c += 128) {
        sBias[c] = p.bias ? p.bias[n0 + c] : 0.f;
        if (p.has_act) {
          const float a = p.alpha[(n0 + c) % p.alpha_mod];
          sAlpha[c] = 2.0f * a;                      // snake_fast takes 2 alpha and 0.5 / (alpha + 1e-9)
          sInv[c] = 0.5f / (a + 1e-9f);
        }
      }
      named_bar_sync(3, 128);
      const float* res_row = (p.residual != nullptr && t < p.T_out)
                                 ? p.residual + static_cast<size_t>(b) * p.res_bstride + static_cast<size_t>(t) * p.Ntot + n0
                                 : nullptr;
      const int chunks = p.N / 32;
      // the residual of chunk 0 is requested before waiting for the accumulator (its latency hides behind the MMAs);
      // later chunks request theirs one chunk ahead
      float4 rr[8];
      if (res_row != nullptr) {
#pragma unroll
        for (int i = 0; i < 8; ++i) rr[i] = *reinterpret_cast<const float4*>(res_row + i * 4);
      }
      mbar_wait(&tfull[acc], acc_phase);
      tc_fence_after_sync();
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(ew * 32) << 16) + static_cast<uint32_t>(acc * 256);
#pragma unroll 1
      for (int ch = 0; ch < chunks; ++ch) {
        uint32_t v[32];
        tmem_ld_32x32b_x32(t_row + ch * 32, v);
        tmem_ld_wait();
        if (ch == chunks - 1) {
          tc_fence_before_sync();
          mbar_arrive(&tempty[acc]);
        }
        float f[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) f[i] = __uint_as_float(v[i]) + sBias[ch * 32 + i];
        if (res_row != nullptr) {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            f[4 * i] += rr[i].x; f[4 * i + 1] += rr[i].y; f[4 * i + 2] += rr[i].z; f[4 * i + 3] += rr[i].w;
          }
          if (ch + 1 < chunks) {
#pragma unroll
            for (int i = 0; i < 8; ++i) rr[i] = *reinterpret_cast<const float4*>(res_row + (ch + 1) * 32 + i * 4);
          }
        }
        // the TMA stores that read the staging tiles of the previous chunk must have drained them
        if (leader) tma_store_wait_read<0>();
        named_bar_sync(1, 128);
        const uint32_t rowp = smem_u32(sOut + row * 128);
        if (p.has_raw) {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            uint4 q;
            q.x = __float_as_uint(f[4 * j]); q.y = __float_as_uint(f[4 * j + 1]);
            q.z = __float_as_uint(f[4 * j + 2]); q.w = __float_as_uint(f[4 * j + 3]);
            sts128(rowp + static_cast<uint32_t>((j ^ (row & 7)) << 4), q);
          }
        }
        if (p.has_act) {
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            // the activated tensor only ever feeds a tf32 MMA: round to nearest here (the pipe would truncate)
            f[i] = round_tf32(snake_fast(f[i], sAlpha[ch * 32 + i], sInv[ch * 32 + i]));
          }
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            uint4 q;
            q.x = __float_as_uint(f[4 * j]); q.y = __float_as_uint(f[4 * j + 1]);
            q.z = __float_as_uint(f[4 * j + 2]); q.w = __float_as_uint(f[4 * j + 3]);
            sts128(rowp + kDtOutStage + static_cast<uint32_t>((j ^ (row & 7)) << 4), q);
          }
        }
        fence_proxy_async_smem();
        named_bar_sync(2, 128);
        if (leader) {
          if (p.has_raw) tma_store_3d(&p.raw_map, sOut, n0 + ch * 32, tt * kDtTile, b);
          if (p.has_act) tma_store_3d(&p.act_map, sOut + kDtOutStage, n0 + ch * 32, tt * kDtTile, b);
          tma_store_commit();
        }
      }
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1u;
    }
    if (leader) tma_store_wait_all<0>();
  }
  __syncwarp();
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

// ------------------------------------------------------------------------------------------------
// final layer: Conv1d(C -> 1, k = 7, pad 3) + tanh on the Snake-activated tensor, fp32 CUDA cores (0.03 % of the
// decoder's FLOPs, HBM-bound): x [B, T, C] (time-major) -> out [B, T].  Block = 128 outputs; the 134 x C input
// window is staged in shared memory.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) dac_final_conv_kernel(const float* __restrict__ x, long long x_bstride,
                                                             const float* __restrict__ w,   // [C][7]
                                                             const float* __restrict__ bias, float* __restrict__ out,
                                                             int T, int C) {
  // row pitch C + 4 floats: rows stay 16-byte aligned and, for C % 32 == 0, lane i of a quarter-warp hits banks
  // 4i .. 4i+3 — conflict-free LDS.128, four channels per load (the C + 1 pitch of the first version needed two
  // shared-memory loads per FMA and made this 0.03 %-of-the-FLOPs layer 6 % of the decoder's time)
  extern __shared__ float sm[];
  const int P = C + 4;
  float* sx = sm;                        // [134][P]
  float* sw = sm + 134 * P;              // [7][C]
  const int b = blockIdx.y, t0 = blockIdx.x * 128;
  const float* xb = x + static_cast<size_t>(b) * x_bstride;
  const int C4 = C >> 2;
  for (int i = threadIdx.x; i < 134 * C4; i += 128) {
    const int r = i / C4, c = (i - r * C4) * 4;
    const int t = t0 - 3 + r;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (t >= 0 && t < T) v = *reinterpret_cast<const float4*>(xb + static_cast<size_t>(t) * C + c);
    *reinterpret_cast<float4*>(sx + r * P + c) = v;
  }
  for (int i = threadIdx.x; i < 7 * C; i += 128) {
    const int k = i / C, c = i - k * C;
    sw[i] = w[c * 7 + k];
  }
  __syncthreads();
  const int t = t0 + threadIdx.x;
  if (t >= T) return;
  float a0 = bias ? bias[0] : 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
  for (int k = 0; k < 7; ++k) {
    const float* xr = sx + (threadIdx.x + k) * P;
    const float* wr = sw + k * C;
    for (int c = 0; c < C; c += 4) {
      const float4 xv = *reinterpret_cast<const float4*>(xr + c);
      const float4 wv = *reinterpret_cast<const float4*>(wr + c);
      a0 = fmaf(xv.x, wv.x, a0);
      a1 = fmaf(xv.y, wv.y, a1);
      a2 = fmaf(xv.z, wv.z, a2);
      a3 = fmaf(xv.w, wv.w, a3);
    }
  }
  out[static_cast<size_t>(b) * T + t] = tanhf((a0 + a1) + (a2 + a3));
}

// first encoder layer: Conv1d(1 -> C, k = 7, pad 3) on the waveform x [B, T] -> raw [B, T, C] and/or the Snake-activated
// tensor (alpha of the first ResidualUnit), fp32 CUDA cores (Cin = 1: nothing for a tensor core to contract)
__global__ void __launch_bounds__(256) dac_first_conv_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                             const float* __restrict__ bias, const float* __restrict__ alpha,
                                                             float* __restrict__ raw, float* __restrict__ act, int B, int T,
                                                             int C) {
  const size_t total = static_cast<size_t>(B) * T * C;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % C);
    const size_t bt = i / C;
    const int t = static_cast<int>(bt % T);
    const float* xb = x + (bt - t);
    float a = bias ? bias[c] : 0.f;
#pragma unroll
    for (int k = 0; k < 7; ++k) {
      const int tt = t + k - 3;
      if (tt >= 0 && tt < T) a = fmaf(w[c * 7 + k], xb[tt], a);
    }
    if (raw) raw[i] = a;
    if (act) {
      const float al = alpha[c];
      act[i] = round_tf32(snake_fast(a, 2.0f * al, 0.5f / (al + 1e-9f)));
    }
  }
}

// [B, C, T] -> [B, T, C] (the decoder's input latent arrives channel-major from `from_codes`, as upstream; the encoder's
// latent goes back the same way with C and T swapped).  round_out: round to tf32 (tensors that feed a tf32 MMA)
__global__ void __launch_bounds__(256) nct_to_ntc_kernel(const float* __restrict__ in, float* __restrict__ out, int C, int T,
                                                         int round_out) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z, c0 = blockIdx.y * 32, t0 = blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int j = ty; j < 32; j += 8) {
    const int c = c0 + j, t = t0 + tx;
    tile[j][tx] = (c < C && t < T) ? in[(static_cast<size_t>(b) * C + c) * T + t] : 0.f;
  }
  __syncthreads();
  for (int j = ty; j < 32; j += 8) {
    const int t = t0 + j, c = c0 + tx;
    if (t < T && c < C) out[(static_cast<size_t>(b) * T + t) * C + c] = round_out ? round_tf32(tile[tx][j]) : tile[tx][j];
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn dt_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
  }
  return fn;
}

// fp32 [B, T, C] with batch stride `bstride` elements; box = 32 channels x `rows` time steps
static int dt_make_ntc_map(CUtensorMap* m, const void* base, int B, int T, int C, long long bstride, int rows) {
  EncodeTiledFn enc = dt_encode_fn();
  FD_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled is not available from the driver");
  FD_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0 && C % 4 == 0 && bstride % 4 == 0,
             "fd_dac_conv_tc: tensors must be 16-byte aligned with channel counts / strides that are multiples of 4");
  cuuint64_t dims[3] = {static_cast<cuuint64_t>(C), static_cast<cuuint64_t>(T), static_cast<cuuint64_t>(B)};
  cuuint64_t strides[2] = {static_cast<cuuint64_t>(C) * 4, static_cast<cuuint64_t>(bstride) * 4};
  cuuint32_t box[3] = {32, static_cast<cuuint32_t>(rows), 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  FD_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(NTC) failed with CUresult %d", (int)r);
  return 0;
}

int device_sm_count();
int current_device();

}  // namespace fd

using namespace fd;

extern "C" int fd_dac_conv_tc(const float* x, int B, int Tin, long long x_bstride, int Cin, const float* wpacked,
                              int Ntot, int ntaps, const int* tap_offsets, const float* bias, const float* residual,
                              long long res_bstride, const float* alpha_next, int alpha_mod, float* out_raw,
                              float* out_act, int Tout_rows, long long out_bstride, cudaStream_t stream) {
  FD_REQUIRE(x != nullptr && wpacked != nullptr && (out_raw != nullptr || out_act != nullptr), "fd_dac_conv_tc: NULL pointer");
  FD_REQUIRE(B >= 1 && Tin >= 1 && Tout_rows >= 1, "fd_dac_conv_tc: bad shape B=%d Tin=%d Tout=%d", B, Tin, Tout_rows);
  FD_REQUIRE(Cin % 32 == 0 && Cin >= 32, "fd_dac_conv_tc: Cin=%d must be a multiple of 32", Cin);
  FD_REQUIRE(ntaps >= 1 && ntaps <= kDtMaxTaps, "fd_dac_conv_tc: ntaps=%d out of range [1,%d]", ntaps, kDtMaxTaps);
  FD_REQUIRE(out_act == nullptr || (alpha_next != nullptr && alpha_mod >= 1), "fd_dac_conv_tc: activated output needs alpha");
  int N = 0;
  for (int c = kDtMaxN; c >= 32; c -= 32)
    if (Ntot % c == 0) { N = c; break; }
  FD_REQUIRE(N > 0, "fd_dac_conv_tc: Ntot=%d must be a multiple of 32", Ntot);
  int lo = tap_offsets[0], hi = tap_offsets[0];
  for (int i = 1; i < ntaps; ++i) { lo = std::min(lo, tap_offsets[i]); hi = std::max(hi, tap_offsets[i]); }
  DacTcParams p;
  memset(&p, 0, sizeof(p));
  p.box_rows = kDtTile + (hi - lo);
  FD_REQUIRE(p.box_rows <= kDtMaxRows, "fd_dac_conv_tc: tap span %d exceeds %d", hi - lo, kDtMaxRows - kDtTile);
  p.box_t0 = lo;
  p.ntaps = ntaps;
  for (int i = 0; i < ntaps; ++i) p.tap_row[i] = tap_offsets[i] - lo;
  p.kslices = Cin / 32;
  p.Cin = Cin;
  p.N = N;
  p.n_tiles = Ntot / N;
  p.t_tiles = (Tout_rows + kDtTile - 1) / kDtTile;
  p.B = B;
  p.T_out = Tout_rows;
  p.Ntot = Ntot;
  p.bias = bias;
  p.alpha = alpha_next;
  p.alpha_mod = alpha_mod > 0 ? alpha_mod : 1;
  p.residual = residual;
  p.res_bstride = res_bstride;
  p.has_raw = out_raw != nullptr;
  p.has_act = out_act != nullptr;
  if (dt_make_ntc_map(&p.a_map, x, B, Tin, Cin, x_bstride, p.box_rows)) return 1;
  {
    EncodeTiledFn enc = dt_encode_fn();
    const int ktot = ntaps * Cin;
    cuuint64_t dims[2] = {static_cast<cuuint64_t>(ktot), static_cast<cuuint64_t>(Ntot)};
    cuuint64_t strides[1] = {static_cast<cuuint64_t>(ktot) * 4};
    cuuint32_t box[2] = {32, static_cast<cuuint32_t>(N)};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(&p.b_map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(wpacked), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    FD_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(dac weights) failed with CUresult %d", (int)r);
  }
  if (p.has_raw && dt_make_ntc_map(&p.raw_map, out_raw, B, Tout_rows, Ntot, out_bstride, kDtTile)) return 1;
  if (p.has_act && dt_make_ntc_map(&p.act_map, out_act, B, Tout_rows, Ntot, out_bstride, kDtTile)) return 1;
  static bool attr_set[kMaxDevices] = {false};
  const int dev = current_device();
  if (!attr_set[dev]) {
    cudaError_t e = cudaFuncSetAttribute(dac_conv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kDtSmem);
    FD_REQUIRE(e == cudaSuccess, "cudaFuncSetAttribute(smem=%d) failed: %s", kDtSmem, cudaGetErrorString(e));
    attr_set[dev] = true;
  }
  const int tiles = B * p.t_tiles * p.n_tiles;
  const int grid = std::min(tiles, device_sm_count());
  dac_conv_tc_kernel<<<grid, 224, kDtSmem, stream>>>(p);
  return check_launch("fd_dac_conv_tc");
}

extern "C" int fd_dac_final_conv(const float* x_act, long long x_bstride, const float* w, const float* bias, float* out,
                                 int B, int T, int C, cudaStream_t stream) {
  const size_t smem = (static_cast<size_t>(134) * (C + 4) + 7 * C) * sizeof(float);
  FD_REQUIRE(C >= 4 && C % 4 == 0 && x_bstride % 4 == 0 && smem <= 200 * 1024,
             "fd_dac_final_conv: C=%d (multiple of 4) needs %zu bytes of shared memory", C, smem);
  static bool attr_set[kMaxDevices] = {false};
  const int dev = current_device();
  if (!attr_set[dev]) {
    cudaFuncSetAttribute(dac_final_conv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    attr_set[dev] = true;
  }
  dac_final_conv_kernel<<<dim3((T + 127) / 128, B), 128, smem, stream>>>(x_act, x_bstride, w, bias, out, T, C);
  return check_launch("fd_dac_final_conv");
}

extern "C" int fd_dac_nct_to_ntc(const float* in, float* out, int B, int C, int T, int round_tf32_out,
                                 cudaStream_t stream) {
  FD_REQUIRE(B >= 1 && B <= 65535, "fd_dac_nct_to_ntc: B=%d", B);
  nct_to_ntc_kernel<<<dim3((T + 31) / 32, (C + 31) / 32, B), 256, 0, stream>>>(in, out, C, T, round_tf32_out);
  return check_launch("fd_dac_nct_to_ntc");
}

extern "C" int fd_dac_first_conv(const float* x, const float* w, const float* bias, const float* alpha, float* raw,
                                 float* act, int B, int T, int C, cudaStream_t stream) {
  FD_REQUIRE(raw != nullptr || act != nullptr, "fd_dac_first_conv: no output");
  FD_REQUIRE(act == nullptr || alpha != nullptr, "fd_dac_first_conv: activated output needs alpha");
  const size_t total = static_cast<size_t>(B) * T * C;
  size_t g = (total + 255) / 256;
  if (g > 148 * 32) g = 148 * 32;
  dac_first_conv_kernel<<<static_cast<int>(g), 256, 0, stream>>>(x, w, bias, alpha, raw, act, B, T, C);
  return check_launch("fd_dac_first_conv");
}
