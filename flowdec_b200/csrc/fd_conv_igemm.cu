// flowdec_b200 — 3x3 / 1x1 convolution as a tcgen05 implicit GEMM (sm_100a).
//
// Replaces, for the NCSN++ backbone of FlowDec, what the reference runs as
// `nn.Conv2d` -> cuDNN (reference: flowdec/backbones/ncsnpp_utils/layers.py:110-134,
// used from layerspp.py:235,243,245 and ncsnpp.py:162,218,230).  64 convolutions per
// backbone forward carry 99.9 % of the arithmetic.
//
// Formulation
//   activations  NHWC bf16   [B, H, W, C]           (H = frequency, W = time)
//   weights      bf16        [Npad, Ktot]  K-major  (packed once at load time)
//   GEMM         D[pixel, cout] = sum_k A[pixel, k] * Wp[cout, k]
//   k runs over "segments": each segment is one source tensor read through 1 or 9
//   taps.  A 3x3 conv is one 9-tap segment; a res-block's second conv plus its
//   1x1 (or identity) skip is a 9-tap segment followed by 1-tap segments, so the
//   residual add happens inside the TMEM accumulator (layerspp.py:278-284).
//
// Kernel structure (one persistent CTA per SM, 192 threads)
//   warp 0      TMA producer: per k-step one 4-D box load of A (128 pixels x 64
//               channels, shifted by the tap; out-of-image coordinates are
//               zero-filled by TMA = the conv's zero padding) and one 2-D box load of
//               the weight slice, into a STAGES-deep 128B-swizzled smem ring.
//   warp 1      MMA issuer: tcgen05.mma.cta_group::1.kind::f16, M=128, N=Npad, K=16,
//               fp32 accumulators in TMEM, double-buffered (2 x Npad columns).
//   warps 2..5  epilogue: tcgen05.ld -> +bias -> bf16 -> swizzled smem -> TMA store
//               (or fp32 direct stores for the 4-channel pyramid convs).
#include <type_traits>

#include "fd_common.cuh"

#include <algorithm>
#include <cstdlib>
#include <cstring>

namespace fd {

constexpr int kMaxSeg = 4;
constexpr int kTileM = 128;   // pixels per tile (= UMMA M)
constexpr int kSliceK = 64;   // channels per k-step (= one 128-byte swizzle span of bf16)
constexpr int kABytes = kTileM * kSliceK * 2;

struct ConvParams {
  CUtensorMap a_map[kMaxSeg];
  CUtensorMap b_map;
  CUtensorMap out_map;
  int nseg;
  int seg_kslices[kMaxSeg];
  int seg_taps[kMaxSeg];
  int B, H, W;
  int bh, bw;            // tile = bh x bw pixels, bh*bw == 128
  int tiles_h, tiles_w;
  int num_tiles;
  const float* bias;     // [Npad] or nullptr
  float* out_f32;        // fp32 NHWC [B,H,W,cout_valid] (OUT_F32 kernels only)
  float* stats;          // optional GroupNorm partials [B, 4*tiles_per_img, N, 2] (bf16-output kernels)
  int cout_valid;
};

template <int N, bool CTA2 = false>
struct ConvCfg {
  // B rows held by one CTA: the pair splits the N weight rows between its two shared memories
  static constexpr int kBRows = CTA2 ? N / 2 : N;
  static constexpr int kBBytes = kBRows * kSliceK * 2;
  // the ring is one stage short of what 227 KB would hold: the ~30 KB left let zero-smem blocks of
  // the HBM-bound GroupNorm / FIR kernels (another micro-batch, another stream) co-reside on the SM
  static constexpr int kStages = (N == 256) ? (CTA2 ? 5 : 3) : (N == 128 ? (CTA2 ? 6 : 5) : 8);
  static_assert(kBBytes % 1024 == 0, "B stage must keep the 1024-byte swizzle alignment");
  static constexpr int kOutBytes = (N >= 64) ? 2 * kTileM * 128 : 0;  // two 64-channel staging tiles
  static constexpr int kTmemCols = (2 * N <= 32) ? 32 : (2 * N <= 64 ? 64 : (2 * N <= 128 ? 128 : (2 * N <= 256 ? 256 : 512)));
  static constexpr int kStatBytes = (N >= 64) ? 4 * 64 * 2 * 4 : 0;   // per-warp column sums of one chunk
  static constexpr int kSmemBytes =
      1024 + kStages * (kABytes + kBBytes) + kOutBytes + N * 4 + 256 + kStatBytes;
};

// One 32-column half of an epilogue chunk for this lane's pixel row: + bias, bf16 pack into the
// 128B-swizzled staging tile (16-byte chunks j0..j0+3 of the row), and optionally the
// GroupNorm partial sums of the 32 columns over the warp's 32 rows.  The column reduction is a
// transpose-reduce butterfly (31 shuffles per statistic instead of 32*5): after the xor-16 step
// every lane keeps 16 columns, ... after xor-1 lane L holds the total of column L.
// Column statistics of the warp's 32 x 32 block.  `scratch` (shared-window address of a 32 x 36-float tile owned by
// this warp, or 0): transpose through shared memory — 8 conflict-free 16-byte stores + 32 conflict-free loads +
// 64 adds/FMAs per lane — instead of the register butterfly (62 shuffles + 124 selects + 62 adds).
#ifndef FD_EPI_STATS_SMEM
#define FD_EPI_STATS_SMEM 1
#endif
__device__ __forceinline__ void epilogue_half(const uint32_t (&v)[32], uint32_t bs_addr, uint32_t rowp_addr,
                                              int row, int j0, float* stat_dst, int lane, uint32_t scratch = 0) {
  float f[32];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float4 b4 = lds_f4(bs_addr + i * 16);      // bias (warp-wide broadcast read)
    const float2 lo = fadd2(make_float2(__uint_as_float(v[4 * i + 0]), __uint_as_float(v[4 * i + 1])),
                            make_float2(b4.x, b4.y));
    const float2 hi = fadd2(make_float2(__uint_as_float(v[4 * i + 2]), __uint_as_float(v[4 * i + 3])),
                            make_float2(b4.z, b4.w));
    f[4 * i + 0] = lo.x; f[4 * i + 1] = lo.y; f[4 * i + 2] = hi.x; f[4 * i + 3] = hi.y;
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    uint4 q;
    q.x = pack_bf16x2(f[8 * j + 0], f[8 * j + 1]);
    q.y = pack_bf16x2(f[8 * j + 2], f[8 * j + 3]);
    q.z = pack_bf16x2(f[8 * j + 4], f[8 * j + 5]);
    q.w = pack_bf16x2(f[8 * j + 6], f[8 * j + 7]);
    sts128(rowp_addr + static_cast<uint32_t>(((j0 + j) ^ (row & 7)) << 4), q);
  }
  if (stat_dst != nullptr && scratch != 0) {
    __syncwarp();                                    // the previous half's column reads are done
#pragma unroll
    for (int i = 0; i < 8; ++i)
      asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(scratch + static_cast<uint32_t>(lane * 144 + i * 16)),
                   "f"(f[4 * i]), "f"(f[4 * i + 1]), "f"(f[4 * i + 2]), "f"(f[4 * i + 3]) : "memory");
    __syncwarp();
    float2 s2 = make_float2(0.f, 0.f), q2 = make_float2(0.f, 0.f);   // (even rows, odd rows)
#pragma unroll
    for (int r = 0; r < 32; r += 2) {
      float2 ab;
      asm volatile("ld.shared.f32 %0, [%1];" : "=f"(ab.x) : "r"(scratch + static_cast<uint32_t>(r * 144 + lane * 4)));
      asm volatile("ld.shared.f32 %0, [%1];" : "=f"(ab.y) : "r"(scratch + static_cast<uint32_t>((r + 1) * 144 + lane * 4)));
      s2 = fadd2(s2, ab);
      q2 = ffma2(ab, ab, q2);
    }
    *reinterpret_cast<float2*>(stat_dst + lane * 2) = make_float2(s2.x + s2.y, q2.x + q2.y);
  } else if (stat_dst != nullptr) {
    float q[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) q[i] = f[i] * f[i];
#define FD_BFLY(OFF, HALF)                                                      \
    {                                                                           \
      const bool up = (lane & OFF) != 0;                                        \
      _Pragma("unroll") for (int i = 0; i < HALF; ++i) {                        \
        const float ks = up ? f[HALF + i] : f[i];                               \
        const float ss = up ? f[i] : f[HALF + i];                               \
        const float kq = up ? q[HALF + i] : q[i];                               \
        const float sq = up ? q[i] : q[HALF + i];                               \
        f[i] = ks + __shfl_xor_sync(0xffffffffu, ss, OFF);                      \
        q[i] = kq + __shfl_xor_sync(0xffffffffu, sq, OFF);                      \
      }                                                                         \
    }
    FD_BFLY(16, 16)
    FD_BFLY(8, 8)
    FD_BFLY(4, 4)
    FD_BFLY(2, 2)
    FD_BFLY(1, 1)
#undef FD_BFLY
    *reinterpret_cast<float2*>(stat_dst + lane * 2) = make_float2(f[0], q[0]);
  }
}

// GroupNorm partial sums, one slab per TILE: every epilogue warp leaves the (sum, sum of squares) of its 32 rows
// for `cols` columns in shared memory (sStat[warp][col][2]); after the chunk's barrier the 128 epilogue threads add
// the four warps in a fixed order ((w0 + w1) + (w2 + w3): deterministic, independent of the batch) and write
// dst[col][2].  4x fewer statistics bytes than one slab per warp, and the consumer reduces them in one kernel.
__device__ __forceinline__ void stats_combine_store(const float* sStat, int cols, int et, float* dst) {
  if (et < 2 * cols) {
    const int col = et >> 1, st = et & 1;
    const float a = sStat[(0 * 64 + col) * 2 + st], b = sStat[(1 * 64 + col) * 2 + st];
    const float c = sStat[(2 * 64 + col) * 2 + st], d = sStat[(3 * 64 + col) * 2 + st];
    dst[col * 2 + st] = (a + b) + (c + d);
  }
}

// fp32-output form (tf32 kernels): one 32-column chunk = one 128-byte staging row of 32 floats.
__device__ __forceinline__ void epilogue_chunk_f32(const uint32_t (&v)[32], uint32_t bs_addr, uint32_t rowp_addr,
                                                   int row, float* stat_dst, int lane, uint32_t scratch) {
  float f[32];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float4 b4 = lds_f4(bs_addr + i * 16);
    const float2 lo = fadd2(make_float2(__uint_as_float(v[4 * i + 0]), __uint_as_float(v[4 * i + 1])),
                            make_float2(b4.x, b4.y));
    const float2 hi = fadd2(make_float2(__uint_as_float(v[4 * i + 2]), __uint_as_float(v[4 * i + 3])),
                            make_float2(b4.z, b4.w));
    f[4 * i + 0] = lo.x; f[4 * i + 1] = lo.y; f[4 * i + 2] = hi.x; f[4 * i + 3] = hi.y;
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    uint4 q;
    q.x = __float_as_uint(f[4 * j + 0]);
    q.y = __float_as_uint(f[4 * j + 1]);
    q.z = __float_as_uint(f[4 * j + 2]);
    q.w = __float_as_uint(f[4 * j + 3]);
    sts128(rowp_addr + static_cast<uint32_t>((j ^ (row & 7)) << 4), q);
  }
  if (stat_dst != nullptr) {
    __syncwarp();
#pragma unroll
    for (int i = 0; i < 8; ++i)
      asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(scratch + static_cast<uint32_t>(lane * 144 + i * 16)),
                   "f"(f[4 * i]), "f"(f[4 * i + 1]), "f"(f[4 * i + 2]), "f"(f[4 * i + 3]) : "memory");
    __syncwarp();
    float2 s2 = make_float2(0.f, 0.f), q2 = make_float2(0.f, 0.f);
#pragma unroll
    for (int r = 0; r < 32; r += 2) {
      float2 ab;
      asm volatile("ld.shared.f32 %0, [%1];" : "=f"(ab.x) : "r"(scratch + static_cast<uint32_t>(r * 144 + lane * 4)));
      asm volatile("ld.shared.f32 %0, [%1];" : "=f"(ab.y) : "r"(scratch + static_cast<uint32_t>((r + 1) * 144 + lane * 4)));
      s2 = fadd2(s2, ab);
      q2 = ffma2(ab, ab, q2);
    }
    *reinterpret_cast<float2*>(stat_dst + lane * 2) = make_float2(s2.x + s2.y, q2.x + q2.y);
  }
}

template <int N, bool OUT_F32, bool CTA2>
__global__ void __launch_bounds__(192, 1) conv_igemm_kernel(const __grid_constant__ ConvParams p) {
  using Cfg = ConvCfg<N, CTA2>;
  static_assert(!(CTA2 && OUT_F32), "the CTA-pair variant is instantiated for the bf16-output tiles only");
  constexpr int STAGES = Cfg::kStages;
  constexpr int B_BYTES = Cfg::kBBytes;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* sA = smem;
  uint8_t* sB = sA + STAGES * kABytes;
  uint8_t* sOut = sB + STAGES * B_BYTES;
  float* sBias = reinterpret_cast<float*>(sOut + Cfg::kOutBytes);
  uint64_t* bars = reinterpret_cast<uint64_t*>(sBias + N);
  uint64_t* full_bar = bars;                  // [STAGES]
  uint64_t* empty_bar = bars + STAGES;        // [STAGES]
  uint64_t* tfull_bar = bars + 2 * STAGES;    // [2]
  uint64_t* tempty_bar = bars + 2 * STAGES + 2;  // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  for (int i = threadIdx.x; i < N; i += blockDim.x) sBias[i] = p.bias ? p.bias[i] : 0.0f;

  // CTA pair: rank 0 (leader) issues the MMAs; its full / tmem-empty barriers collect the
  // arrivals of both CTAs, the empty / tmem-full barriers of both CTAs are signalled by multicast.
  const uint32_t rank = CTA2 ? cluster_ctarank() : 0u;
  const bool leader_cta = (rank == 0);
  constexpr uint32_t kPair = CTA2 ? 2u : 1u;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], kPair);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tfull_bar[a], 1);
      mbar_init(&tempty_bar[a], 128 * kPair);
    }
    fence_mbar_init();
    for (int s = 0; s < p.nseg; ++s) tma_prefetch_desc(&p.a_map[s]);
    tma_prefetch_desc(&p.b_map);
    if (!OUT_F32) tma_prefetch_desc(&p.out_map);
  }
  if (warp == 1) {
    if (CTA2) tmem_alloc_2sm(tmem_slot, Cfg::kTmemCols);
    else tmem_alloc(tmem_slot, Cfg::kTmemCols);
  }
  tc_fence_before_sync();
  if (CTA2) cluster_sync_all(); else __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  // tile walk: a pair takes two consecutive tiles per step (rank r -> tile 2*step + r)
  const int tile_first = CTA2 ? static_cast<int>((blockIdx.x >> 1) * 2 + rank) : static_cast<int>(blockIdx.x);
  const int tile_stride = static_cast<int>(gridDim.x);

  int k_iters = 0;
  for (int s = 0; s < p.nseg; ++s) k_iters += p.seg_kslices[s] * p.seg_taps[s];
  const int tiles_per_img = p.tiles_h * p.tiles_w;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = tile_first; tile < p.num_tiles; tile += tile_stride) {
        const int n = tile / tiles_per_img;
        const int rem = tile - n * tiles_per_img;
        const int h0 = (rem / p.tiles_w) * p.bh;
        const int w0 = (rem % p.tiles_w) * p.bw;
        int kb = 0;
        for (int s = 0; s < p.nseg; ++s) {
          const int taps = p.seg_taps[s];
          const int ksl = p.seg_kslices[s];
          for (int t = 0; t < taps; ++t) {
            const int dh = (taps == 9) ? (t / 3 - 1) : 0;
            const int dw = (taps == 9) ? (t % 3 - 1) : 0;
            for (int ks = 0; ks < ksl; ++ks) {
              mbar_wait(&empty_bar[stage], phase ^ 1u);
              if (CTA2) {
                if (leader_cta) mbar_expect_tx(&full_bar[stage], 2 * (kABytes + B_BYTES));
                else mbar_arrive_remote(&full_bar[stage], 0);
                tma_load_4d_2sm(sA + stage * kABytes, &p.a_map[s], &full_bar[stage], ks * kSliceK,
                                w0 + dw, h0 + dh, n);
                tma_load_2d_2sm(sB + stage * B_BYTES, &p.b_map, &full_bar[stage], kb * kSliceK,
                                static_cast<int>(rank) * Cfg::kBRows);
              } else {
                mbar_expect_tx(&full_bar[stage], kABytes + B_BYTES);
                tma_load_4d(sA + stage * kABytes, &p.a_map[s], &full_bar[stage], ks * kSliceK,
                            w0 + dw, h0 + dh, n);
                tma_load_2d(sB + stage * B_BYTES, &p.b_map, &full_bar[stage], kb * kSliceK, 0);
              }
              ++kb;
              if (++stage == STAGES) {
                stage = 0;
                phase ^= 1u;
              }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // -------------------------------------------------------------- MMA issuer
    // The whole warp walks the loop (warp-uniform control flow and operands, so descriptors live in
    // uniform registers and every tcgen05.mma is issued without a per-thread -> uniform register
    // round trip); one elected lane issues the MMAs and commits.
    if (leader_cta) {
      constexpr uint32_t idesc = umma_idesc_bf16(CTA2 ? 2 * kTileM : kTileM, N);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      int issued = 0;
      for (int tile = tile_first; tile < p.num_tiles; tile += tile_stride) {
        ++issued;
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1u);
        tc_fence_after_sync();
        const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc * N);
        for (int ki = 0; ki < k_iters; ++ki) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after_sync();
          const uint64_t da = umma_desc_k_sw128(smem_u32(sA + stage * kABytes));
          const uint64_t db = umma_desc_k_sw128(smem_u32(sB + stage * B_BYTES));
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < kSliceK / 16; ++k) {
              // +32 bytes per K=16 step inside the 128-byte swizzle span (encoded >>4)
              if (CTA2)
                umma_bf16_2sm(d_tmem, da + static_cast<uint64_t>(k * 2), db + static_cast<uint64_t>(k * 2),
                              idesc, static_cast<uint32_t>((ki | k) != 0));
              else
                umma_bf16(d_tmem, da + static_cast<uint64_t>(k * 2), db + static_cast<uint64_t>(k * 2),
                          idesc, static_cast<uint32_t>((ki | k) != 0));
            }
            if (CTA2) umma_commit_2sm(&empty_bar[stage]); else umma_commit(&empty_bar[stage]);
            if (ki == k_iters - 1) {
              if (CTA2) umma_commit_2sm(&tfull_bar[acc]); else umma_commit(&tfull_bar[acc]);
            }
          }
          __syncwarp();
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1u;
          }
        }
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1u;
      }
      if (CTA2 && issued > 0) {
        // the peer's epilogue arrives remotely on these barriers: drain them before tear-down
        for (int j = (issued >= 2 ? issued - 2 : issued - 1); j < issued; ++j)
          mbar_wait(&tempty_bar[j & 1], static_cast<uint32_t>((j >> 1) & 1));
      }
    }
  } else {
    // ---------------------------------------------------------------- epilogue
    const int ew = warp & 3;               // TMEM lane quarter this warp may read
    const int row = ew * 32 + lane;        // pixel row of the tile
    const bool leader = (threadIdx.x == 64);
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = tile_first; tile < p.num_tiles; tile += tile_stride) {
      const int n = tile / tiles_per_img;
      const int rem = tile - n * tiles_per_img;
      const int h0 = (rem / p.tiles_w) * p.bh;
      const int w0 = (rem % p.tiles_w) * p.bw;
      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after_sync();
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(ew * 32) << 16) +
                             static_cast<uint32_t>(acc * N);
      if constexpr (OUT_F32) {
        const int hl = row / p.bw;
        const int wl = row - hl * p.bw;
        const size_t pix = (static_cast<size_t>(n) * p.H + (h0 + hl)) * p.W + (w0 + wl);
        float* o = p.out_f32 + pix * p.cout_valid;
#pragma unroll
        for (int g = 0; g < N / 16; ++g) {
          uint32_t v[16];
          tmem_ld_32x32b_x16(t_row + g * 16, v);
          tmem_ld_wait();
          if (g == N / 16 - 1) {
            tc_fence_before_sync();
            mbar_arrive(&tempty_bar[acc]);   // OUT_F32 is single-CTA only
          }
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int c = g * 16 + q * 4;
            if (c + 3 < p.cout_valid) {
              float4 r;
              r.x = __uint_as_float(v[q * 4 + 0]) + sBias[c + 0];
              r.y = __uint_as_float(v[q * 4 + 1]) + sBias[c + 1];
              r.z = __uint_as_float(v[q * 4 + 2]) + sBias[c + 2];
              r.w = __uint_as_float(v[q * 4 + 3]) + sBias[c + 3];
              *reinterpret_cast<float4*>(o + c) = r;
            }
          }
        }
      } else {
        constexpr int kChunks = N / 64;
        // GroupNorm partial-sum slab of this tile (the four warps' sums are combined through shared memory)
        float* stat_row = p.stats ? p.stats + ((static_cast<size_t>(n) * tiles_per_img + rem) * N) * 2 : nullptr;
        float* sStat = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + 256);
        float* my_stat = sStat + ew * 128;
#pragma unroll 1
        for (int ch = 0; ch < kChunks; ++ch) {
          uint32_t v0[32], v1[32];
          tmem_ld_32x32b_x32(t_row + ch * 64, v0);
          tmem_ld_32x32b_x32(t_row + ch * 64 + 32, v1);
          tmem_ld_wait();
          if (ch == kChunks - 1) {
            tc_fence_before_sync();
            if (CTA2) mbar_arrive_remote(&tempty_bar[acc], 0); else mbar_arrive(&tempty_bar[acc]);
          }
          uint8_t* stg = sOut + (ch & 1) * (kTileM * 128);
          // the TMA store that last read this staging tile must have drained it
          if (leader) tma_store_wait_read<1>();
          named_bar_sync(1, 128);
          const uint32_t bs = smem_u32(sBias + ch * 64);
          const uint32_t rowp = smem_u32(stg + row * 128);
          epilogue_half(v0, bs, rowp, row, 0, stat_row ? my_stat : nullptr, lane);
          epilogue_half(v1, bs + 128, rowp, row, 4, stat_row ? my_stat + 64 : nullptr, lane);
          fence_proxy_async_smem();
          named_bar_sync(2, 128);
          if (stat_row) stats_combine_store(sStat, 64, static_cast<int>(threadIdx.x) - 64, stat_row + (ch * 64) * 2);
          if (leader) {
            tma_store_4d(&p.out_map, stg, ch * 64, w0, h0, n);
            tma_store_commit();
          }
        }
      }
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1u;
    }
    if (!OUT_F32 && leader) tma_store_wait_all<0>();
  }

  __syncwarp();
  tc_fence_before_sync();
  if (CTA2) cluster_sync_all(); else __syncthreads();
  if (warp == 1) {
    if (CTA2) tmem_dealloc_2sm(tmem_base, Cfg::kTmemCols);
    else tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

// =================================================================================
// "Halo" variant (CTA pairs only): 16 x 8-pixel tiles, ONE A load per k-slice for all nine taps.
//   The A stage is an 18-row x 10-column x 64-channel box (rows of the smem tile = halo pixels at a
//   dense 10-pixel pitch).  Because the UMMA 128B swizzle is address-based (tools/umma_probe.py), the
//   operand of tap (dh, dw) is just a descriptor into that box: start row (1+dh)*10 + (1+dw), stride
//   between 8-row groups (= output rows) 10 rows = 1280 B.  L2->smem traffic of A drops 6.3x.
//   Optional in-shared-memory transform (XF): the box holds the RAW res-block tensor and four helper
//   warps apply GroupNorm affine + SiLU in place ONCE per k-slice (zeroing out-of-image positions =
//   the conv's zero padding of the ACTIVATED tensor) before the 36 MMAs read it.  This removes the
//   separate GroupNorm+SiLU pass and its bf16 round trip through HBM (layerspp.py:253,274 fused into
//   the operand path of Conv_0 / Conv_1).
//   Weights: a separate ring of [N/2 x 64] tiles, one per (tap, k-slice).
// warps: 0 A producer, 1 MMA issuer (leader CTA), 2-5 epilogue, 6 weight producer, 7-14 transform.
// =================================================================================
constexpr int kHaloRows = 18;
constexpr int kHaloCols = 10;
constexpr int kHaloTileH = 16;
constexpr int kHaloTileW = 8;
constexpr int kHaloPix = kHaloRows * kHaloCols;          // 180 smem rows of 128 B
constexpr int kHaloTxBytes = kHaloPix * 128;             // 23040 bytes land per stage
constexpr int kHaloStageBytes = 23552;                   // stage pitch, multiple of 1024

struct HaloParams {
  CUtensorMap a_map[kMaxSeg];
  CUtensorMap b_map;
  CUtensorMap out_map;
  int nseg;
  int seg_kslices[kMaxSeg];
  int seg_taps[kMaxSeg];
  int seg_kbase[kMaxSeg];        // K offset of the segment inside the packed weight
  int seg_cin[kMaxSeg];          // channels consumed by the segment
  const float* seg_ss[kMaxSeg];  // GroupNorm scale/shift [B][ss_pitch][2] (+ channel offset) or nullptr = raw
  int seg_ss_pitch[kMaxSeg];
  int seg_ss_off[kMaxSeg];       // offset of the segment's channels in the smem scale/shift table
  int B, H, W;
  int tiles_h, tiles_w, num_tiles;
  const float* bias;
  float* stats;
  float* out4;                   // OUT4 kernels: fp32 NHWC [B,H,W,4] written directly by the epilogue
  long long* dbg;                // optional [16] cycle counters written by block 0 (tools/halo_dbg.py)
  int xf_mode;                   // experiments (env FD_HALO_XF_MODE): 0 normal, 1 barriers only, 2 loads only, 3 no stores
};

// Cycle counters and ablation modes of the halo kernel (tools/halo_dbg.py) are compiled in only with
// -DFD_HALO_DEBUG=1 (python tools/ab_bench.py --build "-DFD_HALO_DEBUG=1"); the product build carries none of it.
#ifndef FD_HALO_DEBUG
#define FD_HALO_DEBUG 0
#endif
#define FD_DBG (FD_HALO_DEBUG && p.dbg != nullptr)
#define FD_XF_MODE (FD_HALO_DEBUG ? p.xf_mode : 0)

// TF32: fp32 activations / weights in HBM and shared memory, kind::tf32 MMAs (K = 8 per instruction, 32 channels
// per 128-byte k-slice: the byte geometry of boxes, swizzle and descriptors is identical to bf16), fp32 output.
// OUTC (4 or 36): fp32 pixel-major output [B,H,W,OUTC] written directly by the epilogue — the pyramid convs C -> 4
// (ncsnpp.py:218,230), either as a 3x3 conv with N = 16 (4 real columns) or, much cheaper on the tensor pipe, as the
// 1-tap "GEMM first" form with N = 48 (36 = 9 taps x 4 outputs, summed over shifted pixels by fd_pyramid_gather).
template <int N, bool TF32, int OUTC>
struct HaloCfg {
  static constexpr bool OUT4 = OUTC != 0;      // fp32 pixel-major output of OUTC channels (no TMA store, no statistics)
  static constexpr int kStagesA = 3;
  static constexpr int kStagesB = 6;
  static constexpr int kSliceC = TF32 ? 32 : 64;          // channels per k-slice (128 bytes)
  static constexpr int kBBytes = (N / 2) * 128;
  static constexpr int kOutBytes = OUT4 ? 0 : 2 * kTileM * 128;
  static constexpr int kSsFloats = 2 * 512;   // scale/shift of up to 512 transformed channels
  static constexpr int kTmemCols = (2 * N <= 32) ? 32 : ((2 * N <= 64) ? 64 : ((2 * N <= 128) ? 128 : ((2 * N <= 256) ? 256 : 512)));
  static constexpr int kStatScratch = (FD_EPI_STATS_SMEM && !OUT4) ? 4 * 32 * 36 * 4 : 0;   // per epilogue warp: 32 x 36 floats
  static constexpr int kStatBytes = OUT4 ? 0 : 4 * 64 * 2 * 4;                              // per-warp column sums of one chunk
  static constexpr int kSmemBytes = 1024 + kStagesA * kHaloStageBytes + kStagesB * kBBytes + kOutBytes + N * 4 +
                                    kSsFloats * 4 + 512 + kStatScratch + kStatBytes;
  static_assert(kBBytes % 1024 == 0, "B stage must keep the 1024-byte swizzle alignment");
  static_assert(OUTC == 0 || (OUTC == 4 && N == 16) || (OUTC == 36 && N == 48), "fp32 pixel-major forms: 4 of 16 or 36 of 48 columns");
};

// Transform-warp placement.  FD_XF_LAYOUT 0 (default): warps 7..14 (two of them share the MMA warp's scheduler
// partition, warp % 4 == 1); 1: warps {7,8,10,11,12,14,15,16}, i.e. none on that partition.  Measured identical
// on B200 (309.0 us both, 384x128 256->256 x 8 clips; whole step 113.5 vs 113.2 audio-s/s), so the MMA issue
// is not what the transform's ALU work slows down — kept as a build-time switch for tools/bench_variants.py.
#ifndef FD_XF_LAYOUT
#define FD_XF_LAYOUT 0
#endif
constexpr int kHaloXfThreads = FD_XF_LAYOUT ? 544 : 480;

// CL4: clusters of FOUR CTAs = two MMA pairs that share the weight stream: ranks 0 / 1 load their half of every weight
// tile ONCE and multicast it to {0, 2} / {1, 3} (weights are 6.4x the A bytes on the L2 -> SM path: 590 KB vs 92 KB per
// 128-pixel tile at 256 -> 256).  A weight slot is free when BOTH pairs' MMAs have retired it, so the pairs advance in
// lock step on the weight ring (6 stages of slack); everything else stays per pair.
template <int N, bool XF, bool TF32, int OUTC, bool CL4 = false>
__global__ void __launch_bounds__(XF ? kHaloXfThreads : 224, 1) conv_halo_kernel(const __grid_constant__ HaloParams p) {
  using Cfg = HaloCfg<N, TF32, OUTC>;
  constexpr bool OUT4 = OUTC != 0;
  constexpr int SA = Cfg::kStagesA, SB = Cfg::kStagesB, B_BYTES = Cfg::kBBytes;
  constexpr int kSliceC = Cfg::kSliceC;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* sA = smem;
  uint8_t* sB = sA + SA * kHaloStageBytes;
  uint8_t* sOut = sB + SB * B_BYTES;
  float* sBias = reinterpret_cast<float*>(sOut + Cfg::kOutBytes);
  float* sSS = sBias + N;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sSS + Cfg::kSsFloats);
  uint64_t* fullA = bars;                 // [SA] TMA landed (XF: local, consumed by the transform warps)
  uint64_t* readyA = bars + SA;           // [SA] operand ready for the MMA (leader CTA's copy is used)
  uint64_t* emptyA = bars + 2 * SA;       // [SA]
  uint64_t* fullB = bars + 3 * SA;        // [SB] (leader's copy)
  uint64_t* emptyB = bars + 3 * SA + SB;  // [SB]
  uint64_t* tfull_bar = bars + 3 * SA + 2 * SB;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t crank = cluster_ctarank();     // 0..1 (pair) or 0..3 (CL4: two pairs)
  const uint32_t rank = crank & 1u;             // rank inside the MMA pair
  const uint32_t lead = crank & ~1u;            // cluster rank of this pair's leader
  const bool leader_cta = (rank == 0);
  const uint16_t pair_mask = static_cast<uint16_t>(3u << lead);

  for (int i = threadIdx.x; i < N; i += blockDim.x) sBias[i] = p.bias ? p.bias[i] : 0.0f;
  if (threadIdx.x == 0) {
    for (int s = 0; s < SA; ++s) {
      mbar_init(&fullA[s], 1);
      mbar_init(&readyA[s], XF ? 16 : 2);  // XF: 8 transform warps of each CTA; else expect_tx + peer arrive
      mbar_init(&emptyA[s], 1);
    }
    for (int s = 0; s < SB; ++s) {
      mbar_init(&fullB[s], 2);
      mbar_init(&emptyB[s], (CL4 && crank < 2) ? 2 : 1);   // the loading pair waits for both pairs' MMAs
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tfull_bar[a], 1);
      mbar_init(&tempty_bar[a], 256);
    }
    fence_mbar_init();
    for (int s = 0; s < p.nseg; ++s) tma_prefetch_desc(&p.a_map[s]);
    tma_prefetch_desc(&p.b_map);
    if (!OUT4) tma_prefetch_desc(&p.out_map);
  }
  if (warp == 1) tmem_alloc_2sm(tmem_slot, Cfg::kTmemCols);
  tc_fence_before_sync();
  cluster_sync_all();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  const int tile_first = static_cast<int>(blockIdx.x);      // cluster c, rank r -> tile cluster_size * c + r
  const int tile_stride = static_cast<int>(gridDim.x);
  const int tiles_per_img = p.tiles_h * p.tiles_w;

  if (warp == 0) {
    // ---------------------------------------------------------------- A producer (halo boxes)
    // (the weight ring has its own producer thread in warp 6, so neither ring's latency is
    //  coupled to the other's depth)
    if (lane == 0) {
      int sa = 0;
      uint32_t pa = 0;
      for (int tile = tile_first; tile < p.num_tiles; tile += tile_stride) {
        const int n = tile / tiles_per_img;
        const int rem = tile - n * tiles_per_img;
        const int h0 = (rem / p.tiles_w) * kHaloTileH;
        const int w0 = (rem % p.tiles_w) * kHaloTileW;
        for (int s = 0; s < p.nseg; ++s) {
          for (int ks = 0; ks < p.seg_kslices[s]; ++ks) {
            mbar_wait(&emptyA[sa], pa ^ 1u);
            if (XF) {
              mbar_expect_tx(&fullA[sa], kHaloTxBytes);
              tma_load_4d(sA + sa * kHaloStageBytes, &p.a_map[s], &fullA[sa], ks * kSliceC, w0 - 1, h0 - 1, n);
            } else {
              if (leader_cta) mbar_expect_tx(&readyA[sa], 2 * kHaloTxBytes);
              else mbar_arrive_remote(&readyA[sa], lead);
              tma_load_4d_2sm(sA + sa * kHaloStageBytes, &p.a_map[s], &readyA[sa], ks * kSliceC, w0 - 1, h0 - 1, n);
            }
            if (++sa == SA) { sa = 0; pa ^= 1u; }
          }
        }
      }
    }
  } else if (warp == 6) {
    // ---------------------------------------------------------------- B producer (weight tiles)
    if (lane == 0) {
      int sb = 0;
      uint32_t pb = 0;
      for (int tile = tile_first; tile < p.num_tiles; tile += tile_stride) {
        for (int s = 0; s < p.nseg; ++s) {
          const int ntap = p.seg_taps[s];
          for (int ks = 0; ks < p.seg_kslices[s]; ++ks) {
            for (int tap = 0; tap < ntap; ++tap) {
              const int kcol = p.seg_kbase[s] + tap * p.seg_cin[s] + ks * kSliceC;
              mbar_wait(&emptyB[sb], pb ^ 1u);
              if (leader_cta) mbar_expect_tx(&fullB[sb], 2 * B_BYTES);
              else mbar_arrive_remote(&fullB[sb], lead);
              if (!CL4)
                tma_load_2d_2sm(sB + sb * B_BYTES, &p.b_map, &fullB[sb], kcol, static_cast<int>(rank) * (N / 2));
              else if (crank < 2)     // this half of the tile goes to the same-parity CTA of both pairs
                tma_load_2d_2sm_mcast(sB + sb * B_BYTES, &p.b_map, &fullB[sb], kcol, static_cast<int>(rank) * (N / 2),
                                      static_cast<uint16_t>((1u << rank) | (1u << (rank + 2))));
              if (++sb == SB) { sb = 0; pb ^= 1u; }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ---------------------------------------------------------------- MMA issuer (leader CTA)
    // warp-uniform loop, one elected lane issues (descriptors stay in uniform registers)
    if (leader_cta) {
      constexpr uint32_t idesc = TF32 ? umma_idesc_tf32(2 * kTileM, N) : umma_idesc_bf16(2 * kTileM, N);
      int sa = 0, sb = 0;
      uint32_t pa = 0, pb = 0;
      int acc = 0, issued = 0;
      uint32_t acc_phase = 0;
      long long w_tempty = 0, w_ready = 0, w_fullb = 0, tq = 0;
      const long long t_begin = FD_DBG ? clock64() : 0;
      for (int tile = tile_first; tile < p.num_tiles; tile += tile_stride) {
        ++issued;
        if (FD_DBG) tq = clock64();
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1u);
        if (FD_DBG) w_tempty += clock64() - tq;
        tc_fence_after_sync();
        const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc * N);
        uint32_t first = 1;
        for (int s = 0; s < p.nseg; ++s) {
          const int ntap = p.seg_taps[s];
          const bool last_seg = (s == p.nseg - 1);
          for (int ks = 0; ks < p.seg_kslices[s]; ++ks) {
            if (FD_DBG) tq = clock64();
            if (XF) mbar_wait_acquire_cluster(&readyA[sa], pa); else mbar_wait(&readyA[sa], pa);
            if (FD_DBG) w_ready += clock64() - tq;
            tc_fence_after_sync();
            const uint32_t a_base = smem_u32(sA + sa * kHaloStageBytes);
            const bool last_stage = last_seg && (ks == p.seg_kslices[s] - 1);
            for (int tap = 0; tap < ntap; ++tap) {
              const int dh = (ntap == 9) ? tap / 3 - 1 : 0;
              const int dw = (ntap == 9) ? tap % 3 - 1 : 0;
              if (FD_DBG) tq = clock64();
              mbar_wait(&fullB[sb], pb);
              if (FD_DBG) w_fullb += clock64() - tq;
              tc_fence_after_sync();
              // rows of the box are halo pixels at a 10-pixel pitch: tap view = row offset, SBO = one box row
              const uint64_t da = umma_desc_k_sw128_sbo(
                  a_base + static_cast<uint32_t>(((1 + dh) * kHaloCols + (1 + dw)) * 128), kHaloCols * 128);
              const uint64_t db = umma_desc_k_sw128(smem_u32(sB + sb * B_BYTES));
              if (elect_one()) {
#pragma unroll
                for (int k = 0; k < 4; ++k) {     // four 32-byte K steps per 128-byte slice (K = 16 bf16 / 8 tf32)
                  if (TF32)
                    umma_tf32_2sm(d_tmem, da + static_cast<uint64_t>(k * 2), db + static_cast<uint64_t>(k * 2),
                                  idesc, (first && k == 0) ? 0u : 1u);
                  else
                    umma_bf16_2sm(d_tmem, da + static_cast<uint64_t>(k * 2), db + static_cast<uint64_t>(k * 2),
                                  idesc, (first && k == 0) ? 0u : 1u);
                }
                // CL4: the second pair also releases the slot in the loading pair (ranks 0, 1)
                umma_commit_2sm_mask(&emptyB[sb], (CL4 && crank == 2) ? static_cast<uint16_t>(0xF) : pair_mask);
                if (tap == ntap - 1) {
                  umma_commit_2sm_mask(&emptyA[sa], pair_mask);
                  if (last_stage) umma_commit_2sm_mask(&tfull_bar[acc], pair_mask);
                }
              }
              __syncwarp();
              first = 0;
              if (++sb == SB) { sb = 0; pb ^= 1u; }
            }
            if (++sa == SA) { sa = 0; pa ^= 1u; }
          }
        }
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1u;
      }
      if (issued > 0)
        for (int j = (issued >= 2 ? issued - 2 : issued - 1); j < issued; ++j)
          mbar_wait(&tempty_bar[j & 1], static_cast<uint32_t>((j >> 1) & 1));
      if (FD_DBG && blockIdx.x == 0 && lane == 0) {
        p.dbg[0] = clock64() - t_begin;
        p.dbg[1] = w_tempty;
        p.dbg[2] = w_ready;
        p.dbg[3] = w_fullb;
        p.dbg[4] = issued;
      }
    }
  } else if (warp < 6) {
    // ---------------------------------------------------------------- epilogue (as conv_igemm_kernel)
    const int ew = warp & 3;
    const int row = ew * 32 + lane;       // = hl * 8 + wl of the 16 x 8 tile
    const bool leader = (threadIdx.x == 64);
    // per-warp 32 x 36-float tile behind the barrier block (column statistics by shared-memory transpose)
    const uint32_t stat_scratch =
        FD_EPI_STATS_SMEM ? smem_u32(reinterpret_cast<uint8_t*>(bars) + 512) + static_cast<uint32_t>(ew) * 4608u : 0u;
    float* sStat = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + 512 + Cfg::kStatScratch);
    float* my_stat = sStat + ew * 128;
    const int et = static_cast<int>(threadIdx.x) - 64;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = tile_first; tile < p.num_tiles; tile += tile_stride) {
      const int n = tile / tiles_per_img;
      const int rem = tile - n * tiles_per_img;
      const int h0 = (rem / p.tiles_w) * kHaloTileH;
      const int w0 = (rem % p.tiles_w) * kHaloTileW;
      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after_sync();
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(ew * 32) << 16) + static_cast<uint32_t>(acc * N);
      if constexpr (OUT4) {
        // pixel-major fp32 output: the first OUTC accumulator columns of this pixel's row
        const int hl = row >> 3, wl = row & 7;
        float* o = p.out4 + ((static_cast<size_t>(n) * p.H + (h0 + hl)) * p.W + (w0 + wl)) * OUTC;
        constexpr int kGroups = (OUTC + 15) / 16;
#pragma unroll
        for (int g = 0; g < kGroups; ++g) {
          uint32_t v[16];
          tmem_ld_32x32b_x16(t_row + g * 16, v);
          tmem_ld_wait();
          if (g == kGroups - 1) {
            tc_fence_before_sync();
            mbar_arrive_remote(&tempty_bar[acc], lead);
          }
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int c = g * 16 + q * 4;
            if (c < OUTC) {
              float4 r;
              r.x = __uint_as_float(v[q * 4 + 0]) + sBias[c + 0];
              r.y = __uint_as_float(v[q * 4 + 1]) + sBias[c + 1];
              r.z = __uint_as_float(v[q * 4 + 2]) + sBias[c + 2];
              r.w = __uint_as_float(v[q * 4 + 3]) + sBias[c + 3];
              *reinterpret_cast<float4*>(o + c) = r;
            }
          }
        }
      } else if constexpr (TF32) {
        constexpr int kChunks = N / 32;           // 32 fp32 channels = one 128-byte staging row
        float* stat_row = p.stats ? p.stats + ((static_cast<size_t>(n) * tiles_per_img + rem) * N) * 2 : nullptr;
#pragma unroll 1
        for (int ch = 0; ch < kChunks; ++ch) {
          uint8_t* stg = sOut + (ch & 1) * (kTileM * 128);
          if (leader) tma_store_wait_read<1>();
          named_bar_sync(1, 128);
          uint32_t v[32];
          tmem_ld_32x32b_x32(t_row + ch * 32, v);
          tmem_ld_wait();
          if (ch == kChunks - 1) {
            tc_fence_before_sync();
            mbar_arrive_remote(&tempty_bar[acc], lead);
          }
          epilogue_chunk_f32(v, smem_u32(sBias + ch * 32), smem_u32(stg + row * 128), row,
                             stat_row ? my_stat : nullptr, lane, stat_scratch);
          fence_proxy_async_smem();
          named_bar_sync(2, 128);
          if (stat_row) stats_combine_store(sStat, 32, et, stat_row + (ch * 32) * 2);
          if (leader) {
            tma_store_4d(&p.out_map, stg, ch * 32, w0, h0, n);
            tma_store_commit();
          }
        }
      } else {
      constexpr int kChunks = N / 64;
      float* stat_row = p.stats ? p.stats + ((static_cast<size_t>(n) * tiles_per_img + rem) * N) * 2 : nullptr;
#pragma unroll 1
      for (int ch = 0; ch < kChunks; ++ch) {
        uint8_t* stg = sOut + (ch & 1) * (kTileM * 128);
        if (leader) tma_store_wait_read<1>();
        named_bar_sync(1, 128);
        const uint32_t bs = smem_u32(sBias + ch * 64);
        const uint32_t rowp = smem_u32(stg + row * 128);
        {
          // one 32-column half at a time: keeps the register footprint low enough for 480 threads
          uint32_t v[32];
          tmem_ld_32x32b_x32(t_row + ch * 64, v);
          tmem_ld_wait();
          epilogue_half(v, bs, rowp, row, 0, stat_row ? my_stat : nullptr, lane, stat_scratch);
          tmem_ld_32x32b_x32(t_row + ch * 64 + 32, v);
          tmem_ld_wait();
          if (ch == kChunks - 1) {
            tc_fence_before_sync();
            mbar_arrive_remote(&tempty_bar[acc], lead);
          }
          epilogue_half(v, bs + 128, rowp, row, 4, stat_row ? my_stat + 64 : nullptr, lane, stat_scratch);
        }
        fence_proxy_async_smem();
        named_bar_sync(2, 128);
        if (stat_row) stats_combine_store(sStat, 64, et, stat_row + (ch * 64) * 2);
        if (leader) {
          tma_store_4d(&p.out_map, stg, ch * 64, w0, h0, n);
          tma_store_commit();
        }
      }
      }
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1u;
    }
    if (!OUT4 && leader) tma_store_wait_all<0>();
  } else if (XF && warp >= 7 && !(FD_XF_LAYOUT && (warp & 3) == 1)) {
    // ---------------------------------------------------------------- transform warps (8 of them)
    const int tq = FD_XF_LAYOUT ? (warp - 7 - (warp > 9 ? 1 : 0) - (warp > 13 ? 1 : 0)) : (warp - 7);
    const int tx = tq * 32 + lane;          // 0..255
    const int j = tx & 7;                   // logical 16-byte chunk = channels j*8 .. j*8+7 of the slice
    const int g = tx >> 3;                  // rows r = g + 32*i  (r & 7 == g & 7 for all of them)
    const int slot = (j ^ (g & 7)) << 4;    // physical chunk position inside the 128-byte row
    int sa = 0;
    uint32_t pa = 0;
    int cur_n = -1;
    long long x_wait = 0, x_work = 0, xq = 0, x_fence = 0, x_ld = 0;
    for (int tile = tile_first; tile < p.num_tiles; tile += tile_stride) {
      const int n = tile / tiles_per_img;
      const int rem = tile - n * tiles_per_img;
      const int h0 = (rem / p.tiles_w) * kHaloTileH;
      const int w0 = (rem % p.tiles_w) * kHaloTileW;
      if (n != cur_n) {
        // (re)load this sample's GroupNorm scale/shift for every transformed segment
        named_bar_sync(3, 256);
        for (int s = 0; s < p.nseg; ++s) {
          if (p.seg_ss[s] == nullptr) continue;
          const float2* src = reinterpret_cast<const float2*>(p.seg_ss[s]) + static_cast<size_t>(n) * p.seg_ss_pitch[s];
          float2* dst = reinterpret_cast<float2*>(sSS) + p.seg_ss_off[s];
          // halved (SiLU is evaluated from v / 2, silu2_from_half) and laid out per channel pair as
          // (scale_c, scale_c+1, shift_c, shift_c+1): one 16-byte load yields the two fp32x2 operands
          float* dstf = reinterpret_cast<float*>(dst);
          for (int c = tx; c < p.seg_cin[s]; c += 256) {
            const float2 v = src[c];
            dstf[(c >> 1) * 4 + (c & 1)] = 0.5f * v.x;
            dstf[(c >> 1) * 4 + 2 + (c & 1)] = 0.5f * v.y;
          }
        }
        named_bar_sync(3, 256);
        cur_n = n;
      }
      // in-image mask of this thread's box rows r = g + 32 i (depends on the tile only, not on the k-slice)
      constexpr int kItems = (kHaloPix + 31) / 32;       // 6 rows per thread (the last one partial)
      const bool interior = h0 > 0 && h0 + kHaloTileH < p.H && w0 > 0 && w0 + kHaloTileW < p.W;
      uint32_t okmask = 0;
#pragma unroll
      for (int i = 0; i < kItems; ++i) {
        const int r = g + 32 * i;
        const int hh = r / kHaloCols, ww = r - hh * kHaloCols;
        const int hy = h0 - 1 + hh, wx = w0 - 1 + ww;
        if ((r < kHaloPix) && hy >= 0 && hy < p.H && wx >= 0 && wx < p.W) okmask |= 1u << i;
      }
      for (int s = 0; s < p.nseg; ++s) {
        const bool xf = p.seg_ss[s] != nullptr;
        for (int ks = 0; ks < p.seg_kslices[s]; ++ks) {
          if (FD_DBG) xq = clock64();
          mbar_wait(&fullA[sa], pa);
          if (FD_DBG) { const long long now = clock64(); x_wait += now - xq; xq = now; }
          if (xf && FD_XF_MODE != 1) {
            constexpr int kPairs = TF32 ? 2 : 4;    // channel pairs per 16-byte chunk
            float2 sc2[4], sh2[4];                  // (scale, shift) / 2 of channel pairs 2e, 2e+1
            const uint32_t ss_addr = smem_u32(sSS) +
                static_cast<uint32_t>(p.seg_ss_off[s] + ks * kSliceC + j * (2 * kPairs)) * 8u;
#pragma unroll
            for (int e = 0; e < kPairs; ++e) {
              const float4 q = lds_f4(ss_addr + e * 16);
              sc2[e] = make_float2(q.x, q.y);
              sh2[e] = make_float2(q.z, q.w);
            }
            const uint32_t base = smem_u32(sA + sa * kHaloStageBytes) + static_cast<uint32_t>(slot);
            // phase 1: issue every load (independent -> the LSU pipelines them).  Branch-free: out-of-image rows
            // hold TMA zero fill and are transformed like the others, then masked to zero (the conv's padding
            // is zero AFTER the activation); only the partial last item is predicated.
            uint4 raw[kItems];
#pragma unroll
            for (int i = 0; i < kItems; ++i) {
              const int r = g + 32 * i;
              raw[i] = make_uint4(0u, 0u, 0u, 0u);
              if (i < kItems - 1 || r < kHaloPix) raw[i] = lds128(base + static_cast<uint32_t>(r) * 128u);
            }
            if (FD_DBG) { const long long now = clock64(); x_ld += now - xq; }
            // phase 2: affine + SiLU (one MUFU op per element) and store back
            if (FD_XF_MODE == 2) {
              uint32_t acc_x = 0;
#pragma unroll
              for (int i = 0; i < kItems; ++i) acc_x ^= raw[i].x ^ raw[i].y ^ raw[i].z ^ raw[i].w;
              if (acc_x == 0x12345678u) sts128(base, raw[0]);      // keep the loads alive
            } else {
              auto transform_rows = [&](auto masked) {
#pragma unroll
                for (int i = 0; i < kItems; ++i) {
                  const int r = g + 32 * i;
                  if (i < kItems - 1 || r < kHaloPix) {
                    const uint32_t m = decltype(masked)::value ? 0u - ((okmask >> i) & 1u) : 0xffffffffu;
                    const uint32_t w4[4] = {raw[i].x, raw[i].y, raw[i].z, raw[i].w};
                    uint32_t o4[4];
                    if constexpr (TF32) {
#pragma unroll
                      for (int e = 0; e < 2; ++e) {        // four fp32 channels per chunk, rounded (RN) to tf32
                        const float2 x = make_float2(__uint_as_float(w4[2 * e]), __uint_as_float(w4[2 * e + 1]));
                        const float2 y = silu2_from_half(ffma2(x, sc2[e], sh2[e]));
                        o4[2 * e] = m & __float_as_uint(round_tf32(y.x));
                        o4[2 * e + 1] = m & __float_as_uint(round_tf32(y.y));
                      }
                    } else {
#pragma unroll
                      for (int e = 0; e < 4; ++e) {        // two channels per packed fp32x2 operation
                        const float2 y = silu2_from_half(ffma2(unpack_bf16x2(w4[e]), sc2[e], sh2[e]));
                        o4[e] = m & pack_bf16x2(y.x, y.y);
                      }
                    }
                    uint4 q = make_uint4(o4[0], o4[1], o4[2], o4[3]);
                    if (FD_XF_MODE != 3 || q.x == 0x12345678u) sts128(base + static_cast<uint32_t>(r) * 128u, q);
                  }
                }
              };
              // ~90 % of the tiles do not touch the image border: no masking needed there (block-uniform branch)
              if (interior) transform_rows(std::false_type{}); else transform_rows(std::true_type{});
            }
            long long f0 = 0;
            if (FD_DBG) f0 = clock64();
            fence_proxy_async_smem();
            if (FD_DBG) x_fence += clock64() - f0;
          }
          __syncwarp();
          if (lane == 0) mbar_arrive_remote_release_cluster(&readyA[sa], lead);
          if (FD_DBG) x_work += clock64() - xq;
          if (++sa == SA) { sa = 0; pa ^= 1u; }
        }
      }
    }
    if (FD_DBG && blockIdx.x == 0 && tx == 0) {
      p.dbg[14] = x_wait;
      p.dbg[15] = x_work;
      p.dbg[12] = x_ld;
      p.dbg[13] = x_fence;
    }
  }

  __syncwarp();
  tc_fence_before_sync();
  cluster_sync_all();
  if (warp == 1) tmem_dealloc_2sm(tmem_base, Cfg::kTmemCols);
}

// ---------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) ==
            cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
  }
  return fn;
}

// per-device caches (a process may drive several GPUs: `enhance.py --device cuda:1`)
int current_device() {
  int dev = 0;
  cudaGetDevice(&dev);
  return (dev >= 0 && dev < kMaxDevices) ? dev : 0;
}

int device_sm_count() {
  static int sms[kMaxDevices] = {0};
  const int dev = current_device();
  if (!sms[dev]) cudaDeviceGetAttribute(&sms[dev], cudaDevAttrMultiProcessorCount, dev);
  return sms[dev];
}

// NHWC bf16 tensor [B,H,W,Ctot]; the map exposes channels [c_begin, c_begin + c_count)
// (f32: fp32 elements, 32 channels per 128-byte box row — the tf32 kernels)
static int make_nhwc_map(CUtensorMap* m, const void* base, int B, int H, int W, int Ctot,
                         int c_begin, int c_count, int bh, int bw, bool f32 = false) {
  EncodeTiledFn enc = get_encode_fn();
  FD_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled is not available from the driver");
  const size_t el = f32 ? 4 : 2;
  FD_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0 && (c_begin % 8) == 0 && (Ctot % 8) == 0,
             "NHWC tensor must be 16-byte aligned with channel counts that are multiples of 8");
  cuuint64_t dims[4] = {static_cast<cuuint64_t>(c_count), static_cast<cuuint64_t>(W),
                        static_cast<cuuint64_t>(H), static_cast<cuuint64_t>(B)};
  cuuint64_t strides[3] = {static_cast<cuuint64_t>(Ctot) * el,
                           static_cast<cuuint64_t>(W) * Ctot * el,
                           static_cast<cuuint64_t>(H) * W * Ctot * el};
  cuuint32_t box[4] = {static_cast<cuuint32_t>(f32 ? 32 : kSliceK), static_cast<cuuint32_t>(bw),
                       static_cast<cuuint32_t>(bh), 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  void* addr = const_cast<uint8_t*>(static_cast<const uint8_t*>(base)) + static_cast<size_t>(c_begin) * el;
  CUresult r = enc(m, f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, addr, dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  FD_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(NHWC) failed with CUresult %d", (int)r);
  return 0;
}

static int make_weight_map(CUtensorMap* m, const void* base, int npad, int ktot, int box_rows, bool f32 = false) {
  EncodeTiledFn enc = get_encode_fn();
  FD_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled is not available from the driver");
  cuuint64_t dims[2] = {static_cast<cuuint64_t>(ktot), static_cast<cuuint64_t>(npad)};
  cuuint64_t strides[1] = {static_cast<cuuint64_t>(ktot) * (f32 ? 4 : 2)};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(f32 ? 32 : kSliceK), static_cast<cuuint32_t>(box_rows)};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides,
                   box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  FD_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(weights) failed with CUresult %d", (int)r);
  return 0;
}

template <int N, bool OUT_F32, bool CTA2>
static int launch_conv(const ConvParams& p, int max_ctas, cudaStream_t stream) {
  auto kern = conv_igemm_kernel<N, OUT_F32, CTA2>;
  using Cfg = ConvCfg<N, CTA2>;
  static bool attr_set[kMaxDevices] = {false};   // function attributes are per device
  const int dev = current_device();
  if (!attr_set[dev]) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes);
    FD_REQUIRE(e == cudaSuccess, "cudaFuncSetAttribute(smem=%d) failed: %s", Cfg::kSmemBytes,
               cudaGetErrorString(e));
    attr_set[dev] = true;
  }
  int grid = std::min(p.num_tiles, max_ctas > 0 ? max_ctas : device_sm_count());
  if (CTA2) grid &= ~1;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(192);
  cfg.dynamicSmemBytes = Cfg::kSmemBytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CTA2 ? 2 : 1;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, p);
  FD_REQUIRE(e == cudaSuccess, "fd_conv2d_igemm: launch failed: %s", cudaGetErrorString(e));
  return check_launch("fd_conv2d_igemm");
}

// 4-CTA clusters with weight multicast for the bf16 N = 128 / 256 tiles (fd_conv_cluster4).  OFF by default: measured on
// B200 (A/B on one box, 32 x 2 s NFE 6): 123.0 vs 124.4-124.8 audio-s/s for CTA pairs.  Only 33 clusters of 4 (132 of 148
// SMs) are co-resident because GPC SM counts are not multiples of 4, and CTA pairs limited to 132 CTAs lose the same 1 %:
// halving the L2 -> SM weight traffic buys nothing measurable on a power-capped step.
static int g_halo_cl4 = 0;

template <int N, bool XF, bool TF32 = false, int OUTC = 0, bool CL4 = false>
static int launch_halo(const HaloParams& p, int max_ctas, cudaStream_t stream) {
  auto kern = conv_halo_kernel<N, XF, TF32, OUTC, CL4>;
  using Cfg = HaloCfg<N, TF32, OUTC>;
  constexpr int kCluster = CL4 ? 4 : 2;
  static bool attr_set[kMaxDevices] = {false};   // function attributes are per device
  const int dev = current_device();
  if (!attr_set[dev]) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes);
    FD_REQUIRE(e == cudaSuccess, "cudaFuncSetAttribute(smem=%d) failed: %s", Cfg::kSmemBytes,
               cudaGetErrorString(e));
    attr_set[dev] = true;
  }
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.blockDim = dim3(XF ? kHaloXfThreads : 224);
  cfg.dynamicSmemBytes = Cfg::kSmemBytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = kCluster;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  // persistent kernel: the grid is what can be co-resident (a GPC whose SM count is not a multiple of the cluster
  // size leaves SMs out), never more — a second wave would double the tail
  static int max_clusters[kMaxDevices] = {0};
  if (!max_clusters[dev]) {
    cfg.gridDim = dim3(kCluster * 64);
    int n = 0;
    cudaError_t eo = cudaOccupancyMaxActiveClusters(&n, kern, &cfg);
    FD_REQUIRE(eo == cudaSuccess && n > 0, "cudaOccupancyMaxActiveClusters failed: %s", cudaGetErrorString(eo));
    max_clusters[dev] = n;
    if (getenv("FD_DEBUG")) fprintf(stderr, "flowdec_b200: halo<%d,%d,%d,%d> cluster %d: %d co-resident clusters (%d CTAs)\n",
                                    N, (int)XF, (int)TF32, OUTC, kCluster, n, n * kCluster);
  }
  int grid = std::min(p.num_tiles, max_ctas > 0 ? max_ctas : device_sm_count());
  grid = std::min(grid, max_clusters[dev] * kCluster) / kCluster * kCluster;
  FD_REQUIRE(grid >= kCluster, "fd_conv2d_igemm(halo): grid %d smaller than a cluster", grid);
  cfg.gridDim = dim3(grid);
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, p);
  FD_REQUIRE(e == cudaSuccess, "fd_conv2d_igemm(halo): launch failed: %s", cudaGetErrorString(e));
  return check_launch("fd_conv2d_igemm(halo)");
}

}  // namespace fd

// ---------------------------------------------------------------------------------
// C ABI (declared in include/flowdec_b200.h)
// ---------------------------------------------------------------------------------
struct fd_conv_src {
  const void* ptr;           // bf16 NHWC [B,H,W,C]
  int C;                     // channel pitch of the tensor
  int c_begin;               // first channel consumed
  int c_count;               // channels consumed (multiple of 64)
  int taps;                  // 1 or 9
  const float* scale_shift;  // nullptr, or GroupNorm scale/shift [B][ss_pitch][2] of THIS source's first
                             // consumed channel: the kernel applies SiLU(x*scale+shift) to the operand
  int ss_pitch;              // channels per sample in that table (the virtual concat's width)
};

// 1: bf16 halo convs run in 4-CTA clusters with weight multicast when the tile count allows; 0 (default): CTA pairs
extern "C" int fd_conv_cluster4(int on) {
  const int prev = fd::g_halo_cl4;
  fd::g_halo_cl4 = on;
  return prev;
}

extern "C" int fd_conv2d_igemm(const fd_conv_src* srcs, int nsrc, const void* wpacked, int ktot,
                               const float* bias, void* out, int out_is_f32, int cout, int npad,
                               int B, int H, int W, float* stats, int max_ctas, int flags, cudaStream_t stream) {
  using namespace fd;
  FD_REQUIRE(nsrc >= 1 && nsrc <= kMaxSeg, "fd_conv2d_igemm: nsrc=%d out of range [1,%d]", nsrc, kMaxSeg);
  FD_REQUIRE(npad == 16 || npad == 48 || npad == 128 || npad == 256, "fd_conv2d_igemm: npad=%d unsupported", npad);
  FD_REQUIRE(W % 8 == 0 && W >= 8, "fd_conv2d_igemm: W=%d must be a multiple of 8", W);
  int bw = 8;
  while (bw < 128 && W % (bw * 2) == 0) bw *= 2;
  const int bh = kTileM / bw;
  FD_REQUIRE(H % bh == 0, "fd_conv2d_igemm: H=%d not divisible by tile height %d", H, bh);

  const bool tf32 = (flags & 4) != 0;      // fp32 activations + weights, kind::tf32 MMAs (halo kernels only)
  const int slice = tf32 ? 32 : kSliceK;
  int ksum = 0;
  for (int s = 0; s < nsrc; ++s) {
    FD_REQUIRE(srcs[s].taps == 1 || srcs[s].taps == 9, "fd_conv2d_igemm: taps must be 1 or 9");
    FD_REQUIRE(srcs[s].c_count % slice == 0 && srcs[s].c_count > 0,
               "fd_conv2d_igemm: segment channel count %d not a multiple of %d", srcs[s].c_count, slice);
    ksum += srcs[s].c_count * srcs[s].taps;
  }
  FD_REQUIRE(ksum == ktot, "fd_conv2d_igemm: packed K=%d does not match segments (%d)", ktot, ksum);
  bool any_xf = false;
  for (int s = 0; s < nsrc; ++s) any_xf = any_xf || (srcs[s].scale_shift != nullptr);
  {
    // "halo" kernel: 16x8 tiles, one A box per k-slice for all nine taps, optional fused GN+SiLU, optional tf32,
    // optional 4-channel fp32 output (pyramid convs)
    const bool out4 = out_is_f32 && npad == 16 && cout == 4;
    // "GEMM first" pyramid form (1-tap sources): on the halo kernel only when it must be (fused transform, tf32);
    // with a materialised bf16 operand the per-tap kernel below is faster (r2c: 569 vs 452 us per forward)
    const bool out36 = out_is_f32 && npad == 48 && cout == 36 && (tf32 || srcs[0].scale_shift != nullptr);
    const bool geom_ok = (flags & 1) && (flags & 2) && (W % kHaloTileW == 0) && (H % kHaloTileH == 0) &&
                         ((static_cast<long long>(B) * (H / kHaloTileH) * (W / kHaloTileW)) % 2 == 0);
    const bool halo_ok = geom_ok && (out4 || out36 || ((npad == 128 || npad == 256) && (out_is_f32 != 0) == tf32));
    FD_REQUIRE(halo_ok || !any_xf, "fd_conv2d_igemm: fused GroupNorm+SiLU needs the halo kernel "
               "(flags 3, W %% 8 == 0, H %% 16 == 0, even tile count)");
    FD_REQUIRE(halo_ok || !tf32, "fd_conv2d_igemm: tf32 (flags bit 2) needs the halo kernel "
               "(flags 7, fp32 output, npad 128/256 or the 4-channel form, W %% 8 == 0, H %% 16 == 0, even tile count)");
    if (halo_ok) {
      HaloParams hp;
      memset(&hp, 0, sizeof(hp));
      hp.nseg = nsrc;
      int kb = 0, ssoff = 0;
      for (int s = 0; s < nsrc; ++s) {
        hp.seg_kslices[s] = srcs[s].c_count / slice;
        hp.seg_taps[s] = srcs[s].taps;
        hp.seg_kbase[s] = kb;
        hp.seg_cin[s] = srcs[s].c_count;
        hp.seg_ss[s] = srcs[s].scale_shift;
        hp.seg_ss_pitch[s] = srcs[s].ss_pitch;
        hp.seg_ss_off[s] = ssoff;
        if (srcs[s].scale_shift) ssoff += srcs[s].c_count;
        kb += srcs[s].c_count * srcs[s].taps;
        if (make_nhwc_map(&hp.a_map[s], srcs[s].ptr, B, H, W, srcs[s].C, srcs[s].c_begin, srcs[s].c_count,
                          kHaloRows, kHaloCols, tf32))
          return 1;
      }
      FD_REQUIRE(ssoff <= 512, "fd_conv2d_igemm: at most 512 transformed channels (got %d)", ssoff);
      if (make_weight_map(&hp.b_map, wpacked, npad, ktot, npad / 2, tf32)) return 1;
      if (!out4 && !out36) {
        FD_REQUIRE(cout == npad, "fd_conv2d_igemm: NHWC output needs cout == npad");
        if (make_nhwc_map(&hp.out_map, out, B, H, W, cout, 0, cout, kHaloTileH, kHaloTileW, tf32)) return 1;
      } else {
        FD_REQUIRE(stats == nullptr, "fd_conv2d_igemm: no statistics for the 4-channel output");
        hp.out4 = static_cast<float*>(out);
      }
      hp.B = B;
      hp.H = H;
      hp.W = W;
      hp.tiles_h = H / kHaloTileH;
      hp.tiles_w = W / kHaloTileW;
      hp.num_tiles = B * hp.tiles_h * hp.tiles_w;
      hp.bias = bias;
      hp.stats = stats;
      {
        const char* e = getenv("FD_HALO_DBG");   // device pointer (decimal) of an int64[16] buffer
        hp.dbg = e ? reinterpret_cast<long long*>(strtoull(e, nullptr, 10)) : nullptr;
        const char* m = getenv("FD_HALO_XF_MODE");
        hp.xf_mode = m ? atoi(m) : 0;
      }
      if (out4) {
        if (tf32) return any_xf ? launch_halo<16, true, true, 4>(hp, max_ctas, stream)
                                : launch_halo<16, false, true, 4>(hp, max_ctas, stream);
        return any_xf ? launch_halo<16, true, false, 4>(hp, max_ctas, stream)
                      : launch_halo<16, false, false, 4>(hp, max_ctas, stream);
      }
      if (out36) {
        if (tf32) return any_xf ? launch_halo<48, true, true, 36>(hp, max_ctas, stream)
                                : launch_halo<48, false, true, 36>(hp, max_ctas, stream);
        return any_xf ? launch_halo<48, true, false, 36>(hp, max_ctas, stream)
                      : launch_halo<48, false, false, 36>(hp, max_ctas, stream);
      }
      if (tf32) {
        if (npad == 256) return any_xf ? launch_halo<256, true, true>(hp, max_ctas, stream)
                                       : launch_halo<256, false, true>(hp, max_ctas, stream);
        return any_xf ? launch_halo<128, true, true>(hp, max_ctas, stream)
                      : launch_halo<128, false, true>(hp, max_ctas, stream);
      }
      if (g_halo_cl4 && hp.num_tiles % 4 == 0 && hp.num_tiles >= 8) {
        if (npad == 256) return any_xf ? launch_halo<256, true, false, 0, true>(hp, max_ctas, stream)
                                       : launch_halo<256, false, false, 0, true>(hp, max_ctas, stream);
        return any_xf ? launch_halo<128, true, false, 0, true>(hp, max_ctas, stream)
                      : launch_halo<128, false, false, 0, true>(hp, max_ctas, stream);
      }
      if (npad == 256) return any_xf ? launch_halo<256, true>(hp, max_ctas, stream)
                                     : launch_halo<256, false>(hp, max_ctas, stream);
      return any_xf ? launch_halo<128, true>(hp, max_ctas, stream) : launch_halo<128, false>(hp, max_ctas, stream);
    }
  }
  // ---- per-tap kernel (bf16 only)
  ConvParams p;
  memset(&p, 0, sizeof(p));
  p.nseg = nsrc;
  for (int s = 0; s < nsrc; ++s) {
    p.seg_kslices[s] = srcs[s].c_count / kSliceK;
    p.seg_taps[s] = srcs[s].taps;
    if (make_nhwc_map(&p.a_map[s], srcs[s].ptr, B, H, W, srcs[s].C, srcs[s].c_begin,
                      srcs[s].c_count, bh, bw))
      return 1;
  }
  // CTA pairs (cta_group::2, M = 256 per MMA) for the bf16-output tiles when the tile count is even
  const bool pair = !out_is_f32 && ((flags & 1) != 0) && (((B * (H / bh) * (W / bw)) & 1) == 0) &&
                    (B * (H / bh) * (W / bw) >= 2);
  if (make_weight_map(&p.b_map, wpacked, npad, ktot, pair ? npad / 2 : npad)) return 1;
  p.B = B;
  p.H = H;
  p.W = W;
  p.bh = bh;
  p.bw = bw;
  p.tiles_h = H / bh;
  p.tiles_w = W / bw;
  p.num_tiles = B * p.tiles_h * p.tiles_w;
  p.bias = bias;
  p.cout_valid = cout;
  p.stats = stats;
  if (out_is_f32) {
    FD_REQUIRE(stats == nullptr, "fd_conv2d_igemm: stats are produced by the bf16-output kernels only");
    FD_REQUIRE((npad == 16 || npad == 48) && cout >= 4 && cout <= npad && cout % 4 == 0,
               "fd_conv2d_igemm: fp32 output needs npad in {16,48} and cout %% 4 == 0 (got %d/%d)", npad, cout);
    p.out_f32 = static_cast<float*>(out);
    if (npad == 48) return launch_conv<48, true, false>(p, max_ctas, stream);
    return launch_conv<16, true, false>(p, max_ctas, stream);
  }
  FD_REQUIRE(cout == npad && npad >= 128, "fd_conv2d_igemm: bf16 output needs cout == npad in {128,256}");
  if (make_nhwc_map(&p.out_map, out, B, H, W, cout, 0, cout, bh, bw)) return 1;
  if (npad == 256) return pair ? launch_conv<256, false, true>(p, max_ctas, stream)
                               : launch_conv<256, false, false>(p, max_ctas, stream);
  return pair ? launch_conv<128, false, true>(p, max_ctas, stream)
              : launch_conv<128, false, false>(p, max_ctas, stream);
}
