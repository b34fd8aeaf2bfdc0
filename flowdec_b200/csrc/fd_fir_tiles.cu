// flowdec_b200 — GroupNorm affine + SiLU + FIR x2 up / down of the res-block input, tiled through
// shared memory by TMA.
//
// Replaces, for the up / down res-blocks, `act(GroupNorm_0(x))` followed by `upsample_2d` /
// `downsample_2d` on both h and x (/root/reference/flowdec/backbones/ncsnpp_utils/layerspp.py:262-272,
// up_or_down_sampling.py:220-282 -> op/upfirdn2d_kernel.cu:118-218; FIR [1,3,3,1], closed forms in
// SURVEY.md §7).  HBM-bound: 2 B read per input element, 2 x 2 B written per output element
// (activated + raw), i.e. 1.2 GB (down) / 1.8 GB (up) at 768x256x256ch x 8 clips.
//
// The register kernels in fd_elementwise.cu fetch every input pixel 9x (up) / 2.25x (down) through L1
// and re-evaluate the SiLU each time, with <= 16 resident warps per SM to cover HBM latency: 4-7x off
// the HBM roofline.  Here a persistent block streams 10x18-pixel x 64-channel boxes (an 8x16 input
// tile + 1-pixel halo; out-of-image pixels arrive as zeros = the FIR's zero padding) through a
// two-stage TMA ring, activates each box ONCE into an fp32 shared-memory tile, and computes the FIR
// from shared memory:
//   up   : thread = (input column, 4-channel slice) walking down the box rows; horizontal pair per row,
//          vertical blend of consecutive rows -> 16 x 32 output pixels per tile
//   down : thread = (output pixel, 4-channel slice); 4 x 4 taps, horizontal then vertical
// Arithmetic is fp32 on the same values as the register kernels (bf16 rounding only on the outputs).
#include <cuda.h>

#include "fd_common.cuh"

namespace fd {

constexpr int kFtBoxH = 10, kFtBoxW = 18;
constexpr int kFtPix = kFtBoxH * kFtBoxW;            // 180 pixels
constexpr int kFtRawBytes = kFtPix * 128;            // bf16, 64 channels
constexpr int kFtActBytes = kFtPix * 256;            // fp32, 64 channels
constexpr int kFtStages = 2;
constexpr int kFtThreads = 256;
constexpr int kFtSmem = kFtStages * kFtRawBytes + kFtActBytes + 64 + 128;   // + barriers + alignment slack

struct FirTileParams {
  CUtensorMap map[2];          // NHWC bf16 sources of the virtual concat [src1 | src2]
  int C1, C2, B, H, W;         // input geometry
  int tiles_h, tiles_w, groups, num_tiles;
  const float* scale_shift;    // [B, C1 + C2, 2]
  __nv_bfloat16* out;          // activated + resampled [B, Ho, Wo, C]
  __nv_bfloat16* out_raw;      // resampled only
};

__device__ __forceinline__ uint2 lds64(uint32_t addr) {
  uint2 v;
  asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr));
  return v;
}

__device__ __forceinline__ void sts_f4(uint32_t addr, float a, float b, float c, float d) {
  asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

__device__ __forceinline__ void stg64_bf16(__nv_bfloat16* p, const float2 (&v)[2]) {
  uint2 r;
  r.x = pack_bf16x2(v[0].x, v[0].y);
  r.y = pack_bf16x2(v[1].x, v[1].y);
  *reinterpret_cast<uint2*>(p) = r;
}

struct FtTile {
  int b, r0, c0, g;
};

__device__ __forceinline__ FtTile ft_decode(const FirTileParams& p, int t) {
  FtTile q;
  q.g = t % p.groups;
  t /= p.groups;
  const int tw = t % p.tiles_w;
  t /= p.tiles_w;
  q.r0 = (t % p.tiles_h) * 8;
  q.c0 = tw * 16;
  q.b = t / p.tiles_h;
  return q;
}

__device__ __forceinline__ void ft_issue(const FirTileParams& p, int t, void* dst, uint64_t* bar) {
  const FtTile q = ft_decode(p, t);
  const int ch = q.g * 64;
  const bool first = ch < p.C1;
  mbar_expect_tx(bar, kFtRawBytes);
  tma_load_4d(dst, first ? &p.map[0] : &p.map[1], bar, first ? ch : ch - p.C1, q.c0 - 1, q.r0 - 1, q.b);
}

// MODE 1: FIR down (8x16 input tile -> 4x8 outputs); MODE 2: FIR up (8x16 input tile -> 16x32 outputs)
template <int MODE>
__global__ void __launch_bounds__(kFtThreads, 2) fir_tile_kernel(const __grid_constant__ FirTileParams p) {
  extern __shared__ uint8_t smem_raw_[];
  const uint32_t base = (smem_u32(smem_raw_) + 127u) & ~127u;
  uint8_t* gen = smem_raw_ + (base - smem_u32(smem_raw_));
  const uint32_t act_s = base + kFtStages * kFtRawBytes;
  uint64_t* full = reinterpret_cast<uint64_t*>(gen + kFtStages * kFtRawBytes + kFtActBytes);
  const int tid = threadIdx.x;
  const int C = p.C1 + p.C2;

  if (tid == 0) {
    mbar_init(&full[0], 1);
    mbar_init(&full[1], 1);
    fence_mbar_init();
  }
  __syncthreads();
  if (tid == 0) {
    tma_prefetch_desc(&p.map[0]);
    tma_prefetch_desc(&p.map[1]);
    for (int s = 0; s < kFtStages; ++s) {
      const int t = blockIdx.x + s * gridDim.x;
      if (t < p.num_tiles) ft_issue(p, t, gen + s * kFtRawBytes, &full[s]);
    }
  }

  int it = 0;
  for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x, ++it) {
    const int s = it & 1;
    const uint32_t raw_s = base + s * kFtRawBytes;
    const FtTile q = ft_decode(p, t);
    // ---- phase A: box -> SiLU(scale * x + shift) in fp32 (zero outside the image), once per pixel
    {
      const int oct = tid & 7;
      float2 sc2[4], sh2[4];            // (scale, shift) / 2 of channel pairs: SiLU is evaluated from v / 2
      const float4* ss = reinterpret_cast<const float4*>(
          p.scale_shift + (static_cast<size_t>(q.b) * C + q.g * 64 + oct * 8) * 2);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float4 v = __ldg(ss + i);
        sc2[i] = make_float2(0.5f * v.x, 0.5f * v.z);
        sh2[i] = make_float2(0.5f * v.y, 0.5f * v.w);
      }
      mbar_wait(&full[s], (it >> 1) & 1);
#pragma unroll 2
      for (int px = tid >> 3; px < kFtPix; px += kFtThreads / 8) {
        const int pr = px / kFtBoxW, pc = px - pr * kFtBoxW;
        const bool in = static_cast<unsigned>(q.r0 - 1 + pr) < static_cast<unsigned>(p.H) &&
                        static_cast<unsigned>(q.c0 - 1 + pc) < static_cast<unsigned>(p.W);
        const uint4 r = lds128(raw_s + px * 128 + oct * 16);
        const uint32_t w[4] = {r.x, r.y, r.z, r.w};
        float2 a[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {      // packed fp32x2: two channels per FFMA2
          const float2 y = silu2_from_half(ffma2(unpack_bf16x2(w[i]), sc2[i], sh2[i]));
          a[i] = in ? y : make_float2(0.f, 0.f);
        }
        const uint32_t d = act_s + px * 256 + oct * 32;
        sts_f4(d, a[0].x, a[0].y, a[1].x, a[1].y);
        sts_f4(d + 16, a[2].x, a[2].y, a[3].x, a[3].y);
      }
    }
    __syncthreads();
    // ---- phase B
    const int slice = tid & 15;
    const int ch = q.g * 64 + slice * 4;
    if (MODE == 2) {
      const int col = tid >> 4;                               // input column within the tile, 0..15
      const int Ho = 2 * p.H, Wo = 2 * p.W;
      __nv_bfloat16* o_act = p.out + ((static_cast<size_t>(q.b) * Ho + 2 * q.r0) * Wo + 2 * (q.c0 + col)) * C + ch;
      __nv_bfloat16* o_raw = p.out_raw + (o_act - p.out);
      const size_t row_pitch = static_cast<size_t>(Wo) * C;
      // all FIR arithmetic on channel pairs (fp32x2): v[0] = channels 0,1, v[1] = channels 2,3 of the slice
      const float2 k25 = make_float2(0.25f, 0.25f), k75 = make_float2(0.75f, 0.75f);
      float2 pE[2], pF[2], prE[2], prF[2];                    // horizontal pair of the previous box row
#pragma unroll
      for (int br = 0; br < kFtBoxH; ++br) {
        const int px = br * kFtBoxW + col;
        float2 cE[2], cF[2], crE[2], crF[2];
        {
          const float4 l = lds_f4(act_s + px * 256 + slice * 16);
          const float4 m = lds_f4(act_s + (px + 1) * 256 + slice * 16);
          const float4 r = lds_f4(act_s + (px + 2) * 256 + slice * 16);
          const float2 lv[2] = {make_float2(l.x, l.y), make_float2(l.z, l.w)};
          const float2 mv[2] = {make_float2(m.x, m.y), make_float2(m.z, m.w)};
          const float2 rv[2] = {make_float2(r.x, r.y), make_float2(r.z, r.w)};
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            cE[c] = ffma2(k75, mv[c], fmul2(k25, lv[c]));
            cF[c] = ffma2(k75, mv[c], fmul2(k25, rv[c]));
          }
        }
        {
          const uint2 l = lds64(raw_s + px * 128 + slice * 8);
          const uint2 m = lds64(raw_s + (px + 1) * 128 + slice * 8);
          const uint2 r = lds64(raw_s + (px + 2) * 128 + slice * 8);
          const float2 lv[2] = {unpack_bf16x2(l.x), unpack_bf16x2(l.y)};
          const float2 mv[2] = {unpack_bf16x2(m.x), unpack_bf16x2(m.y)};
          const float2 rv[2] = {unpack_bf16x2(r.x), unpack_bf16x2(r.y)};
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            crE[c] = ffma2(k75, mv[c], fmul2(k25, lv[c]));
            crF[c] = ffma2(k75, mv[c], fmul2(k25, rv[c]));
          }
        }
        if (br >= 1 && br <= 8) {       // current row is tile row i = br - 1: out row 2i = .25 prev + .75 cur
          float2 e[2], f[2], re[2], rf[2];
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            e[c] = ffma2(k75, cE[c], fmul2(k25, pE[c]));
            f[c] = ffma2(k75, cF[c], fmul2(k25, pF[c]));
            re[c] = ffma2(k75, crE[c], fmul2(k25, prE[c]));
            rf[c] = ffma2(k75, crF[c], fmul2(k25, prF[c]));
          }
          const size_t off = static_cast<size_t>(2 * (br - 1)) * row_pitch;
          stg64_bf16(o_act + off, e);
          stg64_bf16(o_act + off + C, f);
          stg64_bf16(o_raw + off, re);
          stg64_bf16(o_raw + off + C, rf);
        }
        if (br >= 2) {                  // previous row is tile row i = br - 2: out row 2i+1 = .75 prev + .25 cur
          float2 e[2], f[2], re[2], rf[2];
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            e[c] = ffma2(k75, pE[c], fmul2(k25, cE[c]));
            f[c] = ffma2(k75, pF[c], fmul2(k25, cF[c]));
            re[c] = ffma2(k75, prE[c], fmul2(k25, crE[c]));
            rf[c] = ffma2(k75, prF[c], fmul2(k25, crF[c]));
          }
          const size_t off = static_cast<size_t>(2 * (br - 2) + 1) * row_pitch;
          stg64_bf16(o_act + off, e);
          stg64_bf16(o_act + off + C, f);
          stg64_bf16(o_raw + off, re);
          stg64_bf16(o_raw + off + C, rf);
        }
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          pE[c] = cE[c]; pF[c] = cF[c]; prE[c] = crE[c]; prF[c] = crF[c];
        }
      }
    } else {
      const int Ho = p.H / 2, Wo = p.W / 2;
      const float2 k2[4] = {make_float2(0.125f, 0.125f), make_float2(0.375f, 0.375f), make_float2(0.375f, 0.375f),
                            make_float2(0.125f, 0.125f)};
      const float2 z2 = make_float2(0.f, 0.f);
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int opx = (tid >> 4) + 16 * u;                  // 4 x 8 output pixels
        const int orow = opx >> 3, ocol = opx & 7;
        float2 acc[2] = {z2, z2}, racc[2] = {z2, z2};         // channel pairs (0,1), (2,3): fp32x2 arithmetic
#pragma unroll
        for (int ri = 0; ri < 4; ++ri) {
          const int px = (2 * orow + ri) * kFtBoxW + 2 * ocol;
          float2 h[2] = {z2, z2}, rh[2] = {z2, z2};
#pragma unroll
          for (int ci = 0; ci < 4; ++ci) {
            const float4 a = lds_f4(act_s + (px + ci) * 256 + slice * 16);
            const uint2 r = lds64(raw_s + (px + ci) * 128 + slice * 8);
            h[0] = ffma2(k2[ci], make_float2(a.x, a.y), h[0]);
            h[1] = ffma2(k2[ci], make_float2(a.z, a.w), h[1]);
            rh[0] = ffma2(k2[ci], unpack_bf16x2(r.x), rh[0]);
            rh[1] = ffma2(k2[ci], unpack_bf16x2(r.y), rh[1]);
          }
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            acc[c] = ffma2(k2[ri], h[c], acc[c]);
            racc[c] = ffma2(k2[ri], rh[c], racc[c]);
          }
        }
        const size_t off =
            ((static_cast<size_t>(q.b) * Ho + q.r0 / 2 + orow) * Wo + q.c0 / 2 + ocol) * C + ch;
        stg64_bf16(p.out + off, acc);
        stg64_bf16(p.out_raw + off, racc);
      }
    }
    __syncthreads();                    // everyone is done with stage s and the fp32 tile
    if (tid == 0) {
      const int tn = t + kFtStages * gridDim.x;
      if (tn < p.num_tiles) ft_issue(p, tn, gen + s * kFtRawBytes, &full[s]);
    }
  }
}

// ---------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------
typedef CUresult (*FtEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                               const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                               CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static FtEncodeFn ft_encode_fn() {
  static FtEncodeFn fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<FtEncodeFn>(ptr);
  }
  return fn;
}

// NHWC bf16 [B,H,W,Cs]: boxes of 64 channels x 18 x 10 pixels, plain (unswizzled) layout in shared memory
static int ft_make_map(CUtensorMap* m, const void* base, int B, int H, int W, int Cs) {
  FtEncodeFn enc = ft_encode_fn();
  FD_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled is not available from the driver");
  FD_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0, "fd_gn_act_resample: source must be 16-byte aligned");
  cuuint64_t dims[4] = {static_cast<cuuint64_t>(Cs), static_cast<cuuint64_t>(W), static_cast<cuuint64_t>(H),
                        static_cast<cuuint64_t>(B)};
  cuuint64_t strides[3] = {static_cast<cuuint64_t>(Cs) * 2, static_cast<cuuint64_t>(W) * Cs * 2,
                           static_cast<cuuint64_t>(H) * W * Cs * 2};
  cuuint32_t box[4] = {64, kFtBoxW, kFtBoxH, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  FD_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(FIR tiles) failed with CUresult %d", (int)r);
  return 0;
}

static int g_fir_tiles_on = 1;

int fir_tiles_set(int on) {
  const int prev = g_fir_tiles_on;
  g_fir_tiles_on = on;
  return prev;
}

bool fir_tiles_eligible(int C1, int C2, int H, int W, int mode) {
  return g_fir_tiles_on && (mode == 1 || mode == 2) && C1 > 0 && C1 % 64 == 0 && C2 % 64 == 0 && H % 8 == 0 && W % 16 == 0;
}

int device_sm_count();
int current_device();

int fir_tiles_launch(const void* src1, int C1, const void* src2, int C2, const float* scale_shift, void* out,
                     void* out_raw, int B, int H, int W, int mode, cudaStream_t stream) {
  FirTileParams p;
  memset(&p, 0, sizeof(p));
  if (ft_make_map(&p.map[0], src1, B, H, W, C1)) return 1;
  if (ft_make_map(&p.map[1], C2 > 0 ? src2 : src1, B, H, W, C2 > 0 ? C2 : C1)) return 1;
  p.C1 = C1;
  p.C2 = C2;
  p.B = B;
  p.H = H;
  p.W = W;
  p.tiles_h = H / 8;
  p.tiles_w = W / 16;
  p.groups = (C1 + C2) / 64;
  const long long nt = static_cast<long long>(B) * p.tiles_h * p.tiles_w * p.groups;
  FD_REQUIRE(nt < (1ll << 31), "fd_gn_act_resample: too many tiles");
  p.num_tiles = static_cast<int>(nt);
  p.scale_shift = scale_shift;
  p.out = static_cast<__nv_bfloat16*>(out);
  p.out_raw = static_cast<__nv_bfloat16*>(out_raw);
  static bool attr_set_dev[kMaxDevices] = {false};   // function attributes are per device
  bool& attr_set = attr_set_dev[current_device()];
  if (!attr_set) {
    cudaError_t e1 = cudaFuncSetAttribute(fir_tile_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kFtSmem);
    cudaError_t e2 = cudaFuncSetAttribute(fir_tile_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, kFtSmem);
    FD_REQUIRE(e1 == cudaSuccess && e2 == cudaSuccess, "cudaFuncSetAttribute(smem=%d) failed: %s", kFtSmem,
               cudaGetErrorString(e1 != cudaSuccess ? e1 : e2));
    attr_set = true;
  }
  int grid = 2 * device_sm_count();
  if (grid > p.num_tiles) grid = p.num_tiles;
  if (mode == 1)
    fir_tile_kernel<1><<<grid, kFtThreads, kFtSmem, stream>>>(p);
  else
    fir_tile_kernel<2><<<grid, kFtThreads, kFtSmem, stream>>>(p);
  return check_launch("fd_gn_act_resample(tiles)");
}

}  // namespace fd
