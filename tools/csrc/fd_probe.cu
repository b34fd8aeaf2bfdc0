// flowdec_b200 — hardware-semantics probe (development tool, not on the product path):
// does a K-major SWIZZLE_128B UMMA operand descriptor work when its start address is 128-byte
// aligned but NOT 1024-byte aligned (row-shifted view into a TMA-written tile), with an arbitrary
// stride between 8-row groups (SBO) and which `base_offset` does it need?  tools/umma_probe.py
// compares the result with the row mapping  row(m) = row_off + (m/8)*(sbo/128) + m%8.
#include "fd_common.cuh"

namespace fd {

__global__ void __launch_bounds__(128, 1) umma_probe_kernel(const __grid_constant__ CUtensorMap a_map,
                                                            const __grid_constant__ CUtensorMap b_map,
                                                            float* __restrict__ out, int row_off,
                                                            int sbo_bytes, int base_offset) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* sA = smem;                 // 256 rows x 128 B
  uint8_t* sB = smem + 256 * 128;     // 16 rows x 128 B
  uint64_t* bar = reinterpret_cast<uint64_t*>(sB + 16 * 128);
  uint64_t* done = bar + 1;
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    mbar_init(done, 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc(slot, 32);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = *slot;
  if (threadIdx.x == 0) {
    mbar_expect_tx(bar, 256 * 128 + 16 * 128);
    tma_load_2d(sA, &a_map, bar, 0, 0);
    tma_load_2d(sB, &b_map, bar, 0, 0);
    mbar_wait(bar, 0);
    tc_fence_after_sync();
    const uint32_t a_addr = smem_u32(sA) + static_cast<uint32_t>(row_off) * 128u;
    uint64_t da = umma_desc_k_sw128(a_addr);
    // replace SBO (bits 32..45) and set base_offset (bits 49..51)
    da &= ~(static_cast<uint64_t>(0x3FFF) << 32);
    da |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFF) << 32;
    da |= static_cast<uint64_t>(base_offset & 7) << 49;
    const uint64_t db = umma_desc_k_sw128(smem_u32(sB));
    constexpr uint32_t idesc = umma_idesc_bf16(128, 16);
#pragma unroll
    for (int k = 0; k < 4; ++k)
      umma_bf16(tmem, da + static_cast<uint64_t>(k * 2), db + static_cast<uint64_t>(k * 2), idesc, k != 0);
    umma_commit(done);
  }
  mbar_wait(done, 0);
  tc_fence_after_sync();
  uint32_t v[16];
  tmem_ld_32x32b_x16(tmem + (static_cast<uint32_t>(warp * 32) << 16), v);
  tmem_ld_wait();
  const int row = warp * 32 + lane;
#pragma unroll
  for (int c = 0; c < 16; ++c) out[row * 16 + c] = __uint_as_float(v[c]);
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 32);
}

// throughput probe: `iters` back-to-back M=128 x N=256 x K=16 MMAs whose A descriptor starts at row
// `row_off` of a 256-row tile with 8-row-group stride `sbo_bytes`; reports cycles per MMA.
__global__ void __launch_bounds__(128, 1) umma_rate_kernel(const __grid_constant__ CUtensorMap a_map,
                                                           const __grid_constant__ CUtensorMap b_map,
                                                           long long* __restrict__ cycles, int row_off,
                                                           int sbo_bytes, int iters) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* sA = smem;                  // 256 rows x 128 B
  uint8_t* sB = smem + 256 * 128;      // 256 rows x 128 B (N = 256)
  uint64_t* bar = reinterpret_cast<uint64_t*>(sB + 256 * 128);
  uint64_t* done = bar + 1;
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 2);
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    mbar_init(done, 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc(slot, 256);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = *slot;
  if (threadIdx.x == 0) {
    mbar_expect_tx(bar, 2 * 256 * 128);
    tma_load_2d(sA, &a_map, bar, 0, 0);
    tma_load_2d(sB, &b_map, bar, 0, 0);
    mbar_wait(bar, 0);
    tc_fence_after_sync();
    const uint64_t da = umma_desc_k_sw128_sbo(smem_u32(sA) + static_cast<uint32_t>(row_off) * 128u,
                                              static_cast<uint32_t>(sbo_bytes));
    const uint64_t db = umma_desc_k_sw128(smem_u32(sB));
    constexpr uint32_t idesc = umma_idesc_bf16(128, 256);
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i)
      umma_bf16(tmem, da + static_cast<uint64_t>((i & 3) * 2), db + static_cast<uint64_t>((i & 3) * 2), idesc, i != 0);
    umma_commit(done);
    mbar_wait(done, 0);
    const long long t1 = clock64();
    cycles[0] = t1 - t0;
  }
  __syncthreads();
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 256);
}

typedef CUresult (*EncodeTiledFn2)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

}  // namespace fd

// a: bf16 [256][64], b: bf16 [16][64], out: fp32 [128][16]
extern "C" int fd_umma_probe(const void* a, const void* b, float* out, int row_off, int sbo_bytes,
                             int base_offset, cudaStream_t stream) {
  using namespace fd;
  void* ptr = nullptr;
  cudaDriverEntryPointQueryResult qres;
  FD_REQUIRE(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
                 qres == cudaDriverEntryPointSuccess,
             "cuTensorMapEncodeTiled unavailable");
  EncodeTiledFn2 enc = reinterpret_cast<EncodeTiledFn2>(ptr);
  CUtensorMap ma, mb;
  cuuint64_t dimsa[2] = {64, 256}, dimsb[2] = {64, 16};
  cuuint64_t str[1] = {128};
  cuuint32_t boxa[2] = {64, 256}, boxb[2] = {64, 16}, es[2] = {1, 1};
  FD_REQUIRE(enc(&ma, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(a), dimsa, str, boxa, es,
                 CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                 CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS, "encode a failed");
  FD_REQUIRE(enc(&mb, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(b), dimsb, str, boxb, es,
                 CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                 CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS, "encode b failed");
  const int smem = 1024 + 256 * 128 + 16 * 128 + 64;
  cudaFuncSetAttribute(umma_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  umma_probe_kernel<<<1, 128, smem, stream>>>(ma, mb, out, row_off, sbo_bytes, base_offset);
  return check_launch("fd_umma_probe");
}

// a: bf16 [256][64], b: bf16 [256][64]; cycles: device int64[1]
extern "C" int fd_umma_rate(const void* a, const void* b, long long* cycles, int row_off, int sbo_bytes,
                            int iters, cudaStream_t stream) {
  using namespace fd;
  void* ptr = nullptr;
  cudaDriverEntryPointQueryResult qres;
  FD_REQUIRE(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
                 qres == cudaDriverEntryPointSuccess,
             "cuTensorMapEncodeTiled unavailable");
  EncodeTiledFn2 enc = reinterpret_cast<EncodeTiledFn2>(ptr);
  CUtensorMap ma, mb;
  cuuint64_t dims[2] = {64, 256};
  cuuint64_t str[1] = {128};
  cuuint32_t box[2] = {64, 256}, es[2] = {1, 1};
  FD_REQUIRE(enc(&ma, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(a), dims, str, box, es,
                 CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                 CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS, "encode a failed");
  FD_REQUIRE(enc(&mb, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(b), dims, str, box, es,
                 CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                 CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS, "encode b failed");
  const int smem = 1024 + 2 * 256 * 128 + 64;
  cudaFuncSetAttribute(umma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  umma_rate_kernel<<<1, 128, smem, stream>>>(ma, mb, cycles, row_off, sbo_bytes, iters);
  return check_launch("fd_umma_rate");
}

// ------------------------------------------------------------------------------------------------
// Probe: which mbarrier receives complete_tx for each destination of a cta_group::2 MULTICAST TMA load in a
// cluster of 4 (two CTA pairs)?  Ranks 0 / 1 each multicast a 1 KB box to {r, r+2} with the barrier operand =
// own barrier address with the peer bit cleared (mode 0) or unmodified (mode 1).  Leaders (0, 2) expect 2 KB,
// non-leaders 1 KB; result[rank] = 1 if that CTA's barrier phase completed, dump = the 1 KB each CTA received.
namespace fd {
__global__ void __cluster_dims__(4, 1, 1) __launch_bounds__(32, 1)
    mcast_probe_kernel(const __grid_constant__ CUtensorMap map, int* result, float* dump, int mode) {
  __shared__ __align__(1024) uint8_t buf[1024];
  __shared__ uint64_t bar;
  const uint32_t rank = cluster_ctarank();
  if (threadIdx.x == 0) {
    for (int i = 0; i < 256; ++i) reinterpret_cast<float*>(buf)[i] = -1.0f;
    mbar_init(&bar, 1);
    fence_mbar_init();
  }
  cluster_sync_all();
  if (threadIdx.x == 0) mbar_expect_tx(&bar, (rank & 1) ? 1024u : 2048u);
  cluster_sync_all();                                    // every barrier armed before any copy is issued
  if (threadIdx.x == 0) {
    if (rank < 2) {
      const uint32_t mb = mode == 0 ? (smem_u32(&bar) & kPeerBitMask) : smem_u32(&bar);
      const uint16_t mask = static_cast<uint16_t>((1u << rank) | (1u << (rank + 2)));
      asm volatile(
          "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster"
          " [%0], [%1, {%3, %4}], [%2], %5;"
          ::"r"(smem_u32(buf)), "l"(reinterpret_cast<uint64_t>(&map)), "r"(mb), "r"(0), "r"(static_cast<int>(rank) * 8),
            "h"(mask)
          : "memory");
    }
    int done = 0;
    for (int i = 0; i < 200000 && !done; ++i) done = mbar_test_wait(&bar, 0) ? 1 : 0;
    result[rank] = done;
  }
  __syncwarp();
  // let every in-flight copy land before anyone exits, then dump what arrived
  for (int i = 0; i < 2000; ++i) __nanosleep(100);
  cluster_sync_all();
  if (threadIdx.x == 0)
    for (int i = 0; i < 256; ++i) dump[rank * 256 + i] = reinterpret_cast<float*>(buf)[i];
}
}  // namespace fd

// src: fp32 [16][256] (row r holds the value r everywhere); result int[4]; dump fp32 [4][256]
extern "C" int fd_mcast_probe(const float* src, int* result, float* dump, int mode, cudaStream_t stream) {
  using namespace fd;
  void* ptr = nullptr;
  cudaDriverEntryPointQueryResult qres;
  FD_REQUIRE(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
                 qres == cudaDriverEntryPointSuccess, "cuTensorMapEncodeTiled unavailable");
  EncodeTiledFn2 enc = reinterpret_cast<EncodeTiledFn2>(ptr);
  CUtensorMap m;
  cuuint64_t dims[2] = {32, 16};             // 32 floats (128 B) x 16 rows
  cuuint64_t str[1] = {1024};                // row pitch 256 floats
  cuuint32_t box[2] = {32, 8}, es[2] = {1, 1};
  FD_REQUIRE(enc(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(src), dims, str, box, es,
                 CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                 CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS, "encode failed");
  mcast_probe_kernel<<<4, 32, 0, stream>>>(m, result, dump, mode);
  return check_launch("fd_mcast_probe");
}
