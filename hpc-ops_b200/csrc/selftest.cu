// Bring-up self test for the tcgen05 descriptor conventions used by the product kernels.
// One CTA: copy caller-supplied smem images of A and B, issue `nk` fp8 UMMAs with caller-supplied
// descriptor fields, read D back from TMEM. tests/ sweep these fields against a CPU matmul.
#include "common.cuh"
#include "host_utils.h"

namespace b200 {
namespace selftest {

constexpr int kMaxA = 65536;
constexpr int kMaxB = 32768;

__global__ void __launch_bounds__(128, 1)
    umma_f8_kernel(const uint8_t* __restrict__ a_image, int a_bytes,
                   const uint8_t* __restrict__ b_image, int b_bytes, float* __restrict__ d_out,
                   int ncols, uint32_t idesc, int nk, uint32_t a_lbo, uint32_t a_sbo,
                   uint32_t a_layout, uint32_t a_kstep, uint32_t b_lbo, uint32_t b_sbo,
                   uint32_t b_layout, uint32_t b_kstep) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sa = smem;
  uint8_t* sb = smem + kMaxA;
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + kMaxA + kMaxB);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + kMaxA + kMaxB + 16);

  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  for (int i = tid; i < a_bytes / 16; i += 128) {
    reinterpret_cast<uint4*>(sa)[i] = reinterpret_cast<const uint4*>(a_image)[i];
  }
  for (int i = tid; i < b_bytes / 16; i += 128) {
    reinterpret_cast<uint4*>(sb)[i] = reinterpret_cast<const uint4*>(b_image)[i];
  }
  fence_proxy_async_smem();
  if (tid == 0) {
    mbar_init(bar, 1);
    fence_barrier_init();
  }
  if (warp == 0) {
    tmem_alloc(tmem_slot, 256);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (tid == 0) {
    for (int k = 0; k < nk; k++) {
      const uint64_t ad = make_smem_desc(smem_u32(sa) + k * a_kstep, a_lbo, a_sbo, a_layout);
      const uint64_t bd = make_smem_desc(smem_u32(sb) + k * b_kstep, b_lbo, b_sbo, b_layout);
      umma_f8(tmem_base, ad, bd, idesc, k > 0);
    }
    umma_commit(bar);
  }
  mbar_wait(bar, 0);
  tc_fence_after();
  const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(warp * 32) << 16);
  for (int c0 = 0; c0 < ncols; c0 += 8) {
    uint32_t r[8];
    tmem_ld_x8(lane_addr + c0, r);
    tmem_wait_ld();
#pragma unroll
    for (int i = 0; i < 8; i++) {
      if (c0 + i < ncols) d_out[tid * ncols + c0 + i] = __uint_as_float(r[i]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 256);
}

// ------------------------------------------------------------------------------------------------
// UMMA issue-rate probe (diagnostics): every CTA issues `iters` K blocks of 4 x (M x N x 32) fp8
// UMMAs from resident (zero) shared-memory operands, alternating between two TMEM accumulators.
//   handshake = 0 : back-to-back issue, one commit at the end (the tensor pipe's own rate)
//   handshake = 1 : every K block is committed to a "ready" barrier, a consumer warp answers on a
//                   "drained" barrier, the issuer waits for it before reusing that accumulator
//                   (the grouped GEMM's per-K-block protocol without any data movement)
// kPair: cta_group::2 (M = 256 over a 2-CTA cluster; each CTA holds 128 rows of A and N/2 of B).
// Reports the leader's clock64 span per CTA.
// ------------------------------------------------------------------------------------------------
template <bool kPair>
__global__ void __launch_bounds__(128, 1)
    umma_rate_kernel(int n, int iters, int handshake, long long* __restrict__ cycles) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sa = smem;                 // 128 rows x 128 B
  uint8_t* sb = smem + 16384;         // up to 256 rows x 128 B
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 16384 + 32768);
  uint64_t* done = bars;              // final commit
  uint64_t* ready = bars + 1;         // [2]
  uint64_t* drained = bars + 3;       // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 5);
  const int tid = threadIdx.x, warp = tid >> 5;
  const uint32_t crank = kPair ? cluster_ctarank() : 0u;
  for (int i = tid; i < (16384 + 32768) / 16; i += 128) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  fence_proxy_async_smem();
  if (tid == 0) {
    mbar_init(done, 1);
    for (int i = 0; i < 2; i++) {
      mbar_init(&ready[i], 1);
      mbar_init(&drained[i], kPair ? 2 : 1);
    }
    fence_barrier_init();
  }
  if (warp == 0) {
    if constexpr (kPair) {
      tmem_alloc_2cta(tmem_slot, 512);
      tmem_relinquish_2cta();
    } else {
      tmem_alloc(tmem_slot, 512);
      tmem_relinquish();
    }
  }
  tc_fence_before();
  if constexpr (kPair) {
    cluster_sync_all();
  } else {
    __syncthreads();
  }
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t idesc = make_idesc(kPair ? 256 : 128, n, kFmtE4M3, kFmtE4M3, 0, 0);
  if (tid == 0 && crank == 0) {
    const uint64_t ad = make_smem_desc(smem_u32(sa), 16, 1024, kLayoutSW128);
    const uint64_t bd = make_smem_desc(smem_u32(sb), 16, 1024, kLayoutSW128);
    const long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
      const uint32_t buf = it & 1;
      if (handshake && it >= 2) {
        if constexpr (kPair) {
          mbar_wait_cluster(&drained[buf], ((it >> 1) & 1) ^ 1);
        } else {
          mbar_wait(&drained[buf], ((it >> 1) & 1) ^ 1);
        }
        tc_fence_after();
      }
#pragma unroll
      for (int k = 0; k < 4; k++) {
        if constexpr (kPair) {
          umma_f8_2cta(tmem_base + buf * 256, ad + k * 2, bd + k * 2, idesc, k > 0);
        } else {
          umma_f8(tmem_base + buf * 256, ad + k * 2, bd + k * 2, idesc, k > 0);
        }
      }
      if (handshake) {
        if constexpr (kPair) {
          umma_commit_2cta(&ready[buf], 3);
        } else {
          umma_commit(&ready[buf]);
        }
      }
    }
    if constexpr (kPair) {
      umma_commit_2cta(done, 3);
    } else {
      umma_commit(done);
    }
    mbar_wait(done, 0);
    cycles[blockIdx.x] = clock64() - t0;
  } else if (warp == 1 && handshake) {
    // consumer: answer every "ready" with a "drained" (lane 0), like an epilogue warp would
    for (int it = 0; it < iters; it++) {
      const uint32_t buf = it & 1;
      mbar_wait(&ready[buf], (it >> 1) & 1);
      tc_fence_after();
      tc_fence_before();
      __syncwarp();
      if ((tid & 31) == 0) {
        if constexpr (kPair) {
          mbar_arrive_cluster(map_to_cta(smem_u32(&drained[buf]), 0));
        } else {
          mbar_arrive(&drained[buf]);
        }
      }
    }
  }
  if (kPair && crank == 1 && tid == 0) mbar_wait(done, 0);
  tc_fence_before();
  if constexpr (kPair) {
    cluster_sync_all();
  } else {
    __syncthreads();
  }
  if (warp == 0) {
    if constexpr (kPair) {
      tmem_dealloc_2cta(tmem_base, 512);
    } else {
      tmem_dealloc(tmem_base, 512);
    }
  }
}

}  // namespace selftest
}  // namespace b200

using namespace b200;  // NOLINT

extern "C" int hpc_selftest_umma_f8(const void* a_image, int a_bytes, const void* b_image,
                                    int b_bytes, float* d_out, int ncols, uint32_t idesc, int nk,
                                    uint32_t a_lbo, uint32_t a_sbo, uint32_t a_layout,
                                    uint32_t a_kstep, uint32_t b_lbo, uint32_t b_sbo,
                                    uint32_t b_layout, uint32_t b_kstep, cudaStream_t stream) {
  HPC_REQUIRE(a_bytes > 0 && a_bytes <= selftest::kMaxA && a_bytes % 16 == 0, "bad a_bytes");
  HPC_REQUIRE(b_bytes > 0 && b_bytes <= selftest::kMaxB && b_bytes % 16 == 0, "bad b_bytes");
  HPC_REQUIRE(ncols > 0 && ncols <= 256, "bad ncols");
  const int smem = selftest::kMaxA + selftest::kMaxB + 64;
  static bool configured = false;
  if (!configured) {
    HPC_CUDA_CHECK(cudaFuncSetAttribute(selftest::umma_f8_kernel,
                                        cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured = true;
  }
  selftest::umma_f8_kernel<<<1, 128, smem, stream>>>(
      static_cast<const uint8_t*>(a_image), a_bytes, static_cast<const uint8_t*>(b_image), b_bytes,
      d_out, ncols, idesc, nk, a_lbo, a_sbo, a_layout, a_kstep, b_lbo, b_sbo, b_layout, b_kstep);
  HPC_CUDA_CHECK(cudaGetLastError());
  return HPC_OK;
}

// diagnostics: UMMA issue rate. pair = 0 / 1 (cta_group::1 / ::2), n = MMA N (16..256),
// cycles_out: device int64[grid]. grid = SM count (rounded down to even for pairs).
extern "C" int hpc_selftest_umma_rate(int pair, int n, int iters, int handshake,
                                      long long* cycles_out, cudaStream_t stream) {
  HPC_REQUIRE(n >= 16 && n <= 256 && n % 16 == 0, "bad n");
  const int smem = 16384 + 32768 + 128;
  const int grid = pair ? (sm_count() / 2) * 2 : sm_count();
  if (pair) {
    HPC_CUDA_CHECK(cudaFuncSetAttribute(selftest::umma_rate_kernel<true>,
                                        cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid, 1, 1);
    cfg.blockDim = dim3(128, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    HPC_CUDA_CHECK(cudaLaunchKernelEx(&cfg, selftest::umma_rate_kernel<true>, n, iters, handshake,
                                      cycles_out));
  } else {
    HPC_CUDA_CHECK(cudaFuncSetAttribute(selftest::umma_rate_kernel<false>,
                                        cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    selftest::umma_rate_kernel<false><<<grid, 128, smem, stream>>>(n, iters, handshake, cycles_out);
    HPC_CUDA_CHECK(cudaGetLastError());
  }
  return HPC_OK;
}
