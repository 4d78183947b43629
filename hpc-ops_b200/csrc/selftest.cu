// Bring-up self test for the tcgen05 descriptor conventions used by the product kernels.
// One CTA: copy caller-supplied smem images of A and B, issue `nk` fp8 UMMAs with caller-supplied
// descriptor fields, read D back from TMEM. tests/ sweep these fields against a CPU matmul.
#include "common.cuh"
#include "host_utils.h"

namespace b200 {
namespace selftest {

constexpr int kMaxA = 65536;
constexpr int kMaxB = 32768;

// kF16: kind::f16 (bf16 operands) instead of kind::f8f6f4. K step k reads its operands at
// (k / nk_inner) * kstep2 + (k % nk_inner) * kstep (two-level walk: 64-element swizzle atoms).
template <bool kF16>
__global__ void __launch_bounds__(128, 1)
    umma_kernel(const uint8_t* __restrict__ a_image, int a_bytes,
                   const uint8_t* __restrict__ b_image, int b_bytes, float* __restrict__ d_out,
                   int ncols, uint32_t idesc, int nk, uint32_t a_lbo, uint32_t a_sbo,
                   uint32_t a_layout, uint32_t a_kstep, uint32_t b_lbo, uint32_t b_sbo,
                   uint32_t b_layout, uint32_t b_kstep, int nk_inner, uint32_t a_kstep2,
                   uint32_t b_kstep2) {
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
      const uint32_t ko = k / nk_inner, ki = k % nk_inner;
      const uint64_t ad =
          make_smem_desc(smem_u32(sa) + ko * a_kstep2 + ki * a_kstep, a_lbo, a_sbo, a_layout);
      const uint64_t bd =
          make_smem_desc(smem_u32(sb) + ko * b_kstep2 + ki * b_kstep, b_lbo, b_sbo, b_layout);
      if constexpr (kF16) {
        umma_f16(tmem_base, ad, bd, idesc, k > 0);
      } else {
        umma_f8(tmem_base, ad, bd, idesc, k > 0);
      }
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
// UMMAs from resident shared-memory operands, alternating between two TMEM accumulators, with the
// grouped GEMM's pipeline protocol added piece by piece (`flags`):
//   1  every K block is committed to a "ready" barrier, consumer warps answer on a "drained"
//      barrier, the issuer waits for it before reusing that accumulator
//   2  operand-stage protocol: a producer thread hands out "full" stages (no data is moved), the
//      issuer waits for them and releases each with a second tcgen05.commit ("empty")
//   4  eight consumer warps (the two epilogue warpgroups) instead of one
//   8  pseudo-random operand bytes instead of zeros (realistic switching power)
//   16 consumers drain the accumulator: 4 x tcgen05.ld 32x32b.x32 + 64 FFMA2 per K block (needs 4)
// kPair: cta_group::2 (M = 256 over a 2-CTA cluster; each CTA holds 128 rows of A and N/2 of B).
// Reports the issuer's clock64 span and the globaltimer span (-> effective SM clock) per CTA.
// ------------------------------------------------------------------------------------------------
constexpr int kProbeThreads = 384;
constexpr int kProbeStages = 4;

template <bool kPair>
__global__ void __launch_bounds__(kProbeThreads, 1)
    umma_rate_kernel(int n, int iters, int flags, long long* __restrict__ cycles,
                     float* __restrict__ sink) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sa = smem;                 // 128 rows x 128 B
  uint8_t* sb = smem + 16384;         // up to 256 rows x 128 B
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 16384 + 32768);
  uint64_t* done = bars;              // final commit
  uint64_t* ready = bars + 1;         // [2]
  uint64_t* drained = bars + 3;       // [2]
  uint64_t* full = bars + 5;          // [kProbeStages]
  uint64_t* empty = bars + 5 + kProbeStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 5 + 2 * kProbeStages);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t crank = kPair ? cluster_ctarank() : 0u;
  const bool hs = flags & 1, stage_proto = flags & 2, wide = flags & 4, rnd = flags & 8, drain = flags & 16;
  const int ncons = wide ? 8 : 1;
  for (int i = tid; i < (16384 + 32768) / 16; i += kProbeThreads) {
    uint4 v = make_uint4(0, 0, 0, 0);
    if (rnd) {
      // e4m3 bytes with exponent field <= 9 (|x| <= 7.5): finite, no NaN pattern (0x7f / 0xff)
      uint32_t h = (i * 4 + blockIdx.x * 7919u) * 2654435761u;
      uint32_t w[4];
      for (int j = 0; j < 4; j++) {
        h = h * 1664525u + 1013904223u;
        w[j] = (h >> 3) & 0xCFCFCFCFu;
      }
      v = make_uint4(w[0], w[1], w[2], w[3]);
    }
    reinterpret_cast<uint4*>(smem)[i] = v;
  }
  fence_proxy_async_smem();
  if (tid == 0) {
    mbar_init(done, 1);
    for (int i = 0; i < 2; i++) {
      mbar_init(&ready[i], 1);
      mbar_init(&drained[i], (kPair ? 2 : 1) * ncons);
    }
    for (int i = 0; i < kProbeStages; i++) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
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
  if (warp == 1 && lane == 0 && crank == 0) {
    const uint64_t ad = make_smem_desc(smem_u32(sa), 16, 1024, kLayoutSW128);
    const uint64_t bd = make_smem_desc(smem_u32(sb), 16, 1024, kLayoutSW128);
    unsigned long long g0, g1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g0));
    const long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
      const uint32_t buf = it & 1;
      const uint32_t st = it % kProbeStages;
      if (stage_proto) mbar_wait(&full[st], (it / kProbeStages) & 1);
      if (hs && it >= 2) {
        if constexpr (kPair) {
          mbar_wait_cluster(&drained[buf], ((it >> 1) & 1) ^ 1);
        } else {
          mbar_wait(&drained[buf], ((it >> 1) & 1) ^ 1);
        }
      }
      tc_fence_after();
#pragma unroll
      for (int k = 0; k < 4; k++) {
        if constexpr (kPair) {
          umma_f8_2cta(tmem_base + buf * 256, ad + k * 2, bd + k * 2, idesc, k > 0);
        } else {
          umma_f8(tmem_base + buf * 256, ad + k * 2, bd + k * 2, idesc, k > 0);
        }
      }
      if (stage_proto) {
        if constexpr (kPair) {
          umma_commit_2cta(&empty[st], 1);
        } else {
          umma_commit(&empty[st]);
        }
      }
      if (hs) {
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
    const long long t1 = clock64();
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g1));
    cycles[2 * blockIdx.x] = t1 - t0;
    cycles[2 * blockIdx.x + 1] = static_cast<long long>(g1 - g0);
  } else if (warp == 0 && lane == 0 && crank == 0 && stage_proto) {
    for (int it = 0; it < iters; it++) {
      const uint32_t st = it % kProbeStages;
      mbar_wait(&empty[st], ((it / kProbeStages) & 1) ^ 1);
      mbar_arrive(&full[st]);
    }
  } else if (warp >= 4 && warp < 4 + ncons && hs) {
    // consumers: answer every "ready" with a "drained", optionally after draining the accumulator
    const int quad = warp & 3;
    const int wg = (warp - 4) >> 2;
    const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16);
    float2 acc[64];
    if (drain) {
#pragma unroll
      for (int i = 0; i < 64; i++) acc[i] = make_float2(0.f, 0.f);
    }
    for (int it = 0; it < iters; it++) {
      const uint32_t buf = it & 1;
      mbar_wait(&ready[buf], (it >> 1) & 1);
      tc_fence_after();
      if (drain) {
        const uint64_t ff = pack_f2(1.0009765625f, 1.0009765625f);
        const uint32_t base = lane_addr + buf * 256 + wg * 64;
        uint32_t ra[32], rb[32];
        tmem_ld_x32(base, ra);
        tmem_ld_x32(base + 32, rb);
        tmem_wait_ld();
#pragma unroll
        for (int i = 0; i < 16; i++) {
          uint64_t& a = reinterpret_cast<uint64_t&>(acc[i]);
          a = ffma2(pack_f2(__uint_as_float(ra[2 * i]), __uint_as_float(ra[2 * i + 1])), ff, a);
        }
        tmem_ld_x32(base + 128, ra);
#pragma unroll
        for (int i = 0; i < 16; i++) {
          uint64_t& a = reinterpret_cast<uint64_t&>(acc[16 + i]);
          a = ffma2(pack_f2(__uint_as_float(rb[2 * i]), __uint_as_float(rb[2 * i + 1])), ff, a);
        }
        tmem_ld_x32(base + 128 + 32, rb);
        tmem_wait_ld();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if constexpr (kPair) {
            mbar_arrive_cluster(map_to_cta(smem_u32(&drained[buf]), 0));
          } else {
            mbar_arrive(&drained[buf]);
          }
        }
#pragma unroll
        for (int i = 0; i < 16; i++) {
          uint64_t& a = reinterpret_cast<uint64_t&>(acc[32 + i]);
          a = ffma2(pack_f2(__uint_as_float(ra[2 * i]), __uint_as_float(ra[2 * i + 1])), ff, a);
        }
#pragma unroll
        for (int i = 0; i < 16; i++) {
          uint64_t& a = reinterpret_cast<uint64_t&>(acc[48 + i]);
          a = ffma2(pack_f2(__uint_as_float(rb[2 * i]), __uint_as_float(rb[2 * i + 1])), ff, a);
        }
      } else {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if constexpr (kPair) {
            mbar_arrive_cluster(map_to_cta(smem_u32(&drained[buf]), 0));
          } else {
            mbar_arrive(&drained[buf]);
          }
        }
      }
    }
    if (drain) {
      float t = 0.f;
#pragma unroll
      for (int i = 0; i < 64; i++) t += acc[i].x + acc[i].y;
      if (t == 12345.678f) sink[tid] = t;  // keeps the accumulation alive
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

template <bool kF16>
static int run_umma_selftest(const void* a_image, int a_bytes, const void* b_image, int b_bytes,
                             float* d_out, int ncols, uint32_t idesc, int nk, uint32_t a_lbo,
                             uint32_t a_sbo, uint32_t a_layout, uint32_t a_kstep, uint32_t b_lbo,
                             uint32_t b_sbo, uint32_t b_layout, uint32_t b_kstep, int nk_inner,
                             uint32_t a_kstep2, uint32_t b_kstep2, cudaStream_t stream) {
  HPC_REQUIRE(a_bytes > 0 && a_bytes <= selftest::kMaxA && a_bytes % 16 == 0, "bad a_bytes");
  HPC_REQUIRE(b_bytes > 0 && b_bytes <= selftest::kMaxB && b_bytes % 16 == 0, "bad b_bytes");
  HPC_REQUIRE(ncols > 0 && ncols <= 256, "bad ncols");
  HPC_REQUIRE(nk > 0 && nk_inner > 0, "bad nk");
  const int smem = selftest::kMaxA + selftest::kMaxB + 64;
  static bool configured = false;
  if (!configured) {
    HPC_CUDA_CHECK(cudaFuncSetAttribute(selftest::umma_kernel<kF16>,
                                        cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured = true;
  }
  selftest::umma_kernel<kF16><<<1, 128, smem, stream>>>(
      static_cast<const uint8_t*>(a_image), a_bytes, static_cast<const uint8_t*>(b_image), b_bytes,
      d_out, ncols, idesc, nk, a_lbo, a_sbo, a_layout, a_kstep, b_lbo, b_sbo, b_layout, b_kstep,
      nk_inner, a_kstep2, b_kstep2);
  HPC_CUDA_CHECK(cudaGetLastError());
  return HPC_OK;
}

extern "C" int hpc_selftest_umma_f8(const void* a_image, int a_bytes, const void* b_image,
                                    int b_bytes, float* d_out, int ncols, uint32_t idesc, int nk,
                                    uint32_t a_lbo, uint32_t a_sbo, uint32_t a_layout,
                                    uint32_t a_kstep, uint32_t b_lbo, uint32_t b_sbo,
                                    uint32_t b_layout, uint32_t b_kstep, cudaStream_t stream) {
  return run_umma_selftest<false>(a_image, a_bytes, b_image, b_bytes, d_out, ncols, idesc, nk, a_lbo,
                                  a_sbo, a_layout, a_kstep, b_lbo, b_sbo, b_layout, b_kstep, nk, 0, 0,
                                  stream);
}

// bf16 operands (kind::f16); K step k reads at (k / nk_inner) * kstep2 + (k % nk_inner) * kstep
extern "C" int hpc_selftest_umma_bf16(const void* a_image, int a_bytes, const void* b_image,
                                      int b_bytes, float* d_out, int ncols, uint32_t idesc, int nk,
                                      uint32_t a_lbo, uint32_t a_sbo, uint32_t a_layout,
                                      uint32_t a_kstep, uint32_t b_lbo, uint32_t b_sbo,
                                      uint32_t b_layout, uint32_t b_kstep, int nk_inner,
                                      uint32_t a_kstep2, uint32_t b_kstep2, cudaStream_t stream) {
  return run_umma_selftest<true>(a_image, a_bytes, b_image, b_bytes, d_out, ncols, idesc, nk, a_lbo,
                                 a_sbo, a_layout, a_kstep, b_lbo, b_sbo, b_layout, b_kstep, nk_inner,
                                 a_kstep2, b_kstep2, stream);
}

// diagnostics: UMMA issue rate. pair = 0 / 1 (cta_group::1 / ::2), n = MMA N (16..256), flags: see
// umma_rate_kernel. cycles_out: device int64[2 * grid] (clock64 span, globaltimer ns span).
extern "C" int hpc_selftest_umma_rate(int pair, int n, int iters, int flags,
                                      long long* cycles_out, cudaStream_t stream) {
  HPC_REQUIRE(n >= 16 && n <= 256 && n % 16 == 0, "bad n");
  const int smem = 16384 + 32768 + 256;
  const int grid = pair ? (sm_count() / 2) * 2 : sm_count();
  static float* sink = nullptr;
  if (sink == nullptr) HPC_CUDA_CHECK(cudaMalloc(&sink, 4096));
  if (pair) {
    HPC_CUDA_CHECK(cudaFuncSetAttribute(selftest::umma_rate_kernel<true>,
                                        cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid, 1, 1);
    cfg.blockDim = dim3(selftest::kProbeThreads, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    HPC_CUDA_CHECK(cudaLaunchKernelEx(&cfg, selftest::umma_rate_kernel<true>, n, iters, flags,
                                      cycles_out, sink));
  } else {
    HPC_CUDA_CHECK(cudaFuncSetAttribute(selftest::umma_rate_kernel<false>,
                                        cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    selftest::umma_rate_kernel<false><<<grid, selftest::kProbeThreads, smem, stream>>>(
        n, iters, flags, cycles_out, sink);
    HPC_CUDA_CHECK(cudaGetLastError());
  }
  return HPC_OK;
}
