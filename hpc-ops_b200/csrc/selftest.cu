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
