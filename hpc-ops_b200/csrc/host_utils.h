// Host-side helpers shared by the C-ABI launchers: error reporting, device queries, TMA descriptors.
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdarg>
#include <cstdio>

#include "../../include/hpc_b200.h"

namespace b200 {

void set_last_error(const char* fmt, ...);

#define HPC_CUDA_CHECK(expr)                                                              \
  do {                                                                                    \
    cudaError_t _e = (expr);                                                              \
    if (_e != cudaSuccess) {                                                              \
      b200::set_last_error("%s:%d CUDA error %s: %s", __FILE__, __LINE__, #expr,          \
                           cudaGetErrorString(_e));                                       \
      return HPC_ERR_CUDA;                                                                \
    }                                                                                     \
  } while (0)

#define HPC_REQUIRE(cond, ...)             \
  do {                                     \
    if (!(cond)) {                         \
      b200::set_last_error(__VA_ARGS__);   \
      return HPC_ERR_UNSUPPORTED;          \
    }                                      \
  } while (0)

int sm_count();
// HPC_B200_PDL=0 turns programmatic dependent launch off (plain stream order), default on.
bool pdl_enabled();
// Current CUDA device clamped to [0, 63] (0 when the runtime is unavailable). Launchers keep their
// one-time state (function attributes, scratch pools) per device: one process may drive several GPUs.
int device_slot();
// Two zeroed int32 (work counter, finished-CTA counter) owned by `stream` for kernels that re-arm
// their scheduler themselves (the last CTA resets both): no memset between launches, so PDL chains
// and graph replays are unaffected. One slot per (device, stream), allocated and zeroed on first
// use (not inside a stream capture); launches on one stream are serialised, so the slot is never
// shared by two running kernels.
int* scheduler_counter(cudaStream_t stream);

// Encode a tiled TMA descriptor (uint8 elements). dims/strides innermost-first; strides in bytes for
// dims 1..rank-1. Returns 0 on success.
int encode_tmap_u8(CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                   const uint64_t* strides_bytes, const uint32_t* box, CUtensorMapSwizzle swizzle,
                   CUtensorMapL2promotion promo = CU_TENSOR_MAP_L2_PROMOTION_L2_128B);
// Same for 2-byte and 4-byte element types.
int encode_tmap(CUtensorMap* out, CUtensorMapDataType dtype, int elem_bytes, const void* base,
                int rank, const uint64_t* dims, const uint64_t* strides_bytes, const uint32_t* box,
                CUtensorMapSwizzle swizzle,
                CUtensorMapL2promotion promo = CU_TENSOR_MAP_L2_PROMOTION_L2_128B);

// Launch `kern` so that it may start before the previous kernel in `stream` has finished
// (programmatic dependent launch; the kernel calls pdl_wait() before touching global memory).
// `cluster_x` > 1 additionally launches thread-block clusters of that size.
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem,
                              cudaStream_t stream, int cluster_x, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  int n = 0;
  if (pdl_enabled()) {
    attr[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[n].val.programmaticStreamSerializationAllowed = 1;
    n++;
  }
  if (cluster_x > 1) {
    attr[n].id = cudaLaunchAttributeClusterDimension;
    attr[n].val.clusterDim.x = static_cast<unsigned>(cluster_x);
    attr[n].val.clusterDim.y = 1;
    attr[n].val.clusterDim.z = 1;
    n++;
  }
  cfg.attrs = attr;
  cfg.numAttrs = n;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

}  // namespace b200
