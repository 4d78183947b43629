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
// Current CUDA device clamped to [0, 63] (0 when the runtime is unavailable). Launchers keep their
// one-time state (function attributes, scratch pools) per device: one process may drive several GPUs.
int device_slot();
// A zeroed int32 in device memory for one launch (dynamic tile / work counters): taken from a
// rotating per-device pool of 256 and cleared with cudaMemsetAsync on `stream`. The pool is
// allocated on the first call per device, which therefore must not happen inside a stream capture.
int* launch_counter(cudaStream_t stream);

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

}  // namespace b200
