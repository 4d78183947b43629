#include "host_utils.h"

#include <cstdlib>
#include <cstring>
#include <mutex>

namespace b200 {

static thread_local char g_last_error[1024] = {0};

void set_last_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_last_error, sizeof(g_last_error), fmt, ap);
  va_end(ap);
}

int sm_count() {
  // cached per device (B200: 148)
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 0;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return 0;
    cached[dev] = n;
  }
  return cached[dev];
}

bool pdl_enabled() {
  static const bool on = [] {
    const char* e = std::getenv("HPC_B200_PDL");
    return !(e != nullptr && e[0] == '0');
  }();
  return on;
}

int device_slot() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0) return 0;
  return dev < 64 ? dev : 63;
}

int* scheduler_counter(cudaStream_t stream) {
  constexpr int kSlots = 128;
  struct Entry {
    cudaStream_t stream;
    bool used;
  };
  static int* pool[64] = {nullptr};
  static Entry table[64][kSlots] = {};
  static std::mutex mu;
  const int dev = device_slot();
  std::lock_guard<std::mutex> lock(mu);
  if (pool[dev] == nullptr) {
    if (cudaMalloc(&pool[dev], kSlots * 2 * sizeof(int)) != cudaSuccess ||
        cudaMemset(pool[dev], 0, kSlots * 2 * sizeof(int)) != cudaSuccess) {
      set_last_error("scheduler_counter: allocation failed: %s", cudaGetErrorString(cudaGetLastError()));
      pool[dev] = nullptr;
      return nullptr;
    }
  }
  int free_slot = -1;
  for (int i = 0; i < kSlots; i++) {
    if (table[dev][i].used && table[dev][i].stream == stream) return pool[dev] + 2 * i;
    if (!table[dev][i].used && free_slot < 0) free_slot = i;
  }
  if (free_slot < 0) free_slot = static_cast<int>((reinterpret_cast<uintptr_t>(stream) >> 4) % kSlots);
  table[dev][free_slot].stream = stream;
  table[dev][free_slot].used = true;
  return pool[dev] + 2 * free_slot;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) ==
            cudaSuccess &&
        qres == cudaDriverEntryPointSuccess) {
      fn = reinterpret_cast<EncodeTiledFn>(p);
    }
  });
  return fn;
}

int encode_tmap(CUtensorMap* out, CUtensorMapDataType dtype, int elem_bytes, const void* base,
                int rank, const uint64_t* dims, const uint64_t* strides_bytes, const uint32_t* box,
                CUtensorMapSwizzle swizzle, CUtensorMapL2promotion promo) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) {
    set_last_error("cuTensorMapEncodeTiled entry point unavailable (no CUDA driver?)");
    return HPC_ERR_DRIVER;
  }
  cuuint64_t gdim[5];
  cuuint64_t gstr[5];
  cuuint32_t bx[5];
  cuuint32_t es[5];
  for (int i = 0; i < rank; i++) {
    gdim[i] = dims[i];
    bx[i] = box[i];
    es[i] = 1;
    if (i > 0) gstr[i - 1] = strides_bytes[i - 1];
  }
  (void)elem_bytes;
  CUresult r = fn(out, dtype, static_cast<cuuint32_t>(rank), const_cast<void*>(base), gdim, gstr,
                  bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle, promo,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_last_error(
        "cuTensorMapEncodeTiled failed (%d): rank %d dims [%llu %llu %llu %llu] strides [%llu %llu "
        "%llu] box [%u %u %u %u] base %p",
        static_cast<int>(r), rank, (unsigned long long)dims[0],
        (unsigned long long)(rank > 1 ? dims[1] : 0), (unsigned long long)(rank > 2 ? dims[2] : 0),
        (unsigned long long)(rank > 3 ? dims[3] : 0),
        (unsigned long long)(rank > 1 ? strides_bytes[0] : 0),
        (unsigned long long)(rank > 2 ? strides_bytes[1] : 0),
        (unsigned long long)(rank > 3 ? strides_bytes[2] : 0), box[0], rank > 1 ? box[1] : 0,
        rank > 2 ? box[2] : 0, rank > 3 ? box[3] : 0, base);
    return HPC_ERR_DRIVER;
  }
  return HPC_OK;
}

int encode_tmap_u8(CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                   const uint64_t* strides_bytes, const uint32_t* box, CUtensorMapSwizzle swizzle,
                   CUtensorMapL2promotion promo) {
  return encode_tmap(out, CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, base, rank, dims, strides_bytes, box,
                     swizzle, promo);
}

}  // namespace b200

extern "C" const char* hpc_last_error() { return b200::g_last_error; }

extern "C" int hpc_sm_count() { return b200::sm_count(); }
