// Stand-alone activation / quantisation helpers of the MoE path (B200 / sm_100a) — pure HBM-bound
// streaming kernels, 16-byte vector loads, grid = multiple of the SM count.
//
//   act_mul_and_quant : out[r, c] = e4m3( silu(gate[r, c]) * up[r, c] * scale )      gate_up bf16 [R, 2C]
//                       (reference src/activation/activation.cu:19-136, launcher :528-625; in the
//                       fused MoE pipeline this work is the epilogue of the Gate-Up GEMM instead)
//   scaled_fp8_quant  : out[i] = e4m3( in[i] * (1 / scale) )                          in f32 / f16 / bf16
//                       (reference src/activation/activation.cu:461-500, launcher :783-804)
//   gather_rows       : dst[i, :] = src[row_indices[i], :]   (the A-operand gather of the reference's
//                       scatter grouped GEMM, src/group_gemm/cp_async/group_gemm_fp8_scatter.cu:20-46,
//                       done here as a streaming pre-pass in front of the TMA-fed grouped GEMM)
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "common.cuh"
#include "group_gemm.h"
#include "host_utils.h"

namespace b200 {
namespace act {

constexpr int kThreads = 256;

__device__ __forceinline__ float silu_exact(float x) { return x / (1.f + expf(-x)); }

// one thread = 8 output columns of one row: 16 B of gate + 16 B of up in, 8 B out
__global__ void __launch_bounds__(kThreads)
    act_mul_and_quant_kernel(uint8_t* __restrict__ out, const __nv_bfloat16* __restrict__ in,
                             const float* __restrict__ scale_ptr, long long num_row, int half_col,
                             int use_bf16_mul) {
  const float scale = __ldg(scale_ptr);
  const int vec_per_row = half_col >> 3;
  const long long total = num_row * vec_per_row;
  for (long long v = static_cast<long long>(blockIdx.x) * kThreads + threadIdx.x; v < total;
       v += static_cast<long long>(gridDim.x) * kThreads) {
    const long long r = v / vec_per_row;
    const int c = static_cast<int>(v - r * vec_per_row) << 3;
    const __nv_bfloat16* g = in + r * (2LL * half_col) + c;
    const uint4 gv = ld_nc_v4(g);
    const uint4 uv = ld_nc_v4(g + half_col);
    const __nv_bfloat162* g2 = reinterpret_cast<const __nv_bfloat162*>(&gv);
    const __nv_bfloat162* u2 = reinterpret_cast<const __nv_bfloat162*>(&uv);
    float m[8];
#pragma unroll
    for (int i = 0; i < 4; i++) {
      const float2 gf = __bfloat1622float2(g2[i]);
      const float s0 = silu_exact(gf.x), s1 = silu_exact(gf.y);
      if (use_bf16_mul) {
        // bf16(silu(gate)) * up with a bf16-rounded product, as the reference kernel and its test do
        const __nv_bfloat162 p = __hmul2(__floats2bfloat162_rn(s0, s1), u2[i]);
        const float2 pf = __bfloat1622float2(p);
        m[2 * i] = pf.x;
        m[2 * i + 1] = pf.y;
      } else {
        const float2 uf = __bfloat1622float2(u2[i]);
        m[2 * i] = s0 * uf.x;
        m[2 * i + 1] = s1 * uf.y;
      }
    }
    uint2 w;
    w.x = cvt_e4m3x4(m[0] * scale, m[1] * scale, m[2] * scale, m[3] * scale);
    w.y = cvt_e4m3x4(m[4] * scale, m[5] * scale, m[6] * scale, m[7] * scale);
    *reinterpret_cast<uint2*>(out + r * half_col + c) = w;
  }
}

template <typename T>
__device__ __forceinline__ float to_f32(T v);
template <>
__device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <>
__device__ __forceinline__ float to_f32<__half>(__half v) { return __half2float(v); }
template <>
__device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }

// one thread = one 16-byte input vector (4 floats or 8 halves)
template <typename T>
__global__ void __launch_bounds__(kThreads)
    scaled_fp8_quant_kernel(uint8_t* __restrict__ out, const T* __restrict__ in,
                            const float* __restrict__ scale_ptr, long long numel) {
  constexpr int kVec = 16 / sizeof(T);
  const float inv = 1.0f / __ldg(scale_ptr);
  const long long nvec = numel / kVec;
  for (long long v = static_cast<long long>(blockIdx.x) * kThreads + threadIdx.x; v < nvec;
       v += static_cast<long long>(gridDim.x) * kThreads) {
    const uint4 raw = ld_nc_v4(in + v * kVec);
    const T* e = reinterpret_cast<const T*>(&raw);
    float f[kVec];
#pragma unroll
    for (int i = 0; i < kVec; i++) f[i] = to_f32<T>(e[i]) * inv;
    uint32_t w[kVec / 4];
#pragma unroll
    for (int i = 0; i < kVec / 4; i++) w[i] = cvt_e4m3x4(f[4 * i], f[4 * i + 1], f[4 * i + 2], f[4 * i + 3]);
    if constexpr (kVec == 4) {
      *reinterpret_cast<uint32_t*>(out + v * 4) = w[0];
    } else {
      *reinterpret_cast<uint2*>(out + v * 8) = make_uint2(w[0], w[1]);
    }
  }
  // ragged tail (numel not a multiple of the vector width)
  if (blockIdx.x == 0) {
    for (long long i = nvec * kVec + threadIdx.x; i < numel; i += kThreads) {
      const float f = to_f32<T>(in[i]) * inv;
      out[i] = static_cast<uint8_t>(cvt_e4m3x4(f, 0.f, 0.f, 0.f) & 0xffu);
    }
  }
}

// one warp per destination row, 16-byte copies
__global__ void __launch_bounds__(kThreads)
    gather_rows_kernel(uint4* __restrict__ dst, const uint4* __restrict__ src,
                       const int* __restrict__ row_indices, int num_rows, int src_rows,
                       int vec_per_row) {
  const int warps_per_block = kThreads / 32;
  const int lane = threadIdx.x & 31;
  for (int r = blockIdx.x * warps_per_block + (threadIdx.x >> 5); r < num_rows;
       r += gridDim.x * warps_per_block) {
    int s = __ldg(row_indices + r);
    s = s < 0 ? 0 : (s >= src_rows ? src_rows - 1 : s);  // never read outside the pool
    const uint4* sp = src + static_cast<long long>(s) * vec_per_row;
    uint4* dp = dst + static_cast<long long>(r) * vec_per_row;
    for (int c = lane; c < vec_per_row; c += 32) dp[c] = ld_nc_v4(sp + c);
  }
}

static int stream_grid(long long work_items) {
  int sms = sm_count();
  if (sms <= 0) sms = 148;
  long long blocks = (work_items + kThreads - 1) / kThreads;
  const long long cap = static_cast<long long>(sms) * 8;  // 8 resident blocks of 256 threads per SM
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return static_cast<int>(blocks);
}

}  // namespace act
}  // namespace b200

using namespace b200;  // NOLINT

// replaces reference src/activation/activation.h:15-17 (act_mul_and_quant_async, bf16 in, e4m3 out).
// num_col = 2 * C (gate columns then up columns).
extern "C" int hpc_act_mul_and_quant_async(void* y_ptr, const void* x_ptr, const float* scale_ptr,
                                           int num_row, int num_col, int use_bf16_mul,
                                           cudaStream_t stream) {
  HPC_REQUIRE(num_col > 0 && num_col % 16 == 0,
              "act_mul_and_quant: last dim (%d) must be a multiple of 16", num_col);
  HPC_REQUIRE((reinterpret_cast<uintptr_t>(x_ptr) & 15) == 0 &&
                  (reinterpret_cast<uintptr_t>(y_ptr) & 7) == 0,
              "act_mul_and_quant: input must be 16-byte and output 8-byte aligned");
  if (num_row <= 0) return HPC_OK;
  const int half = num_col / 2;
  const long long items = static_cast<long long>(num_row) * (half / 8);
  act::act_mul_and_quant_kernel<<<act::stream_grid(items), act::kThreads, 0, stream>>>(
      static_cast<uint8_t*>(y_ptr), static_cast<const __nv_bfloat16*>(x_ptr), scale_ptr, num_row,
      half, use_bf16_mul);
  HPC_CUDA_CHECK(cudaGetLastError());
  return HPC_OK;
}

// replaces reference src/activation/activation.h:46-53 (scaled_fp8_quant_async overloads).
// in_dtype: 0 = float32, 1 = float16, 2 = bfloat16.
extern "C" int hpc_scaled_fp8_quant_async(void* y_ptr, const void* x_ptr, const float* scale_ptr,
                                          int64_t numel, int in_dtype, cudaStream_t stream) {
  HPC_REQUIRE(in_dtype >= 0 && in_dtype <= 2, "scaled_fp8_quant: bad input dtype code %d", in_dtype);
  HPC_REQUIRE((reinterpret_cast<uintptr_t>(x_ptr) & 15) == 0 &&
                  (reinterpret_cast<uintptr_t>(y_ptr) & 7) == 0,
              "scaled_fp8_quant: input must be 16-byte and output 8-byte aligned");
  if (numel <= 0) return HPC_OK;
  uint8_t* y = static_cast<uint8_t*>(y_ptr);
  if (in_dtype == 0) {
    act::scaled_fp8_quant_kernel<float><<<act::stream_grid(numel / 4), act::kThreads, 0, stream>>>(
        y, static_cast<const float*>(x_ptr), scale_ptr, numel);
  } else if (in_dtype == 1) {
    act::scaled_fp8_quant_kernel<__half><<<act::stream_grid(numel / 8), act::kThreads, 0, stream>>>(
        y, static_cast<const __half*>(x_ptr), scale_ptr, numel);
  } else {
    act::scaled_fp8_quant_kernel<__nv_bfloat16>
        <<<act::stream_grid(numel / 8), act::kThreads, 0, stream>>>(
            y, static_cast<const __nv_bfloat16*>(x_ptr), scale_ptr, numel);
  }
  HPC_CUDA_CHECK(cudaGetLastError());
  return HPC_OK;
}

// replaces reference src/group_gemm/cp_async/group_gemm.h:11-16 (group_gemm_fp8_multistage_async):
// the small-M cp.async kernel of the reference is served by the same tcgen05 grouped GEMM
// (per-group scalar y_scale). tiles / cu_tiles / task_map: accepted, ignored.
extern "C" int hpc_group_gemm_fp8_multistage_async(
    void* y_ptr, const void* x_ptr, const void* w_ptr, const void* y_scale_ptr,
    const void* seqlens_ptr, const void* cu_seqlens_ptr, const void* tiles_ptr,
    const void* cu_tiles_ptr, const void* task_map_ptr, int task_map_len, int m, int n, int k,
    int num_group, int num_seq_per_group_avg, int use_pdl, cudaStream_t stream) {
  (void)tiles_ptr; (void)cu_tiles_ptr; (void)task_map_ptr; (void)task_map_len; (void)use_pdl;
  return ggemm::run(0, x_ptr, w_ptr, static_cast<const int*>(seqlens_ptr),
                    static_cast<const int*>(cu_seqlens_ptr), nullptr,
                    static_cast<const float*>(y_scale_ptr), nullptr, y_ptr, nullptr, nullptr,
                    num_group, m, n, k, 0, 0, ggemm::scale_tile_from_avg(num_seq_per_group_avg), 0,
                    stream);
}

// replaces reference src/group_gemm/cp_async/group_gemm.h:18-24 (group_gemm_fp8_scatter_async):
// row i of the compact problem is row row_indices[i] of the pool x [pool_rows, k]. `gather_ptr`
// is an e4m3 scratch of m * k bytes (extra argument: the reference gathers inside its kernel).
extern "C" int hpc_group_gemm_fp8_scatter_async(
    void* y_ptr, const void* x_ptr, const void* w_ptr, const void* y_scale_ptr,
    const void* row_indices_ptr, const void* seqlens_ptr, const void* cu_seqlens_ptr,
    const void* tiles_ptr, const void* cu_tiles_ptr, const void* task_map_ptr, int task_map_len,
    int m, int n, int k, int num_group, int num_seq_per_group_avg, int use_pdl, void* gather_ptr,
    int pool_rows, cudaStream_t stream) {
  HPC_REQUIRE(k > 0 && k % 16 == 0, "scatter group gemm: k (%d) must be a multiple of 16", k);
  HPC_REQUIRE(gather_ptr != nullptr && pool_rows > 0, "scatter group gemm: gather scratch / pool rows missing");
  HPC_REQUIRE((reinterpret_cast<uintptr_t>(x_ptr) & 15) == 0 &&
                  (reinterpret_cast<uintptr_t>(gather_ptr) & 15) == 0,
              "scatter group gemm: x and scratch must be 16-byte aligned");
  if (m <= 0) return HPC_OK;
  const int warps_per_block = act::kThreads / 32;
  int blocks = (m + warps_per_block - 1) / warps_per_block;
  int sms = sm_count();
  if (sms <= 0) sms = 148;
  if (blocks > sms * 8) blocks = sms * 8;
  act::gather_rows_kernel<<<blocks, act::kThreads, 0, stream>>>(
      static_cast<uint4*>(gather_ptr), static_cast<const uint4*>(x_ptr),
      static_cast<const int*>(row_indices_ptr), m, pool_rows, k / 16);
  HPC_CUDA_CHECK(cudaGetLastError());
  return hpc_group_gemm_fp8_multistage_async(y_ptr, gather_ptr, w_ptr, y_scale_ptr, seqlens_ptr,
                                             cu_seqlens_ptr, tiles_ptr, cu_tiles_ptr, task_map_ptr,
                                             task_map_len, m, n, k, num_group,
                                             num_seq_per_group_avg, use_pdl, stream);
}
