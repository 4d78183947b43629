// Fused AllReduce + residual-add + RMSNorm over NVLink-5 / NVSwitch (B200), written from scratch.
//
// High-throughput path (replaces reference src/allreduce/fuse_allreduce_rmsnorm_high_throughput.cu:15-154):
//   every rank owns a contiguous token slice. ONE kernel per rank does, per owned token row,
//     reduce   : multimem.ld_reduce (in-switch NVLS reduction, bf16x2 inputs, f32 accumulate) of the
//                row over all ranks' symmetric input buffers     -- or P2P loads from every peer
//     fuse     : + residual -> residual_out (bf16), RMS over the row, * gamma
//     broadcast: multimem.st of the normalised row into every rank's symmetric output buffer
//                                                                 -- or P2P stores to every peer
//   bracketed by a per-block cross-GPU signal-pad barrier (entry: all inputs written; exit: all
//   outputs visible), CAS 0->1 post / 1->0 consume so that CUDA-graph replays are safe.
//
// Low-latency path (replaces reference src/allreduce/fuse_allreduce_rmsnorm_low_latency.cu:16-453):
//   Lamport two-shot in ONE kernel (the reference uses two, PDL-chained): token t is owned by rank
//   t % W. Each rank scatters its row to the owner (P2P 16-B stores, -0.0 is the "not yet written"
//   sentinel), the owner reduces the W rows in rank order and broadcasts the sum (multimem.st or
//   P2P), every rank then adds the residual and normalises locally. Triple-buffered workspace with
//   clear-ahead, state in `buffer_flags` exactly as laid out by the reference test
//   (tests/test_fuse_allreduce_rmsnorm_low_latency.py:54-76).
#include "common.cuh"
#include "host_utils.h"

namespace b200 {
namespace ar {

constexpr int kMaxRanks = 16;
constexpr int kMaxVecPerThread = 4;  // 16-B vectors of a row held per thread

// ---- system-scope signalling -----------------------------------------------------------------
__device__ __forceinline__ uint32_t cas_sys_release(uint32_t* addr, uint32_t cmp, uint32_t val) {
  uint32_t old;
  asm volatile("atom.global.release.sys.cas.b32 %0, [%1], %2, %3;"
               : "=r"(old)
               : "l"(addr), "r"(cmp), "r"(val)
               : "memory");
  return old;
}
__device__ __forceinline__ uint32_t cas_sys_acquire(uint32_t* addr, uint32_t cmp, uint32_t val) {
  uint32_t old;
  asm volatile("atom.global.acquire.sys.cas.b32 %0, [%1], %2, %3;"
               : "=r"(old)
               : "l"(addr), "r"(cmp), "r"(val)
               : "memory");
  return old;
}
#ifndef B200_AR_SPIN_LIMIT
#define B200_AR_SPIN_LIMIT (1u << 28)
#endif
// Post "I arrived" into slot [block][my rank] of every peer's signal pad, then consume the W posts
// in my own pad. signal_ptrs[r] = base of rank r's pad (uint32 slots). `phase` selects one of two
// slot sets so that the entry and exit barriers of one launch never alias.
__device__ __forceinline__ void block_barrier(uint64_t* const* signal_ptrs_unused,
                                              const long long* signal_ptrs, int rank, int world,
                                              int block, int nblocks, int phase) {
  (void)signal_ptrs_unused;
  __syncthreads();
  if (threadIdx.x < world) {
    const int peer = threadIdx.x;
    const int slot_base = (phase * nblocks + block) * world;
    uint32_t* post = reinterpret_cast<uint32_t*>(signal_ptrs[peer]) + slot_base + rank;
    uint32_t spins = 0;
    while (cas_sys_release(post, 0u, 1u) != 0u) {
      if (++spins > B200_AR_SPIN_LIMIT) {
        printf("allreduce barrier post timeout rank %d block %d peer %d\n", rank, block, peer);
        __trap();
      }
    }
    uint32_t* mine = reinterpret_cast<uint32_t*>(signal_ptrs[rank]) + slot_base + peer;
    spins = 0;
    while (cas_sys_acquire(mine, 1u, 0u) != 1u) {
      if (++spins > B200_AR_SPIN_LIMIT) {
        printf("allreduce barrier wait timeout rank %d block %d peer %d\n", rank, block, peer);
        __trap();
      }
    }
  }
  __syncthreads();
}

// ---- NVLS (multimem) 16-byte accesses ----------------------------------------------------------
__device__ __forceinline__ uint4 multimem_ld_reduce_bf16x8(const void* mc_ptr) {
  uint4 r;
  asm volatile(
      "multimem.ld_reduce.relaxed.sys.global.add.acc::f32.v4.bf16x2 {%0,%1,%2,%3}, [%4];"
      : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
      : "l"(mc_ptr)
      : "memory");
  return r;
}
__device__ __forceinline__ void multimem_st_v4(void* mc_ptr, uint4 v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(mc_ptr),
               "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w)
               : "memory");
}
__device__ __forceinline__ uint4 ld_sys_v4(const void* p) {
  uint4 r;
  asm volatile("ld.global.relaxed.sys.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p)
               : "memory");
  return r;
}
__device__ __forceinline__ uint4 ld_volatile_v4(const void* p) {
  uint4 r;
  asm volatile("ld.volatile.global.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p)
               : "memory");
  return r;
}
__device__ __forceinline__ void st_sys_v4(void* p, uint4 v) {
  asm volatile("st.global.relaxed.sys.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y),
               "r"(v.z), "r"(v.w)
               : "memory");
}

__device__ __forceinline__ void unpack8(uint4 v, float* f) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
  for (int i = 0; i < 4; i++) {
    const float2 t = __bfloat1622float2(h[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}
__device__ __forceinline__ uint4 pack8(const float* f) {
  uint4 v;
  __nv_bfloat162 a = __floats2bfloat162_rn(f[0], f[1]);
  __nv_bfloat162 b = __floats2bfloat162_rn(f[2], f[3]);
  __nv_bfloat162 c = __floats2bfloat162_rn(f[4], f[5]);
  __nv_bfloat162 d = __floats2bfloat162_rn(f[6], f[7]);
  v.x = *reinterpret_cast<uint32_t*>(&a);
  v.y = *reinterpret_cast<uint32_t*>(&b);
  v.z = *reinterpret_cast<uint32_t*>(&c);
  v.w = *reinterpret_cast<uint32_t*>(&d);
  return v;
}

__device__ __forceinline__ float block_sum(float v, float* smem, int nwarps) {
  v = warp_sum_f32(v);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __syncthreads();  // smem reuse across rows
  if (lane == 0) smem[warp] = v;
  __syncthreads();
  float t = 0.f;
  for (int w = 0; w < nwarps; w++) t += smem[w];
  return t;
}

// residual add + RMSNorm of one row held as `nvec` 16-B vectors per thread.
//   sum[] in : reduced x (float), out: nothing. Writes residual_out (bf16) and returns y vectors.
__device__ __forceinline__ void fuse_row(float (*s)[8], int nvec, int vec0, int vstride, int nv_row,
                                         const __nv_bfloat16* residual_row,
                                         __nv_bfloat16* res_out_row, const __nv_bfloat16* weight,
                                         float eps, int hidden, float* smem, int nwarps,
                                         uint4* y_out) {
  float sq = 0.f;
#pragma unroll
  for (int j = 0; j < kMaxVecPerThread; j++) {
    const int v = vec0 + j * vstride;
    if (j < nvec && v < nv_row) {
      float r[8];
      unpack8(ld_nc_v4(residual_row + v * 8), r);
      float t[8];
#pragma unroll
      for (int i = 0; i < 8; i++) t[i] = s[j][i] + r[i];
      const uint4 packed = pack8(t);  // residual_out is bf16; the norm sees the rounded values
      *reinterpret_cast<uint4*>(res_out_row + v * 8) = packed;
      unpack8(packed, s[j]);
#pragma unroll
      for (int i = 0; i < 8; i++) sq += s[j][i] * s[j][i];
    }
  }
  const float tot = block_sum(sq, smem, nwarps);
  const float rstd = rsqrtf(tot / static_cast<float>(hidden) + eps);
#pragma unroll
  for (int j = 0; j < kMaxVecPerThread; j++) {
    const int v = vec0 + j * vstride;
    if (j < nvec && v < nv_row) {
      float w[8], n[8];
      unpack8(ld_nc_v4(weight + v * 8), w);
      // (x * rstd) rounded to bf16, then * gamma in bf16 (reference test rmsnorm():16-19)
#pragma unroll
      for (int i = 0; i < 8; i++) n[i] = __bfloat162float(__float2bfloat16_rn(s[j][i] * rstd)) * w[i];
      y_out[j] = pack8(n);
    }
  }
}

struct HtParams {
  const __nv_bfloat16* x;        // local slice (used when world == 1)
  const void* mc_x;              // multicast address of the slice, or NULL
  long long peer_x[kMaxRanks];   // P2P: every rank's slice address (when mc_x == NULL)
  const __nv_bfloat16* residual;
  const __nv_bfloat16* weight;
  __nv_bfloat16* out_x;          // local slice of the output
  void* mc_out_x;                // multicast address of the output slice, or NULL
  long long peer_out[kMaxRanks];
  __nv_bfloat16* out_residual;
  const long long* signal_ptrs;  // [world] device array of signal-pad addresses
  int rank, world, num_tokens, hidden;
  float eps;
};

// reduce one row's vectors over the ranks (fp32) and fetch the matching residual vectors
__device__ __forceinline__ void ht_load_row(const HtParams& p, int row, int nvec, int vstride,
                                            int nv_row, float (*s)[8], uint4* res) {
  const long long roff = static_cast<long long>(row) * p.hidden;
#pragma unroll
  for (int j = 0; j < kMaxVecPerThread; j++) {
    const int v = threadIdx.x + j * vstride;
    if (j < nvec && v < nv_row) {
      if (p.world == 1) {
        unpack8(ld_nc_v4(p.x + roff + v * 8), s[j]);
      } else if (p.mc_x != nullptr) {
        unpack8(multimem_ld_reduce_bf16x8(static_cast<const __nv_bfloat16*>(p.mc_x) + roff + v * 8),
                s[j]);
      } else {
#pragma unroll
        for (int i = 0; i < 8; i++) s[j][i] = 0.f;
        for (int r = 0; r < p.world; r++) {
          float t[8];
          unpack8(ld_sys_v4(reinterpret_cast<const __nv_bfloat16*>(p.peer_x[r]) + roff + v * 8), t);
#pragma unroll
          for (int i = 0; i < 8; i++) s[j][i] += t[i];
        }
      }
      res[j] = ld_nc_v4(p.residual + roff + v * 8);
    }
  }
}

template <int NVEC>
__global__ void __launch_bounds__(1024)
    ar_rmsnorm_ht_kernel(const HtParams p) {
  __shared__ float s_red[2][32];
  const int nv_row = p.hidden / 8;
  const int vstride = blockDim.x;
  const int nwarps = blockDim.x / 32;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

  if (p.world > 1) block_barrier(nullptr, p.signal_ptrs, p.rank, p.world, blockIdx.x, gridDim.x, 0);

  // software pipeline over this block's rows: the (NVLS) loads of row i+1 are in flight while
  // row i is normalised and broadcast
  float s_cur[NVEC][8], s_nxt[NVEC][8];
  uint4 r_cur[NVEC], r_nxt[NVEC];
  int row = blockIdx.x;
  if (row < p.num_tokens) ht_load_row(p, row, NVEC, vstride, nv_row, s_cur, r_cur);
  int it = 0;
  for (; row < p.num_tokens; row += gridDim.x, it++) {
    const int nrow = row + gridDim.x;
    if (nrow < p.num_tokens) ht_load_row(p, nrow, NVEC, vstride, nv_row, s_nxt, r_nxt);
    const long long roff = static_cast<long long>(row) * p.hidden;
    float sq = 0.f;
#pragma unroll
    for (int j = 0; j < NVEC; j++) {
      const int v = threadIdx.x + j * vstride;
      if (v < nv_row) {
        float r[8], t[8];
        unpack8(r_cur[j], r);
#pragma unroll
        for (int i = 0; i < 8; i++) t[i] = s_cur[j][i] + r[i];
        const uint4 packed = pack8(t);  // residual_out is bf16; the norm sees the rounded values
        *reinterpret_cast<uint4*>(p.out_residual + roff + v * 8) = packed;
        unpack8(packed, s_cur[j]);
#pragma unroll
        for (int i = 0; i < 8; i++) sq += s_cur[j][i] * s_cur[j][i];
      }
    }
    sq = warp_sum_f32(sq);
    float* red = s_red[it & 1];  // double buffered: one __syncthreads per row
    if (lane == 0) red[warp] = sq;
    __syncthreads();
    float tot = 0.f;
    for (int w = 0; w < nwarps; w++) tot += red[w];
    const float rstd = rsqrtf(tot / static_cast<float>(p.hidden) + p.eps);
#pragma unroll
    for (int j = 0; j < NVEC; j++) {
      const int v = threadIdx.x + j * vstride;
      if (v < nv_row) {
        float w[8], n[8];
        unpack8(ld_nc_v4(p.weight + v * 8), w);
        // (x * rstd) rounded to bf16, then * gamma in bf16 (reference test rmsnorm():16-19)
#pragma unroll
        for (int i = 0; i < 8; i++)
          n[i] = __bfloat162float(__float2bfloat16_rn(s_cur[j][i] * rstd)) * w[i];
        const uint4 y = pack8(n);
        if (p.world == 1) {
          *reinterpret_cast<uint4*>(p.out_x + roff + v * 8) = y;
        } else if (p.mc_out_x != nullptr) {
          multimem_st_v4(static_cast<__nv_bfloat16*>(p.mc_out_x) + roff + v * 8, y);
        } else {
          for (int r = 0; r < p.world; r++) {
            st_sys_v4(reinterpret_cast<__nv_bfloat16*>(p.peer_out[r]) + roff + v * 8, y);
          }
        }
      }
    }
#pragma unroll
    for (int j = 0; j < NVEC; j++) {
#pragma unroll
      for (int i = 0; i < 8; i++) s_cur[j][i] = s_nxt[j][i];
      r_cur[j] = r_nxt[j];
    }
  }
  if (p.world > 1) {
    __threadfence_system();
    block_barrier(nullptr, p.signal_ptrs, p.rank, p.world, blockIdx.x, gridDim.x, 1);
  }
}

// ------------------------------------------------------------------------------------------------
// Low-latency Lamport two-shot, one kernel.
// workspace (per rank, symmetric): 3 buffers of `buf_bytes`; inside a buffer
//   stage 0 (scatter)  : [ceil(T/W)][W][H] bf16   rows owned by this rank, one per source rank
//   stage 1 (broadcast): [T_pad][H] bf16          reduced rows of all tokens
// buffer_flags u32[9]: {cur, dirty, bytes_per_buffer, dirty_num_stages, clear[4], arrive_counter}
// ------------------------------------------------------------------------------------------------
constexpr uint32_t kNegZero = 0x80000000u;

__device__ __forceinline__ bool vec_ready(uint4 v) {
  return v.x != kNegZero && v.y != kNegZero && v.z != kNegZero && v.w != kNegZero;
}
// a bf16 pair equal to (-0.0, +0.0) / any word equal to 0x80000000 would read as "not written"
__device__ __forceinline__ uint4 scrub_neg_zero(uint4 v) {
  v.x = v.x == kNegZero ? 0u : v.x;
  v.y = v.y == kNegZero ? 0u : v.y;
  v.z = v.z == kNegZero ? 0u : v.z;
  v.w = v.w == kNegZero ? 0u : v.w;
  return v;
}

struct LlParams {
  const __nv_bfloat16* x;          // local [T, H]
  const long long* peer_ws;        // [world] workspace base addresses (P2P)
  void* mc_ws;                     // multicast address of the workspace, or NULL
  uint32_t* flags;                 // buffer_flags
  const __nv_bfloat16* residual;
  const __nv_bfloat16* weight;
  __nv_bfloat16* out;
  __nv_bfloat16* out_residual;
  int rank, world, num_tokens, hidden;
  float eps;
};

__global__ void __launch_bounds__(1024)
    ar_rmsnorm_ll_kernel(const LlParams p) {
  __shared__ float s_red[32];
  __shared__ uint32_t s_last;
  const int W = p.world;
  const int H = p.hidden;
  const int nv_row = H / 8;
  const int vstride = blockDim.x;
  const int nvec = (nv_row + vstride - 1) / vstride;
  const int nwarps = blockDim.x / 32;
  const uint32_t cur = p.flags[0];
  const uint32_t buf_bytes = p.flags[2];
  const int tpr = (p.num_tokens + W - 1) / W;          // tokens per rank (owned)
  const long long stage1_off = static_cast<long long>(tpr) * W * H * 2;  // bytes
  const long long buf_off = static_cast<long long>(cur) * buf_bytes;
  uint8_t* my_ws = reinterpret_cast<uint8_t*>(p.peer_ws[p.rank]);

  // clear-ahead: the buffer used two calls ago (every rank has finished with it)
  {
    const uint32_t clr = (cur + 1) % 3;
    uint4 sent = make_uint4(kNegZero, kNegZero, kNegZero, kNegZero);
    uint4* dst = reinterpret_cast<uint4*>(my_ws + static_cast<long long>(clr) * buf_bytes);
    const long long nv = (stage1_off + static_cast<long long>(tpr) * W * H * 2) / 16;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < nv;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
      dst[i] = sent;
    }
  }

  for (int t = blockIdx.x; t < p.num_tokens; t += gridDim.x) {
    const int owner = t % W;
    const int lrow = t / W;
    // ---- shot 1: my row of token t -> owner's stage-0 slot [lrow][rank] ----
    {
      uint8_t* dst = reinterpret_cast<uint8_t*>(p.peer_ws[owner]) + buf_off +
                     (static_cast<long long>(lrow) * W + p.rank) * H * 2;
      const __nv_bfloat16* src = p.x + static_cast<long long>(t) * H;
      for (int v = threadIdx.x; v < nv_row; v += vstride) {
        st_sys_v4(dst + v * 16, scrub_neg_zero(ld_nc_v4(src + v * 8)));
      }
    }
    // ---- owner: reduce the W rows in rank order, broadcast the sum into stage 1 of every rank ----
    if (owner == p.rank) {
      const uint8_t* base = my_ws + buf_off + static_cast<long long>(lrow) * W * H * 2;
      for (int v = threadIdx.x; v < nv_row; v += vstride) {
        float acc[8];
#pragma unroll
        for (int i = 0; i < 8; i++) acc[i] = 0.f;
        for (int r = 0; r < W; r++) {
          const uint8_t* src = base + static_cast<long long>(r) * H * 2 + v * 16;
          uint4 d = ld_volatile_v4(src);
          uint32_t spins = 0;
          while (!vec_ready(d)) {
            d = ld_volatile_v4(src);
            if (++spins > B200_AR_SPIN_LIMIT) {
              printf("allreduce LL scatter timeout rank %d token %d src %d\n", p.rank, t, r);
              __trap();
            }
          }
          float f[8];
          unpack8(d, f);
#pragma unroll
          for (int i = 0; i < 8; i++) acc[i] += f[i];
        }
        const uint4 sum = scrub_neg_zero(pack8(acc));
        const long long off = buf_off + stage1_off + static_cast<long long>(t) * H * 2 + v * 16;
        if (p.mc_ws != nullptr) {
          multimem_st_v4(static_cast<uint8_t*>(p.mc_ws) + off, sum);
        } else {
          for (int r = 0; r < W; r++) {
            st_sys_v4(reinterpret_cast<uint8_t*>(p.peer_ws[r]) + off, sum);
          }
        }
      }
    }
    // ---- shot 2 consumer: reduced row of token t, + residual, RMSNorm ----
    {
      const uint8_t* src = my_ws + buf_off + stage1_off + static_cast<long long>(t) * H * 2;
      float s[kMaxVecPerThread][8];
#pragma unroll
      for (int j = 0; j < kMaxVecPerThread; j++) {
        const int v = threadIdx.x + j * vstride;
        if (j < nvec && v < nv_row) {
          uint4 d = ld_volatile_v4(src + v * 16);
          uint32_t spins = 0;
          while (!vec_ready(d)) {
            d = ld_volatile_v4(src + v * 16);
            if (++spins > B200_AR_SPIN_LIMIT) {
              printf("allreduce LL broadcast timeout rank %d token %d\n", p.rank, t);
              __trap();
            }
          }
          unpack8(d, s[j]);
        }
      }
      const long long roff = static_cast<long long>(t) * H;
      uint4 y[kMaxVecPerThread];
      fuse_row(s, nvec, threadIdx.x, vstride, nv_row, p.residual + roff, p.out_residual + roff,
               p.weight, p.eps, H, s_red, nwarps, y);
#pragma unroll
      for (int j = 0; j < kMaxVecPerThread; j++) {
        const int v = threadIdx.x + j * vstride;
        if (j < nvec && v < nv_row) *reinterpret_cast<uint4*>(p.out + roff + v * 8) = y[j];
      }
    }
  }

  // last block out rotates the buffer index (replay-safe: state lives on the device)
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    s_last = atomicAdd(&p.flags[8], 1u);
  }
  __syncthreads();
  if (s_last == gridDim.x - 1 && threadIdx.x == 0) {
    p.flags[8] = 0;
    p.flags[1] = cur;            // dirty = the buffer just used
    p.flags[0] = (cur + 1) % 3;  // next call
    __threadfence();
  }
}

}  // namespace ar
}  // namespace b200

using namespace b200;  // NOLINT

static int pick_threads(int hidden) {
  const int nv = hidden / 8;
  int th = (nv + 31) / 32 * 32;
  if (th > 1024) th = 1024;
  if (th < 64) th = 64;
  return th;
}

// replaces reference src/allreduce/fuse_allreduce_rmsnorm_high_throughput.h:12-18 (same arguments).
// `signal_ptr` = device int64[world_size] of every rank's signal-pad address.
// With mc_input_ptr / mc_output_ptr == NULL and world_size > 1 the P2P variant below must be used.
extern "C" int hpc_fuse_allreduce_rmsnorm_high_throughput_p2p_async(
    const void* input_ptr, const void* mc_input_ptr, const void* in_res_ptr, const void* weight_ptr,
    void* output_ptr, void* mc_output_ptr, void* out_res_ptr, void* signal_ptr,
    const int64_t* peer_input_ptrs_host, const int64_t* peer_output_ptrs_host, int64_t rank,
    int64_t world_size, int64_t num_max_blocks, double rms_norm_eps, int num_tokens,
    int hidden_size, cudaStream_t stream) {
  HPC_REQUIRE(hidden_size % 8 == 0 && hidden_size > 0 && hidden_size <= 8 * 1024 * ar::kMaxVecPerThread,
              "allreduce: hidden_size %d unsupported (multiple of 8, <= %d)", hidden_size,
              8 * 1024 * ar::kMaxVecPerThread);
  HPC_REQUIRE(world_size >= 1 && world_size <= ar::kMaxRanks, "allreduce: world_size %lld",
              (long long)world_size);
  HPC_REQUIRE(rank >= 0 && rank < world_size, "allreduce: bad rank");
  HPC_REQUIRE(num_max_blocks >= 1, "allreduce: num_max_blocks must be >= 1");
  if (world_size > 1) {
    HPC_REQUIRE(signal_ptr != nullptr, "allreduce: signal pointers required");
    HPC_REQUIRE((mc_input_ptr != nullptr && mc_output_ptr != nullptr) ||
                    (peer_input_ptrs_host != nullptr && peer_output_ptrs_host != nullptr),
                "allreduce: need multicast pointers or peer pointer tables");
  }
  ar::HtParams p;
  p.x = static_cast<const __nv_bfloat16*>(input_ptr);
  p.mc_x = mc_input_ptr;
  p.residual = static_cast<const __nv_bfloat16*>(in_res_ptr);
  p.weight = static_cast<const __nv_bfloat16*>(weight_ptr);
  p.out_x = static_cast<__nv_bfloat16*>(output_ptr);
  p.mc_out_x = mc_output_ptr;
  p.out_residual = static_cast<__nv_bfloat16*>(out_res_ptr);
  p.signal_ptrs = static_cast<const long long*>(signal_ptr);
  for (int r = 0; r < ar::kMaxRanks; r++) {
    p.peer_x[r] = (peer_input_ptrs_host && r < world_size) ? peer_input_ptrs_host[r] : 0;
    p.peer_out[r] = (peer_output_ptrs_host && r < world_size) ? peer_output_ptrs_host[r] : 0;
  }
  p.rank = static_cast<int>(rank);
  p.world = static_cast<int>(world_size);
  p.num_tokens = num_tokens;
  p.hidden = hidden_size;
  p.eps = static_cast<float>(rms_norm_eps);
  // The grid must be identical on every rank (the barrier pairs block b with block b), so it
  // depends only on num_max_blocks.
  const int grid = static_cast<int>(num_max_blocks);
  const int threads = pick_threads(hidden_size);
  const int nvec = (hidden_size / 8 + threads - 1) / threads;
  switch (nvec) {
    case 1: ar::ar_rmsnorm_ht_kernel<1><<<grid, threads, 0, stream>>>(p); break;
    case 2: ar::ar_rmsnorm_ht_kernel<2><<<grid, threads, 0, stream>>>(p); break;
    default: ar::ar_rmsnorm_ht_kernel<4><<<grid, threads, 0, stream>>>(p); break;
  }
  HPC_CUDA_CHECK(cudaGetLastError());
  return HPC_OK;
}

extern "C" int hpc_fuse_allreduce_rmsnorm_high_throughput_async(
    const void* input_ptr, const void* mc_input_ptr, const void* in_res_ptr, const void* weight_ptr,
    void* output_ptr, void* mc_output_ptr, void* out_res_ptr, void* signal_ptr, int64_t rank,
    int64_t world_size, int64_t num_max_blocks, double rms_norm_eps, int num_tokens,
    int hidden_size, cudaStream_t stream) {
  return hpc_fuse_allreduce_rmsnorm_high_throughput_p2p_async(
      input_ptr, mc_input_ptr, in_res_ptr, weight_ptr, output_ptr, mc_output_ptr, out_res_ptr,
      signal_ptr, nullptr, nullptr, rank, world_size, num_max_blocks, rms_norm_eps, num_tokens,
      hidden_size, stream);
}

// replaces reference src/allreduce/fuse_allreduce_rmsnorm_low_latency.h:29-49,503-504
// (AllReduceFusionParams flattened into plain arguments).
extern "C" int hpc_fuse_allreduce_rmsnorm_low_latency_async(
    int n_ranks, int rank, int num_tokens, int token_dim, void** buffer_ptrs_dev,
    void* buffer_ptr_local, void* multicast_ptr, uint32_t* buffer_flags, int rmsnorm_fusion,
    int launch_with_pdl, const void* input, const void* residual_in, const void* gamma,
    double epsilon, void* residual_out, void* output, int num_max_blocks, cudaStream_t stream) {
  (void)buffer_ptr_local;
  (void)launch_with_pdl;
  HPC_REQUIRE(rmsnorm_fusion, "allreduce LL: only the fused RMSNorm mode is implemented");
  HPC_REQUIRE(token_dim % 8 == 0 && token_dim > 0 && token_dim <= 8 * 1024 * ar::kMaxVecPerThread,
              "allreduce LL: hidden_size %d unsupported", token_dim);
  HPC_REQUIRE(n_ranks >= 1 && n_ranks <= ar::kMaxRanks && rank >= 0 && rank < n_ranks,
              "allreduce LL: bad rank/world");
  HPC_REQUIRE(buffer_ptrs_dev != nullptr && buffer_flags != nullptr, "allreduce LL: null workspace");
  if (num_tokens <= 0) return HPC_OK;
  ar::LlParams p;
  p.x = static_cast<const __nv_bfloat16*>(input);
  p.peer_ws = reinterpret_cast<const long long*>(buffer_ptrs_dev);
  p.mc_ws = multicast_ptr;
  p.flags = buffer_flags;
  p.residual = static_cast<const __nv_bfloat16*>(residual_in);
  p.weight = static_cast<const __nv_bfloat16*>(gamma);
  p.out = static_cast<__nv_bfloat16*>(output);
  p.out_residual = static_cast<__nv_bfloat16*>(residual_out);
  p.rank = rank;
  p.world = n_ranks;
  p.num_tokens = num_tokens;
  p.hidden = token_dim;
  p.eps = static_cast<float>(epsilon);
  // every token needs its block to be resident on all ranks at about the same time: the loop is
  // in increasing token order on every rank, so any grid size is deadlock-free
  int grid = num_tokens;
  const int cap = num_max_blocks > 0 ? num_max_blocks : sm_count();
  if (grid > cap) grid = cap;
  ar::ar_rmsnorm_ll_kernel<<<grid, pick_threads(token_dim), 0, stream>>>(p);
  HPC_CUDA_CHECK(cudaGetLastError());
  return HPC_OK;
}
