// Fused AllReduce + residual-add + RMSNorm over NVLink-5 / NVSwitch (B200), written from scratch.
//
// High-throughput path (replaces reference src/allreduce/fuse_allreduce_rmsnorm_high_throughput.cu:15-154):
//   every rank owns a contiguous token slice. ONE kernel per rank does, per owned token row,
//     reduce   : multimem.ld_reduce (in-switch NVLS reduction, bf16x2 inputs, f32 accumulate) of the
//                row over all ranks' symmetric input buffers     -- or P2P loads from every peer
//     fuse     : + residual -> residual_out (bf16), RMS over the row, * gamma
//     broadcast: multimem.st of the normalised row into every rank's symmetric output buffer
//                                                                 -- or P2P stores to every peer
//   Blocks are 256 threads holding one row as packed bf16 (4 x 16 B per thread at hidden 8192), several
//   blocks per SM so that many rows are in flight per SM; the row-invariant gamma vectors are loaded
//   once per block.
//   Cross-GPU ordering is ONE rank-level barrier at entry and one at exit, built on monotonic epoch
//   words in the signal pads (no CAS round trips, independent of the grid size, CUDA-graph replay
//   safe because the epoch lives on the device): block 0 posts "rank r entered launch e" into every
//   peer's pad, every block polls its OWN pad until all peers have posted; the last block to finish
//   posts / awaits the exit epoch, so the kernel only completes once every peer's stores have landed.
//
// Low-latency path (replaces reference src/allreduce/fuse_allreduce_rmsnorm_low_latency.cu:16-453):
//   Lamport protocol (-0.0 = "not yet written"), triple-buffered workspace, state in `buffer_flags`
//   as laid out by the reference test (tests/test_fuse_allreduce_rmsnorm_low_latency.py:54-76).
//   two-shot (one kernel; the reference chains two with PDL): token t is owned by rank t % W; each
//     rank scatters its row to the owner, the owner reduces the W rows in rank order and broadcasts
//     the sum (multimem.st or P2P), every rank adds the residual and normalises locally.
//   one-shot (small batches, when the workspace is large enough): every rank multicasts its row
//     into slot [t][rank] of every rank's buffer, then reduces the W slots locally in rank order:
//     one NVLink traversal instead of two.
//   A call clears the buffer dirtied by the previous call, using the byte count that call recorded
//   in buffer_flags[4] (the reference's clearDirtyLamportBuf bookkeeping), so batches of varying
//   size never leave stale rows behind.
#include "common.cuh"
#include "host_utils.h"

namespace b200 {
namespace ar {

constexpr int kMaxRanks = 16;

// ---- signal pad layout (uint32 words, zeroed when the symmetric buffer is created) ---------------
constexpr int kPadEntry = 0;    // [0, 16)  : entry epoch posted by peer r
constexpr int kPadExit = 16;    // [16, 32) : exit epoch posted by peer r
constexpr int kPadEpoch = 32;   // local    : launches completed on this pad
constexpr int kPadDone = 33;    // local    : blocks of the running launch that have finished

#ifndef B200_AR_SPIN_LIMIT
#define B200_AR_SPIN_LIMIT (1u << 28)
#endif

__device__ __forceinline__ void st_release_sys_u32(uint32_t* p, uint32_t v) {
  asm volatile("st.global.release.sys.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys_u32(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.global.acquire.sys.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ uint32_t ld_volatile_u32(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// spin until *p has reached `target` (epochs are monotonic; compared modulo 2^32)
__device__ __forceinline__ void wait_epoch(const uint32_t* p, uint32_t target, int rank, int peer,
                                           const char* what) {
  uint32_t spins = 0;
  while (static_cast<int32_t>(ld_acquire_sys_u32(p) - target) < 0) {
    if (++spins > B200_AR_SPIN_LIMIT) {
      printf("allreduce %s barrier timeout: rank %d waits for peer %d (target %u)\n", what, rank,
             peer, target);
      __trap();
    }
  }
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

// (reduced x + residual) -> bf16 (that is residual_out; the norm sees the rounded values);
// returns the packed sum and accumulates its squares
__device__ __forceinline__ uint4 add_residual(uint4 x, uint4 r, float& sq) {
  float a[8], b[8];
  unpack8(x, a);
  unpack8(r, b);
#pragma unroll
  for (int i = 0; i < 8; i++) a[i] += b[i];
  const uint4 packed = pack8(a);
  unpack8(packed, a);
#pragma unroll
  for (int i = 0; i < 8; i++) sq += a[i] * a[i];
  return packed;
}
// (x * rstd) rounded to bf16, then * gamma in bf16 (reference test rmsnorm():16-19)
__device__ __forceinline__ uint4 normalise(uint4 x, uint4 gamma, float rstd) {
  float a[8], w[8];
  unpack8(x, a);
  unpack8(gamma, w);
#pragma unroll
  for (int i = 0; i < 8; i++) a[i] = __bfloat162float(__float2bfloat16_rn(a[i] * rstd)) * w[i];
  return pack8(a);
}

struct HtParams {
  const __nv_bfloat16* x;        // local slice
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

// kMode: 0 = single rank (pure residual + RMSNorm), 1 = NVLS multimem, 2 = P2P on two GPUs,
// 3 = P2P for any world size (fabrics without a multicast mapping)
template <int NVEC, int THREADS, int kMode>
__global__ void __launch_bounds__(THREADS, (THREADS == 256 && NVEC <= 4) ? (kMode == 0 ? 3 : (kMode <= 2 ? 2 : 1)) : 1)
    ar_rmsnorm_ht_kernel(const HtParams p) {
  constexpr int kWarps = THREADS / 32;
  __shared__ float s_red[2][kWarps];
  __shared__ uint32_t s_last;
  const int nv_row = p.hidden / 8;
  const int tid = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5;

  uint32_t* pad = nullptr;
  uint32_t epoch = 0;
  if constexpr (kMode != 0) {
    pad = reinterpret_cast<uint32_t*>(p.signal_ptrs[p.rank]);
    epoch = ld_volatile_u32(pad + kPadEpoch);
    if (blockIdx.x == 0 && tid < p.world) {
      // release: this rank's input rows (written by earlier work on the stream) are visible
      st_release_sys_u32(reinterpret_cast<uint32_t*>(p.signal_ptrs[tid]) + kPadEntry + p.rank,
                         2 * epoch + 1);
    }
  }
  // gamma is row-invariant: one load per block, overlapping the entry wait
  uint4 wv[NVEC];
#pragma unroll
  for (int j = 0; j < NVEC; j++) {
    const int v = tid + j * THREADS;
    wv[j] = v < nv_row ? ld_nc_v4(p.weight + v * 8) : make_uint4(0, 0, 0, 0);
  }
  if constexpr (kMode != 0) {
    if (tid < p.world) wait_epoch(pad + kPadEntry + tid, 2 * epoch + 1, p.rank, tid, "entry");
    __syncthreads();
  }

  int it = 0;
  for (int row = blockIdx.x; row < p.num_tokens; row += gridDim.x, it++) {
    const long long roff = static_cast<long long>(row) * p.hidden;
    uint4 xv[NVEC], rv[NVEC];
#pragma unroll
    for (int j = 0; j < NVEC; j++) {
      const int v = tid + j * THREADS;
      if (v < nv_row) rv[j] = ld_nc_v4(p.residual + roff + v * 8);
    }
    if constexpr (kMode == 0) {
#pragma unroll
      for (int j = 0; j < NVEC; j++) {
        const int v = tid + j * THREADS;
        if (v < nv_row) xv[j] = ld_nc_v4(p.x + roff + v * 8);
      }
    } else if constexpr (kMode == 1) {
#pragma unroll
      for (int j = 0; j < NVEC; j++) {
        const int v = tid + j * THREADS;
        if (v < nv_row) {
          xv[j] = multimem_ld_reduce_bf16x8(static_cast<const __nv_bfloat16*>(p.mc_x) + roff + v * 8);
        }
      }
    } else if constexpr (kMode == 2) {
      // P2P on two GPUs (the transport of choice at W=2): both ranks' vectors of the row are
      // requested before any is consumed -- a peer load takes ~2 us, so the bytes in flight decide
      // the link utilisation -- and summed pairwise (rank 0 + rank 1, rounded to bf16 like NVLS).
      uint4 r0[NVEC], r1[NVEC];
#pragma unroll
      for (int j = 0; j < NVEC; j++) {
        const int v = tid + j * THREADS;
        if (v < nv_row) {
          r0[j] = ld_sys_v4(reinterpret_cast<const __nv_bfloat16*>(p.peer_x[0]) + roff + v * 8);
          r1[j] = ld_sys_v4(reinterpret_cast<const __nv_bfloat16*>(p.peer_x[1]) + roff + v * 8);
        }
      }
#pragma unroll
      for (int j = 0; j < NVEC; j++) {
        const int v = tid + j * THREADS;
        if (v < nv_row) {
          float a[8], b[8];
          unpack8(r0[j], a);
          unpack8(r1[j], b);
#pragma unroll
          for (int i = 0; i < 8; i++) a[i] += b[i];
          xv[j] = pack8(a);
        }
      }
    } else {
      // P2P, any world size (fabrics without an NVLS mapping): fp32 sum in rank order, two ranks'
      // loads in flight at a time
      float acc[NVEC][8];
#pragma unroll
      for (int j = 0; j < NVEC; j++)
#pragma unroll
        for (int i = 0; i < 8; i++) acc[j][i] = 0.f;
      for (int r0 = 0; r0 < p.world; r0 += 2) {
        uint4 raw[2][NVEC];
#pragma unroll
        for (int rr = 0; rr < 2; rr++) {
#pragma unroll
          for (int j = 0; j < NVEC; j++) {
            const int v = tid + j * THREADS;
            if (r0 + rr < p.world && v < nv_row) {
              raw[rr][j] = ld_sys_v4(reinterpret_cast<const __nv_bfloat16*>(p.peer_x[r0 + rr]) + roff + v * 8);
            }
          }
        }
#pragma unroll
        for (int rr = 0; rr < 2; rr++) {
#pragma unroll
          for (int j = 0; j < NVEC; j++) {
            const int v = tid + j * THREADS;
            if (r0 + rr < p.world && v < nv_row) {
              float t[8];
              unpack8(raw[rr][j], t);
#pragma unroll
              for (int i = 0; i < 8; i++) acc[j][i] += t[i];
            }
          }
        }
      }
#pragma unroll
      for (int j = 0; j < NVEC; j++) xv[j] = pack8(acc[j]);
    }
    float sq = 0.f;
#pragma unroll
    for (int j = 0; j < NVEC; j++) {
      const int v = tid + j * THREADS;
      if (v < nv_row) {
        xv[j] = add_residual(xv[j], rv[j], sq);
        *reinterpret_cast<uint4*>(p.out_residual + roff + v * 8) = xv[j];
      }
    }
    sq = warp_sum_f32(sq);
    float* red = s_red[it & 1];  // double buffered: one __syncthreads per row
    if (lane == 0) red[warp] = sq;
    __syncthreads();
    float tot = 0.f;
#pragma unroll
    for (int w = 0; w < kWarps; w++) tot += red[w];
    const float rstd = rsqrtf(tot / static_cast<float>(p.hidden) + p.eps);
#pragma unroll
    for (int j = 0; j < NVEC; j++) {
      const int v = tid + j * THREADS;
      if (v < nv_row) {
        const uint4 y = normalise(xv[j], wv[j], rstd);
        if constexpr (kMode == 0) {
          *reinterpret_cast<uint4*>(p.out_x + roff + v * 8) = y;
        } else if constexpr (kMode == 1) {
          multimem_st_v4(static_cast<__nv_bfloat16*>(p.mc_out_x) + roff + v * 8, y);
        } else {
          for (int r = 0; r < p.world; r++) {
            st_sys_v4(reinterpret_cast<__nv_bfloat16*>(p.peer_out[r]) + roff + v * 8, y);
          }
        }
      }
    }
  }

  if constexpr (kMode != 0) {
    __threadfence_system();  // this thread's broadcast stores have been performed system-wide
    __syncthreads();
    if (tid == 0) s_last = (atomicAdd(pad + kPadDone, 1u) == gridDim.x - 1) ? 1u : 0u;
    __syncthreads();
    if (s_last) {
      // the whole rank is done: tell the peers, and do not finish before they all are
      if (tid < p.world) {
        __threadfence_system();
        st_release_sys_u32(reinterpret_cast<uint32_t*>(p.signal_ptrs[tid]) + kPadExit + p.rank,
                           2 * epoch + 2);
        wait_epoch(pad + kPadExit + tid, 2 * epoch + 2, p.rank, tid, "exit");
      }
      __syncthreads();
      if (tid == 0) {
        pad[kPadDone] = 0;
        pad[kPadEpoch] = epoch + 1;
        __threadfence();
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Low-latency Lamport path, one kernel.
// workspace (per rank, symmetric): 3 buffers of `buf_bytes`; inside a buffer
//   two-shot  stage 0 (scatter)  : [ceil(T/W)][W][H] bf16   rows owned by this rank, one per source
//             stage 1 (broadcast): [T][H] bf16              reduced rows of all tokens
//   one-shot  [T][W][H] bf16                                every rank's row of every token
// buffer_flags u32[9]: {cur, dirty, bytes_per_buffer, -, dirty_bytes, -, -, -, arrive_counter}
//   cur   = buffer this call uses; dirty = buffer the previous call used (cleared by this call,
//   for the `dirty_bytes` that call recorded).
// ------------------------------------------------------------------------------------------------
constexpr uint32_t kNegZero = 0x80000000u;
constexpr long long kOneShotMaxBytes = 2ll << 20;  // per-rank receive volume up to which one-shot wins

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
__device__ __forceinline__ uint4 wait_vec(const uint8_t* src, int rank, int t, const char* what) {
  uint4 d = ld_volatile_v4(src);
  uint32_t spins = 0;
  while (!vec_ready(d)) {
    d = ld_volatile_v4(src);
    if (++spins > B200_AR_SPIN_LIMIT) {
      printf("allreduce LL %s timeout rank %d token %d\n", what, rank, t);
      __trap();
    }
  }
  return d;
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
  int mode;                        // 0 = choose by size, 1 = force two-shot
  float eps;
};

template <int NVEC, int THREADS>
__global__ void __launch_bounds__(THREADS)
    ar_rmsnorm_ll_kernel(const LlParams p) {
  constexpr int kWarps = THREADS / 32;
  __shared__ float s_red[2][kWarps];
  __shared__ uint32_t s_last;
  const int W = p.world;
  const int H = p.hidden;
  const int nv_row = H / 8;
  const int tid = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5;
  const uint32_t cur = p.flags[0] % 3u;
  const uint32_t dirty = p.flags[1] % 3u;
  const uint32_t buf_bytes = p.flags[2];
  const uint32_t dirty_bytes = p.flags[4];
  const int tpr = (p.num_tokens + W - 1) / W;          // tokens per rank (owned), two-shot
  const long long row_bytes = static_cast<long long>(H) * 2;
  const long long stage1_off = static_cast<long long>(tpr) * W * row_bytes;
  const long long one_shot_bytes = static_cast<long long>(p.num_tokens) * W * row_bytes;
  const long long two_shot_bytes = stage1_off + static_cast<long long>(p.num_tokens) * row_bytes;
  // identical on every rank (same T, W, H and workspace size)
  const bool one_shot = p.mode == 0 && one_shot_bytes <= buf_bytes && one_shot_bytes <= kOneShotMaxBytes;
  const long long buf_off = static_cast<long long>(cur) * buf_bytes;
  uint8_t* my_ws = reinterpret_cast<uint8_t*>(p.peer_ws[p.rank]);

  // clear the buffer the previous call dirtied (every rank has finished with it: that call
  // completed only after all of its rows had arrived), for exactly the bytes it recorded
  if (dirty != cur) {
    const uint4 sent = make_uint4(kNegZero, kNegZero, kNegZero, kNegZero);
    uint4* dst = reinterpret_cast<uint4*>(my_ws + static_cast<long long>(dirty) * buf_bytes);
    const long long nv = (static_cast<long long>(dirty_bytes < buf_bytes ? dirty_bytes : buf_bytes) + 15) / 16;
    for (long long i = static_cast<long long>(blockIdx.x) * THREADS + tid; i < nv;
         i += static_cast<long long>(gridDim.x) * THREADS) {
      dst[i] = sent;
    }
  }
  uint4 wv[NVEC];
#pragma unroll
  for (int j = 0; j < NVEC; j++) {
    const int v = tid + j * THREADS;
    wv[j] = v < nv_row ? ld_nc_v4(p.weight + v * 8) : make_uint4(0, 0, 0, 0);
  }

  int it = 0;
  for (int t = blockIdx.x; t < p.num_tokens; t += gridDim.x, it++) {
    const long long roff = static_cast<long long>(t) * H;
    uint4 xv[NVEC], rv[NVEC];
#pragma unroll
    for (int j = 0; j < NVEC; j++) {
      const int v = tid + j * THREADS;
      if (v < nv_row) {
        xv[j] = scrub_neg_zero(ld_nc_v4(p.x + roff + v * 8));
        rv[j] = ld_nc_v4(p.residual + roff + v * 8);
      }
    }
    if (one_shot) {
      // ---- my row of token t -> slot [t][rank] of every rank ----
      const long long off = buf_off + (static_cast<long long>(t) * W + p.rank) * row_bytes;
#pragma unroll
      for (int j = 0; j < NVEC; j++) {
        const int v = tid + j * THREADS;
        if (v < nv_row) {
          if (p.mc_ws != nullptr) {
            multimem_st_v4(static_cast<uint8_t*>(p.mc_ws) + off + v * 16, xv[j]);
          } else {
            for (int r = 0; r < W; r++) st_sys_v4(reinterpret_cast<uint8_t*>(p.peer_ws[r]) + off + v * 16, xv[j]);
          }
        }
      }
      // ---- reduce the W slots of token t in rank order ----
      const uint8_t* base = my_ws + buf_off + static_cast<long long>(t) * W * row_bytes;
#pragma unroll
      for (int j = 0; j < NVEC; j++) {
        const int v = tid + j * THREADS;
        if (v < nv_row) {
          float acc[8];
#pragma unroll
          for (int i = 0; i < 8; i++) acc[i] = 0.f;
          for (int r = 0; r < W; r++) {
            float f[8];
            unpack8(wait_vec(base + r * row_bytes + v * 16, p.rank, t, "one-shot"), f);
#pragma unroll
            for (int i = 0; i < 8; i++) acc[i] += f[i];
          }
          xv[j] = pack8(acc);
        }
      }
    } else {
      const int owner = t % W;
      const int lrow = t / W;
      // ---- shot 1: my row of token t -> owner's stage-0 slot [lrow][rank] ----
      {
        uint8_t* dst = reinterpret_cast<uint8_t*>(p.peer_ws[owner]) + buf_off +
                       (static_cast<long long>(lrow) * W + p.rank) * row_bytes;
#pragma unroll
        for (int j = 0; j < NVEC; j++) {
          const int v = tid + j * THREADS;
          if (v < nv_row) st_sys_v4(dst + v * 16, xv[j]);
        }
      }
      // ---- owner: reduce the W rows in rank order, broadcast the sum into stage 1 of every rank ----
      if (owner == p.rank) {
        const uint8_t* base = my_ws + buf_off + static_cast<long long>(lrow) * W * row_bytes;
#pragma unroll
        for (int j = 0; j < NVEC; j++) {
          const int v = tid + j * THREADS;
          if (v < nv_row) {
            float acc[8];
#pragma unroll
            for (int i = 0; i < 8; i++) acc[i] = 0.f;
            for (int r = 0; r < W; r++) {
              float f[8];
              unpack8(wait_vec(base + r * row_bytes + v * 16, p.rank, t, "scatter"), f);
#pragma unroll
              for (int i = 0; i < 8; i++) acc[i] += f[i];
            }
            const uint4 sum = scrub_neg_zero(pack8(acc));
            const long long off = buf_off + stage1_off + static_cast<long long>(t) * row_bytes + v * 16;
            if (p.mc_ws != nullptr) {
              multimem_st_v4(static_cast<uint8_t*>(p.mc_ws) + off, sum);
            } else {
              for (int r = 0; r < W; r++) st_sys_v4(reinterpret_cast<uint8_t*>(p.peer_ws[r]) + off, sum);
            }
          }
        }
      }
      // ---- shot 2 consumer: reduced row of token t ----
      const uint8_t* src = my_ws + buf_off + stage1_off + static_cast<long long>(t) * row_bytes;
#pragma unroll
      for (int j = 0; j < NVEC; j++) {
        const int v = tid + j * THREADS;
        if (v < nv_row) xv[j] = wait_vec(src + v * 16, p.rank, t, "broadcast");
      }
    }
    // ---- + residual, RMSNorm ----
    float sq = 0.f;
#pragma unroll
    for (int j = 0; j < NVEC; j++) {
      const int v = tid + j * THREADS;
      if (v < nv_row) {
        xv[j] = add_residual(xv[j], rv[j], sq);
        *reinterpret_cast<uint4*>(p.out_residual + roff + v * 8) = xv[j];
      }
    }
    sq = warp_sum_f32(sq);
    float* red = s_red[it & 1];
    if (lane == 0) red[warp] = sq;
    __syncthreads();
    float tot = 0.f;
#pragma unroll
    for (int w = 0; w < kWarps; w++) tot += red[w];
    const float rstd = rsqrtf(tot / static_cast<float>(H) + p.eps);
#pragma unroll
    for (int j = 0; j < NVEC; j++) {
      const int v = tid + j * THREADS;
      if (v < nv_row) *reinterpret_cast<uint4*>(p.out + roff + v * 8) = normalise(xv[j], wv[j], rstd);
    }
  }

  // last block out rotates the buffers (replay-safe: the state lives on the device)
  __syncthreads();
  if (tid == 0) {
    __threadfence();
    s_last = atomicAdd(&p.flags[8], 1u);
  }
  __syncthreads();
  if (s_last == gridDim.x - 1 && tid == 0) {
    p.flags[8] = 0;
    p.flags[1] = cur;            // dirty = the buffer just used ...
    p.flags[4] = static_cast<uint32_t>(one_shot ? one_shot_bytes : two_shot_bytes);  // ... this far
    p.flags[0] = (cur + 1) % 3;  // next call
    __threadfence();
  }
}

}  // namespace ar
}  // namespace b200

using namespace b200;  // NOLINT

// block size / vectors per thread for a row of `hidden` bf16: 256-thread blocks holding up to 8
// 16-B vectors per thread (hidden <= 16384), 1024 threads beyond that
template <template <int, int> class L, typename P>
static int dispatch_row_kernel(const P& p, int hidden, int grid, cudaStream_t stream) {
  const int nv = hidden / 8;
  if (nv <= 32) return L<1, 32>::run(p, grid, stream);
  if (nv <= 64) return L<1, 64>::run(p, grid, stream);
  if (nv <= 128) return L<1, 128>::run(p, grid, stream);
  if (nv <= 256) return L<1, 256>::run(p, grid, stream);
  if (nv <= 512) return L<2, 256>::run(p, grid, stream);
  if (nv <= 1024) return L<4, 256>::run(p, grid, stream);
  if (nv <= 2048) return L<8, 256>::run(p, grid, stream);
  return L<4, 1024>::run(p, grid, stream);
}
static int row_threads(int hidden) {
  const int nv = hidden / 8;
  if (nv <= 32) return 32;
  if (nv <= 64) return 64;
  if (nv <= 128) return 128;
  if (nv <= 2048) return 256;
  return 1024;
}
// low-latency path: as many threads per row as there are vectors (one block per token, T is small)
template <template <int, int> class L, typename P>
static int dispatch_row_kernel_wide(const P& p, int hidden, int grid, cudaStream_t stream) {
  const int nv = hidden / 8;
  if (nv <= 32) return L<1, 32>::run(p, grid, stream);
  if (nv <= 64) return L<1, 64>::run(p, grid, stream);
  if (nv <= 128) return L<1, 128>::run(p, grid, stream);
  if (nv <= 256) return L<1, 256>::run(p, grid, stream);
  if (nv <= 512) return L<1, 512>::run(p, grid, stream);
  if (nv <= 1024) return L<1, 1024>::run(p, grid, stream);
  if (nv <= 2048) return L<2, 1024>::run(p, grid, stream);
  return L<4, 1024>::run(p, grid, stream);
}
static int row_threads_wide(int hidden) {
  const int nv = hidden / 8;
  int th = 32;
  while (th < nv && th < 1024) th *= 2;
  return th;
}
template <int NVEC, int THREADS>
struct HtLaunch {
  static int run(const ar::HtParams& p, int grid, cudaStream_t stream) {
    if (p.world == 1) {
      ar::ar_rmsnorm_ht_kernel<NVEC, THREADS, 0><<<grid, THREADS, 0, stream>>>(p);
    } else if (p.mc_x != nullptr) {
      ar::ar_rmsnorm_ht_kernel<NVEC, THREADS, 1><<<grid, THREADS, 0, stream>>>(p);
    } else if (p.world == 2) {
      ar::ar_rmsnorm_ht_kernel<NVEC, THREADS, 2><<<grid, THREADS, 0, stream>>>(p);
    } else {
      ar::ar_rmsnorm_ht_kernel<NVEC, THREADS, 3><<<grid, THREADS, 0, stream>>>(p);
    }
    HPC_CUDA_CHECK(cudaGetLastError());
    return HPC_OK;
  }
};
template <int NVEC, int THREADS>
struct LlLaunch {
  static int run(const ar::LlParams& p, int grid, cudaStream_t stream) {
    ar::ar_rmsnorm_ll_kernel<NVEC, THREADS><<<grid, THREADS, 0, stream>>>(p);
    HPC_CUDA_CHECK(cudaGetLastError());
    return HPC_OK;
  }
};

// replaces reference src/allreduce/fuse_allreduce_rmsnorm_high_throughput.h:12-18 (same arguments).
// `signal_ptr` = device int64[world_size] of every rank's signal-pad address.
// With mc_input_ptr / mc_output_ptr == NULL and world_size > 1 the P2P variant below must be used.
extern "C" int hpc_fuse_allreduce_rmsnorm_high_throughput_p2p_async(
    const void* input_ptr, const void* mc_input_ptr, const void* in_res_ptr, const void* weight_ptr,
    void* output_ptr, void* mc_output_ptr, void* out_res_ptr, void* signal_ptr,
    const int64_t* peer_input_ptrs_host, const int64_t* peer_output_ptrs_host, int64_t rank,
    int64_t world_size, int64_t num_max_blocks, double rms_norm_eps, int num_tokens,
    int hidden_size, cudaStream_t stream) {
  HPC_REQUIRE(hidden_size % 8 == 0 && hidden_size > 0 && hidden_size <= 32768,
              "allreduce: hidden_size %d unsupported (multiple of 8, <= 32768)", hidden_size);
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
  // num_max_blocks is the caller's cap in units of full (1024-thread) CTAs, as in the reference;
  // the barrier does not depend on the grid, so ranks may even use different grids. An empty slice
  // still launches one block: the rank has to take part in the barriers.
  const int threads = row_threads(hidden_size);
  long long grid = static_cast<long long>(num_max_blocks) * (1024 / threads);
  const long long resident = static_cast<long long>(sm_count()) * (threads >= 1024 ? 1 : 8);
  if (grid > resident) grid = resident;
  if (grid > num_tokens) grid = num_tokens;
  if (grid < 1) {
    if (world_size == 1) return HPC_OK;
    grid = 1;
  }
  return dispatch_row_kernel<HtLaunch>(p, hidden_size, static_cast<int>(grid), stream);
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
// (AllReduceFusionParams flattened into plain arguments). `launch_with_pdl` is accepted for
// signature compatibility (one kernel: nothing to chain); `num_max_blocks` <= 0 means "one block
// per token up to the SM count". `protocol`: 0 = one-shot when the batch is small and the workspace
// holds [T][W][H] (decided on the device, identically on every rank), 1 = always two-shot.
extern "C" int hpc_fuse_allreduce_rmsnorm_low_latency_ex_async(
    int n_ranks, int rank, int num_tokens, int token_dim, void** buffer_ptrs_dev,
    void* buffer_ptr_local, void* multicast_ptr, uint32_t* buffer_flags, int rmsnorm_fusion,
    int launch_with_pdl, const void* input, const void* residual_in, const void* gamma,
    double epsilon, void* residual_out, void* output, int num_max_blocks, int protocol,
    cudaStream_t stream) {
  (void)buffer_ptr_local;
  (void)launch_with_pdl;
  HPC_REQUIRE(rmsnorm_fusion, "allreduce LL: only the fused RMSNorm mode is implemented");
  HPC_REQUIRE(token_dim % 8 == 0 && token_dim > 0 && token_dim <= 32768,
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
  p.mode = protocol;
  p.eps = static_cast<float>(epsilon);
  // every token needs its block to be resident on all ranks at about the same time: the loop is
  // in increasing token order on every rank, so any grid size is deadlock-free
  const int threads = row_threads_wide(token_dim);
  long long grid = num_tokens;
  const long long cap = num_max_blocks > 0 ? static_cast<long long>(num_max_blocks) * (1024 / threads)
                                           : static_cast<long long>(sm_count()) * (threads >= 1024 ? 1 : 2);
  if (grid > cap) grid = cap;
  return dispatch_row_kernel_wide<LlLaunch>(p, token_dim, static_cast<int>(grid), stream);
}

extern "C" int hpc_fuse_allreduce_rmsnorm_low_latency_async(
    int n_ranks, int rank, int num_tokens, int token_dim, void** buffer_ptrs_dev,
    void* buffer_ptr_local, void* multicast_ptr, uint32_t* buffer_flags, int rmsnorm_fusion,
    int launch_with_pdl, const void* input, const void* residual_in, const void* gamma,
    double epsilon, void* residual_out, void* output, int num_max_blocks, cudaStream_t stream) {
  return hpc_fuse_allreduce_rmsnorm_low_latency_ex_async(
      n_ranks, rank, num_tokens, token_dim, buffer_ptrs_dev, buffer_ptr_local, multicast_ptr,
      buffer_flags, rmsnorm_fusion, launch_with_pdl, input, residual_in, gamma, epsilon,
      residual_out, output, num_max_blocks, 1, stream);
}
