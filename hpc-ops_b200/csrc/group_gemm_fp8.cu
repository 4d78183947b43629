// Grouped FP8 GEMM for MoE (B200 / sm_100a), written from scratch.
//
//   Y[rows of group g, :] = X[rows of group g, :] . W[g]^T        X: [M, K] e4m3, W: [G, N, K] e4m3
//
// scale modes
//   blockwise : per 128-wide K block kb   acc += (X_kb . W_kb^T) * xscale[kb, row] * wscale[g, n/128, kb]
//               (reference src/group_gemm/kernels.cuh:532-892, semantics of
//                tests/test_fuse_moe_blockwise.py:84-138)
//   per-tensor: acc = (X . W^T) * yscale[g]   (reference src/group_gemm/kernels.cuh:215-530)
// epilogues
//   plain     : bf16 Y[M, N]
//   fused act : N = 2*I (gate rows then up rows of W): q = silu(bf16(gate)) * bf16(up), then either
//               128-column block quant (scale = amax/448, q/(scale+1e-8) -> e4m3, scales transposed)
//               (reference src/activation/activation.cu:282-356) or per-tensor quant q / act_scale
//               (reference src/activation/activation.cu:19-136). The bf16 Gate-Up matrix never
//               goes to HBM.
//
// Kernel: persistent, one CTA per SM, 384 threads:
//   warp 0      TMA producer: A tile 128 rows x 128 B, B tile 256 rows x 128 B per K block
//               (4 stages x 48 KB, 128B swizzle)
//   warp 1      tcgen05 issuer: 4 x UMMA M=128 N=256 K=32 (kind::f8f6f4) per K block into one of
//               two TMEM accumulators (2 x 256 columns = all of TMEM)
//   warps 4-11  two epilogue warpgroups. Blockwise: every K block is drained TMEM -> registers
//               and promoted with the block scales (fp32 FFMA2), so MMA of block kb+1 overlaps the
//               promotion of block kb. Thread = output row: the quant amax over a row is
//               thread-local. Warpgroup w owns columns [64w, 64w+64) of both 128-column halves,
//               so in the fused epilogue each thread holds matching gate/up pairs.
#include <cstdlib>

#include "common.cuh"
#include "host_utils.h"

namespace b200 {
namespace ggemm {

constexpr int kBM = 128;
constexpr int kBN = 256;
constexpr int kBK = 128;
constexpr int kStages = 4;
constexpr int kABytes = kBM * kBK;
constexpr int kBBytes = kBN * kBK;
constexpr int kStageBytes = kABytes + kBBytes;
constexpr int kThreads = 384;
constexpr int kMaxGroups = 512;
constexpr int kEpiBar = 2;
constexpr int kTileQ = 4;
constexpr int kXsSlots = kStages + 2;  // activation-scale ring (see the producer)
constexpr int kMaxKB = 128;            // K blocks per tile whose weight scales are staged (K <= 16384)

struct Params {
  const int* seqlens;      // [G] rows per group
  const int* cu_seqlens;   // [G+1] first row of each group in X / Y
  const float* xscale_t;   // blockwise: [K/128, m_pad] (column of (g, i) = pad_base[g] + i)
  const float* wscale;     // blockwise: [G, N/128, kpad4]; per-tensor: yscale [G]
  const float* act_scale;  // per-tensor fused: [1]
  __nv_bfloat16* y;        // plain: [M, N]
  uint8_t* q_out;          // fused: e4m3 [M, N/2]
  float* q_scale_t;        // fused blockwise: [N/2/128, m_pad]
  int num_group;
  int m_total;
  int n;       // rows of W per group
  int k;
  int m_pad;
  int kpad4;
  int scale_tile;  // column padding granule of the transposed activation-scale layout
  int use_bf16_mul;
  int* tile_counter;  // dynamic tile scheduler: [0] next tile, [1] CTAs done (self-resetting)
  int debug;  // HPC_B200_MOE_DEBUG diagnostics (timing experiments only; results are wrong when set)
  long long* debug_out;  // debug & 8: per CTA {cycles, K blocks, ns, tiles} of the MMA thread
};

__device__ __forceinline__ void ffma2(float2& acc, float a0, float a1, float2 f) {
  uint64_t& accu = reinterpret_cast<uint64_t&>(acc);
  uint64_t av, fv;
  asm("mov.b64 %0, {%1, %2};" : "=l"(av) : "f"(a0), "f"(a1));
  asm("mov.b64 %0, {%1, %2};" : "=l"(fv) : "f"(f.x), "f"(f.y));
  asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(accu) : "l"(av), "l"(fv));
}

__device__ __forceinline__ float silu_f(float x) {
  // x / (1 + exp(-x)) with ex2.approx / approximate division (relative error ~1e-6, far below
  // the bf16 / e4m3 granularity around it; reference src/utils/utils.cuh:300-330 does the same).
  // An IEEE expf + division here costs ~35 % of a Gate-Up tile's MMA time in the epilogue.
  return __fdividef(x, 1.f + __expf(-x));
}

__device__ __forceinline__ float bf16_round(float x) {
  return __bfloat162float(__float2bfloat16_rn(x));
}

// debug & 8: cycles a role spends blocked in a wait (attribution of the pipeline's bubbles)
#define HPC_TIMED(acc_var, stmt)                 \
  do {                                           \
    if (p.debug & 8) {                           \
      const long long _t0 = clock64();           \
      stmt;                                      \
      acc_var += clock64() - _t0;                \
    } else {                                     \
      stmt;                                      \
    }                                            \
  } while (0)

struct TileInfo {
  int g, mt, nt;
  int row0;    // first row (global) of the tile
  int nvalid;  // valid rows in the tile
  int scol0;   // first activation-scale column of the tile
};

// smem-resident schedule: cu_tiles (prefix of m-tiles*NT), pad_base (scale column base) per group
struct Sched {
  int* cu_tiles;   // [G+1]
  int* pad_base;   // [G]
  int* rows;       // [G]
  int* row_start;  // [G]
  int num_group;
  int nt_count;
};

// dynamic shared memory: operand stages, schedule arrays, amax exchange, scale rings, barriers,
// tile queue, TMEM slot
constexpr int kSmemBytes = kStages * kStageBytes + (kMaxGroups + 4 + 3 * kMaxGroups) * 4 + 256 * 4 +
                           kXsSlots * kBM * 4 + kTileQ * 2 * kMaxKB * 4 +
                           (2 * kStages + 4 + 2 * kTileQ) * 8 + kTileQ * 4 + 16;

__device__ __forceinline__ bool decode_tile(const Sched& s, int tile, TileInfo& t) {
  if (tile >= s.cu_tiles[s.num_group]) return false;
  int lo = 0, hi = s.num_group - 1;
  while (lo < hi) {
    int mid = (lo + hi + 1) >> 1;
    if (s.cu_tiles[mid] <= tile) {
      lo = mid;
    } else {
      hi = mid - 1;
    }
  }
  const int g = lo;
  const int local = tile - s.cu_tiles[g];
  const int mtiles = (s.rows[g] + kBM - 1) / kBM;
  t.g = g;
  t.nt = local / mtiles;
  t.mt = local - t.nt * mtiles;
  t.row0 = s.row_start[g] + t.mt * kBM;
  const int left = s.rows[g] - t.mt * kBM;
  t.nvalid = left < kBM ? left : kBM;
  t.scol0 = s.pad_base[g] + t.mt * kBM;
  return true;
}

template <bool kBlockwise, bool kFused>
__global__ void __launch_bounds__(kThreads, 1)
    group_gemm_fp8_kernel(const __grid_constant__ CUtensorMap tmap_a,
                          const __grid_constant__ CUtensorMap tmap_b, const Params p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* stages = smem;
  int* s_cu_tiles = reinterpret_cast<int*>(smem + kStages * kStageBytes);
  int* s_pad_base = s_cu_tiles + (kMaxGroups + 4);  // keeps everything after 16-B aligned
  int* s_rows = s_pad_base + kMaxGroups;
  int* s_row_start = s_rows + kMaxGroups;
  float* s_amax = reinterpret_cast<float*>(s_row_start + kMaxGroups);  // [2][128]
  float* s_xs = s_amax + 256;                        // [kXsSlots][128] activation scales of a K block
  float* s_ws = s_xs + kXsSlots * kBM;               // [kTileQ][2][kMaxKB] weight scales of a tile
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_ws + kTileQ * 2 * kMaxKB);
  uint64_t* full = bars;
  uint64_t* empty = bars + kStages;
  uint64_t* part_full = bars + 2 * kStages;
  uint64_t* part_empty = part_full + 2;
  uint64_t* tq_full = part_empty + 2;   // tile-id queue (kTileQ slots): producer -> consumers
  uint64_t* tq_empty = tq_full + kTileQ;
  int* s_tileq = reinterpret_cast<int*>(tq_empty + kTileQ);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(s_tileq + kTileQ);

  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int lane = tid & 31;
  const int KB = (p.k + kBK - 1) / kBK;  // a ragged last K block is zero-filled by TMA
  const int nt_count = kFused ? (p.n / 2 + 127) / 128 : (p.n + kBN - 1) / kBN;
  const int nblk_per_group = p.n / 128;

  // ---- set-up that touches no global memory (overlaps the previous kernel under PDL) ----
  if (warp == 0 && lane == 0) {
    prefetch_tensormap(&tmap_a);
    prefetch_tensormap(&tmap_b);
    for (int i = 0; i < kStages; i++) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < 2; i++) {
      mbar_init(&part_full[i], 1);
      mbar_init(&part_empty[i], 8);  // one arrive per epilogue warp
    }
    for (int i = 0; i < kTileQ; i++) {
      mbar_init(&tq_full[i], 1);
      mbar_init(&tq_empty[i], 9);  // MMA thread + 8 epilogue warps
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  pdl_launch_dependents();
  pdl_wait();  // routing (seqlens, gathered rows, scales) / the previous GEMM's output are complete

  // ---- schedule prologue: every CTA derives the same tile list from seqlens ----
  if (warp == 2) {
    // one warp scans <= 512 groups
    int carry_tiles = 0, carry_pad = 0;
    for (int g0 = 0; g0 < p.num_group; g0 += 32) {
      const int g = g0 + lane;
      const int r = g < p.num_group ? p.seqlens[g] : 0;
      const int mt = (r + kBM - 1) / kBM;
      const int pd = (r + p.scale_tile - 1) / p.scale_tile * p.scale_tile;
      int it = mt * nt_count, ip = pd;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int a = __shfl_up_sync(0xffffffffu, it, o);
        const int b = __shfl_up_sync(0xffffffffu, ip, o);
        if (lane >= o) {
          it += a;
          ip += b;
        }
      }
      if (g < p.num_group) {
        s_cu_tiles[g] = carry_tiles + it - mt * nt_count;
        s_pad_base[g] = carry_pad + ip - pd;
        s_rows[g] = r;
        s_row_start[g] = p.cu_seqlens[g];
      }
      carry_tiles += __shfl_sync(0xffffffffu, it, 31);
      carry_pad += __shfl_sync(0xffffffffu, ip, 31);
    }
    if (lane == 0) s_cu_tiles[p.num_group] = carry_tiles;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  Sched sched{s_cu_tiles, s_pad_base, s_rows, s_row_start, p.num_group, nt_count};

  if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
    if (warp == 0) {
      // =========================== TMA producer ==========================================
      // The whole warp runs this loop with warp-uniform values (tile fields are broadcast with
      // shfl) and ONE elected lane issues: TMA / mbarrier / tcgen05 instructions take their operands
      // from uniform registers, and for values the compiler cannot prove uniform it wraps every
      // such instruction in an elect + R2UR.BROADCAST "waterfall" loop (~70 cycles per instruction:
      // measured 600-700 cycles of issue work per K block with single-lane roles).
      // A rows of a group are re-read by every n-tile -> keep them in L2. Weight tiles are shared
      // by the 2-3 m-tiles of the same (group, n-tile), which run on neighbouring CTAs at about
      // the same time -> default policy (evict_first made every m-tile re-read HBM: 2.35x traffic).
      const uint64_t pol_a = make_policy_evict_last();
      uint32_t it = 0;
      uint32_t xsl = 0;  // it % kXsSlots
      uint32_t tq = 0;
      TileInfo t;
      long long w_empty = 0, w_tq = 0;
      const long long pc0 = clock64();
      // Dynamic scheduler: tiles are claimed from a global counter, so tiles with neighbouring ids
      // (the m-tiles sharing one weight tile) start within a short window on different CTAs and
      // share that weight tile through L2. The id of the next tile is claimed one tile ahead.
      auto claim = [&]() -> int {
        int v = 0;
        if (lane == 0) v = atomicAdd(p.tile_counter, 1);
        return __shfl_sync(0xffffffffu, v, 0);
      };
      int next_tile = claim();
      while (true) {
        const int tile = next_tile;
        bool valid = decode_tile(sched, tile, t);
        valid = __shfl_sync(0xffffffffu, static_cast<int>(valid), 0) != 0;
        t.g = __shfl_sync(0xffffffffu, t.g, 0);
        t.nt = __shfl_sync(0xffffffffu, t.nt, 0);
        t.row0 = __shfl_sync(0xffffffffu, t.row0, 0);
        t.nvalid = __shfl_sync(0xffffffffu, t.nvalid, 0);
        t.scol0 = __shfl_sync(0xffffffffu, t.scol0, 0);
        const uint32_t qs = tq % kTileQ;
        HPC_TIMED(w_tq, mbar_wait(&tq_empty[qs], ((tq / kTileQ) & 1) ^ 1));
        const int nb0 = kFused ? t.nt : t.nt * 2;
        int nb1 = kFused ? nblk_per_group / 2 + t.nt : t.nt * 2 + 1;
        if (nb1 >= nblk_per_group) nb1 = nb0;
        if (elect_one()) {
          s_tileq[qs] = valid ? tile : -1;
          if (kBlockwise && valid) {
            // the tile's weight scales (KB floats for each 128-column half) travel with its queue slot
            const uint32_t bytes = static_cast<uint32_t>(p.kpad4) * 4u;
            mbar_arrive_expect_tx(&tq_full[qs], 2 * bytes);
            const float* ws = p.wscale + static_cast<long long>(t.g) * nblk_per_group * p.kpad4;
            bulk_load_1d(s_ws + (qs * 2 + 0) * kMaxKB, ws + static_cast<long long>(nb0) * p.kpad4, bytes, &tq_full[qs]);
            bulk_load_1d(s_ws + (qs * 2 + 1) * kMaxKB, ws + static_cast<long long>(nb1) * p.kpad4, bytes, &tq_full[qs]);
          } else {
            mbar_arrive(&tq_full[qs]);
          }
        }
        __syncwarp();
        tq++;
        if (!valid) break;
        next_tile = claim();
        const int nrow0 = kFused ? t.nt * 128 : t.nt * kBN;
        const int nrow1 = kFused ? p.n / 2 + t.nt * 128 : t.nt * kBN + 128;
        // activation scales of the tile's rows: 16-byte granules covering the valid rows
        const uint32_t xs_bytes = static_cast<uint32_t>((t.nvalid + 3) / 4) * 16u;
        for (int kb = 0; kb < KB; kb++, it++) {
          const uint32_t s = it % kStages;
          uint8_t* a_dst = stages + s * kStageBytes;
          uint8_t* b_dst = a_dst + kABytes;
          HPC_TIMED(w_empty, mbar_wait(&empty[s], ((it / kStages) & 1) ^ 1));
          if (elect_one()) {
            uint32_t extra = 0;
            if constexpr (kBlockwise) {
              // The K block's 128 activation scales ride on the stage's barrier. Ring of kStages + 2
              // slots with no "empty" barrier: K block `it` is loaded once stage s is free, i.e.
              // MMA(it - kStages) has completed, which was issued only after the epilogue finished
              // with K block it - kStages - 2 -- the previous user of this slot. The epilogue reads
              // the slot after it has seen the K block's accumulator (committed after the MMA, which
              // was issued after `full[s]` completed): no barrier wait of its own.
              extra = xs_bytes;
            }
            if (p.debug & 3) {  // diagnostics: leave out the weight (1) / activation (2) tile loads
              const uint32_t bytes = ((p.debug & 1) ? 0 : kBBytes) + ((p.debug & 2) ? 0 : kABytes) + extra;
              if (bytes == 0) {
                mbar_arrive(&full[s]);
              } else {
                mbar_arrive_expect_tx(&full[s], bytes);
              }
              if constexpr (kBlockwise) {
                bulk_load_1d(s_xs + xsl * kBM, p.xscale_t + static_cast<long long>(kb) * p.m_pad + t.scol0,
                             xs_bytes, &full[s]);
              }
              if (!(p.debug & 2)) tma_load_2d_hint(a_dst, &tmap_a, &full[s], kb * kBK, t.row0, pol_a);
              if (!(p.debug & 1)) {
                tma_load_3d(b_dst, &tmap_b, &full[s], kb * kBK, nrow0, t.g);
                tma_load_3d(b_dst + kBBytes / 2, &tmap_b, &full[s], kb * kBK, nrow1, t.g);
              }
            } else {
              mbar_arrive_expect_tx(&full[s], kStageBytes + extra);
              if constexpr (kBlockwise) {
                bulk_load_1d(s_xs + xsl * kBM, p.xscale_t + static_cast<long long>(kb) * p.m_pad + t.scol0,
                             xs_bytes, &full[s]);
              }
              tma_load_2d_hint(a_dst, &tmap_a, &full[s], kb * kBK, t.row0, pol_a);
              tma_load_3d(b_dst, &tmap_b, &full[s], kb * kBK, nrow0, t.g);
              tma_load_3d(b_dst + kBBytes / 2, &tmap_b, &full[s], kb * kBK, nrow1, t.g);
              // (Tried: TMA L2 prefetch of the weight tiles 6 K blocks ahead -- 6 % slower at C3, the
              // operands are not what the MMA warp waits for)
            }
          }
          if (++xsl == kXsSlots) xsl = 0;
          __syncwarp();
        }
      }
      if ((p.debug & 8) && lane == 0) {
        long long* o = p.debug_out + (kFused ? 0 : 16 * 256) + 16 * blockIdx.x;
        o[8] = clock64() - pc0;
        o[9] = w_empty;
        o[10] = w_tq;
      }
      // the last CTA out re-arms the scheduler for the next launch (no memset between launches:
      // a memset node would break a PDL chain, and a graph replay needs nothing else)
      if (lane == 0) {
        __threadfence();
        if (atomicAdd(p.tile_counter + 1, 1) == static_cast<int>(gridDim.x) - 1) {
          p.tile_counter[0] = 0;
          p.tile_counter[1] = 0;
          __threadfence();
        }
      }
    } else if (warp == 1) {
      // =========================== tcgen05 issuer (whole warp, one elected lane issues) =======
      constexpr uint32_t idesc = make_idesc(kBM, kBN, kFmtE4M3, kFmtE4M3, 0, 0);
      const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
      const uint64_t adesc0 = make_smem_desc(smem_u32(stages), 16, 1024, kLayoutSW128);
      const uint64_t bdesc0 = make_smem_desc(smem_u32(stages) + kABytes, 16, 1024, kLayoutSW128);
      uint32_t it = 0;   // K-block counter (smem ring)
      uint32_t acc_it = 0;  // accumulator-buffer use counter
      uint32_t tq = 0;
      uint32_t ntiles = 0;
      unsigned long long g0 = 0;
      long long c0 = 0, w_full = 0, w_pempty = 0, w_tqm = 0;
      if (p.debug & 8) {
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g0));
        c0 = clock64();
      }
      while (true) {
        const uint32_t qs = tq % kTileQ;
        HPC_TIMED(w_tqm, mbar_wait(&tq_full[qs], (tq / kTileQ) & 1));
        const int tile = __shfl_sync(0xffffffffu, s_tileq[qs], 0);
        __syncwarp();
        if (lane == 0) mbar_arrive(&tq_empty[qs]);
        tq++;
        if (tile < 0) break;
        ntiles++;
        for (int kb = 0; kb < KB; kb++, it++) {
          const uint32_t s = it % kStages;
          const bool new_acc = kBlockwise || kb == 0;
          const uint32_t buf = acc_it & 1;
          HPC_TIMED(w_full, mbar_wait(&full[s], (it / kStages) & 1));
          if (new_acc) HPC_TIMED(w_pempty, mbar_wait(&part_empty[buf], ((acc_it >> 1) & 1) ^ 1));
          tc_fence_after();
          const uint64_t ad = adesc0 + static_cast<uint64_t>(s * (kStageBytes >> 4));
          const uint64_t bd = bdesc0 + static_cast<uint64_t>(s * (kStageBytes >> 4));
          const uint32_t d = tmem_u + buf * kBN;
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < 4; k++) {
              umma_f8(d, ad + k * 2, bd + k * 2, idesc, (k > 0) || !new_acc);
            }
            // The accumulator hand-over is the critical path (MMA -> drain -> next MMA into this
            // buffer); the stage release is not (the producer runs kStages ahead). Commits are
            // processed in order, so "accumulator ready" goes first (tools/umma_rate.py: 725 ->
            // 563 cycles per K block).
            if (kBlockwise || kb == KB - 1) umma_commit(&part_full[buf]);
            umma_commit(&empty[s]);
          }
          __syncwarp();
          if (kBlockwise || kb == KB - 1) acc_it++;
        }
      }
      if ((p.debug & 8) && lane == 0) {
        unsigned long long g1;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g1));
        long long* o = p.debug_out + (kFused ? 0 : 16 * 256) + 16 * blockIdx.x;
        o[0] = clock64() - c0;
        o[1] = it;
        o[2] = static_cast<long long>(g1 - g0);
        o[3] = ntiles;
        o[4] = w_full;
        o[5] = w_pempty;
        o[6] = w_tqm;
      }
    }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 232;");
    // =========================== epilogue warpgroups ======================================
    const int wg = (warp - 4) >> 2;
    const int quad = warp & 3;
    const int row_local = quad * 32 + lane;
    const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16);

    uint32_t acc_it = 0;
    uint32_t xsl = 0;  // acc_it % kXsSlots (blockwise)
    uint32_t tq = 0;
    TileInfo t;
    long long w_xs = 0, w_pfull = 0, w_tqe = 0, t_epi = 0;
    const long long ec0 = clock64();
    while (true) {
      const uint32_t qs = tq % kTileQ;
      HPC_TIMED(w_tqe, mbar_wait(&tq_full[qs], (tq / kTileQ) & 1));
      const int tile = s_tileq[qs];
      tq++;
      if (tile < 0) {
        __syncwarp();
        if (lane == 0) mbar_arrive(&tq_empty[qs]);
        break;
      }
      decode_tile(sched, tile, t);
      const bool row_valid = row_local < t.nvalid;

      float2 acc[2][32];  // [half][pair]: 64 columns of each 128-column half
#pragma unroll
      for (int h = 0; h < 2; h++)
#pragma unroll
        for (int i = 0; i < 32; i++) acc[h][i] = make_float2(0.f, 0.f);

      if constexpr (kBlockwise) {
        const float* ws0 = s_ws + (qs * 2 + 0) * kMaxKB;
        const float* ws1 = s_ws + (qs * 2 + 1) * kMaxKB;
        for (int kb = 0; kb < KB; kb++, acc_it++) {
          const uint32_t buf = acc_it & 1;
          HPC_TIMED(w_pfull, mbar_wait(&part_full[buf], (acc_it >> 1) & 1));
          tc_fence_after();
          // block scales from shared memory (staged by the producer with the operands of this K
          // block, which the MMA has consumed by now): no global load latency, no extra barrier
          const float xs = row_valid ? s_xs[xsl * kBM + row_local] : 0.f;
          const float w0 = ws0[kb], w1 = ws1[kb];
          if (++xsl == kXsSlots) xsl = 0;
          if (p.debug & 4) {  // diagnostics: no TMEM drain / promotion
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&part_empty[buf]);
            continue;
          }
          const float f0 = xs * w0, f1 = xs * w1;
          const float2 ff0 = make_float2(f0, f0), ff1 = make_float2(f1, f1);
          const uint32_t base = lane_addr + buf * kBN + wg * 64;
          // 4 chunks of 32 columns: (half 0: c0, c1), (half 1: c2, c3). TMEM loads are issued two
          // deep so their latency overlaps the FFMA2s; the accumulator buffer is handed back to
          // the MMA warp as soon as its last column is in registers.
          uint32_t ra[32], rb[32];
          tmem_ld_x32(base, ra);
          tmem_ld_x32(base + 32, rb);
          tmem_wait_ld();
#pragma unroll
          for (int i = 0; i < 16; i++)
            ffma2(acc[0][i], __uint_as_float(ra[2 * i]), __uint_as_float(ra[2 * i + 1]), ff0);
          tmem_ld_x32(base + 128, ra);
#pragma unroll
          for (int i = 0; i < 16; i++)
            ffma2(acc[0][16 + i], __uint_as_float(rb[2 * i]), __uint_as_float(rb[2 * i + 1]), ff0);
          tmem_ld_x32(base + 128 + 32, rb);
          tmem_wait_ld();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&part_empty[buf]);
#pragma unroll
          for (int i = 0; i < 16; i++)
            ffma2(acc[1][i], __uint_as_float(ra[2 * i]), __uint_as_float(ra[2 * i + 1]), ff1);
#pragma unroll
          for (int i = 0; i < 16; i++)
            ffma2(acc[1][16 + i], __uint_as_float(rb[2 * i]), __uint_as_float(rb[2 * i + 1]), ff1);
        }
      } else {
        const uint32_t buf = acc_it & 1;
        mbar_wait(&part_full[buf], (acc_it >> 1) & 1);
        tc_fence_after();
        const float ys = __ldg(p.wscale + t.g);
#pragma unroll
        for (int h = 0; h < 2; h++) {
#pragma unroll
          for (int c = 0; c < 2; c++) {
            uint32_t r[32];
            tmem_ld_x32(lane_addr + buf * kBN + h * 128 + wg * 64 + c * 32, r);
            tmem_wait_ld();
#pragma unroll
            for (int i = 0; i < 16; i++) {
              acc[h][c * 16 + i] =
                  make_float2(__uint_as_float(r[2 * i]) * ys, __uint_as_float(r[2 * i + 1]) * ys);
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&part_empty[buf]);
        acc_it++;
      }
      // the queue slot (tile id + its weight scales) is free once the K loop is through
      __syncwarp();
      if (lane == 0) mbar_arrive(&tq_empty[qs]);

      // ---------------- tile epilogue ----------------
      const long long te0 = (p.debug & 8) ? clock64() : 0;
      const long long grow = static_cast<long long>(t.row0) + row_local;
      if constexpr (!kFused) {
        if (row_valid) {
#pragma unroll
          for (int h = 0; h < 2; h++) {
            const int col0 = t.nt * kBN + h * 128 + wg * 64;
            if (col0 < p.n) {
              __nv_bfloat16* dst = p.y + grow * p.n + col0;
#pragma unroll
              for (int v = 0; v < 8; v++) {
                uint4 w;
                __nv_bfloat162 b0 = __floats2bfloat162_rn(acc[h][v * 4 + 0].x, acc[h][v * 4 + 0].y);
                __nv_bfloat162 b1 = __floats2bfloat162_rn(acc[h][v * 4 + 1].x, acc[h][v * 4 + 1].y);
                __nv_bfloat162 b2 = __floats2bfloat162_rn(acc[h][v * 4 + 2].x, acc[h][v * 4 + 2].y);
                __nv_bfloat162 b3 = __floats2bfloat162_rn(acc[h][v * 4 + 3].x, acc[h][v * 4 + 3].y);
                w.x = *reinterpret_cast<uint32_t*>(&b0);
                w.y = *reinterpret_cast<uint32_t*>(&b1);
                w.z = *reinterpret_cast<uint32_t*>(&b2);
                w.w = *reinterpret_cast<uint32_t*>(&b3);
                *reinterpret_cast<uint4*>(dst + v * 8) = w;
              }
            }
          }
        }
      } else {
        // silu(gate) * up on the bf16-rounded Gate-Up values (the reference materialises the
        // Gate-Up output as bf16 before the activation: src/fuse_moe/entry.cc:565-566)
        float v[64];
        float amax = 0.f;
#pragma unroll
        for (int i = 0; i < 32; i++) {
          const float g0 = bf16_round(acc[0][i].x), g1 = bf16_round(acc[0][i].y);
          const float u0 = bf16_round(acc[1][i].x), u1 = bf16_round(acc[1][i].y);
          float a0 = silu_f(g0), a1 = silu_f(g1);
          if (!kBlockwise && p.use_bf16_mul) {
            a0 = bf16_round(bf16_round(a0) * u0);
            a1 = bf16_round(bf16_round(a1) * u1);
          } else {
            a0 *= u0;
            a1 *= u1;
          }
          v[2 * i] = a0;
          v[2 * i + 1] = a1;
          amax = fmaxf(amax, fmaxf(fabsf(a0), fabsf(a1)));
        }
        float inv;
        if constexpr (kBlockwise) {
          s_amax[wg * 128 + row_local] = amax;
          named_bar_sync(kEpiBar, 256);
          amax = fmaxf(amax, s_amax[(wg ^ 1) * 128 + row_local]);
          const float scale = amax / 448.f;
          inv = 1.f / (scale + 1e-8f);
          if (row_valid && wg == 0) {
            p.q_scale_t[static_cast<long long>(t.nt) * p.m_pad + t.scol0 + row_local] = scale;
          }
        } else {
          inv = __ldg(p.act_scale);  // per-tensor: the scale is a multiplier (activation.cu:19-136)
        }
        if (row_valid) {
          uint8_t* dst = p.q_out + grow * (p.n / 2) + t.nt * 128 + wg * 64;
#pragma unroll
          for (int q = 0; q < 4; q++) {
            if (t.nt * 128 + wg * 64 + q * 16 >= p.n / 2) break;
            uint4 w;
            w.x = cvt_e4m3x4(v[q * 16 + 0] * inv, v[q * 16 + 1] * inv, v[q * 16 + 2] * inv, v[q * 16 + 3] * inv);
            w.y = cvt_e4m3x4(v[q * 16 + 4] * inv, v[q * 16 + 5] * inv, v[q * 16 + 6] * inv, v[q * 16 + 7] * inv);
            w.z = cvt_e4m3x4(v[q * 16 + 8] * inv, v[q * 16 + 9] * inv, v[q * 16 + 10] * inv, v[q * 16 + 11] * inv);
            w.w = cvt_e4m3x4(v[q * 16 + 12] * inv, v[q * 16 + 13] * inv, v[q * 16 + 14] * inv, v[q * 16 + 15] * inv);
            *reinterpret_cast<uint4*>(dst + q * 16) = w;
          }
        }
        if constexpr (kBlockwise) named_bar_sync(kEpiBar, 256);  // s_amax reuse by the next tile
      }
      if (p.debug & 8) t_epi += clock64() - te0;
    }
    if ((p.debug & 8) && warp == 4 && lane == 0) {
      long long* o = p.debug_out + (kFused ? 0 : 16 * 256) + 16 * blockIdx.x;
      o[11] = clock64() - ec0;
      o[12] = w_xs;
      o[13] = w_pfull;
      o[14] = w_tqe;
      o[15] = t_epi;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

static long long* g_dbg_buf[64] = {nullptr};  // HPC_B200_MOE_DEBUG & 8: [2][256][16] int64 per device

template <bool kBlockwise, bool kFused>
static int launch(const CUtensorMap& ta, const CUtensorMap& tb, const Params& p, cudaStream_t stream) {
  auto kern = group_gemm_fp8_kernel<kBlockwise, kFused>;
  static bool configured[64] = {false};
  const int dev = device_slot();
  if (!configured[dev]) {
    HPC_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    configured[dev] = true;
  }
  Params pp = p;
  static const int dbg = [] {
    const char* e = std::getenv("HPC_B200_MOE_DEBUG");
    return e ? std::atoi(e) : 0;
  }();
  pp.debug = dbg;
  pp.debug_out = nullptr;
  if (dbg & 8) {
    if (g_dbg_buf[dev] == nullptr) {
      HPC_CUDA_CHECK(cudaMalloc(&g_dbg_buf[dev], 2 * 16 * 256 * sizeof(long long)));
      HPC_CUDA_CHECK(cudaMemset(g_dbg_buf[dev], 0, 2 * 16 * 256 * sizeof(long long)));
    }
    pp.debug_out = g_dbg_buf[dev];
  }
  pp.tile_counter = scheduler_counter(stream);  // self-resetting pair of ints, one per stream
  if (pp.tile_counter == nullptr) return HPC_ERR_CUDA;
  HPC_CUDA_CHECK(launch_pdl(kern, dim3(sm_count()), dim3(kThreads), kSmemBytes, stream, 1, ta, tb, pp));
  return HPC_OK;
}

// Common launcher. mode bits: 1 = blockwise scales, 2 = fused activation epilogue.
int run(int mode, const void* x, const void* w, const int* seqlens, const int* cu_seqlens,
        const float* xscale_t, const float* wscale, const float* act_scale, void* y, void* q_out,
        float* q_scale_t, int num_group, int m, int n, int k, int m_pad, int kpad4, int scale_tile,
        int use_bf16_mul, cudaStream_t stream) {
  HPC_REQUIRE(num_group > 0 && num_group <= kMaxGroups, "group gemm: num_group %d not in (0, %d]",
              num_group, kMaxGroups);
  if (mode & 1) {
    HPC_REQUIRE(k % 128 == 0 && k >= 128, "group gemm: k (%d) must be a multiple of 128", k);
    HPC_REQUIRE(n % 128 == 0, "group gemm: n (%d) must be a multiple of 128", n);
    if (mode & 2) HPC_REQUIRE(n % 256 == 0, "fused act: gate_up rows (%d) must be a multiple of 256", n);
    HPC_REQUIRE(kpad4 % 4 == 0 && kpad4 * 128 >= k && kpad4 <= kMaxKB,
                "group gemm: weight-scale row length %d must be a multiple of 4 covering k/128 (<= %d)",
                kpad4, kMaxKB);
    HPC_REQUIRE(m_pad % 4 == 0 && scale_tile % 4 == 0 &&
                    (reinterpret_cast<uintptr_t>(xscale_t) & 15) == 0 &&
                    (reinterpret_cast<uintptr_t>(wscale) & 15) == 0,
                "group gemm: activation-scale columns (%d, tile %d) must be multiples of 4 and the "
                "scale tensors 16-byte aligned", m_pad, scale_tile);
  } else {
    HPC_REQUIRE(k % 16 == 0 && k >= 16, "group gemm: k (%d) must be a multiple of 16", k);
    HPC_REQUIRE(n % 64 == 0, "group gemm: n (%d) must be a multiple of 64", n);
    if (mode & 2) HPC_REQUIRE(n % 32 == 0, "fused act: gate_up rows (%d) must be a multiple of 32", n);
  }
  HPC_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(w) & 15) == 0,
              "group gemm: x and weight must be 16-byte aligned");
  if (m <= 0) return HPC_OK;
  HPC_REQUIRE(scale_tile > 0, "group gemm: scale_tile must be positive");

  CUtensorMap ta, tb;
  {
    uint64_t dims[2] = {static_cast<uint64_t>(k), static_cast<uint64_t>(m)};
    uint64_t strides[1] = {static_cast<uint64_t>(k)};
    uint32_t box[2] = {128, 128};
    int rc = encode_tmap_u8(&ta, x, 2, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B,
                            CU_TENSOR_MAP_L2_PROMOTION_L2_256B);
    if (rc) return rc;
  }
  {
    uint64_t dims[3] = {static_cast<uint64_t>(k), static_cast<uint64_t>(n),
                        static_cast<uint64_t>(num_group)};
    uint64_t strides[2] = {static_cast<uint64_t>(k), static_cast<uint64_t>(k) * n};
    uint32_t box[3] = {128, 128, 1};
    int rc = encode_tmap_u8(&tb, w, 3, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B,
                            CU_TENSOR_MAP_L2_PROMOTION_L2_256B);
    if (rc) return rc;
  }
  Params p;
  p.seqlens = seqlens;
  p.cu_seqlens = cu_seqlens;
  p.xscale_t = xscale_t;
  p.wscale = wscale;
  p.act_scale = act_scale;
  p.y = static_cast<__nv_bfloat16*>(y);
  p.q_out = static_cast<uint8_t*>(q_out);
  p.q_scale_t = q_scale_t;
  p.num_group = num_group;
  p.m_total = m;
  p.n = n;
  p.k = k;
  p.m_pad = m_pad;
  p.kpad4 = kpad4;
  p.scale_tile = scale_tile;
  p.use_bf16_mul = use_bf16_mul;
  switch (mode & 3) {
    case 0: return launch<false, false>(ta, tb, p, stream);
    case 1: return launch<true, false>(ta, tb, p, stream);
    case 2: return launch<false, true>(ta, tb, p, stream);
    default: return launch<true, true>(ta, tb, p, stream);
  }
}

// activation-scale column granule from the average rows per group
// (reference src/fuse_moe/entry.cc:525-543; replicated in group_gemm_blockwise_fp8.cu:384-456)
int scale_tile_from_avg(int avg) {
  if (avg <= 8) return 8;
  if (avg <= 16) return 16;
  if (avg <= 32) return 32;
  if (avg <= 48) return 48;
  if (avg <= 64) return 64;
  if (avg <= 96) return 48;
  if (avg <= 128) return 32;
  if (avg <= 144) return 48;
  return 64;
}

}  // namespace ggemm
}  // namespace b200

using namespace b200;  // NOLINT

// diagnostics (HPC_B200_MOE_DEBUG=8): copy the MMA threads' per-CTA counters of the last Gate-Up
// (fused) and Down / plain launches to the host: out[2][256][16] (see the kernel's debug_out stores)
extern "C" int hpc_group_gemm_debug_counters(long long* out_host) {
  const int dev = device_slot();
  HPC_REQUIRE(ggemm::g_dbg_buf[dev] != nullptr, "no debug counters recorded (HPC_B200_MOE_DEBUG=8?)");
  HPC_CUDA_CHECK(cudaMemcpy(out_host, ggemm::g_dbg_buf[dev], 2 * 16 * 256 * sizeof(long long),
                            cudaMemcpyDeviceToHost));
  return HPC_OK;
}

// replaces reference src/group_gemm/group_gemm.h:22-29 (group_gemm_blockwise_fp8_async). The
// tmas / tiles / cu_tiles / task_map / num_waves / update_tma / use_pdl arguments are accepted for
// signature compatibility; this build derives its tile schedule inside the kernel.
extern "C" int hpc_group_gemm_blockwise_fp8_async(
    void* y_ptr, const void* x_ptr, const void* w_ptr, const void* seqlens_ptr,
    const void* cu_seqlens_ptr, const void* xscale_ptr, const void* wscale_ptr, void* tmas_ptr,
    void* tiles_ptr, void* cu_tiles_ptr, void* task_map_ptr, int num_waves, int num_group, int m,
    int n, int k, int m_pad, int num_block_k_pad4, int num_seq_per_group_avg, int update_tma,
    int use_pdl, cudaStream_t stream) {
  (void)tmas_ptr; (void)tiles_ptr; (void)cu_tiles_ptr; (void)task_map_ptr; (void)num_waves;
  (void)update_tma; (void)use_pdl;
  return ggemm::run(1, x_ptr, w_ptr, static_cast<const int*>(seqlens_ptr),
                    static_cast<const int*>(cu_seqlens_ptr), static_cast<const float*>(xscale_ptr),
                    static_cast<const float*>(wscale_ptr), nullptr, y_ptr, nullptr, nullptr,
                    num_group, m, n, k, m_pad, num_block_k_pad4,
                    ggemm::scale_tile_from_avg(num_seq_per_group_avg), 0, stream);
}

// replaces reference src/group_gemm/group_gemm.h:12-20 (group_gemm_fp8_async, per-group y_scale)
extern "C" int hpc_group_gemm_fp8_async(void* y_ptr, const void* x_ptr, const void* w_ptr,
                                        const void* seqlens_ptr, const void* cu_seqlens_ptr,
                                        const void* y_scale, void* tmas_ptr, void* tiles_ptr,
                                        void* cu_tiles_ptr, void* task_map_ptr, int num_waves,
                                        int num_group, int m, int n, int k,
                                        int num_seq_per_group_avg, int update_tma, int use_pdl,
                                        cudaStream_t stream) {
  (void)tmas_ptr; (void)tiles_ptr; (void)cu_tiles_ptr; (void)task_map_ptr; (void)num_waves;
  (void)update_tma; (void)use_pdl; (void)num_seq_per_group_avg;
  return ggemm::run(0, x_ptr, w_ptr, static_cast<const int*>(seqlens_ptr),
                    static_cast<const int*>(cu_seqlens_ptr), nullptr,
                    static_cast<const float*>(y_scale), nullptr, y_ptr, nullptr, nullptr, num_group,
                    m, n, k, 0, 0, 64, 0, stream);
}
