// BF16 paged decode attention, dynamic split-k task map. Same skeleton as the
// fp8 kernel (decode_attn_fp8.cu), re-dimensioned for 2-byte operands:
//
//   * persistent grid = num_total_ctas; each CTA walks its bin of the task map (128-key tiles)
//   * warp 0  : TMA producer  - a K tile and a V tile are two [128 keys x 64 dims] halves each
//                               (128-byte swizzled rows); a page of 16 / 32 / 64 tokens lands as a
//                               [page x 64] box at row offset page_index * page in each half;
//                               3 stages of 64 KB
//   * warp 1  : tcgen05 issuer - S^T[128 keys, NQ] = K_tile . Q^T   (kind::f16, K-major, 8 x K=16)
//                                O^T[128 d, NQ]   = V_tile^T . P^T  (MN-major A: V as stored, the
//                                two 64-dim halves are the two MN atoms, LBO = half stride)
//   * warps 2-5: softmax      - one thread per key, log2-domain online softmax, P -> bf16 -> smem
//   accumulators in TMEM (S^T and O^T double buffered, 4*NQ columns).
//
// Semantics follow reference
//   src/attention/decode/sm90/dynamic/smallm_bf16_dim128_dynamic_splitk_kernels.cuh:29
//   src/attention/decode/sm90/util_kernels.cuh:280-305 (mask), :332-436 (online softmax)
// and the entry contract of src/attention/entry.cc:411-520 (attention_decode_bf16_entry).
#include <cstdlib>

#include "decode_common.cuh"

namespace b200 {
namespace decode_bf16 {

using decode::kD;
using decode::kSoftmaxBar;
using decode::kTaskStride;
using decode::kTileN;
using decode::load_task;
using decode::Params;
using decode::Task;

constexpr int kThreads = 192;
constexpr int kHalfBytes = kTileN * 128;       // [128 rows x 64 bf16]: 16 KB
constexpr int kTileBytes = 2 * kHalfBytes;     // one K or V tile: 32 KB
constexpr int kStageBytes = 2 * kTileBytes;    // K tile + V tile
constexpr int kNumStages = 3;
constexpr int kQHalfBytes = 4096;              // [32 query rows x 64 bf16]
constexpr int kQBytes = 2 * kQHalfBytes;

template <int NQ>
struct Smem {
  static constexpr int kPPlanes = NQ / 8;                  // 8 queries (16 B) per plane
  static constexpr int kPBytes = kPPlanes * kTileN * 16;   // one P^T buffer
  static constexpr int kOffStages = 0;
  static constexpr int kOffQ = kNumStages * kStageBytes;
  static constexpr int kOffP = kOffQ + 2 * kQBytes;
  static constexpr int kOffMax = kOffP + 2 * kPBytes;
  static constexpr int kOffBar = kOffMax + 2 * 4 * 32 * 4;
  static constexpr int kNumBars = 3 * kNumStages + 10;
  static constexpr int kOffTmem = kOffBar + kNumBars * 8;
  static constexpr int kTotal = kOffTmem + 16;
};

// Barrier protocol: identical to the fp8 kernel (see the comment there): k_full / v_full /
// stage_empty per stage, q_full / q_empty, s_full, p_full (4 warp arrivals), o_full.
template <int NQ, int RL>
__global__ void __launch_bounds__(kThreads, 1)
    decode_attn_bf16_kernel(const __grid_constant__ CUtensorMap tmap_q,
                            const __grid_constant__ CUtensorMap tmap_k,
                            const __grid_constant__ CUtensorMap tmap_v, const Params p,
                            const int page_log2) {
  using L = Smem<NQ>;
  extern __shared__ __align__(1024) uint8_t smem[];

  uint8_t* stages = smem + L::kOffStages;
  uint8_t* q_smem = smem + L::kOffQ;
  uint8_t* p_smem = smem + L::kOffP;
  float* smax = reinterpret_cast<float*>(smem + L::kOffMax);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::kOffBar);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + L::kOffTmem);

  uint64_t* k_full = bars;
  uint64_t* v_full = bars + kNumStages;
  uint64_t* stage_empty = bars + 2 * kNumStages;
  uint64_t* q_full = bars + 3 * kNumStages;
  uint64_t* q_empty = q_full + 2;
  uint64_t* s_full = q_full + 4;
  uint64_t* p_full = q_full + 6;
  uint64_t* o_full = q_full + 8;

  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int lane = tid & 31;

  {
    // zero both Q buffers: query rows beyond group * num_seq_q stay exact zeros
    uint4 z = make_uint4(0, 0, 0, 0);
    for (int i = tid; i < 2 * kQBytes / 16; i += kThreads) {
      reinterpret_cast<uint4*>(q_smem)[i] = z;
    }
    fence_proxy_async_smem();
  }
  if (warp == 0 && lane == 0) {
    prefetch_tensormap(&tmap_q);
    prefetch_tensormap(&tmap_k);
    prefetch_tensormap(&tmap_v);
    for (int i = 0; i < kNumStages; i++) {
      mbar_init(&k_full[i], 1);
      mbar_init(&v_full[i], 1);
      mbar_init(&stage_empty[i], 1);
    }
    for (int i = 0; i < 2; i++) {
      mbar_init(&q_full[i], 1);
      mbar_init(&q_empty[i], 1);
      mbar_init(&s_full[i], 1);
      mbar_init(&p_full[i], 4);
      mbar_init(&o_full[i], 1);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 128);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  pdl_launch_dependents();
  pdl_wait();

  const int ntpc1 = p.task_map[0];
  const int* bin = p.task_map + (1 + static_cast<long long>(blockIdx.x) * ntpc1) * kTaskStride;
  const int page = 1 << page_log2;
  const int ppt = kTileN >> page_log2;  // pages per 128-key tile: 8, 4 or 2

  if (warp == 0) {
    // =========================== TMA producer (whole warp, one elected lane issues) =========
    const uint64_t pol_stream = make_policy_evict_first();
    const int tiles_per_group = 32 / ppt;  // page ids fetched 32 at a time, one per lane
    const uint32_t page_bytes = static_cast<uint32_t>(page) * 128u;  // rows of one page in a half
    uint32_t n = 0;
    uint32_t qcnt = 0;
    Task t;
    for (const int* row = bin; load_task(row, t); row += kTaskStride) {
      t.ihead_kv = __shfl_sync(0xffffffffu, t.ihead_kv, 0);
      t.ibatch = __shfl_sync(0xffffffffu, t.ibatch, 0);
      t.num_tile_kv = __shfl_sync(0xffffffffu, t.num_tile_kv, 0);
      {
        const int qb = qcnt & 1;
        mbar_wait(&q_empty[qb], ((qcnt >> 1) & 1) ^ 1);
        if (elect_one()) {
          mbar_arrive_expect_tx(&q_full[qb], p.num_seq_q * p.group * kD * 2);
          uint8_t* qd = q_smem + qb * kQBytes;
          tma_load_3d(qd, &tmap_q, &q_full[qb], 0, t.ihead_kv * p.group, t.ibatch * p.num_seq_q);
          tma_load_3d(qd + kQHalfBytes, &tmap_q, &q_full[qb], 64, t.ihead_kv * p.group,
                      t.ibatch * p.num_seq_q);
        }
        __syncwarp();
      }
      qcnt++;
      const int nblk = (t.num_seqkv + page - 1) >> page_log2;
      const int* ids = p.block_ids + static_cast<long long>(t.ibatch) * p.num_seq_max_blocks +
                       (t.iseq_start >> page_log2);
      const int ntiles = t.num_tile_kv;
      const int kc1 = p.k_head_first ? t.ihead_kv : 0;
      const int kc2 = p.k_head_first ? 0 : t.ihead_kv;
      const int vc1 = p.v_head_first ? t.ihead_kv : 0;
      const int vc2 = p.v_head_first ? 0 : t.ihead_kv;
      for (int g0 = 0; g0 < ntiles; g0 += tiles_per_group) {
        int bi = g0 * ppt + lane;
        bi = bi < nblk ? bi : nblk - 1;  // pages past the end of the last tile re-read its last page
        const int my_id = __ldg(ids + bi);
        const int gt = (ntiles - g0) < tiles_per_group ? (ntiles - g0) : tiles_per_group;
        for (int tt = 0; tt < gt; tt++) {
          const uint32_t st = n % kNumStages;
          mbar_wait(&stage_empty[st], ((n / kNumStages) & 1) ^ 1);
          uint8_t* dst = stages + st * kStageBytes;
          if (elect_one()) {
            mbar_arrive_expect_tx(&k_full[st], kTileBytes);
            mbar_arrive_expect_tx(&v_full[st], kTileBytes);
          }
          __syncwarp();
          for (int j = 0; j < ppt; j++) {
            const int id = __shfl_sync(0xffffffffu, my_id, tt * ppt + j);
            if (elect_one()) {
              uint8_t* d = dst + j * page_bytes;
              tma_load_4d_hint(d, &tmap_k, &k_full[st], 0, kc1, kc2, id, pol_stream);
              tma_load_4d_hint(d + kHalfBytes, &tmap_k, &k_full[st], 64, kc1, kc2, id, pol_stream);
            }
            __syncwarp();
          }
          for (int j = 0; j < ppt; j++) {
            const int id = __shfl_sync(0xffffffffu, my_id, tt * ppt + j);
            if (elect_one()) {
              uint8_t* d = dst + kTileBytes + j * page_bytes;
              tma_load_4d_hint(d, &tmap_v, &v_full[st], 0, vc1, vc2, id, pol_stream);
              tma_load_4d_hint(d + kHalfBytes, &tmap_v, &v_full[st], 64, vc1, vc2, id, pol_stream);
            }
            __syncwarp();
          }
          n++;
        }
      }
    }
  } else if (warp == 1) {
    // =========================== tcgen05 issuer (whole warp, one elected lane issues) ========
    const uint32_t tmem_u = __shfl_sync(0xffffffffu, *tmem_slot, 0);
    constexpr uint32_t idesc_qk = make_idesc(128, NQ, kFmtBF16, kFmtBF16, 0, 0);
    constexpr uint32_t idesc_pv = make_idesc(128, NQ, kFmtBF16, kFmtBF16, 1, 1);
    //  K tile / Q rows : K-major, 128B swizzle, 8-row groups 1024 B apart; one MMA consumes 16
    //                    dims = 32 B of a row; dims 64..127 are the second half (+16 KB / +4 KB)
    //  V tile          : MN-major (d contiguous as stored), 128B swizzle: the MN extent of 128 dims
    //                    is two 64-dim atoms kHalfBytes apart (LBO), 8-key groups 1024 B apart
    //                    (SBO); one MMA consumes 16 keys = 2048 B
    //  P^T             : MN-major, no swizzle, [plane][key][8 queries]: 8-key core matrices 128 B
    //                    apart (LBO), planes kTileN*16 B apart (SBO); one MMA consumes 256 B
    const uint64_t kdesc0 = make_smem_desc(smem_u32(stages), 16, 1024, kLayoutSW128);
    const uint64_t vdesc0 =
        make_smem_desc(smem_u32(stages) + kTileBytes, kHalfBytes, 1024, kLayoutSW128);
    const uint64_t qdesc0 = make_smem_desc(smem_u32(q_smem), 16, 1024, kLayoutSW128);
    const uint64_t pdesc0 = make_smem_desc(smem_u32(p_smem), 128, kTileN * 16, kLayoutNone);

    auto issue_pv = [&](uint32_t m) {
      const uint32_t st = m % kNumStages;
      const uint32_t buf = m & 1;
      mbar_wait(&p_full[buf], (m >> 1) & 1);
      mbar_wait(&v_full[st], (m / kNumStages) & 1);
      tc_fence_after();
      const uint64_t ad = vdesc0 + static_cast<uint64_t>(st * (kStageBytes >> 4));
      const uint64_t bd = pdesc0 + static_cast<uint64_t>(buf * (L::kPBytes >> 4));
      const uint32_t d = tmem_u + 2 * NQ + buf * NQ;
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < 8; k++) {
          umma_f16(d, ad + k * (2048 >> 4), bd + k * (256 >> 4), idesc_pv, k > 0);
        }
        umma_commit(&o_full[buf]);
        umma_commit(&stage_empty[st]);
      }
      __syncwarp();
    };

    uint32_t n = 0;
    uint32_t qcnt = 0;
    Task t;
    for (const int* row = bin; load_task(row, t); row += kTaskStride) {
      const int qb = qcnt & 1;
      mbar_wait(&q_full[qb], (qcnt >> 1) & 1);
      const uint64_t bd = qdesc0 + static_cast<uint64_t>(qb * (kQBytes >> 4));
      const int ntiles = __shfl_sync(0xffffffffu, t.num_tile_kv, 0);
      for (int tt = 0; tt < ntiles; tt++) {
        const uint32_t st = n % kNumStages;
        const uint32_t buf = n & 1;
        mbar_wait(&k_full[st], (n / kNumStages) & 1);
        tc_fence_after();
        const uint64_t ad = kdesc0 + static_cast<uint64_t>(st * (kStageBytes >> 4));
        const uint32_t d = tmem_u + buf * NQ;
        if (elect_one()) {
#pragma unroll
          for (int h = 0; h < 2; h++) {
#pragma unroll
            for (int k = 0; k < 4; k++) {
              umma_f16(d, ad + h * (kHalfBytes >> 4) + k * (32 >> 4),
                       bd + h * (kQHalfBytes >> 4) + k * (32 >> 4), idesc_qk, (h | k) > 0);
            }
          }
          umma_commit(&s_full[buf]);
          if (tt == ntiles - 1) umma_commit(&q_empty[qb]);
        }
        __syncwarp();
        if (n > 0) issue_pv(n - 1);
        n++;
      }
      qcnt++;
    }
    if (n > 0) issue_pv(n - 1);
  } else {
    // =========================== softmax / epilogue warps =================================
    const int quad = warp & 3;
    const int row_in_tile = quad * 32 + lane;  // key index (S^T) and d index (O^T)
    const int sw = warp - 2;
    const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16);
    const int* chunk_table = p.task_map + kTaskStride * (ntpc1 * p.task_map[1] + 1);

    uint32_t n = 0;
    Task t;
    for (const int* row = bin; load_task(row, t); row += kTaskStride) {
      float mrun[RL], lrun[RL], alpha_pend[RL];
      float acc[RL];
#pragma unroll
      for (int r = 0; r < RL; r++) {
        mrun[r] = -INFINITY;
        lrun[r] = 0.f;
        alpha_pend[r] = 1.f;
        acc[r] = 0.f;
      }
      const float c = p.softmax_scale_log2;
      const int lim_len = t.num_seqkv;
      const int lim_causal = t.num_seqkvcache;
      const int ntiles = t.num_tile_kv;

      auto consume_o = [&](uint32_t m) {
        const uint32_t buf = m & 1;
        mbar_wait(&o_full[buf], (m >> 1) & 1);
        tc_fence_after();
        uint32_t o[NQ];
        if constexpr (NQ == 16) {
          tmem_ld_x16(lane_addr + 2 * NQ + buf * NQ, o);
        } else {
          tmem_ld_x32(lane_addr + 2 * NQ + buf * NQ, o);
        }
        tmem_wait_ld();
#pragma unroll
        for (int r = 0; r < RL; r++) {
          acc[r] = acc[r] * alpha_pend[r] + __uint_as_float(o[r]);
        }
      };

      for (int tt = 0; tt < ntiles; tt++) {
        const uint32_t buf = n & 1;
        const uint32_t ph = (n >> 1) & 1;
        mbar_wait(&s_full[buf], ph);
        tc_fence_after();
        uint32_t sraw[NQ];
        if constexpr (NQ == 16) {
          tmem_ld_x16(lane_addr + buf * NQ, sraw);
        } else {
          tmem_ld_x32(lane_addr + buf * NQ, sraw);
        }
        tmem_wait_ld();

        const int key = tt * kTileN + row_in_tile;
        const int lim_min = lim_len < lim_causal ? lim_len : lim_causal;
        const bool need_mask = (tt + 1) * kTileN > lim_min;
        float x[RL];
        float* mx = smax + (buf * 4 + sw) * 32;
#pragma unroll
        for (int r = 0; r < RL; r++) {
          float v = __uint_as_float(sraw[r]) * c;
          if (need_mask) {
            const int sq = r / p.group;
            const bool dead = (key >= lim_len) || (key > lim_causal + sq);
            v = dead ? -INFINITY : v;
          }
          x[r] = v;
          const float wm = warp_max_f32(v);
          if (lane == 0) mx[r] = wm;
        }
        named_bar_sync(kSoftmaxBar, 128);
        const float* mall = smax + buf * 4 * 32;
        float pv[RL];
#pragma unroll
        for (int r = 0; r < RL; r++) {
          const float tm = fmaxf(fmaxf(mall[r], mall[32 + r]), fmaxf(mall[64 + r], mall[96 + r]));
          const float mold = mrun[r];
          const float mnew = fmaxf(mold, tm);
          float a = 1.f, e = 0.f;
          if (mnew != -INFINITY) {
            a = exp2_approx(mold - mnew);
            e = exp2_approx(x[r] - mnew);
          }
          mrun[r] = mnew;
          lrun[r] = lrun[r] * a + e;
          pv[r] = e;
          x[r] = a;  // applied when the O tile of this key tile is consumed
        }
        // ---- P^T row of this key -> smem (bf16), 8 queries per plane ----
        {
          uint8_t* pb = p_smem + buf * L::kPBytes + row_in_tile * 16;
#pragma unroll
          for (int pl = 0; pl < L::kPPlanes; pl++) {
            float f[8];
#pragma unroll
            for (int i = 0; i < 8; i++) {
              const int r = pl * 8 + i;
              f[i] = (r < RL) ? pv[r < RL ? r : 0] : 0.f;
            }
            uint4 w;
            w.x = pack_bf16x2(f[0], f[1]);
            w.y = pack_bf16x2(f[2], f[3]);
            w.z = pack_bf16x2(f[4], f[5]);
            w.w = pack_bf16x2(f[6], f[7]);
            *reinterpret_cast<uint4*>(pb + pl * kTileN * 16) = w;
          }
        }
        fence_proxy_async_smem();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_full[buf]);

        if (tt > 0) consume_o(n - 1);
#pragma unroll
        for (int r = 0; r < RL; r++) alpha_pend[r] = x[r];
        n++;
      }
      if (ntiles > 0) consume_o(n - 1);

      // ---- task epilogue: 1/sum, partial O and LSE out (or bf16 y when the pair has one chunk) --
      float* red = smax;
      named_bar_sync(kSoftmaxBar, 128);
#pragma unroll
      for (int r = 0; r < RL; r++) {
        const float ws = warp_sum_f32(lrun[r]);
        if (lane == 0) red[sw * 32 + r] = ws;
      }
      named_bar_sync(kSoftmaxBar, 128);
      const long long chunk_row = static_cast<long long>(t.ibatch) * p.max_splitk + t.ichunk;
      const bool single =
          p.y != nullptr && __ldg(chunk_table + t.ihead_kv * p.num_batch + t.ibatch) == 1;
#pragma unroll
      for (int r = 0; r < RL; r++) {
        const int sq = r / p.group;
        const int g = r - sq * p.group;
        if (sq < p.num_seq_q) {
          const float tot = red[r] + red[32 + r] + red[64 + r] + red[96 + r];
          const float inv = tot != 0.f ? rcp_approx(tot) : 0.f;
          if (single) {
            p.y[(static_cast<long long>(t.ibatch) * p.num_seq_q + sq) * p.ld_y +
                (t.ihead_kv * p.group + g) * kD + row_in_tile] = __float2bfloat16_rn(acc[r] * inv);
            continue;
          }
          const long long orow =
              (chunk_row * p.num_seq_q + sq) * p.num_head_q + t.ihead_kv * p.group + g;
          p.split_out[orow * kD + row_in_tile] = acc[r] * inv;
          if (row_in_tile == r) {
            const float l = (mrun[r] == -INFINITY) ? -INFINITY : mrun[r] + log2_approx(tot);
            p.lse[((chunk_row * p.num_head_kv + t.ihead_kv) * p.num_seq_q + sq) * p.lse_pad + g] = l;
          }
        }
      }
      named_bar_sync(kSoftmaxBar, 128);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tmem_dealloc(tmem_base, 128);
  }
}

template <int NQ, int RL>
static int launch_attn(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv,
                       const Params& p, int page_log2, int grid, cudaStream_t stream) {
  using L = Smem<NQ>;
  auto kern = decode_attn_bf16_kernel<NQ, RL>;
  static bool configured[64] = {false};
  const int dev = device_slot();
  if (!configured[dev]) {
    HPC_CUDA_CHECK(
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::kTotal));
    configured[dev] = true;
  }
  HPC_CUDA_CHECK(
      launch_pdl(kern, dim3(grid), dim3(kThreads), L::kTotal, stream, 1, tq, tk, tv, p, page_log2));
  return HPC_OK;
}

}  // namespace decode_bf16
}  // namespace b200

using namespace b200;  // NOLINT

// Replaces reference src/attention/decode/decode.h attention_decode_bf16_async (entry
// src/attention/entry.cc:411-520). Strides in ELEMENTS (bf16), as the torch tensors report them.
extern "C" int hpc_attention_decode_bf16_async(
    void* y_ptr, void* lse_ptr, void* split_out_ptr, const int* task_map_ptr, const void* q_ptr,
    void* kcache_ptr, void* vcache_ptr, const int* block_ids_ptr, const int* num_seq_kvcache_ptr,
    int* split_flag_ptr, int new_kv_included, int splitk, int num_batch, int num_seq_q,
    int num_head_q, int num_head_k, int num_head_v, int num_dim_qk, int num_dim_v,
    int num_kvcache_blocks, int block_size, int num_seq_max_blocks, int ldY, int ldQ,
    int64_t kcache_block_stride, int64_t kcache_token_stride, int64_t kcache_head_stride,
    int64_t vcache_block_stride, int64_t vcache_token_stride, int64_t vcache_head_stride,
    cudaStream_t stream) {
  (void)num_seq_kvcache_ptr;  // lengths come from the task map
  (void)split_flag_ptr;
  (void)new_kv_included;
  HPC_REQUIRE(task_map_ptr != nullptr, "attention_decode_bf16: a task_map is required on sm_100");
  HPC_REQUIRE(num_dim_qk == 128 && num_dim_v == 128, "head dim must be 128");
  HPC_REQUIRE(block_size == 16 || block_size == 32 || block_size == 64,
              "kvcache paged blocksize must be 16, 32 or 64");
  HPC_REQUIRE(num_head_k == num_head_v && num_head_k > 0 && num_head_q % num_head_k == 0,
              "bad head counts q=%d k=%d v=%d", num_head_q, num_head_k, num_head_v);
  const int group = num_head_q / num_head_k;
  const int rows = group * num_seq_q;
  HPC_REQUIRE(rows >= 1 && rows <= 32 && group <= 16,
              "heads_per_group * num_seq_q = %d not in [1, 32]", rows);
  HPC_REQUIRE(splitk >= 1, "splitk (max chunks) must be >= 1");
  HPC_REQUIRE((reinterpret_cast<uintptr_t>(q_ptr) & 15) == 0 && (ldQ % 8) == 0,
              "q must be 16-byte aligned");
  HPC_REQUIRE((kcache_block_stride % 8) == 0 && (kcache_token_stride % 8) == 0 &&
                  (kcache_head_stride % 8) == 0 && (vcache_block_stride % 8) == 0 &&
                  (vcache_token_stride % 8) == 0 && (vcache_head_stride % 8) == 0 &&
                  (reinterpret_cast<uintptr_t>(kcache_ptr) & 15) == 0 &&
                  (reinterpret_cast<uintptr_t>(vcache_ptr) & 15) == 0,
              "kv cache base and strides must be multiples of 16 bytes");
  const int page_log2 = block_size == 16 ? 4 : block_size == 32 ? 5 : 6;

  CUtensorMap tq, tk, tv;
  {
    uint64_t dims[3] = {128, static_cast<uint64_t>(num_head_q),
                        static_cast<uint64_t>(num_batch) * num_seq_q};
    uint64_t strides[2] = {256, static_cast<uint64_t>(ldQ) * 2};
    uint32_t box[3] = {64, static_cast<uint32_t>(group), static_cast<uint32_t>(num_seq_q)};
    int rc = encode_tmap(&tq, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, q_ptr, 3, dims, strides, box,
                         CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
  }
  auto encode_cache = [&](CUtensorMap* tm, const void* base, int heads, int64_t blk_stride,
                          int64_t tok_stride, int64_t head_stride, int* head_first) -> int {
    // strides ascending: (d, head, token, blk) for NHD caches, (d, token, head, blk) for HND;
    // the smem image of a [page x 64] box is the same
    *head_first = head_stride <= tok_stride ? 1 : 0;
    uint64_t dims[4];
    uint64_t strides[3];
    uint32_t box[4];
    dims[0] = 128;
    box[0] = 64;
    if (*head_first) {
      dims[1] = static_cast<uint64_t>(heads);
      dims[2] = static_cast<uint64_t>(block_size);
      strides[0] = static_cast<uint64_t>(head_stride) * 2;
      strides[1] = static_cast<uint64_t>(tok_stride) * 2;
      box[1] = 1;
      box[2] = static_cast<uint32_t>(block_size);
    } else {
      dims[1] = static_cast<uint64_t>(block_size);
      dims[2] = static_cast<uint64_t>(heads);
      strides[0] = static_cast<uint64_t>(tok_stride) * 2;
      strides[1] = static_cast<uint64_t>(head_stride) * 2;
      box[1] = static_cast<uint32_t>(block_size);
      box[2] = 1;
    }
    dims[3] = static_cast<uint64_t>(num_kvcache_blocks);
    strides[2] = static_cast<uint64_t>(blk_stride) * 2;
    box[3] = 1;
    // promotion no wider than one head's contiguous run (see the fp8 kernel)
    const CUtensorMapL2promotion promo =
        (tok_stride == 128) ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B : CU_TENSOR_MAP_L2_PROMOTION_L2_128B;
    return encode_tmap(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, base, 4, dims, strides, box,
                       CU_TENSOR_MAP_SWIZZLE_128B, promo);
  };
  int k_head_first = 1, v_head_first = 1;
  {
    int rc = encode_cache(&tk, kcache_ptr, num_head_k, kcache_block_stride, kcache_token_stride,
                          kcache_head_stride, &k_head_first);
    if (rc) return rc;
    rc = encode_cache(&tv, vcache_ptr, num_head_v, vcache_block_stride, vcache_token_stride,
                      vcache_head_stride, &v_head_first);
    if (rc) return rc;
  }

  decode::Params p = {};
  p.task_map = task_map_ptr;
  p.block_ids = block_ids_ptr;
  p.split_out = static_cast<float*>(split_out_ptr);
  p.lse = static_cast<float*>(lse_ptr);
  p.y = static_cast<__nv_bfloat16*>(y_ptr);
  p.ld_y = ldY;
  p.num_batch = num_batch;
  p.num_seq_q = num_seq_q;
  p.num_head_q = num_head_q;
  p.num_head_kv = num_head_k;
  p.group = group;
  p.num_seq_max_blocks = num_seq_max_blocks;
  p.max_splitk = splitk;
  p.lse_pad = (group + 7) / 8 * 8;
  p.k_head_first = k_head_first;
  p.v_head_first = v_head_first;
  p.softmax_scale_log2 = 1.4426950408889634f / sqrtf(static_cast<float>(num_dim_qk));

  const int grid = splitk;  // == num_total_ctas of the task map
  int rc;
  if (rows <= 4) {
    rc = decode_bf16::launch_attn<16, 4>(tq, tk, tv, p, page_log2, grid, stream);
  } else if (rows <= 8) {
    rc = decode_bf16::launch_attn<16, 8>(tq, tk, tv, p, page_log2, grid, stream);
  } else if (rows <= 12) {
    rc = decode_bf16::launch_attn<16, 12>(tq, tk, tv, p, page_log2, grid, stream);
  } else if (rows <= 16) {
    rc = decode_bf16::launch_attn<16, 16>(tq, tk, tv, p, page_log2, grid, stream);
  } else if (rows <= 24) {
    rc = decode_bf16::launch_attn<32, 24>(tq, tk, tv, p, page_log2, grid, stream);
  } else {
    rc = decode_bf16::launch_attn<32, 32>(tq, tk, tv, p, page_log2, grid, stream);
  }
  if (rc) return rc;
  HPC_CUDA_CHECK(decode::launch_combine(static_cast<__nv_bfloat16*>(y_ptr), p.split_out, p.lse,
                                        task_map_ptr, num_batch, num_seq_q, num_head_q, num_head_k,
                                        group, splitk, p.lse_pad, ldY, stream));
  return HPC_OK;
}
