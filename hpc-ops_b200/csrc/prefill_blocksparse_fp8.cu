// FP8 block-sparse (and dense) causal prefill attention over a paged KV cache (B200 / sm_100a).
//
// Replaces reference src/attention/prefill/kernels.cuh:1978-2554 (q per-token/head, k/v per-tensor)
// and :2558-3150 (q per-token/head, k per-token/head in-cache scales, v per-head), their launcher
// src/attention/prefill/warp_spec_with_kvcache_blocksparse_fp8_dim128.cu:19-253 and the varlen TMA
// patch kernel (kernels.cuh:169-215, not needed here: Q is addressed with a 3-D descriptor).
//
// Work item = (batch, q head, 128-row Q tile); it visits only the 128x128 KV tiles that are set in
// block_mask (absolute KV tile index, Q tile index relative to the request's first new token):
//   active = { j < min(num_tile_kv, Kb) : mask[b, hq, mq, j] } U { Kb if Kb < num_tile_kv }
// Items are handed out heaviest-first (largest mq first) by a global atomic counter.
//
// CTA = 256 threads, two CTAs resident per SM:
//   warp 0     : TMA producer (Q tile, K pages -> 2-slot ring, V pages -> 2-slot ring, per-token k
//                scales), builds the active-tile list of each item with ballot compaction
//   warp 1     : tcgen05 issuer (one thread). S[128 q, 128 keys] = Q . K^T (K-major A/B);
//                O[128 q, 128 d] += P . V with P K-major from smem and V MN-major exactly as it
//                lies in the cache (no software transpose, unlike wgmma fp8: utils.cuh:461-521).
//                QK(n+1) is issued as soon as the softmax threads hold S(n) in registers, i.e. it
//                runs under softmax(n); PV(n) follows when P(n) is in smem.
//   warps 4-7  : softmax, one thread per query row (row max / sum are thread-local): the whole S
//                row (128 fp32) is pulled TMEM -> registers at once, then scale, causal/length
//                mask, base-2 online softmax with a lazy reference maximum, P*256 -> e4m3 into
//                128B-swizzled smem. O stays in TMEM across the tiles of an item (rare rescale by
//                tcgen05.ld/st) and is read once by the epilogue.
// TMEM: 256 columns per CTA (S at +0, O at +128).
#include <cstdlib>
#include <type_traits>

#include "common.cuh"
#include "host_utils.h"
#include "prefill_common.cuh"

namespace b200 {
namespace prefill {

constexpr int kThreads = 256;

// Two CTAs are resident per SM (each 256 threads, ~103 KB smem, 256 TMEM columns): the second CTA
// fills the issue slots and the tensor pipe while the first one waits on a barrier.
template <bool kKPerToken>
struct Smem {
  static constexpr int kOffK = 0;                           // kStages x 16 KB
  static constexpr int kOffV = kOffK + kStages * kTileBytes;  // kStages x 16 KB
  static constexpr int kOffQ = kOffV + kStages * kTileBytes;  // 16 KB
  static constexpr int kOffP = kOffQ + kTileBytes;          // 16 KB
  static constexpr int kOffKs = kOffP + kTileBytes;         // kKsBufs x 128 floats
  static constexpr int kOffList = kOffKs + kKsBufs * 128 * 4;
  static constexpr int kListStride = kMaxKvTiles + 8;  // int16 entries
  static constexpr int kOffBar = kOffList + kListStride * 2;
  static constexpr int kNumBars = 4 * kStages + 5;
  static constexpr int kOffTmem = kOffBar + kNumBars * 8;
  static constexpr int kTotal = kOffTmem + 64;
};

// Protocol (all mbarriers; "phase" = use-count parity), n = running KV-tile counter of the CTA:
//   q_full      producer: item published (list + work id in smem, Q tile landed)
//   q_empty     MMA thread (commit after the item's last QK) + 128 softmax threads (item finished)
//   k_full/k_empty[slot]   K ring; the slot is released by the commit after QK(n)
//   v_full/v_empty[slot]   V ring; released by the commit after PV(n). v_empty also tells the
//                          softmax threads that PV(n) is complete (P buffer free, O consistent)
//   s_full      commit after QK(n)                            -> softmax
//   s_free      the 4 softmax warps hold S(n) in registers      -> MMA thread may issue QK(n+1)
//   p_full      the 4 softmax warps wrote P(n) (and rescaled O) -> MMA thread issues PV(n)
// The k-scale buffer of tile n+2 was last read by softmax(n-2), which every softmax thread
// finished before arriving on s_free(n-1), which precedes QK(n) and so the release of K slot n%2.
// kPoly: exponentials per group of 8 scores that take the polynomial path (0, 2 or 4)
template <bool kKPerToken, int kPoly>
__global__ void __launch_bounds__(kThreads, 2)
    prefill_blocksparse_fp8_kernel(const __grid_constant__ CUtensorMap tmap_q,
                                   const __grid_constant__ CUtensorMap tmap_k,
                                   const __grid_constant__ CUtensorMap tmap_v, const Params p) {
  using L = Smem<kKPerToken>;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* k_smem = smem + L::kOffK;
  uint8_t* v_smem = smem + L::kOffV;
  uint8_t* q_smem = smem + L::kOffQ;
  uint8_t* p_smem = smem + L::kOffP;
  float* ks_smem = reinterpret_cast<float*>(smem + L::kOffKs);
  int16_t* list = reinterpret_cast<int16_t*>(smem + L::kOffList);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::kOffBar);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + L::kOffTmem);
  int* s_work = reinterpret_cast<int*>(tmem_slot + 4);  // work id, number of active tiles

  uint64_t* k_full = bars;
  uint64_t* k_empty = bars + kStages;
  uint64_t* v_full = bars + 2 * kStages;
  uint64_t* v_empty = bars + 3 * kStages;
  uint64_t* q_full = bars + 4 * kStages;
  uint64_t* q_empty = q_full + 1;
  uint64_t* s_full = q_full + 2;
  uint64_t* s_free = q_full + 3;
  uint64_t* p_full = q_full + 4;

  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int lane = tid & 31;

  if (warp == 0 && lane == 0) {
    prefetch_tensormap(&tmap_q);
    prefetch_tensormap(&tmap_k);
    prefetch_tensormap(&tmap_v);
    for (int i = 0; i < kStages; i++) {
      mbar_init(&k_full[i], 1);
      mbar_init(&k_empty[i], 1);
      mbar_init(&v_full[i], 1);
      mbar_init(&v_empty[i], 1);
    }
    mbar_init(q_full, 1);
    mbar_init(q_empty, 1 + 4);  // MMA commit + one arrive per softmax warp
    mbar_init(s_full, 1);
    mbar_init(s_free, 4);       // one arrive per softmax warp (128 per-thread arrives serialise on the barrier)
    mbar_init(p_full, 4);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 256);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 48;");
    if (warp == 0) {
      // =========================== producer ==============================================
      const uint64_t pol_kv = make_policy_evict_last();  // KV of a request is re-read by its q-heads
      uint32_t n = 0;     // kv tile counter (rings)
      uint32_t item = 0;  // item counter
      while (true) {
        int w = 0;
        if (lane == 0) w = atomicAdd(p.work_counter, 1);
        w = __shfl_sync(0xffffffffu, w, 0);
        Work k;
        const bool alive = decode_work(p, w, k);
        if (alive && k.rows == 0) continue;  // Q tile beyond this request
        mbar_wait(q_empty, (item & 1) ^ 1);
        int nact = 0;
        if (alive) {
          // ---- active tile list (ballot compaction, ascending j) ----
          const int kb = p.block_mask ? p.mask_kb : k.num_tile_kv;
          const int lim = k.num_tile_kv < kb ? k.num_tile_kv : kb;
          const uint8_t* mrow =
              p.block_mask
                  ? p.block_mask + ((static_cast<long long>(k.b) * p.num_head_q + k.hq) * p.mask_mq +
                                    (k.mq < p.mask_mq ? k.mq : p.mask_mq - 1)) *
                                       p.mask_kb
                  : nullptr;
          for (int j0 = 0; j0 < lim; j0 += 32) {
            const int j = j0 + lane;
            const bool on = (j < lim) && (mrow == nullptr || mrow[j] != 0);
            const unsigned m = __ballot_sync(0xffffffffu, on);
            if (on) list[nact + __popc(m & ((1u << lane) - 1))] = static_cast<int16_t>(j);
            nact += __popc(m);
          }
          if (p.block_mask && kb < k.num_tile_kv) {  // one unconditional tile past the mask width
            if (lane == 0) list[nact] = static_cast<int16_t>(kb);
            nact++;
          }
        }
        __syncwarp();
        // warp-uniform copies of what the elected lane's TMA instructions take as operands (the
        // compiler cannot prove values loaded from memory uniform and would wrap every TMA /
        // mbarrier instruction in an elect + R2UR.BROADCAST loop)
        const int u_hq = __shfl_sync(0xffffffffu, k.hq, 0);
        const int u_q0 = __shfl_sync(0xffffffffu, k.q0, 0);
        if (elect_one()) {
          s_work[0] = alive ? w : -1;
          s_work[1] = nact;
          if (alive) {
            mbar_arrive_expect_tx(q_full, kTileBytes);
            tma_load_3d(q_smem, &tmap_q, q_full, 0, u_hq, u_q0);
          } else {
            mbar_arrive(q_full);
          }
        }
        __syncwarp();
        item++;
        if (!alive) {
          // the last CTA out re-arms the work counter for the next launch (no memset node between
          // launches; one counter pair per stream, see host_utils.h scheduler_counter)
          if (lane == 0) {
            __threadfence();
            if (atomicAdd(p.work_counter + 1, 1) == static_cast<int>(gridDim.x) - 1) {
              p.work_counter[0] = 0;
              p.work_counter[1] = 0;
              __threadfence();
            }
          }
          break;
        }
        // ---- K/V tiles of the active list ----
        const int hkv = u_hq / p.group;
        const int nblk = (k.seq_kv + kPage - 1) / kPage;
        const int* ids = p.block_ids + static_cast<long long>(k.b) * p.max_blocks;
        const int kc1 = p.k_head_first ? hkv : 0, kc2 = p.k_head_first ? 0 : hkv;
        const int vc1 = p.v_head_first ? hkv : 0, vc2 = p.v_head_first ? 0 : hkv;
        for (int i0 = 0; i0 < nact; i0 += 16) {
          // lanes 2t / 2t+1 fetch the two page ids of tile i0+t
          const int ti = i0 + (lane >> 1);
          int id = 0;
          if (ti < nact) {
            int blk = static_cast<int>(list[ti]) * 2 + (lane & 1);
            blk = blk < nblk ? blk : nblk - 1;  // a missing 2nd page re-reads the 1st (keys masked)
            id = __ldg(ids + blk);
          }
          const int cnt = (nact - i0) < 16 ? (nact - i0) : 16;
          for (int t = 0; t < cnt; t++) {
            const int id0 = __shfl_sync(0xffffffffu, id, 2 * t);
            const int id1 = __shfl_sync(0xffffffffu, id, 2 * t + 1);
            const uint32_t st = n % kStages;
            const uint32_t ph = ((n / kStages) & 1) ^ 1;
            mbar_wait(&k_empty[st], ph);  // QK(n - 2) finished
            if (elect_one()) {
              uint8_t* kd8 = k_smem + st * kTileBytes;
              mbar_arrive_expect_tx(&k_full[st], kTileBytes + (kKPerToken ? 512 : 0));
              tma_load_4d_hint(kd8, &tmap_k, &k_full[st], 0, kc1, kc2, id0, pol_kv);
              tma_load_4d_hint(kd8 + kTileBytes / 2, &tmap_k, &k_full[st], 0, kc1, kc2, id1, pol_kv);
              if constexpr (kKPerToken) {
                // scales of token t of a page: kscale[page, t / 32, hkv, t % 32]
                float* kd = ks_smem + (n % kKsBufs) * 128;
                const float* s0 = p.kscale + id0 * p.ks_stride_blk + hkv * p.ks_stride_head;
                const float* s1 = p.kscale + id1 * p.ks_stride_blk + hkv * p.ks_stride_head;
                bulk_load_1d(kd, s0, 128, &k_full[st]);
                bulk_load_1d(kd + 32, s0 + p.ks_stride_grp, 128, &k_full[st]);
                bulk_load_1d(kd + 64, s1, 128, &k_full[st]);
                bulk_load_1d(kd + 96, s1 + p.ks_stride_grp, 128, &k_full[st]);
              }
            }
            __syncwarp();
            mbar_wait(&v_empty[st], ph);  // PV(n - 2) finished
            if (elect_one()) {
              uint8_t* vd8 = v_smem + st * kTileBytes;
              mbar_arrive_expect_tx(&v_full[st], kTileBytes);
              tma_load_4d_hint(vd8, &tmap_v, &v_full[st], 0, vc1, vc2, id0, pol_kv);
              tma_load_4d_hint(vd8 + kTileBytes / 2, &tmap_v, &v_full[st], 0, vc1, vc2, id1, pol_kv);
            }
            __syncwarp();
            n++;
          }
        }
      }
    } else if (warp == 1) {
      // =========================== tcgen05 issuer (whole warp, one elected lane issues) =======
      const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);  // warp-uniform copy
      constexpr uint32_t idesc_qk = make_idesc(128, 128, kFmtE4M3, kFmtE4M3, 0, 0);
      constexpr uint32_t idesc_pv = make_idesc(128, 128, kFmtE4M3, kFmtE4M3, 0, 1);
      const uint64_t kdesc0 = make_smem_desc(smem_u32(k_smem), 16, 1024, kLayoutSW128);
      const uint64_t vdesc0 = make_smem_desc(smem_u32(v_smem), 16, 1024, kLayoutSW128);
      const uint64_t qdesc = make_smem_desc(smem_u32(q_smem), 16, 1024, kLayoutSW128);
      const uint64_t pdesc = make_smem_desc(smem_u32(p_smem), 16, 1024, kLayoutSW128);
      // S(m) = Q . K(m)^T; needs K(m) in smem and S(m-1) drained into the softmax registers
      auto issue_qk = [&](const uint32_t m, const bool last_of_item) {
        const uint32_t st = m % kStages;
        mbar_wait(&k_full[st], (m / kStages) & 1);
        if (m > 0) mbar_wait(s_free, (m - 1) & 1);
        tc_fence_after();
        const uint64_t kd = kdesc0 + static_cast<uint64_t>(st * (kTileBytes >> 4));
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < 4; k++) umma_f8(tmem_base, qdesc + k * 2, kd + k * 2, idesc_qk, k > 0);
          umma_commit(s_full);
          umma_commit(&k_empty[st]);
          if (last_of_item) umma_commit(q_empty);  // Q tile / list reusable after this QK
        }
        __syncwarp();
      };
      uint32_t n = 0;
      uint32_t item = 0;
      while (true) {
        mbar_wait(q_full, item & 1);
        const int w = __shfl_sync(0xffffffffu, s_work[0], 0);
        const int nact = __shfl_sync(0xffffffffu, s_work[1], 0);
        if (w < 0) break;
        if (nact == 0) {
          if (elect_one()) umma_commit(q_empty);
          __syncwarp();
        } else {
          issue_qk(n, nact == 1);
          for (int i = 0; i < nact; i++) {
            if (i + 1 < nact) issue_qk(n + 1, i + 2 == nact);  // runs under softmax(n)
            const uint32_t st = n % kStages;
            mbar_wait(&v_full[st], (n / kStages) & 1);
            mbar_wait(p_full, n & 1);
            tc_fence_after();
            const uint64_t vd = vdesc0 + static_cast<uint64_t>(st * (kTileBytes >> 4));
            if (elect_one()) {
#pragma unroll
              for (int k = 0; k < 4; k++) {
                // A = P (K-major, 32 B per MMA); B = V as stored: MN-major, 32 keys = 4096 B per MMA
                umma_f8(tmem_base + 128, pdesc + k * 2, vd + k * (4096 >> 4), idesc_pv,
                        (k > 0) || (i > 0));
              }
              umma_commit(&v_empty[st]);
            }
            __syncwarp();
            n++;
          }
        }
        item++;
      }
    }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 208;");
    // =========================== softmax / epilogue =======================================
    const int quad = warp & 3;
    const int row = quad * 32 + lane;  // query row of the tile == TMEM lane
    const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16);
    const float ks_tensor = kKPerToken ? 1.f : p.kscale[0];

    uint32_t n = 0;
    uint32_t item = 0;
    while (true) {
      mbar_wait(q_full, item & 1);
      const int w = s_work[0];
      const int nact = s_work[1];
      if (w < 0) break;
      Work k;
      decode_work(p, w, k);
      const int hkv = k.hq / p.group;
      const bool row_ok = row < k.rows;
      const float qs = row_ok ? __ldg(p.qscale + (static_cast<long long>(k.b) * p.num_head_q + k.hq) *
                                                      p.qscale_ld + k.mq * kTile + row)
                              : 0.f;
      // > 0 so that a masked score (-inf) stays -inf after scaling (an all-zero q row has qs = 0,
      // and then every raw score is 0 as well)
      const float cq = fmaxf(qs * ks_tensor * p.softmax_scale_log2, 1e-30f);
      // kv positions visible to this row: pos <= row_lim and pos < seq_kv
      const int row_lim = k.seq_kv - k.seq_q + k.mq * kTile + row;
      const int tile_lim_min = k.seq_kv - k.seq_q + k.mq * kTile;  // row 0
      // Online softmax with a LAZY reference maximum `mref` (log2 units): P = 256 * 2^(s - mref).
      // mref only moves when the tile maximum exceeds it by more than 0.75 (P would pass ~430 and
      // approach the e4m3 limit 448) or when the row has not seen a key yet; only then are the
      // row's O values in TMEM rescaled. With causal / block-sparse attention that happens on the
      // first tiles of an item and rarely afterwards.
      float mref = -INFINITY, lrun = 0.f;
      uint8_t* prow = p_smem + row * 128;

      auto softmax_tile = [&](auto mask_tag, const int i, const int key0, const float* ksr) {
        constexpr bool kMask = decltype(mask_tag)::value;
        // ---- S row -> registers; S is then free for QK(n+1) ----
        uint32_t sr[128];
#pragma unroll
        for (int c = 0; c < 4; c++) tmem_ld_x32(lane_addr + c * 32, sr + c * 32);
        tmem_wait_ld();
#pragma unroll
        for (int c = 0; c < 8; c++) tmem_anchor16(sr + c * 16);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(s_free);
        // ---- pass 1: dequantised (and masked) scores in place, row maximum ----
        float vmax[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
        for (int e0 = 0; e0 < 128; e0 += 8) {
          float v[8];
#pragma unroll
          for (int t = 0; t < 8; t += 2) {
            const int e = e0 + t;
            v[t] = __uint_as_float(sr[e]);
            v[t + 1] = __uint_as_float(sr[e + 1]);
            if constexpr (kKPerToken) {
              const float2 ks2 = *reinterpret_cast<const float2*>(ksr + e);
              unpack_f2(fmul2(pack_f2(v[t], v[t + 1]), pack_f2(ks2.x, ks2.y)), v[t], v[t + 1]);
            }
            if constexpr (kMask) {
              const int pos = key0 + e;
              v[t] = (pos > row_lim || pos >= k.seq_kv) ? -INFINITY : v[t];
              v[t + 1] = (pos + 1 > row_lim || pos + 1 >= k.seq_kv) ? -INFINITY : v[t + 1];
            }
            if constexpr (kKPerToken || kMask) {
              sr[e] = __float_as_uint(v[t]);
              sr[e + 1] = __float_as_uint(v[t + 1]);
            }
          }
#pragma unroll
          for (int t = 0; t < 4; t++) vmax[t] = fmaxf(vmax[t], fmaxf(v[t], v[t + 4]));
        }
        float tmax = fmaxf(fmaxf(vmax[0], vmax[1]), fmaxf(vmax[2], vmax[3])) * cq;
        if (!row_ok) tmax = -INFINITY;
        const bool update = (mref == -INFINITY) || (tmax > mref + 0.75f);
        const float mnew = update ? fmaxf(mref, tmax) : mref;
        const bool dead = (mnew == -INFINITY);
        const float alpha =
            (dead || mref == -INFINITY) ? (dead ? 1.f : 0.f) : exp2_approx(mref - mnew);
        mref = mnew;
        if (i > 0) {
          // PV(n-1) complete: the P buffer may be overwritten and O is consistent
          mbar_wait(&v_empty[(n - 1) % kStages], ((n - 1) / kStages) & 1);
          const bool any_scale = __any_sync(0xffffffffu, alpha != 1.f);
          if (any_scale) {
            tc_fence_after();
#pragma unroll 1
            for (int c = 0; c < 8; c++) {
              uint32_t o[16];
              tmem_ld_x16(lane_addr + 128 + c * 16, o);
              tmem_wait_ld();
              tmem_anchor16(o);
#pragma unroll
              for (int e = 0; e < 16; e++) o[e] = __float_as_uint(__uint_as_float(o[e]) * alpha);
              tmem_st_x16(lane_addr + 128 + c * 16, o);
            }
            tmem_wait_st();
          }
        }
        // ---- pass 2: P = 256 * 2^(s - mref) -> e4m3, row sum ----
        const uint64_t bias2 = dead ? pack_f2(-INFINITY, -INFINITY) : pack_f2(8.f - mnew, 8.f - mnew);
        const uint64_t cq2 = pack_f2(cq, cq);
        uint64_t psum2[2] = {0ull, 0ull};
#pragma unroll
        for (int c = 0; c < 8; c++) {
          uint32_t packed[4];
#pragma unroll
          for (int q4 = 0; q4 < 4; q4++) {
            float e4[4];
#pragma unroll
            for (int t = 0; t < 4; t += 2) {
              const int e = c * 16 + q4 * 4 + t;
              const uint64_t x2 =
                  ffma2(pack_f2(__uint_as_float(sr[e]), __uint_as_float(sr[e + 1])), cq2, bias2);
              const bool poly = (t == 2) && (kPoly == 4 || (kPoly == 2 && (q4 & 1)));
              if (poly) {
                exp2_poly_pair(x2, e4[t], e4[t + 1]);
              } else {
                float x0, x1;
                unpack_f2(x2, x0, x1);
                e4[t] = exp2_approx(x0);
                e4[t + 1] = exp2_approx(x1);
              }
              psum2[t >> 1] = fadd2(psum2[t >> 1], pack_f2(e4[t], e4[t + 1]));
            }
            packed[q4] = cvt_e4m3x4(e4[0], e4[1], e4[2], e4[3]);
          }
          // 16 keys = 16-B chunk c of this row, 128B swizzle: chunk ^ (row & 7)
          *reinterpret_cast<uint4*>(prow + ((c ^ (row & 7)) << 4)) =
              make_uint4(packed[0], packed[1], packed[2], packed[3]);
        }
        float s0, s1, s2, s3;
        unpack_f2(psum2[0], s0, s1);
        unpack_f2(psum2[1], s2, s3);
        lrun = lrun * alpha + ((s0 + s1) + (s2 + s3));
      };

      for (int i = 0; i < nact; i++) {
        const int j = list[i];
        mbar_wait(s_full, n & 1);
        tc_fence_after();
        const int key0 = j * kTile;
        const bool need_mask = (key0 + kTile - 1 > tile_lim_min) || (key0 + kTile > k.seq_kv);
        const float* ksr = ks_smem + (n % kKsBufs) * 128;
        if (need_mask) {
          softmax_tile(std::true_type{}, i, key0, ksr);
        } else {
          softmax_tile(std::false_type{}, i, key0, ksr);
        }
        fence_proxy_async_smem();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(p_full);
        n++;
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(q_empty);  // the list / work slot may be refilled by the producer
      // ---- epilogue: O / sum * vscale -> bf16 row ----
      if (nact > 0) {
        mbar_wait(&v_empty[(n - 1) % kStages], ((n - 1) / kStages) & 1);  // last PV of the item
        tc_fence_after();
      }
      {
        const float vs = kKPerToken ? __ldg(p.vscale + hkv) : p.vscale[0];  // (v/256)/(sum/256)
        // a row whose every visible tile was skipped has sum 0 -> NaN, as documented for the
        // reference (hpc/attention.py:274-277)
        const float inv = vs / lrun;
        __nv_bfloat16* dst = p.out + static_cast<long long>(k.q0 + row) * p.ld_out + k.hq * kD;
#pragma unroll 1
        for (int c = 0; c < 8; c++) {
          uint32_t o[16];
          if (nact > 0) {
            tmem_ld_x16(lane_addr + 128 + c * 16, o);
            tmem_wait_ld();
            tmem_anchor16(o);
          } else {
#pragma unroll
            for (int e = 0; e < 16; e++) o[e] = 0u;
          }
          if (row_ok) {
            uint4 w0, w1;
            __nv_bfloat162 b[8];
#pragma unroll
            for (int e = 0; e < 8; e++)
              b[e] = __floats2bfloat162_rn(__uint_as_float(o[2 * e]) * inv,
                                           __uint_as_float(o[2 * e + 1]) * inv);
            w0.x = *reinterpret_cast<uint32_t*>(&b[0]);
            w0.y = *reinterpret_cast<uint32_t*>(&b[1]);
            w0.z = *reinterpret_cast<uint32_t*>(&b[2]);
            w0.w = *reinterpret_cast<uint32_t*>(&b[3]);
            w1.x = *reinterpret_cast<uint32_t*>(&b[4]);
            w1.y = *reinterpret_cast<uint32_t*>(&b[5]);
            w1.z = *reinterpret_cast<uint32_t*>(&b[6]);
            w1.w = *reinterpret_cast<uint32_t*>(&b[7]);
            *reinterpret_cast<uint4*>(dst + c * 16) = w0;
            *reinterpret_cast<uint4*>(dst + c * 16 + 8) = w1;
          }
        }
      }
      tc_fence_before();  // orders the O reads before this thread's next p_full arrive
      item++;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 256);
}

}  // namespace prefill
}  // namespace b200

using namespace b200;  // NOLINT

constexpr int kDefaultPoly = 2;

static int encode_cache_map(CUtensorMap* tm, const void* base, int heads, int num_blocks,
                            int64_t blk_stride, int64_t tok_stride, int64_t head_stride,
                            int* head_first) {
  *head_first = head_stride <= tok_stride ? 1 : 0;
  uint64_t dims[4];
  uint64_t strides[3];
  uint32_t box[4];
  dims[0] = 128;
  box[0] = 128;
  if (*head_first) {
    dims[1] = static_cast<uint64_t>(heads);
    dims[2] = 64;
    strides[0] = static_cast<uint64_t>(head_stride);
    strides[1] = static_cast<uint64_t>(tok_stride);
    box[1] = 1;
    box[2] = 64;
  } else {
    dims[1] = 64;
    dims[2] = static_cast<uint64_t>(heads);
    strides[0] = static_cast<uint64_t>(tok_stride);
    strides[1] = static_cast<uint64_t>(head_stride);
    box[1] = 64;
    box[2] = 1;
  }
  dims[3] = static_cast<uint64_t>(num_blocks);
  strides[2] = static_cast<uint64_t>(blk_stride);
  box[3] = 1;
  const CUtensorMapL2promotion promo = (tok_stride == 128) ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B
                                                           : CU_TENSOR_MAP_L2_PROMOTION_L2_128B;
  return encode_tmap_u8(tm, base, 4, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B, promo);
}

// Common launcher of the two quant schemes.
static int prefill_launch(bool k_per_token, void* y_ptr, const void* q_ptr, const void* kcache_ptr,
                          const void* vcache_ptr, const float* qscale_ptr, const float* kscale_ptr,
                          const float* vscale_ptr, const int* cu_seqlens_q_ptr,
                          const int* block_ids_ptr, const int* seqlens_kv_ptr,
                          const uint8_t* block_mask_ptr, int num_batch, int total_seq_q,
                          int max_seq_q, int num_head_q, int num_head_kv, int num_dim,
                          int num_kvcache_blocks, int block_size, int max_blocks, int qscale_ld,
                          int mask_mq, int mask_kb, int ldY, int ldQ, int64_t k_blk, int64_t k_tok,
                          int64_t k_head, int64_t v_blk, int64_t v_tok, int64_t v_head,
                          int64_t ks_blk, int64_t ks_grp, int64_t ks_head, cudaStream_t stream) {
  HPC_REQUIRE(num_dim == 128, "blocksparse prefill: head dim must be 128");
  HPC_REQUIRE(block_size == 64, "blocksparse prefill: paged block size must be 64 in this build");
  HPC_REQUIRE(num_head_kv > 0 && num_head_q % num_head_kv == 0, "bad head counts");
  HPC_REQUIRE(num_batch > 0 && max_seq_q > 0, "bad batch / max_seqlens_q");
  HPC_REQUIRE((ldQ % 16) == 0 && (reinterpret_cast<uintptr_t>(q_ptr) & 15) == 0, "q alignment");
  if (total_seq_q <= 0) return HPC_OK;

  CUtensorMap tq, tk, tv;
  {
    uint64_t dims[3] = {128, static_cast<uint64_t>(num_head_q), static_cast<uint64_t>(total_seq_q)};
    uint64_t strides[2] = {128, static_cast<uint64_t>(ldQ)};
    uint32_t box[3] = {128, 1, 128};
    int rc = encode_tmap_u8(&tq, q_ptr, 3, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B,
                            CU_TENSOR_MAP_L2_PROMOTION_L2_128B);
    if (rc) return rc;
  }
  int khf = 1, vhf = 1;
  int rc = encode_cache_map(&tk, kcache_ptr, num_head_kv, num_kvcache_blocks, k_blk, k_tok, k_head, &khf);
  if (rc) return rc;
  rc = encode_cache_map(&tv, vcache_ptr, num_head_kv, num_kvcache_blocks, v_blk, v_tok, v_head, &vhf);
  if (rc) return rc;

  int* counter = scheduler_counter(stream);  // work-item counter {next item, CTAs done}, self-resetting
  if (counter == nullptr) return HPC_ERR_CUDA;

  prefill::Params p;
  p.cu_seqlens_q = cu_seqlens_q_ptr;
  p.seqlens_kv = seqlens_kv_ptr;
  p.block_ids = block_ids_ptr;
  p.block_mask = block_mask_ptr;
  p.qscale = qscale_ptr;
  p.kscale = kscale_ptr;
  p.vscale = vscale_ptr;
  p.out = static_cast<__nv_bfloat16*>(y_ptr);
  p.work_counter = counter;
  p.ks_stride_blk = ks_blk;
  p.ks_stride_grp = ks_grp;
  p.ks_stride_head = ks_head;
  p.num_batch = num_batch;
  p.num_head_q = num_head_q;
  p.num_head_kv = num_head_kv;
  p.group = num_head_q / num_head_kv;
  p.max_q_tiles = (max_seq_q + prefill::kTile - 1) / prefill::kTile;
  p.mask_mq = mask_mq;
  p.mask_kb = mask_kb;
  p.max_blocks = max_blocks;
  p.qscale_ld = qscale_ld;
  p.ld_out = ldY;
  p.k_head_first = khf;
  p.v_head_first = vhf;
  p.softmax_scale_log2 = 1.4426950408889634f / sqrtf(128.f);

  const int grid = 2 * sm_count();  // two resident CTAs per SM
  // share of exponentials on the FMA pipe: HPC_B200_PREFILL_POLY = 0 | 2 | 4 (per 8 scores)
  static const int poly = [] {
    const char* e = std::getenv("HPC_B200_PREFILL_POLY");
    const int v = e ? std::atoi(e) : kDefaultPoly;
    return (v == 0 || v == 2 || v == 4) ? v : kDefaultPoly;
  }();
  using KernelFn = void (*)(CUtensorMap, CUtensorMap, CUtensorMap, prefill::Params);
  KernelFn kern;
  int smem_bytes;
  if (k_per_token) {
    smem_bytes = prefill::Smem<true>::kTotal;
    kern = poly == 4   ? prefill::prefill_blocksparse_fp8_kernel<true, 4>
           : poly == 2 ? prefill::prefill_blocksparse_fp8_kernel<true, 2>
                       : prefill::prefill_blocksparse_fp8_kernel<true, 0>;
  } else {
    smem_bytes = prefill::Smem<false>::kTotal;
    kern = poly == 4   ? prefill::prefill_blocksparse_fp8_kernel<false, 4>
           : poly == 2 ? prefill::prefill_blocksparse_fp8_kernel<false, 2>
                       : prefill::prefill_blocksparse_fp8_kernel<false, 0>;
  }
  HPC_CUDA_CHECK(cudaFuncSetAttribute(reinterpret_cast<const void*>(kern),
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
  kern<<<grid, prefill::kThreads, smem_bytes, stream>>>(tq, tk, tv, p);
  HPC_CUDA_CHECK(cudaGetLastError());
  return HPC_OK;
}

#define PREFILL_PARAMS                                                                             \
  void *y_ptr, const void *q_ptr, const void *kcache_ptr, const void *vcache_ptr,                 \
      const float *qscale_ptr, const float *kscale_ptr, const float *vscale_ptr,                  \
      const int *cu_seqlens_q_ptr, const int *block_ids_ptr, const int *seqlens_kv_ptr,           \
      const uint8_t *block_mask_ptr, int num_batch, int total_seq_q, int max_seq_q,               \
      int num_head_q, int num_head_kv, int num_dim, int num_kvcache_blocks, int block_size,       \
      int max_blocks, int qscale_ld, int mask_mq, int mask_kb, int ldY, int ldQ, int64_t k_blk,   \
      int64_t k_tok, int64_t k_head, int64_t v_blk, int64_t v_tok, int64_t v_head
#define PREFILL_ARGS                                                                               \
  y_ptr, q_ptr, kcache_ptr, vcache_ptr, qscale_ptr, kscale_ptr, vscale_ptr, cu_seqlens_q_ptr,     \
      block_ids_ptr, seqlens_kv_ptr, block_mask_ptr, num_batch, total_seq_q, max_seq_q,           \
      num_head_q, num_head_kv, num_dim, num_kvcache_blocks, block_size, max_blocks, qscale_ld,    \
      mask_mq, mask_kb, ldY, ldQ, k_blk, k_tok, k_head, v_blk, v_tok, v_head

// replaces reference src/attention/prefill/prefill.h:46-54
// (attention_with_kvcache_blocksparse_prefill_qpertoken_perhead_kvpertensor_fp8_async)
extern "C" int hpc_attention_blocksparse_prefill_qpertoken_perhead_kvpertensor_fp8_async(
    PREFILL_PARAMS, cudaStream_t stream) {
  return prefill_launch(false, PREFILL_ARGS, 0, 0, 0, stream);
}

// replaces reference src/attention/prefill/prefill.h:55-63
// (attention_with_kvcache_blocksparse_prefill_qkpertoken_perhead_vperhead_fp8_async);
// ks_* = strides in floats of the k-scale tensor [blocks, block/32, Hkv, 32]
extern "C" int hpc_attention_blocksparse_prefill_qkpertoken_perhead_vperhead_fp8_async(
    PREFILL_PARAMS, int64_t ks_blk, int64_t ks_grp, int64_t ks_head, cudaStream_t stream) {
  HPC_REQUIRE(ks_grp % 4 == 0 && ks_head % 4 == 0 && ks_blk % 4 == 0,
              "k scale strides must be multiples of 4 floats (16-byte bulk copies)");
  return prefill_launch(true, PREFILL_ARGS, ks_blk, ks_grp, ks_head, stream);
}
