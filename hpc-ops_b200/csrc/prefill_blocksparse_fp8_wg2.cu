// Experimental variant of the FP8 block-sparse prefill kernel (prefill_blocksparse_fp8.cu) with TWO
// softmax warpgroups per CTA (HPC_B200_PREFILL_WG2=1; off by default, NOT yet validated on
// hardware). The one-warpgroup kernel is bound by the issue latency of its softmax: with two CTAs
// per SM each scheduler hosts only two softmax warps (profiles/r1_prefill_v5_final.md: issue slots
// 53 %, no pipe saturated). Here every query row is shared by two threads -- warps 4-7 own key
// columns 0..63 of each tile, warps 8-11 columns 64..127 -- so four softmax warps per scheduler hide
// each other's latency. Per tile the two threads of a row exchange their partial row maximum
// through shared memory (one named barrier); row sums are combined once per item; each thread
// rescales / writes its 64 of the 128 O columns. Producer and MMA roles, the mbarrier protocol and
// the numerics are those of the one-warpgroup kernel (arrival counts 256 instead of 128).
#include <cstdlib>
#include <type_traits>

#include "common.cuh"
#include "host_utils.h"
#include "prefill_common.cuh"

namespace b200 {
namespace prefill {

constexpr int kThreads2 = 384;
constexpr int kXchBar = 2;  // named barrier of the 256 softmax threads

// Two CTAs are resident per SM (each 384 threads, ~106 KB smem, 256 TMEM columns).
template <bool kKPerToken>
struct Smem2 {
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
  // exchange between the two threads of a row: [2 tile-parity slots + 1 row-sum slot][2 halves][128 rows]
  static constexpr int kOffXch = kOffTmem + 64;
  static constexpr int kTotal = kOffXch + 3 * 2 * 128 * 4;
};

// Protocol (all mbarriers; "phase" = use-count parity), n = running KV-tile counter of the CTA:
//   q_full      producer: item published (list + work id in smem, Q tile landed)
//   q_empty     MMA thread (commit after the item's last QK) + 256 softmax threads (item finished)
//   k_full/k_empty[slot]   K ring; the slot is released by the commit after QK(n)
//   v_full/v_empty[slot]   V ring; released by the commit after PV(n). v_empty also tells the
//                          softmax threads that PV(n) is complete (P buffer free, O consistent)
//   s_full      commit after QK(n)                            -> softmax
//   s_free      256 softmax threads hold S(n) in registers     -> MMA thread may issue QK(n+1)
//   p_full      256 softmax threads wrote P(n) (and rescaled O) -> MMA thread issues PV(n)
// The k-scale buffer of tile n+2 was last read by softmax(n-2), which every softmax thread
// finished before arriving on s_free(n-1), which precedes QK(n) and so the release of K slot n%2.
template <bool kKPerToken>
__global__ void __launch_bounds__(kThreads2, 2)
    prefill_blocksparse_fp8_wg2_kernel(const __grid_constant__ CUtensorMap tmap_q,
                                       const __grid_constant__ CUtensorMap tmap_k,
                                       const __grid_constant__ CUtensorMap tmap_v, const Params p) {
  using L = Smem2<kKPerToken>;
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
    mbar_init(q_empty, 1 + 256);
    mbar_init(s_full, 1);
    mbar_init(s_free, 256);
    mbar_init(p_full, 256);
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
    asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
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
        if (lane == 0) {
          s_work[0] = alive ? w : -1;
          s_work[1] = nact;
          if (alive) {
            mbar_arrive_expect_tx(q_full, kTileBytes);
            tma_load_3d(q_smem, &tmap_q, q_full, 0, k.hq, k.q0);
          } else {
            mbar_arrive(q_full);
          }
        }
        __syncwarp();
        item++;
        if (!alive) break;
        // ---- K/V tiles of the active list ----
        const int hkv = k.hq / p.group;
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
            if (lane == 0) {
              const uint32_t st = n % kStages;
              const uint32_t ph = ((n / kStages) & 1) ^ 1;
              mbar_wait(&k_empty[st], ph);  // QK(n - 2) finished
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
              mbar_wait(&v_empty[st], ph);  // PV(n - 2) finished
              uint8_t* vd8 = v_smem + st * kTileBytes;
              mbar_arrive_expect_tx(&v_full[st], kTileBytes);
              tma_load_4d_hint(vd8, &tmap_v, &v_full[st], 0, vc1, vc2, id0, pol_kv);
              tma_load_4d_hint(vd8 + kTileBytes / 2, &tmap_v, &v_full[st], 0, vc1, vc2, id1, pol_kv);
            }
            n++;
          }
        }
      }
    } else if (warp == 1 && lane == 0) {
      // =========================== tcgen05 issuer (one thread) ============================
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
#pragma unroll
        for (int k = 0; k < 4; k++) umma_f8(tmem_base, qdesc + k * 2, kd + k * 2, idesc_qk, k > 0);
        umma_commit(s_full);
        umma_commit(&k_empty[st]);
        if (last_of_item) umma_commit(q_empty);  // Q tile / list reusable after this QK
      };
      uint32_t n = 0;
      uint32_t item = 0;
      while (true) {
        mbar_wait(q_full, item & 1);
        const int w = s_work[0];
        const int nact = s_work[1];
        if (w < 0) break;
        if (nact == 0) {
          umma_commit(q_empty);
        } else {
          issue_qk(n, nact == 1);
          for (int i = 0; i < nact; i++) {
            if (i + 1 < nact) issue_qk(n + 1, i + 2 == nact);  // runs under softmax(n)
            const uint32_t st = n % kStages;
            mbar_wait(&v_full[st], (n / kStages) & 1);
            mbar_wait(p_full, n & 1);
            tc_fence_after();
            const uint64_t vd = vdesc0 + static_cast<uint64_t>(st * (kTileBytes >> 4));
#pragma unroll
            for (int k = 0; k < 4; k++) {
              // A = P (K-major, 32 B per MMA); B = V as stored: MN-major, 32 keys = 4096 B per MMA
              umma_f8(tmem_base + 128, pdesc + k * 2, vd + k * (4096 >> 4), idesc_pv,
                      (k > 0) || (i > 0));
            }
            umma_commit(&v_empty[st]);
            n++;
          }
        }
        item++;
      }
    }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 104;");
    // =========================== softmax / epilogue (two warpgroups) =======================
    // Warps 4-7 take key columns 0..63 of every tile, warps 8-11 columns 64..127; the two threads of
    // a row exchange their partial row maximum (per tile) and row sum (per item) through smem.
    const int half = (warp - 4) >> 2;
    const int quad = warp & 3;
    const int row = quad * 32 + lane;  // query row of the tile == TMEM lane
    const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16);
    const float ks_tensor = kKPerToken ? 1.f : p.kscale[0];
    float* xch = reinterpret_cast<float*>(smem + L::kOffXch);

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
      const float cq = fmaxf(qs * ks_tensor * p.softmax_scale_log2, 1e-30f);
      const int row_lim = k.seq_kv - k.seq_q + k.mq * kTile + row;
      const int tile_lim_min = k.seq_kv - k.seq_q + k.mq * kTile;  // row 0
      // lazy reference maximum, as in the one-warpgroup kernel; both threads of a row take the same
      // decisions because they see the same exchanged maximum
      float mref = -INFINITY, lrun = 0.f;  // lrun: sum over THIS thread's columns only
      uint8_t* prow = p_smem + row * 128;

      auto softmax_tile = [&](auto mask_tag, const int i, const int key0, const float* ksr) {
        constexpr bool kMask = decltype(mask_tag)::value;
        uint32_t sr[64];
        tmem_ld_x32(lane_addr + half * 64, sr);
        tmem_ld_x32(lane_addr + half * 64 + 32, sr + 32);
        tmem_wait_ld();
#pragma unroll
        for (int c = 0; c < 4; c++) tmem_anchor16(sr + c * 16);
        tc_fence_before();
        mbar_arrive(s_free);
        // ---- pass 1 over this thread's 64 columns ----
        float vmax[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
        for (int e0 = 0; e0 < 64; e0 += 8) {
          float v[8];
#pragma unroll
          for (int t = 0; t < 8; t += 2) {
            const int e = e0 + t;
            v[t] = __uint_as_float(sr[e]);
            v[t + 1] = __uint_as_float(sr[e + 1]);
            if constexpr (kKPerToken) {
              const float2 ks2 = *reinterpret_cast<const float2*>(ksr + half * 64 + e);
              unpack_f2(fmul2(pack_f2(v[t], v[t + 1]), pack_f2(ks2.x, ks2.y)), v[t], v[t + 1]);
            }
            if constexpr (kMask) {
              const int pos = key0 + half * 64 + e;
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
        float hmax = fmaxf(fmaxf(vmax[0], vmax[1]), fmaxf(vmax[2], vmax[3]));
        // ---- row maximum = max of the two halves (double-buffered exchange, one barrier) ----
        float* xb = xch + (n & 1) * 256;
        xb[half * 128 + row] = hmax;
        named_bar_sync(kXchBar, 256);
        float tmax = fmaxf(hmax, xb[(half ^ 1) * 128 + row]) * cq;
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
            for (int c = 0; c < 4; c++) {  // this thread's 64 of the 128 O columns
              uint32_t o[16];
              tmem_ld_x16(lane_addr + 128 + half * 64 + c * 16, o);
              tmem_wait_ld();
              tmem_anchor16(o);
#pragma unroll
              for (int e = 0; e < 16; e++) o[e] = __float_as_uint(__uint_as_float(o[e]) * alpha);
              tmem_st_x16(lane_addr + 128 + half * 64 + c * 16, o);
            }
            tmem_wait_st();
          }
        }
        // ---- pass 2: P = 256 * 2^(s - mref) -> e4m3 (this thread's four 16-byte chunks) ----
        const uint64_t bias2 = dead ? pack_f2(-INFINITY, -INFINITY) : pack_f2(8.f - mnew, 8.f - mnew);
        const uint64_t cq2 = pack_f2(cq, cq);
        uint64_t psum2[2] = {0ull, 0ull};
#pragma unroll
        for (int c = 0; c < 4; c++) {
          uint32_t packed[4];
#pragma unroll
          for (int q4 = 0; q4 < 4; q4++) {
            float e4[4];
#pragma unroll
            for (int t = 0; t < 4; t += 2) {
              const int e = c * 16 + q4 * 4 + t;
              float x0, x1;
              unpack_f2(ffma2(pack_f2(__uint_as_float(sr[e]), __uint_as_float(sr[e + 1])), cq2, bias2),
                        x0, x1);
              e4[t] = exp2_approx(x0);
              e4[t + 1] = exp2_approx(x1);
              psum2[t >> 1] = fadd2(psum2[t >> 1], pack_f2(e4[t], e4[t + 1]));
            }
            packed[q4] = cvt_e4m3x4(e4[0], e4[1], e4[2], e4[3]);
          }
          // 16 keys = 16-B chunk (half * 4 + c) of this row, 128B swizzle: chunk ^ (row & 7)
          *reinterpret_cast<uint4*>(prow + (((half * 4 + c) ^ (row & 7)) << 4)) =
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
        mbar_arrive(p_full);
        n++;
      }
      mbar_arrive(q_empty);  // the list / work slot may be refilled by the producer
      // ---- row sum = sum of the two halves. Its own slot (the tile-parity slots may already be
      // rewritten by a thread that has moved on to the next item); the first barrier makes sure
      // every thread has read the previous item's sums before they are overwritten ----
      float* lb = xch + 2 * 256;
      named_bar_sync(kXchBar, 256);
      lb[half * 128 + row] = lrun;
      named_bar_sync(kXchBar, 256);
      const float lsum = lrun + lb[(half ^ 1) * 128 + row];
      // ---- epilogue: O / sum * vscale -> bf16, this thread's 64 of the 128 output columns ----
      if (nact > 0) {
        mbar_wait(&v_empty[(n - 1) % kStages], ((n - 1) / kStages) & 1);  // last PV of the item
        tc_fence_after();
      }
      {
        const float vs = kKPerToken ? __ldg(p.vscale + hkv) : p.vscale[0];
        const float inv = vs / lsum;  // all-skipped row: 0 / 0 -> NaN, as documented for the reference
        __nv_bfloat16* dst =
            p.out + static_cast<long long>(k.q0 + row) * p.ld_out + k.hq * kD + half * 64;
#pragma unroll 1
        for (int c = 0; c < 4; c++) {
          uint32_t o[16];
          if (nact > 0) {
            tmem_ld_x16(lane_addr + 128 + half * 64 + c * 16, o);
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

// Launcher used by prefill_launch() when HPC_B200_PREFILL_WG2=1 (tensor maps and Params are built
// there; `p.work_counter` is already set).
int prefill_wg2_launch(bool k_per_token, const CUtensorMap& tq, const CUtensorMap& tk,
                       const CUtensorMap& tv, const prefill::Params& p, cudaStream_t stream) {
  const int grid = 2 * sm_count();
  if (k_per_token) {
    auto kern = prefill::prefill_blocksparse_fp8_wg2_kernel<true>;
    HPC_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        prefill::Smem2<true>::kTotal));
    kern<<<grid, prefill::kThreads2, prefill::Smem2<true>::kTotal, stream>>>(tq, tk, tv, p);
  } else {
    auto kern = prefill::prefill_blocksparse_fp8_wg2_kernel<false>;
    HPC_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        prefill::Smem2<false>::kTotal));
    kern<<<grid, prefill::kThreads2, prefill::Smem2<false>::kTotal, stream>>>(tq, tk, tv, p);
  }
  HPC_CUDA_CHECK(cudaGetLastError());
  return HPC_OK;
}
