// Dynamic split-k task map for paged decode attention (B200 build).
//
// Replaces (same bytes out, different algorithm):
//   reference src/attention/decode/assign_task.cu:41-329  (CUDA: one thread per CTA replays a
//                                                          serial greedy walk + cross-CTA spin flags)
//   reference src/attention/decode/assign_task.cu:362-492 (CPU serial walk)
//   reference src/attention/entry.cc:727-778              (CPU packing of the host task map)
//
// B200 formulation: lay all KV tiles of all (kv-head, batch) pairs on one line (head outer, batch
// inner). Bin `icta` owns the tile interval [icta*P, (icta+1)*P). A task is the intersection of a
// pair's interval with a bin's interval, so every 48-byte row of the map is a pure function of
// (icta, slot) given the per-batch exclusive prefix sums -- no serial replay, no inter-CTA flags.
// One thread block per bin recomputes the (<= 2048 element) scan in shared memory and then its
// threads emit that bin's rows independently.
#include <algorithm>
#include <cstring>
#include <vector>

#include "common.cuh"
#include "host_utils.h"

namespace b200 {
namespace taskmap {

constexpr int kTaskStride = 12;  // ints per row (48 B), reference sched_task_info.h:17-33
constexpr int kMaxNumBatch = 2048;

struct Row {
  int v[kTaskStride];
};

struct Geometry {
  int num_batch;
  int num_head_kv;
  int num_seq_q;
  int tilen;
  int num_total_ctas;
  int tiles_per_head;    // TB
  int nonzero_per_head;  // NZB
  int num_tile_per_cta;  // P
};

// Pair (rank-th non-empty pair in walk order) -> global start tile, tile count, batch.
struct PairInfo {
  long long start;  // global tile index of the pair's first tile
  int ntile;
  int batch;
  int head;
  int nseq;
};

__host__ __device__ inline PairInfo pair_of_rank(const Geometry& g, long long rank, const int* nzb,
                                                 const int* tile_prefix, const int* ntile,
                                                 const int* nseq) {
  PairInfo p;
  int h = static_cast<int>(rank / g.nonzero_per_head);
  int j = static_cast<int>(rank % g.nonzero_per_head);
  int b = nzb[j];
  p.head = h;
  p.batch = b;
  p.ntile = ntile[b];
  p.nseq = nseq[b];
  p.start = static_cast<long long>(h) * g.tiles_per_head + tile_prefix[b];
  return p;
}

// rank of the non-empty pair covering global tile x (0 <= x < H*TB)
__host__ __device__ inline long long cover_rank(const Geometry& g, long long x, const int* nzb,
                                                const int* tile_prefix) {
  int h = static_cast<int>(x / g.tiles_per_head);
  int xr = static_cast<int>(x % g.tiles_per_head);
  int lo = 0, hi = g.nonzero_per_head - 1;  // largest j with tile_prefix[nzb[j]] <= xr
  while (lo < hi) {
    int mid = (lo + hi + 1) >> 1;
    if (tile_prefix[nzb[mid]] <= xr) {
      lo = mid;
    } else {
      hi = mid - 1;
    }
  }
  return static_cast<long long>(h) * g.nonzero_per_head + lo;
}

__host__ __device__ inline int chunk_seqkv(const Geometry& g, const PairInfo& p,
                                           long long start_tiles, long long add_tiles) {
  long long a = add_tiles * g.tilen;
  long long b = static_cast<long long>(p.nseq) - start_tiles * g.tilen;
  return static_cast<int>(a < b ? a : b);
}

// Row `slot` of bin `icta`. Returns false (and a terminator row) when the bin has no such task.
__host__ __device__ inline bool make_row(const Geometry& g, int icta, int slot, const int* nzb,
                                         const int* tile_prefix, const int* ntile, const int* nseq,
                                         Row* out) {
#pragma unroll
  for (int i = 0; i < kTaskStride; i++) out->v[i] = 0;
  out->v[0] = -1;
  out->v[1] = -1;

  const long long P = g.num_tile_per_cta;
  const long long total = static_cast<long long>(g.num_head_kv) * g.tiles_per_head;
  const long long lo = static_cast<long long>(icta) * P;
  if (lo >= total || g.nonzero_per_head == 0) return false;
  const long long hi = (lo + P < total) ? lo + P : total;

  const long long total_pairs = static_cast<long long>(g.num_head_kv) * g.nonzero_per_head;
  const long long rank = cover_rank(g, lo, nzb, tile_prefix) + slot;
  if (rank >= total_pairs) return false;
  PairInfo p = pair_of_rank(g, rank, nzb, tile_prefix, ntile, nseq);
  if (p.start >= hi) return false;

  const long long beg = p.start > lo ? p.start : lo;
  const long long pend = p.start + p.ntile;
  const long long end = pend < hi ? pend : hi;
  const long long start_tiles = beg - p.start;
  const long long add_tiles = end - beg;

  int num_seqkv = chunk_seqkv(g, p, start_tiles, add_tiles);
  int num_seqkvcache = num_seqkv;
  int is_causal = 0;
  int num_tile_full = num_seqkvcache / g.tilen;

  if (end == pend) {
    // last chunk of the pair: the num_seq_q new tokens attend causally
    is_causal = 1;
    num_seqkvcache -= g.num_seq_q;
    int q = num_seqkvcache / g.tilen;
    num_tile_full = q > 0 ? q : 0;
  } else {
    // if the pair's next (and last) chunk is shorter than num_seq_q, the causal window spills
    // back into this chunk (reference assign_task.cu:239-252,286-300)
    long long nstart = end - p.start;
    long long nadd = pend - end;
    if (nadd <= P) {
      int nseqkv = chunk_seqkv(g, p, nstart, nadd);
      int overflow = nseqkv - g.num_seq_q;
      if (overflow < 0) {
        is_causal = 1;
        num_seqkvcache += overflow;
        int q = num_seqkvcache / g.tilen;
        num_tile_full = q > 0 ? q : 0;
      }
    }
  }

  out->v[0] = p.head;
  out->v[1] = p.batch;
  out->v[2] = icta - static_cast<int>(p.start / P);  // ichunk
  out->v[3] = static_cast<int>(start_tiles) * g.tilen;  // iseq_start
  out->v[4] = num_seqkv;
  out->v[5] = num_seqkvcache;
  out->v[6] = (num_seqkv + g.tilen - 1) / g.tilen;  // num_tile_kv
  out->v[7] = num_tile_full;
  out->v[8] = is_causal;
  return true;
}

__host__ __device__ inline int pair_num_chunks(const Geometry& g, int h, int b,
                                               const int* tile_prefix, const int* ntile) {
  if (ntile[b] <= 0) return 0;
  long long P = g.num_tile_per_cta;
  long long s = static_cast<long long>(h) * g.tiles_per_head + tile_prefix[b];
  long long e = s + ntile[b] - 1;
  return static_cast<int>(e / P - s / P + 1);
}

// ----------------------------------------------------------------------------------------------
// CUDA: grid = num_total_ctas bins, 256 threads each.
// ----------------------------------------------------------------------------------------------
constexpr int kThreads = 256;
constexpr int kItems = kMaxNumBatch / kThreads;  // 8 batches per thread

__global__ void __launch_bounds__(kThreads)
    assign_task_kernel(int* __restrict__ task_map, const int* __restrict__ num_seq_kvcache,
                       int num_batch, int num_head_kv, int num_seq_q, int new_kv_included,
                       int min_process_len, int num_total_ctas, int tilen) {
  __shared__ int s_nseq[kMaxNumBatch];
  __shared__ int s_ntile[kMaxNumBatch];
  __shared__ int s_prefix[kMaxNumBatch];
  __shared__ int s_nzb[kMaxNumBatch];
  __shared__ int s_warp_tiles[kThreads / 32];
  __shared__ int s_warp_nz[kThreads / 32];
  __shared__ int s_red[kThreads / 32];

  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int warp = tid >> 5;
  const int icta = blockIdx.x;

  // PDL: the previous kernel in the stream (e.g. the last step's combine) may still be reading the
  // task map this kernel rewrites; the next kernel (attention) may set itself up meanwhile
  pdl_wait();
  pdl_launch_dependents();

  // ---- per-batch tile counts + block-wide exclusive scan (tiles and non-empty count) ----
  int loc_tiles[kItems];
  int loc_nz[kItems];
  int sum_tiles = 0, sum_nz = 0;
#pragma unroll
  for (int i = 0; i < kItems; i++) {
    int b = tid * kItems + i;
    int ns = 0;
    if (b < num_batch) {
      ns = num_seq_kvcache[b] + (new_kv_included ? 0 : num_seq_q);
    }
    int nt = (ns + tilen - 1) / tilen;
    if (b < num_batch) {
      s_nseq[b] = ns;
      s_ntile[b] = nt;
    } else {
      nt = 0;
    }
    loc_tiles[i] = sum_tiles;
    loc_nz[i] = sum_nz;
    sum_tiles += nt;
    sum_nz += (nt > 0) ? 1 : 0;
  }
  int inc_tiles = sum_tiles, inc_nz = sum_nz;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int t = __shfl_up_sync(0xffffffffu, inc_tiles, o);
    int z = __shfl_up_sync(0xffffffffu, inc_nz, o);
    if (lane >= o) {
      inc_tiles += t;
      inc_nz += z;
    }
  }
  if (lane == 31) {
    s_warp_tiles[warp] = inc_tiles;
    s_warp_nz[warp] = inc_nz;
  }
  __syncthreads();
  int base_tiles = 0, base_nz = 0, tot_tiles = 0, tot_nz = 0;
#pragma unroll
  for (int w = 0; w < kThreads / 32; w++) {
    if (w < warp) {
      base_tiles += s_warp_tiles[w];
      base_nz += s_warp_nz[w];
    }
    tot_tiles += s_warp_tiles[w];
    tot_nz += s_warp_nz[w];
  }
  base_tiles += inc_tiles - sum_tiles;
  base_nz += inc_nz - sum_nz;
#pragma unroll
  for (int i = 0; i < kItems; i++) {
    int b = tid * kItems + i;
    if (b < num_batch) {
      s_prefix[b] = base_tiles + loc_tiles[i];
      if (s_ntile[b] > 0) s_nzb[base_nz + loc_nz[i]] = b;
    }
  }
  __syncthreads();

  Geometry g;
  g.num_batch = num_batch;
  g.num_head_kv = num_head_kv;
  g.num_seq_q = num_seq_q;
  g.tilen = tilen;
  g.num_total_ctas = num_total_ctas;
  g.tiles_per_head = tot_tiles;
  g.nonzero_per_head = tot_nz;
  long long total = static_cast<long long>(tot_tiles) * num_head_kv;
  int per_cta = static_cast<int>((total + num_total_ctas - 1) / num_total_ctas);
  int floor_tiles = min_process_len / tilen;
  g.num_tile_per_cta = per_cta > floor_tiles ? per_cta : floor_tiles;
  const int P = g.num_tile_per_cta;

  const int max_num_batch = task_map[3];
  const int chunk_entries = max_num_batch * num_head_kv;
  const int chunk_pad = (chunk_entries + kTaskStride - 1) / kTaskStride * kTaskStride;
  const int cta_pad = (num_total_ctas + kTaskStride - 1) / kTaskStride * kTaskStride;
  int* chunk_table = task_map + kTaskStride * ((P + 1) * num_total_ctas + 1);
  int* finish_flags = chunk_table + chunk_pad;
  int* num_task_table = finish_flags + cta_pad;

  // ---- this bin's rows (tasks, then terminators up to and including slot P) ----
  int* bin = task_map + (1 + icta * (P + 1)) * kTaskStride;
  int my_tasks = 0;
  for (int slot = tid; slot <= P; slot += kThreads) {
    Row r;
    bool valid = (P > 0) && (slot < P) &&
                 make_row(g, icta, slot, s_nzb, s_prefix, s_ntile, s_nseq, &r);
    if (!valid) {
#pragma unroll
      for (int i = 0; i < kTaskStride; i++) r.v[i] = 0;
      r.v[0] = -1;
      r.v[1] = -1;
    } else {
      my_tasks++;
    }
    int4* dst = reinterpret_cast<int4*>(bin + slot * kTaskStride);
    dst[0] = make_int4(r.v[0], r.v[1], r.v[2], r.v[3]);
    dst[1] = make_int4(r.v[4], r.v[5], r.v[6], r.v[7]);
    dst[2] = make_int4(r.v[8], r.v[9], r.v[10], r.v[11]);
  }
  // block-reduce the task count of this bin
  my_tasks = __reduce_add_sync(0xffffffffu, my_tasks);
  if (lane == 0) s_red[warp] = my_tasks;
  __syncthreads();
  if (tid == 0) {
    int n = 0;
#pragma unroll
    for (int w = 0; w < kThreads / 32; w++) n += s_red[w];
    num_task_table[icta] = n;
    finish_flags[icta] = 0;
  }

  // ---- chunk table: entries are strided over the whole grid; bin 0 also writes the header ----
  for (int e = icta * kThreads + tid; e < chunk_entries; e += num_total_ctas * kThreads) {
    int v = 0;
    if (e < num_batch * num_head_kv) {
      v = pair_num_chunks(g, e / num_batch, e % num_batch, s_prefix, s_ntile);
    }
    chunk_table[e] = v;
  }
  if (icta == 0) {
    int mx = 0;
    for (int e = tid; e < num_batch * num_head_kv; e += kThreads) {
      int v = pair_num_chunks(g, e / num_batch, e % num_batch, s_prefix, s_ntile);
      mx = v > mx ? v : mx;
    }
    mx = __reduce_max_sync(0xffffffffu, mx);
    __syncthreads();
    if (lane == 0) s_red[warp] = mx;
    __syncthreads();
    if (tid == 0) {
      int m = 0;
#pragma unroll
      for (int w = 0; w < kThreads / 32; w++) m = s_red[w] > m ? s_red[w] : m;
      task_map[0] = P + 1;
      task_map[1] = num_total_ctas;
      task_map[5] = m;
      // tiles per kv head on the task line (rotated bin walk, decode_common.cuh). A pad word of the
      // reference's header; the packed host map of the CPU scheduler leaves it zero like the
      // reference, hpc/attention.py writes it when it splices a host map into the workspace.
      task_map[6] = tot_tiles;
    }
  }
}

// ----------------------------------------------------------------------------------------------
// Host (CPU) path: same row function, packed like the reference CPU entry (entry.cc:750-776):
//   row 0 header {P+1, ctas, 0, 0, 0, max_chunks}, then ctas*(P+1) task rows, then num_chunks
//   with stride num_batch padded to whole rows.
// ----------------------------------------------------------------------------------------------
struct HostPlan {
  Geometry g;
  std::vector<int> nseq, ntile, prefix, nzb;
};

static HostPlan make_plan(const int* lens, int num_total_ctas, int num_batch, int num_head_kv,
                          int num_seq_q, int tilen, bool new_kv_included, int min_process_len) {
  HostPlan p;
  p.nseq.resize(num_batch);
  p.ntile.resize(num_batch);
  p.prefix.resize(num_batch);
  int tiles = 0;
  for (int b = 0; b < num_batch; b++) {
    int ns = new_kv_included ? lens[b] : lens[b] + num_seq_q;
    int nt = (ns + tilen - 1) / tilen;
    p.nseq[b] = ns;
    p.ntile[b] = nt;
    p.prefix[b] = tiles;
    tiles += nt;
    if (nt > 0) p.nzb.push_back(b);
  }
  p.g.num_batch = num_batch;
  p.g.num_head_kv = num_head_kv;
  p.g.num_seq_q = num_seq_q;
  p.g.tilen = tilen;
  p.g.num_total_ctas = num_total_ctas;
  p.g.tiles_per_head = tiles;
  p.g.nonzero_per_head = static_cast<int>(p.nzb.size());
  long long total = static_cast<long long>(tiles) * num_head_kv;
  int per_cta = static_cast<int>((total + num_total_ctas - 1) / num_total_ctas);
  p.g.num_tile_per_cta = std::max(per_cta, min_process_len / tilen);
  if (p.nzb.empty()) p.nzb.push_back(0);
  return p;
}

}  // namespace taskmap
}  // namespace b200

using namespace b200;           // NOLINT
using namespace b200::taskmap;  // NOLINT

extern "C" int64_t hpc_assign_attention_decode_task_host_bytes(
    const int* num_seq_kvcache, int num_total_ctas, int num_batch, int num_head_kv, int num_seq_q,
    int tilen, int new_kv_included, int min_process_len) {
  if (num_batch <= 0 || num_total_ctas <= 0 || tilen <= 0 || num_head_kv <= 0) return -1;
  HostPlan p = make_plan(num_seq_kvcache, num_total_ctas, num_batch, num_head_kv, num_seq_q, tilen,
                         new_kv_included != 0, min_process_len);
  int64_t num_task = static_cast<int64_t>(num_total_ctas) * (p.g.num_tile_per_cta + 1);
  int64_t chunk_bytes = static_cast<int64_t>(num_head_kv) * num_batch * 4;
  int64_t rows = 1 + num_task + (chunk_bytes + 47) / 48;
  return rows * 48;
}

extern "C" int hpc_assign_attention_decode_task_sync(const int* num_seq_kvcache,
                                                     int num_total_ctas, int num_batch,
                                                     int num_head_kv, int num_seq_q, int tilen,
                                                     int new_kv_included, int min_process_len,
                                                     void* task_map_host, int64_t task_map_bytes) {
  HPC_REQUIRE(num_batch > 0 && num_total_ctas > 0 && num_head_kv > 0, "bad task-map geometry");
  HPC_REQUIRE(tilen == 64 || tilen == 128, "tilen must be 64 or 128, got %d", tilen);
  int64_t need = hpc_assign_attention_decode_task_host_bytes(
      num_seq_kvcache, num_total_ctas, num_batch, num_head_kv, num_seq_q, tilen, new_kv_included,
      min_process_len);
  HPC_REQUIRE(task_map_bytes >= need, "host task map too small: %lld < %lld",
              (long long)task_map_bytes, (long long)need);
  HostPlan p = make_plan(num_seq_kvcache, num_total_ctas, num_batch, num_head_kv, num_seq_q, tilen,
                         new_kv_included != 0, min_process_len);
  const int P = p.g.num_tile_per_cta;
  int* out = static_cast<int*>(task_map_host);
  std::memset(out, 0, static_cast<size_t>(need));
  for (int icta = 0; icta < num_total_ctas; icta++) {
    int* bin = out + (1 + static_cast<int64_t>(icta) * (P + 1)) * kTaskStride;
    for (int slot = 0; slot <= P; slot++) {
      Row r;
      bool valid = (slot < P) && make_row(p.g, icta, slot, p.nzb.data(), p.prefix.data(),
                                          p.ntile.data(), p.nseq.data(), &r);
      if (!valid) {
        std::memset(r.v, 0, sizeof(r.v));
        r.v[0] = -1;
        r.v[1] = -1;
      }
      std::memcpy(bin + slot * kTaskStride, r.v, sizeof(r.v));
    }
  }
  int* chunk_table = out + kTaskStride * (static_cast<int64_t>(P + 1) * num_total_ctas + 1);
  int mx = 0;
  for (int h = 0; h < num_head_kv; h++) {
    for (int b = 0; b < num_batch; b++) {
      int v = pair_num_chunks(p.g, h, b, p.prefix.data(), p.ntile.data());
      chunk_table[h * num_batch + b] = v;
      mx = std::max(mx, v);
    }
  }
  out[0] = P + 1;
  out[1] = num_total_ctas;
  out[5] = mx;
  return HPC_OK;
}

extern "C" int hpc_assign_attention_decode_task_async(int* task_map, const int* num_seq_kvcache,
                                                      int num_total_ctas, int num_batch,
                                                      int num_head_kv, int num_seq_q, int tilen,
                                                      int new_kv_included, int min_process_len,
                                                      cudaStream_t stream) {
  HPC_REQUIRE(num_batch > 0 && num_batch <= kMaxNumBatch,
              "assign_attention_decode_task: batch %d outside (0, %d]", num_batch, kMaxNumBatch);
  HPC_REQUIRE(tilen == 64 || tilen == 128, "tilen must be 64 or 128, got %d", tilen);
  HPC_REQUIRE(num_total_ctas > 0 && num_head_kv > 0, "bad task-map geometry");
  HPC_CUDA_CHECK(launch_pdl(assign_task_kernel, dim3(num_total_ctas), dim3(kThreads), 0, stream, 1,
                            task_map, num_seq_kvcache, num_batch, num_head_kv, num_seq_q,
                            new_kv_included, min_process_len, num_total_ctas, tilen));
  return HPC_OK;
}
