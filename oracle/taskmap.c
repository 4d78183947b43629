/*
 * ORACLE — test infrastructure only. Not part of the product; only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load this.
 *
 * Plain-C restatement of the reference's CPU decode-attention task scheduler:
 *   /root/reference/src/attention/decode/assign_task.cu:362-492  (assign_attention_decode_task_sync:
 *       serial greedy walk, head outer / batch inner, bins of num_tile_per_cta tiles)
 *   /root/reference/src/attention/entry.cc:727-778               (packing into the host task map)
 *   /root/reference/src/attention/decode/sched_task_info.h:17-33 (48-byte TaskScheduleInfo row)
 *
 * It deliberately keeps the reference's serial control flow (per-pair progress arrays, bucket
 * loop, last_cta/last_task back-patching) so it is an independent check of the product's
 * closed-form interval formulation in hpc-ops_b200/csrc/decode_taskmap.cu.
 *
 * Pinned against the real reference code compiled from /root/reference (oracle/_ref, see Makefile)
 * and against tests/golden/taskmap_*.npz generated from it.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ROW 12 /* ints per TaskScheduleInfo */

enum { F_HEAD = 0, F_BATCH, F_CHUNK, F_SEQ_START, F_NUM_SEQKV, F_NUM_SEQKVCACHE, F_NUM_TILE_KV,
       F_NUM_TILE_FULL, F_IS_CAUSAL };

static int imax(int a, int b) { return a > b ? a : b; }
static int imin(int a, int b) { return a < b ? a : b; }

/* number of bytes of the packed host task map (entry.cc:756-758) */
int64_t oracle_taskmap_bytes(const int* num_seq_kvcache, int num_total_ctas, int num_batch,
                             int num_head_kv, int num_seq_q, int tilen, int new_kv_included,
                             int min_process_len) {
  int64_t total_tiles_per_head = 0;
  for (int b = 0; b < num_batch; b++) {
    int n = new_kv_included ? num_seq_kvcache[b] : num_seq_kvcache[b] + num_seq_q;
    total_tiles_per_head += (n + tilen - 1) / tilen;
  }
  int64_t total = total_tiles_per_head * num_head_kv;
  int ntpc = imax((int)((total + num_total_ctas - 1) / num_total_ctas), min_process_len / tilen);
  int64_t num_task = (int64_t)num_total_ctas * (ntpc + 1);
  int64_t chunk_bytes = (int64_t)num_head_kv * num_batch * 4;
  return (1 + num_task + (chunk_bytes + 47) / 48) * 48;
}

int oracle_assign_attention_decode_task(const int* num_seq_kvcache, int num_total_ctas,
                                        int num_batch, int num_head_kv, int num_seq_q, int tilen,
                                        int new_kv_included, int min_process_len, int* out,
                                        int64_t out_bytes) {
  int64_t need = oracle_taskmap_bytes(num_seq_kvcache, num_total_ctas, num_batch, num_head_kv,
                                      num_seq_q, tilen, new_kv_included, min_process_len);
  if (out_bytes < need) return 1;
  memset(out, 0, (size_t)need);

  /* assign_task.cu:367-381 */
  int* num_seqkvs = (int*)calloc((size_t)num_batch, sizeof(int));
  int* num_tiles = (int*)calloc((size_t)num_batch, sizeof(int));
  int64_t total_tiles_per_head = 0;
  for (int b = 0; b < num_batch; b++) {
    int n = new_kv_included ? num_seq_kvcache[b] : num_seq_kvcache[b] + num_seq_q;
    num_seqkvs[b] = n;
    num_tiles[b] = (n + tilen - 1) / tilen;
    total_tiles_per_head += num_tiles[b];
  }
  int64_t total_all = total_tiles_per_head * num_head_kv;
  int ntpc = imax((int)((total_all + num_total_ctas - 1) / num_total_ctas), min_process_len / tilen);

  int* tasks = out + ROW; /* row 0 is the header */
  size_t npairs = (size_t)num_batch * num_head_kv;
  int* num_chunks = (int*)calloc(npairs + 1, sizeof(int));
  int* start_tiles = (int*)calloc(npairs, sizeof(int));
  int* chunks_in_progress = (int*)calloc(npairs, sizeof(int));
  int* num_tiles_left = (int*)calloc(npairs, sizeof(int));
  for (int h = 0; h < num_head_kv; h++)
    for (int b = 0; b < num_batch; b++) num_tiles_left[(size_t)h * num_batch + b] = num_tiles[b];

  int ihead_kv = 0, ibatch = 0;
  int last_cta = 0, last_task = 0;

  /* assign_task.cu:402-478 */
  for (int icta = 0; icta < num_total_ctas; icta++) {
    int bucket = ntpc;
    int itask = 0;
    int* bin = tasks + (size_t)icta * (ntpc + 1) * ROW;
    while (bucket > 0 && ihead_kv < num_head_kv) {
      size_t idx = (size_t)ihead_kv * num_batch + ibatch;
      int num_tile = num_tiles_left[idx];
      if (num_tile <= 0) { /* skip empty pairs */
        ibatch++;
        if (ibatch >= num_batch) {
          ibatch = 0;
          ihead_kv++;
          if (ihead_kv >= num_head_kv) break;
        }
        continue;
      }
      int add_tiles = imin(num_tile, bucket);
      int num_seqkv = num_seqkvs[ibatch];
      if (chunks_in_progress[idx] == num_total_ctas - 1) add_tiles = num_tile;

      int* t = bin + (size_t)itask * ROW;
      memset(t, 0, ROW * sizeof(int));
      t[F_HEAD] = ihead_kv;
      t[F_BATCH] = ibatch;
      t[F_CHUNK] = chunks_in_progress[idx];
      t[F_SEQ_START] = start_tiles[idx] * tilen;
      t[F_NUM_SEQKV] = imin(add_tiles * tilen, num_seqkv - t[F_SEQ_START]);
      t[F_NUM_SEQKVCACHE] = t[F_NUM_SEQKV];
      t[F_NUM_TILE_KV] = (t[F_NUM_SEQKV] + tilen - 1) / tilen;
      t[F_NUM_TILE_FULL] = t[F_NUM_SEQKVCACHE] / tilen;
      t[F_IS_CAUSAL] = 0;

      itask++;
      chunks_in_progress[idx]++;
      start_tiles[idx] += add_tiles;
      num_tiles_left[idx] -= add_tiles;
      bucket -= add_tiles;

      if (num_tiles_left[idx] <= 0) { /* last chunk of this (head, batch) */
        int* cur = bin + (size_t)(itask - 1) * ROW;
        cur[F_IS_CAUSAL] = 1;
        cur[F_NUM_SEQKVCACHE] -= num_seq_q;
        cur[F_NUM_TILE_FULL] = imax(cur[F_NUM_SEQKVCACHE] / tilen, 0);
        num_chunks[idx] = chunks_in_progress[idx];
        if (cur[F_NUM_SEQKVCACHE] < 0) { /* causal window spills into the previous task */
          int* prev = tasks + ((size_t)last_cta * (ntpc + 1) + last_task) * ROW;
          prev[F_IS_CAUSAL] = 1;
          prev[F_NUM_SEQKVCACHE] += cur[F_NUM_SEQKVCACHE];
          prev[F_NUM_TILE_FULL] = imax(prev[F_NUM_SEQKVCACHE] / tilen, 0);
        }
        ibatch++;
        if (ibatch >= num_batch) {
          ibatch = 0;
          ihead_kv++;
        }
      }
      last_task = itask - 1;
    }
    last_cta = icta;
    for (int slot = itask; slot <= ntpc; slot++) { /* terminators, assign_task.cu:480-486 */
      int* t = bin + (size_t)slot * ROW;
      if (slot > itask) memset(t, 0, ROW * sizeof(int));
      else memset(t + 2, 0, (ROW - 2) * sizeof(int));
      t[F_HEAD] = -1;
      t[F_BATCH] = -1;
    }
  }

  /* entry.cc:750-776 */
  int64_t num_task = (int64_t)num_total_ctas * (ntpc + 1);
  int* chunk_dst = out + ROW * (num_task + 1);
  int max_chunks = 0;
  for (size_t r = 0; r < npairs; r++) {
    chunk_dst[r] = num_chunks[r];
    if (num_chunks[r] > max_chunks) max_chunks = num_chunks[r];
  }
  out[0] = ntpc + 1;
  out[1] = num_total_ctas;
  out[5] = max_chunks;

  free(num_seqkvs); free(num_tiles); free(num_chunks); free(start_tiles);
  free(chunks_in_progress); free(num_tiles_left);
  return 0;
}
