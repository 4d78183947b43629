// Definitions shared by the paged decode attention kernels (fp8: decode_attn_fp8.cu, bf16:
// decode_attn_bf16.cu): task-map rows, kernel parameters, the split-k combine launcher.
#pragma once
#include "common.cuh"
#include "host_utils.h"

namespace b200 {
namespace decode {

constexpr int kTileN = 128;  // keys per tile == UMMA M
constexpr int kPage = 64;    // paged block size (tokens) of the fp8 caches
constexpr int kD = 128;      // head dim
constexpr int kTaskStride = 12;
constexpr int kSoftmaxBar = 1;

struct Params {
  const int* task_map;
  const int* block_ids;
  const float* qscale;
  const float* kscale;
  const float* vscale;
  float* split_out;
  float* lse;
  __nv_bfloat16* y;  // final output: tasks that are the only chunk of their (batch, kv head) write it directly
  int ld_y;
  int num_batch;
  int num_seq_q;
  int num_head_q;
  int num_head_kv;
  int group;
  int num_seq_max_blocks;
  int qscale_stride;
  int max_splitk;
  int lse_pad;
  int k_head_first;  // TMA dim order of the cache maps: (d, head, token, blk) or (d, token, head, blk)
  int v_head_first;
  float softmax_scale_log2;
  // k-per-token variant: in-cache scale rows: float index
  //   blk * ks_blk + (t / 32) * ks_row + head * ks_head + t % 32     (t = token slot in the page)
  long long ks_blk, ks_row, ks_head;
};

struct Task {
  int ihead_kv, ibatch, ichunk, iseq_start;
  int num_seqkv, num_seqkvcache, num_tile_kv, num_tile_full;
  int is_causal;
};

__device__ __forceinline__ bool load_task(const int* row, Task& t) {
  int4 a = *reinterpret_cast<const int4*>(row);
  if (a.x < 0 || a.y < 0) return false;
  int4 b = *reinterpret_cast<const int4*>(row + 4);
  int c = row[8];
  t.ihead_kv = a.x;
  t.ibatch = a.y;
  t.ichunk = a.z;
  t.iseq_start = a.w;
  t.num_seqkv = b.x;
  t.num_seqkvcache = b.y;
  t.num_tile_kv = b.z;
  t.num_tile_full = b.w;
  t.is_causal = c;
  return true;
}

// split-k combine (decode_attn_fp8.cu): y = sum_c 2^(lse_c - m) O_c / sum_c 2^(lse_c - m) -> bf16;
// (batch, kv head) pairs with a single chunk were written by the attention kernel itself.
cudaError_t launch_combine(__nv_bfloat16* y, const float* split_out, const float* lse,
                           const int* task_map, int num_batch, int num_seq_q, int num_head_q,
                           int num_head_kv, int group, int max_splitk, int lse_pad, int ldY,
                           cudaStream_t stream);

}  // namespace decode
}  // namespace b200
