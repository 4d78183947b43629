// ORACLE — test infrastructure only.
// Thin C wrapper around the REAL reference CPU scheduler, compiled in place from
// /root/reference/src/attention/decode/assign_task.cu (see Makefile; output oracle/_ref/).
// Only this shim is ours: it packs the returned vectors the way the reference's torch CPU entry
// does (/root/reference/src/attention/entry.cc:750-776) and stubs the one helper the launcher
// side of that file references.
#include <cstdint>
#include <cstring>
#include <utility>
#include <vector>

#include "src/attention/decode/decode.h"

namespace hpc {
int get_sm_count() { return 0; }  // only used by the (never called) CUDA launcher
}  // namespace hpc

extern "C" int64_t ref_assign_attention_decode_task(const int* num_seq_kvcache, int num_total_ctas,
                                                    int num_batch, int num_head_kv, int num_seq_q,
                                                    int tilen, int new_kv_included,
                                                    int min_process_len, int* out,
                                                    int64_t out_bytes) {
  auto pr = hpc::attention::decode::assign_attention_decode_task_sync(
      num_seq_kvcache, num_total_ctas, num_batch, num_head_kv, num_seq_q, tilen,
      new_kv_included != 0, min_process_len);
  auto& tasks = pr.first;
  auto& num_chunks = pr.second;
  int num_tile_per_cta = num_chunks[num_head_kv * num_batch];
  constexpr int kTaskInfoSize = sizeof(hpc::attention::decode::dynamic::TaskScheduleInfo);
  int64_t num_task = static_cast<int64_t>(tasks.size());
  int64_t chunk_bytes = static_cast<int64_t>(num_head_kv) * num_batch * sizeof(int);
  int64_t rows = 1 + num_task + (chunk_bytes + kTaskInfoSize - 1) / kTaskInfoSize;
  int64_t need = rows * kTaskInfoSize;
  if (out == nullptr || out_bytes < need) return need;
  std::memset(out, 0, need);
  auto* p = reinterpret_cast<uint8_t*>(out);
  std::memcpy(p, &num_tile_per_cta, sizeof(int));
  std::memcpy(p + sizeof(int), &num_total_ctas, sizeof(int));
  std::memcpy(p + kTaskInfoSize, tasks.data(), kTaskInfoSize * num_task);
  std::memcpy(p + kTaskInfoSize * (num_task + 1), num_chunks.data(), chunk_bytes);
  int max_num_chunks = 0;
  for (int r = 0; r < num_head_kv * num_batch; r++) {
    if (num_chunks[r] > max_num_chunks) max_num_chunks = num_chunks[r];
  }
  std::memcpy(p + 5 * sizeof(int), &max_num_chunks, sizeof(int));
  return need;
}
