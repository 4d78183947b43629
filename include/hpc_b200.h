/*
 * hpc_b200.h — C ABI of the B200 (sm_100a) build of the HPC-Ops quantized-inference hot path.
 *
 * Every entry point is `extern "C"`, takes raw device/host pointers, plain integer sizes and
 * ELEMENT strides, float scalars and a cudaStream_t last — the same shape as the reference's
 * L1 host launchers (`*_async(void*…, cudaStream_t)`), which is what the reference's torch op
 * entries (src/<op>/entry.cc) bind.  No torch types cross this boundary.
 *
 * Return value: HPC_OK (0) or an HPC_ERR_* code; `hpc_last_error()` returns a thread-local,
 * human-readable message for the last failure (the Python layer raises RuntimeError with it,
 * mirroring the reference's TORCH_CHECK behaviour).
 *
 * Nothing here falls back to the CPU: a launcher either enqueues sm_100a kernels on `stream`
 * or fails loudly.
 */
#ifndef INCLUDE_HPC_B200_H_
#define INCLUDE_HPC_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#ifndef __CUDA_RUNTIME_API_H__
typedef struct CUstream_st* cudaStream_t;
#endif

enum {
  HPC_OK = 0,
  HPC_ERR_UNSUPPORTED = 1, /* shape / dtype / argument rejected (reference: TORCH_CHECK) */
  HPC_ERR_CUDA = 2,        /* a CUDA runtime call failed                                  */
  HPC_ERR_DRIVER = 3       /* a CUDA driver entry point (TMA descriptor encode) failed    */
};

/* ---- library ------------------------------------------------------------------------------- */
const char* hpc_last_error(void);
int hpc_sm_count(void);
/* replaces torch.ops.hpc.version / built_json (reference src/C/version.cc, src/C/built_json.cu) */
const char* hpc_version(void);
const char* hpc_built_json(void);

/* ---- decode attention: dynamic split-k task map ---------------------------------------------
 * replaces reference src/attention/decode/decode.h:39-46
 *   assign_attention_decode_task_sync  (CPU)  / assign_attention_decode_task_async (CUDA)
 * The CPU variant writes the packed host map of reference src/attention/entry.cc:750-776:
 *   row 0 = {num_tile_per_cta+1, num_total_ctas, 0,0,0, max_num_chunks}, then
 *   num_total_ctas*(num_tile_per_cta+1) 48-byte task rows, then num_chunks[h*num_batch+b].
 * The CUDA variant fills a workspace from get_attention_decode_task_workspace (hpc/attention.py).
 */
int64_t hpc_assign_attention_decode_task_host_bytes(const int* num_seq_kvcache,
                                                    int num_total_ctas, int num_batch,
                                                    int num_head_kv, int num_seq_q, int tilen,
                                                    int new_kv_included, int min_process_len);
int hpc_assign_attention_decode_task_sync(const int* num_seq_kvcache, int num_total_ctas,
                                          int num_batch, int num_head_kv, int num_seq_q, int tilen,
                                          int new_kv_included, int min_process_len,
                                          void* task_map_host, int64_t task_map_bytes);
int hpc_assign_attention_decode_task_async(int* task_map, const int* num_seq_kvcache,
                                           int num_total_ctas, int num_batch, int num_head_kv,
                                           int num_seq_q, int tilen, int new_kv_included,
                                           int min_process_len, cudaStream_t stream);

/* ---- decode attention: FP8 paged KV, split-k partials + combine -----------------------------
 * replaces reference src/attention/decode/decode.h:28-37 (attention_decode_fp8_async);
 * argument order and meaning are identical (strides in elements == bytes for fp8).
 *   lse        f32 [num_batch, splitk, num_head_k, num_seq_q, pad8(heads_per_group)]
 *   split_out  f32 [num_batch, splitk, num_seq_q, num_head_q, num_dim_v]
 *   splitk     = num_total_ctas of the task map (max chunks per (batch, kv head))
 */
int hpc_attention_decode_fp8_async(
    void* y_ptr, void* lse_ptr, void* split_out_ptr, const int* task_map_ptr, const void* q_ptr,
    void* kcache_ptr, void* vcache_ptr, const int* block_ids_ptr, const int* num_seq_kvcache_ptr,
    const float* qscale_ptr, const float* kscale_ptr, const float* vscale_ptr, int* split_flag_ptr,
    int new_kv_included, int splitk, int splitk_min_len, int consumers, int quant_type,
    int num_batch, int num_seq_q, int num_head_q, int num_head_k, int num_head_v, int num_dim_qk,
    int num_dim_v, int num_kvcache_blocks, int block_size, int num_seq_max_blocks,
    int qscale_pad_stride, int ldY, int ldQ, int64_t kcache_block_stride,
    int64_t kcache_token_stride, int64_t kcache_head_stride, int64_t vcache_block_stride,
    int64_t vcache_token_stride, int64_t vcache_head_stride, cudaStream_t stream);

/* The two stages of hpc_attention_decode_fp8_async on their own, same argument list: the split-k
 * attention kernel (writes lse / split_out) and the LSE combine (reads them, writes y). The
 * reference launches them back to back inside one launcher
 * (src/attention/decode/sm90/dynamic/...dynamic.cu:44-195); they are exposed separately so the
 * dominant kernel can be timed alone. */
int hpc_attention_decode_fp8_partial_async(
    void* y_ptr, void* lse_ptr, void* split_out_ptr, const int* task_map_ptr, const void* q_ptr,
    void* kcache_ptr, void* vcache_ptr, const int* block_ids_ptr, const int* num_seq_kvcache_ptr,
    const float* qscale_ptr, const float* kscale_ptr, const float* vscale_ptr, int* split_flag_ptr,
    int new_kv_included, int splitk, int splitk_min_len, int consumers, int quant_type,
    int num_batch, int num_seq_q, int num_head_q, int num_head_k, int num_head_v, int num_dim_qk,
    int num_dim_v, int num_kvcache_blocks, int block_size, int num_seq_max_blocks,
    int qscale_pad_stride, int ldY, int ldQ, int64_t kcache_block_stride,
    int64_t kcache_token_stride, int64_t kcache_head_stride, int64_t vcache_block_stride,
    int64_t vcache_token_stride, int64_t vcache_head_stride, cudaStream_t stream);
int hpc_attention_decode_fp8_combine_async(
    void* y_ptr, void* lse_ptr, void* split_out_ptr, const int* task_map_ptr, const void* q_ptr,
    void* kcache_ptr, void* vcache_ptr, const int* block_ids_ptr, const int* num_seq_kvcache_ptr,
    const float* qscale_ptr, const float* kscale_ptr, const float* vscale_ptr, int* split_flag_ptr,
    int new_kv_included, int splitk, int splitk_min_len, int consumers, int quant_type,
    int num_batch, int num_seq_q, int num_head_q, int num_head_k, int num_head_v, int num_dim_qk,
    int num_dim_v, int num_kvcache_blocks, int block_size, int num_seq_max_blocks,
    int qscale_pad_stride, int ldY, int ldQ, int64_t kcache_block_stride,
    int64_t kcache_token_stride, int64_t kcache_head_stride, int64_t vcache_block_stride,
    int64_t vcache_token_stride, int64_t vcache_head_stride, cudaStream_t stream);

/* ---- bring-up self test: one CTA, nk tcgen05.mma (kind::f8f6f4) with caller-supplied smem
 * images and descriptor fields; D[128, ncols] fp32 is copied out of TMEM. Used by tests to pin
 * the UMMA descriptor conventions the kernels rely on. */
int hpc_selftest_umma_f8(const void* a_image, int a_bytes, const void* b_image, int b_bytes,
                         float* d_out, int ncols, uint32_t idesc, int nk, uint32_t a_lbo,
                         uint32_t a_sbo, uint32_t a_layout, uint32_t a_kstep, uint32_t b_lbo,
                         uint32_t b_sbo, uint32_t b_layout, uint32_t b_kstep, cudaStream_t stream);

#ifdef __cplusplus
}
#endif

#endif /* INCLUDE_HPC_B200_H_ */
