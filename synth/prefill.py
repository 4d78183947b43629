"""Synthetic inputs for FP8 block-sparse prefill (distributions and mask generator of reference
tests/test_attention_blocksparse_qpertoken_perhead_kvpertensor_fp8.py:21-35,135-190)."""
import math

import torch


def generate_block_sparse_mask(batch, heads, nrow, ncol, skip_ratio, causal=True, gen=None,
                               device="cpu"):
    """Block-level mask, True = attend; the causal diagonal is always kept
    (reference ...kvpertensor_fp8.py:21-35)."""
    mask = torch.rand(batch, heads, nrow, ncol, generator=gen, device=device) >= skip_ratio
    row_idx = torch.arange(nrow, device=device).view(nrow, 1)
    col_idx = torch.arange(ncol, device=device).view(1, ncol)
    if causal:
        causal_boundary = row_idx + (ncol - nrow)
        mask = mask & (col_idx <= causal_boundary)
        diag_col = torch.clamp(causal_boundary, max=ncol - 1)
        mask = mask | (col_idx == diag_col)
    return mask


def make_inputs(q_lens, kv_lens, Hq, Hkv, skip_ratio, k_per_token, seed=10086, layout="nhd",
                device="cpu", mask_cols=None):
    """Seeded inputs with the distributions of the reference tests (:135-190)."""
    dev = torch.device(device)
    g = torch.Generator(device=dev).manual_seed(seed)
    B, D, bs = len(q_lens), 128, 64
    total = sum(q_lens)
    max_q = max(q_lens)
    pad = (max_q + 127) // 128 * 128
    Q = (torch.randn(total, Hq, D, generator=g, device=dev) / math.sqrt(D)).to(torch.float8_e4m3fn)
    qscale = torch.randn(B, Hq, pad, generator=g, device=dev).abs() / 10 + 1e-3
    nblk = [(L + bs - 1) // bs for L in kv_lens]
    max_blocks = sum(nblk) * 2 + 2
    kv = torch.empty(max_blocks, 2, bs, Hkv, D, dtype=torch.float8_e4m3fn, device=dev)
    for b0 in range(0, max_blocks, 256):
        n = min(256, max_blocks - b0)
        kv[b0:b0 + n] = torch.randn(n, 2, bs, Hkv, D, generator=g, device=dev).to(torch.float8_e4m3fn)
    if layout == "hnd":
        kc = kv[:, 0].transpose(1, 2).contiguous().transpose(1, 2)
        vc = kv[:, 1].transpose(1, 2).contiguous().transpose(1, 2)
    else:
        kc, vc = kv[:, 0], kv[:, 1]
    perm = torch.randperm(max_blocks, generator=g, device=dev)[: sum(nblk)].to(torch.int32)
    block_ids = torch.zeros(B, max(nblk), dtype=torch.int32, device=dev)
    cu = 0
    for i in range(B):
        block_ids[i, : nblk[i]] = perm[cu:cu + nblk[i]]
        cu += nblk[i]
    if k_per_token:
        kscale = torch.randn(max_blocks, bs // 32, Hkv, D // 4, generator=g, device=dev).abs() + 0.05
        vscale = torch.randn(Hkv, generator=g, device=dev).abs() + 0.05
    else:
        kscale = torch.rand(1, generator=g, device=dev) + 0.5
        vscale = torch.randn(1, generator=g, device=dev)
    cu_q = torch.zeros(B + 1, dtype=torch.int32, device=dev)
    cu_q[1:] = torch.cumsum(torch.tensor(q_lens, device=dev), 0)
    mask = None
    if skip_ratio is not None:
        nrow = (max_q + 127) // 128
        ncol = mask_cols if mask_cols is not None else (max(kv_lens) + 127) // 128
        mask = generate_block_sparse_mask(B, Hq, nrow, ncol, skip_ratio, True, g, dev)
    return dict(q=Q, kcache=kc, vcache=vc, qscale=qscale, kscale=kscale, vscale=vscale,
                cu_seqlens_q=cu_q, seqlens_kv=torch.tensor(kv_lens, dtype=torch.int32, device=dev),
                block_ids=block_ids, block_mask=mask, max_q=max_q)
