"""ORACLE — test infrastructure only (see oracle/__init__.py).

Torch restatements of the reference's block-sparse FP8 prefill references:
  kv-per-tensor : /root/reference/tests/test_attention_blocksparse_qpertoken_perhead_kvpertensor_fp8.py:21-109
  k-per-token   : /root/reference/tests/test_attention_blocksparse_qkpertoken_perhead_vperhead_fp8.py:17-106
generalised to ragged requests (q length != kv length per request). Pinned by
tests/golden/prefill_*.npz (reference functions executed on CPU by tests/golden/make_golden.py).
"""
import math

import torch

BSA_BLOCK = 128


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


def blocksparse_prefill(q, kcache, vcache, qscale, kscale, vscale, cu_seqlens_q, seqlens_kv,
                        block_ids, block_mask=None, k_per_token=False):
    """q [total, Hq, D] e4m3; caches [blocks, bs, Hkv, D]; qscale [B, Hq, pad]; block_mask bool
    [B, Hq, nrow, ncol] or None. Returns bf16 [total, Hq, D]."""
    total, Hq, D = q.shape
    bs, Hkv = kcache.shape[1], kcache.shape[2]
    g = Hq // Hkv
    B = seqlens_kv.shape[0]
    out = torch.empty((total, Hq, D), dtype=torch.bfloat16)
    for i in range(B):
        s0, s1 = int(cu_seqlens_q[i]), int(cu_seqlens_q[i + 1])
        nq, nkv = s1 - s0, int(seqlens_kv[i])
        if nq == 0:
            continue
        nblk = (nkv + bs - 1) // bs
        ids = block_ids[i, :nblk].long()
        BQ = q[s0:s1].transpose(0, 1).float()
        BK = kcache[ids].reshape(-1, Hkv, D).transpose(0, 1)[:, :nkv].repeat_interleave(g, 0).float()
        BV = vcache[ids].reshape(-1, Hkv, D).transpose(0, 1)[:, :nkv].repeat_interleave(g, 0).float()
        scale = qscale[i, :, :nq].unsqueeze(-1)
        scores = torch.matmul(BQ, BK.transpose(-2, -1)) / math.sqrt(D)
        if k_per_token:
            BKS = (kscale[ids].permute(0, 1, 3, 2).reshape(-1, Hkv).transpose(0, 1)[:, :nkv]
                   .repeat_interleave(g, 0)).float()
            scores = scores * scale * BKS.unsqueeze(1)
        else:
            scores = scores * scale * kscale[0]
        if block_mask is not None:
            bm = block_mask[i].bool()
            em = bm.repeat_interleave(BSA_BLOCK, dim=-2)[:, :nq, :]
            em = em.repeat_interleave(BSA_BLOCK, dim=-1)
            if em.shape[-1] < nkv:  # tiles past the mask width: only the first one is visited
                pad = torch.zeros(em.shape[0], nq, nkv - em.shape[-1], dtype=torch.bool)
                pad[:, :, :BSA_BLOCK] = True
                em = torch.cat([em, pad], dim=-1)
            scores = scores.masked_fill(~em[:, :, :nkv], float("-inf"))
        cm = torch.tril(torch.ones(nkv, nkv, dtype=torch.bool))[nkv - nq:, :].unsqueeze(0)
        scores = scores.masked_fill(~cm, float("-inf"))
        w = torch.exp(scores - scores.max(dim=-1, keepdim=True)[0])
        gsum = w.sum(dim=-1, keepdim=True)
        w = (w * 256.0).to(torch.float8_e4m3fn).float()
        o = torch.matmul(w, BV) / gsum
        if k_per_token:
            o = o * (vscale[:, None, None].repeat_interleave(g, 0) / 256.0)
        else:
            o = o * (vscale[0] / 256.0)
        out[s0:s1] = o.transpose(0, 1).to(torch.bfloat16)
    return out


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
