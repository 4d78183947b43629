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


# mask generator and synthetic inputs live in synth/ (neutral code), re-exported for the tests
from synth.prefill import generate_block_sparse_mask, make_inputs  # noqa: E402,F401


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
