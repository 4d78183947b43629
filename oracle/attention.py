"""ORACLE — test infrastructure only (see oracle/__init__.py).

Torch restatements (device-agnostic, run on CPU) of the reference's in-test attention references.
Each function cites the reference lines it follows. They are pinned by tests/golden/*.npz, which
tests/golden/make_golden.py produced by executing the reference's own functions (AST-extracted
from /root/reference/tests) on the same seeded inputs.
"""
import math

import torch


def _gather_kv(cache, blk_ids, seqlen, num_head_kv, head_dim, head_per_group):
    # cache: [blocks, block_size, Hkv, D] (any strides). -> [Hq, seqlen, D] float32
    x = cache[blk_ids.long()].reshape(-1, num_head_kv, head_dim).transpose(0, 1)[:, :seqlen, :]
    return x.repeat_interleave(head_per_group, dim=0).float()


def _causal_mask(sq, seqlen, device):
    # rows = new tokens, cols = all kv; new token i sees kv positions <= seqlen - sq + i
    # (reference tests/test_attention_decode_bf16.py:47-53)
    head = torch.ones(sq, seqlen - sq, device=device, dtype=torch.bool)
    tail = torch.tril(torch.ones(sq, sq, device=device, dtype=torch.bool))
    return torch.cat([head, tail], dim=-1).unsqueeze(0)


def decode_bf16(q, kcache, vcache, block_ids, kv_lens_total, num_seq_q):
    """bf16 paged decode attention.

    Follows reference tests/test_attention_decode_bf16.py:15-59
    (`ref_attn_with_paged_kvcache_func`), with K/V passed as two cache views.
      q [B*Sq, Hq, D] bf16; kcache/vcache [blocks, bs, Hkv, D]; block_ids [B, max_blocks];
      kv_lens_total [B] = tokens in cache including the Sq new ones.
    """
    num_batch = kv_lens_total.shape[0]
    num_head_q, head_dim = q.shape[1], q.shape[2]
    num_head_kv, block_size = kcache.shape[2], kcache.shape[1]
    g = num_head_q // num_head_kv
    qv = q.reshape(num_batch, num_seq_q, num_head_q, head_dim)
    out = torch.empty_like(qv)
    for bi in range(num_batch):
        seqlen = int(kv_lens_total[bi])
        nblk = (seqlen + block_size - 1) // block_size
        ids = block_ids[bi, :nblk]
        qb = qv[bi].transpose(0, 1).float()
        kb = _gather_kv(kcache, ids, seqlen, num_head_kv, head_dim, g)
        vb = _gather_kv(vcache, ids, seqlen, num_head_kv, head_dim, g)
        p = qb @ kb.transpose(-1, -2) / math.sqrt(head_dim)
        p = p.masked_fill(~_causal_mask(num_seq_q, seqlen, q.device), float("-inf"))
        w = torch.softmax(p, dim=-1)
        out[bi] = (w @ vb).transpose(0, 1).to(out.dtype)
    return out.reshape(-1, num_head_q, head_dim)


def decode_fp8_kvpertensor(q, kcache, vcache, block_ids, kv_lens_total, q_scale, k_scale, v_scale,
                           num_seq_q, per_token_qscale=True):
    """FP8 paged decode attention, q per-token/per-head scale, k/v per-tensor scale.

    Follows reference tests/test_attention_decode_qpertoken_perhead_kvpertensor_fp8.py:14-79:
      S = Q K^T / sqrt(D) * q_scale * k_scale; causal mask; P = exp(S - max);
      sum over unquantised P; P*256 -> e4m3 -> fp32; Y = (P V) / sum * (v_scale / 256); bf16.
    `per_token_qscale=True` indexes q_scale by token (kernel semantics, reference
    ...dynamic_splitk_kernels.cuh:280-293); False reproduces the test's `q_scale[bi]` indexing
    (identical when num_seq_q == 1).
      q [B*Sq, Hq, D] e4m3; q_scale [B*Sq, Hq] f32; k_scale, v_scale [1] f32.
    """
    num_batch = kv_lens_total.shape[0]
    num_head_q, head_dim = q.shape[1], q.shape[2]
    num_head_kv, block_size = kcache.shape[2], kcache.shape[1]
    g = num_head_q // num_head_kv
    qv = q.reshape(num_batch, num_seq_q, num_head_q, head_dim)
    qs = q_scale.reshape(-1, num_head_q)
    out = torch.empty(qv.shape, dtype=torch.bfloat16, device=q.device)
    for bi in range(num_batch):
        seqlen = int(kv_lens_total[bi])
        nblk = (seqlen + block_size - 1) // block_size
        ids = block_ids[bi, :nblk]
        qb = qv[bi].transpose(0, 1).float()  # [Hq, Sq, D]
        kb = _gather_kv(kcache, ids, seqlen, num_head_kv, head_dim, g)
        vb = _gather_kv(vcache, ids, seqlen, num_head_kv, head_dim, g)
        p = qb @ kb.transpose(-1, -2)
        if per_token_qscale:
            sc = qs[bi * num_seq_q:(bi + 1) * num_seq_q].transpose(0, 1)[:, :, None]  # [Hq,Sq,1]
        else:
            sc = qs[bi][:, None, None]
        p = p / math.sqrt(head_dim) * sc * k_scale
        p = p.masked_fill(~_causal_mask(num_seq_q, seqlen, q.device), float("-inf"))
        w = torch.exp(p - p.max(dim=-1)[0][:, :, None])
        gsum = w.sum(dim=-1)[:, :, None]
        w = (w * 256.0).to(torch.float8_e4m3fn).float()
        y = torch.matmul(w, vb) / gsum * (v_scale / 256.0)
        out[bi] = y.transpose(0, 1).to(torch.bfloat16)
    return out.reshape(-1, num_head_q, head_dim)


def decode_fp8_kpertoken(q, kcache, vcache, block_ids, kv_lens_total, q_scale, k_scale, v_scale,
                         num_seq_q, per_token_qscale=True):
    """FP8 paged decode attention, q and k per-token/per-head scales, v per-head scale.

    Follows reference tests/test_attention_decode_qkpertoken_perhead_vperhead_fp8.py:55-132:
      S = Q K^T / sqrt(D) * k_scale[token, head] * q_scale; causal mask; P = exp(S - max); sum over
      unquantised P; P*256 -> e4m3; Y = (P V) / sum * v_scale[head] / 256; bf16.
      k_scale: the cache's scale rows, fp8 view or f32, [blocks, bs/32, Hkv, D or D/4];
      v_scale [Hkv] f32. q_scale indexing as in decode_fp8_kvpertensor.
    """
    num_batch = kv_lens_total.shape[0]
    num_head_q, head_dim = q.shape[1], q.shape[2]
    num_head_kv, block_size = kcache.shape[2], kcache.shape[1]
    g = num_head_q // num_head_kv
    qv = q.reshape(num_batch, num_seq_q, num_head_q, head_dim)
    qs = q_scale.reshape(-1, num_head_q)
    ks = k_scale.contiguous()
    if ks.element_size() == 1:
        ks = ks.view(torch.float32)
    out = torch.empty(qv.shape, dtype=torch.bfloat16, device=q.device)
    for bi in range(num_batch):
        seqlen = int(kv_lens_total[bi])
        nblk = (seqlen + block_size - 1) // block_size
        ids = block_ids[bi, :nblk]
        qb = qv[bi].transpose(0, 1).float()
        kb = _gather_kv(kcache, ids, seqlen, num_head_kv, head_dim, g)
        vb = _gather_kv(vcache, ids, seqlen, num_head_kv, head_dim, g)
        ksb = (ks[ids.long()].permute(0, 1, 3, 2).reshape(-1, num_head_kv).transpose(0, 1)[:, :seqlen]
               .repeat_interleave(g, dim=0)).float()
        p = qb @ kb.transpose(-1, -2)
        if per_token_qscale:
            sc = qs[bi * num_seq_q:(bi + 1) * num_seq_q].transpose(0, 1)[:, :, None]
        else:
            sc = qs[bi][:, None, None]
        p = p / math.sqrt(head_dim) * ksb.unsqueeze(1) * sc
        p = p.masked_fill(~_causal_mask(num_seq_q, seqlen, q.device), float("-inf"))
        w = torch.exp(p - p.max(dim=-1)[0][:, :, None])
        gsum = w.sum(dim=-1)[:, :, None]
        w = (w * 256.0).to(torch.float8_e4m3fn).float()
        y = torch.matmul(w, vb) / gsum
        y = y * v_scale[:, None, None].repeat_interleave(g, dim=0) / 256.0
        out[bi] = y.transpose(0, 1).to(torch.bfloat16)
    return out.reshape(-1, num_head_q, head_dim)


# synthetic input builders: live in synth/ (neutral code), re-exported for the tests
from synth.decode import (make_decode_bf16_inputs, make_decode_fp8_inputs,  # noqa: E402,F401
                          make_decode_fp8_kpt_inputs)
