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


# ---------------------------------------------------------------------------------------------
# synthetic input builders shared by tests and bench (distributions of the reference's tests /
# benchmark: reference benchmark/attention_decode/bench_attention_decode_fp8.py:135-186)
# ---------------------------------------------------------------------------------------------
def make_decode_fp8_inputs(num_batch, num_seq_q, kv_lens_total, num_head_kv, num_head_q,
                           head_dim=128, block_size=64, seed=41, layout="NHD", device="cpu",
                           extra_blocks=8, dtype=torch.float8_e4m3fn):
    """Seeded inputs for FP8 decode. kv_lens_total includes the num_seq_q new tokens.
    Unused slots of each request's last block are zero (API contract, hpc/attention.py:364)."""
    dev = torch.device(device)
    gen = torch.Generator(device=dev).manual_seed(seed)
    kv_lens_total = torch.as_tensor(kv_lens_total, dtype=torch.int32).cpu()
    nblocks = (kv_lens_total + block_size - 1) // block_size
    total_blocks = int(nblocks.sum())
    num_blocks = int(total_blocks * 1.2) + num_batch + extra_blocks

    q = torch.randn((num_batch * num_seq_q, num_head_q, head_dim), generator=gen, device=dev)
    q = q / math.sqrt(head_dim)
    q_scale = q.abs().amax(-1).clamp_min(1e-6) / 10
    q8 = (q / q_scale[:, :, None]).to(torch.float8_e4m3fn)
    kvcache = torch.empty((num_blocks, 2, block_size, num_head_kv, head_dim),
                          dtype=torch.float8_e4m3fn, device=dev)
    step = 256  # generate in slabs: the fp32 staging buffer stays small
    for b0 in range(0, num_blocks, step):
        n = min(step, num_blocks - b0)
        slab = torch.randn((n, 2, block_size, num_head_kv, head_dim), generator=gen, device=dev)
        slab[:, 0] /= math.sqrt(head_dim)
        kvcache[b0:b0 + n] = slab.to(torch.float8_e4m3fn)
    k_scale = torch.rand(1, generator=gen, device=dev).clamp_min(0.05)
    v_scale = torch.rand(1, generator=gen, device=dev).clamp_min(0.05)

    perm = torch.randperm(num_blocks, generator=gen, device=dev)[:total_blocks].to(torch.int32).cpu()
    max_blocks = int(nblocks.max())
    block_ids = torch.zeros((num_batch, max_blocks), dtype=torch.int32)
    cu = 0
    kv_u8 = kvcache.view(torch.uint8)
    for i in range(num_batch):
        nb = int(nblocks[i])
        block_ids[i, :nb] = perm[cu:cu + nb]
        cu += nb
        tail = int(kv_lens_total[i]) % block_size
        if tail:
            kv_u8[int(block_ids[i, nb - 1]), :, tail:] = 0
    if layout == "HND":
        kvcache = kvcache.permute(0, 1, 3, 2, 4).contiguous().permute(0, 1, 3, 2, 4)
    return dict(q=q8, q_scale=q_scale.float(), kvcache=kvcache, k_scale=k_scale, v_scale=v_scale,
                block_ids=block_ids.to(dev), kv_lens_total=kv_lens_total.to(dev))
