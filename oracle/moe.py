"""ORACLE — test infrastructure only (see oracle/__init__.py).

Vectorised torch restatements (CPU) of the reference's in-test FusedMoE / grouped-GEMM references:
  blockwise : /root/reference/tests/test_fuse_moe_blockwise.py:23-262
  per-tensor: /root/reference/tests/test_fuse_moe_pertensor.py:23-151 (+ use_bf16_mul of
              tests/test_fuse_moe_cp_async.py)
  group gemm: /root/reference/tests/test_group_gemm_blockwise.py:21-47,
              /root/reference/tests/test_group_gemm_pertensor.py:20-44
Pinned by tests/golden/moe_*.npz (the reference's own functions executed on CPU by
tests/golden/make_golden.py).
"""
import torch


def gather_expert_inputs(x, x_scale, topk_ids, num_expert_local, rank_ep):
    """naive_gather_expert_inputs (test_fuse_moe_blockwise.py:23-81): rows of a local expert are
    consecutive, in flattened (token, k) order. Returns y, y_scale, topk_pos, counts, cu_counts."""
    T, K = topk_ids.shape
    flat = topk_ids.reshape(-1).long()
    lo = rank_ep * num_expert_local
    local = flat - lo
    valid = (local >= 0) & (local < num_expert_local)
    counts = torch.bincount(local[valid], minlength=num_expert_local).to(torch.int32)
    cu = torch.zeros(num_expert_local + 1, dtype=torch.int32)
    cu[1:] = torch.cumsum(counts, 0)
    # stable sort by expert keeps token order inside an expert
    idx = torch.nonzero(valid).reshape(-1)
    order = torch.argsort(local[idx], stable=True)
    src = idx[order]
    pos = torch.full((T * K,), -1, dtype=torch.int32)
    pos[src] = torch.arange(src.numel(), dtype=torch.int32)
    tok = (src // K)
    y = torch.zeros((T * K, x.shape[1]), dtype=x.dtype)
    y[: src.numel()] = x[tok]
    ys = None
    if x_scale is not None:
        ys = torch.zeros((T * K, x_scale.shape[1]), dtype=torch.float32)
        ys[: src.numel()] = x_scale[tok]
    return y, ys, pos.reshape(T, K), counts, cu


def group_gemm_blockwise(x, w, counts, cu, xscale, wscale):
    """naive_group_gemm (test_fuse_moe_blockwise.py:84-138): per 128x128 block fp32 product,
    tmp += prod * xscale[row, kb] * wscale[nb, kb]; bf16 out. xscale row-major [m, k/128]."""
    m, k = x.shape
    G, n, _ = w.shape
    KB = k // 128
    y = torch.zeros((m, n), dtype=torch.bfloat16)
    for g in range(G):
        c = int(counts[g])
        if c == 0:
            continue
        s = int(cu[g])
        xg = x[s:s + c].float().reshape(c, KB, 128)
        wg = w[g].float().reshape(n, KB, 128)
        ws = wscale[g][:, :KB].float().repeat_interleave(128, dim=0)  # [n, KB]
        xs = xscale[s:s + c].float()  # [c, KB]
        out = torch.zeros((c, n), dtype=torch.float32)
        for kb in range(KB):
            out += (xg[:, kb] @ wg[:, kb].t()) * xs[:, kb, None] * ws[None, :, kb]
        y[s:s + c] = out.to(torch.bfloat16)
    return y


def act_mul_and_blockwise_quant(gate_up_out):
    """naive_act_mul_and_blockwise_quant (test_fuse_moe_blockwise.py:141-193)."""
    gu = gate_up_out.float()
    gate, up = torch.chunk(gu, 2, dim=1)
    out = gate / (1 + (-gate).exp()) * up
    m, n = out.shape
    blk = out.reshape(m, n // 128, 128)
    scale = blk.abs().amax(dim=2) / 448.0
    q = (blk * (1.0 / (scale + 1e-8))[:, :, None]).to(torch.float8_e4m3fn).reshape(m, n)
    return q, scale


def reduce(x_bf16, topk_pos, topk_scale, shared_output=None):
    """naive_reduce (test_fuse_moe_blockwise.py:196-211): fp32 accumulate in k order."""
    T, K = topk_pos.shape
    acc = torch.zeros((T, x_bf16.shape[1]), dtype=torch.float32)
    for j in range(K):
        p = topk_pos[:, j].long()
        ok = p >= 0
        acc[ok] += x_bf16[p[ok]].float() * topk_scale[ok, j].float()[:, None]
    if shared_output is not None:
        acc += shared_output.float()
    return acc.to(torch.bfloat16)


def fuse_moe_blockwise(x, x_scale, gate_up_weight, gate_up_weight_scale, down_weight,
                       down_weight_scale, topk_ids, topk_scale, rank_ep, shared_output=None):
    """naive_fuse_moe_blockwise_fp8 (test_fuse_moe_blockwise.py:214-262)."""
    E = gate_up_weight.size(0)
    gi, gis, pos, counts, cu = gather_expert_inputs(x, x_scale, topk_ids, E, rank_ep)
    gu = group_gemm_blockwise(gi, gate_up_weight, counts, cu, gis, gate_up_weight_scale)
    di, dis = act_mul_and_blockwise_quant(gu)
    do = group_gemm_blockwise(di, down_weight, counts, cu, dis, down_weight_scale)
    return reduce(do, pos, topk_scale, shared_output)


def group_gemm_pertensor(x, w, cu, scale):
    """naive_group_gemm (test_fuse_moe_pertensor.py:72-91)."""
    m = x.shape[0]
    G, n, _ = w.shape
    y = torch.zeros((m, n), dtype=torch.bfloat16)
    for g in range(G):
        s, e = int(cu[g]), int(cu[g + 1])
        if e > s:
            y[s:e] = ((x[s:e].float() @ w[g].float().t()) * scale[g].float()).to(torch.bfloat16)
    return y


def act_mul_and_quant(gate_up, scale, use_bf16_mul=True):
    """naive_act_mul_and_quant (test_fuse_moe_pertensor.py:94-103): bf16 product, * scale -> e4m3."""
    gate, up = torch.chunk(gate_up.float(), 2, dim=1)
    a = gate / (1 + (-gate).exp())
    if use_bf16_mul:
        out = (a.to(torch.bfloat16) * up.to(torch.bfloat16)).float() * scale
    else:
        out = a * up * scale
    return out.to(torch.float8_e4m3fn)


def fuse_moe_pertensor(x, gate_up_weight, down_weight, gate_up_scale, down_scale, act_scale,
                       topk_ids, topk_scale, rank_ep, shared_output=None, use_bf16_mul=True):
    """naive_fuse_moe_pertensor_fp8 (test_fuse_moe_pertensor.py:118-151)."""
    E = gate_up_weight.size(0)
    gi, _, pos, counts, cu = gather_expert_inputs(x, None, topk_ids, E, rank_ep)
    gu = group_gemm_pertensor(gi, gate_up_weight, cu, gate_up_scale)
    di = act_mul_and_quant(gu, act_scale, use_bf16_mul)
    do = group_gemm_pertensor(di, down_weight, cu, down_scale)
    return reduce(do, pos, topk_scale, shared_output)


def group_gemm_blockwise_standalone(x, w, seqlens, cu_seqlens, xscale_t, wscale, pad_per_group):
    """naive_group_gemm of tests/test_group_gemm_blockwise.py:21-47 (scales applied to bf16-rounded
    operands, bf16 matmul). xscale_t [k/128, G*pad_per_group]."""
    m, k = x.shape
    G, n, _ = w.shape
    xs = xscale_t.repeat_interleave(128, dim=0).permute(1, 0).reshape(G, pad_per_group, k)
    wsf = wscale.repeat_interleave(128, dim=1).repeat_interleave(128, dim=2)[:, :, :k]
    y = torch.zeros((m, n), dtype=torch.bfloat16)
    for g in range(G):
        c, s = int(seqlens[g]), int(cu_seqlens[g])
        if c == 0:
            continue
        xg = (x[s:s + c].to(torch.bfloat16) * xs[g, :c].to(torch.bfloat16)).to(torch.bfloat16)
        wg = (w[g].to(torch.bfloat16) * wsf[g].to(torch.bfloat16)).to(torch.bfloat16)
        y[s:s + c] = (xg.float() @ wg.float().t()).to(torch.bfloat16)
    return y


# synthetic inputs live in synth/ (neutral code), re-exported for the tests
from synth.moe import make_moe_blockwise_inputs  # noqa: E402,F401
