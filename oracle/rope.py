"""ORACLE — test infrastructure only (see oracle/__init__.py).

Restatement of the reference's in-test RoPE + QK-norm + paged KV store reference
(/root/reference/tests/test_rope.py:39-119 `rope_norm_ref` with its helpers :35-46), vectorised per
request, plus the FP8 quantisation rules of the reference kernel (src/rope/rope.cu:655-667 dynamic
Q scale = amax / upper_max and q * rcp(scale); :683,729-761 K, V stored as x * rcp(static scale)).
Pinned by tests/golden/rope_*.npz (the reference's own `rope_norm_ref` executed on CPU by
tests/golden/make_golden.py).
"""
import torch


def rms_norm(x, weight, eps=1e-6):
    return x * torch.rsqrt(x.pow(2).mean(-1, keepdim=True) + eps) * weight


def rotary_neox(x, cos_sin):
    h = x.shape[-1] // 2
    x1, x2 = x[..., :h], x[..., h:]
    c = cos_sin[:, :h].unsqueeze(1)
    s = cos_sin[:, h:].unsqueeze(1)
    return torch.cat([x1 * c - x2 * s, x2 * c + x1 * s], dim=-1)


def _split_rows(kcache, vcache, qkv, num_seqlen_per_req, q_index):
    num_kv, qk_dim, v_dim = kcache.shape[2], kcache.shape[3], vcache.shape[3]
    num_q = (qkv.shape[1] - num_kv * qk_dim - num_kv * v_dim) // qk_dim
    num_rows = int(q_index[-1])
    q = qkv[:num_rows, : num_q * qk_dim].float().view(num_rows, num_q, qk_dim)
    k = qkv[:num_rows, num_q * qk_dim:(num_q + num_kv) * qk_dim].float().view(num_rows, num_kv, qk_dim)
    v = qkv[:num_rows, (num_q + num_kv) * qk_dim:].view(num_rows, num_kv, v_dim)
    return q, k, v, num_rows


def _positions(num_seqlen_per_req, q_index):
    """absolute position and request of every row"""
    pos, req = [], []
    for i in range(num_seqlen_per_req.shape[0]):
        sl = int(num_seqlen_per_req[i])
        ql = int(q_index[i + 1]) - int(q_index[i])
        pos += list(range(sl - ql, sl))
        req += [i] * ql
    return torch.tensor(pos, dtype=torch.long), torch.tensor(req, dtype=torch.long)


def rope_norm_qk(kcache, vcache, qkv, cos_sin, num_seqlen_per_req, q_index, q_norm_weight,
                 k_norm_weight, qk_norm_policy):
    """fp32 rotated (and normalised) q, k and the untouched v rows, plus (pos, req) per row."""
    q, k, v, num_rows = _split_rows(kcache, vcache, qkv, num_seqlen_per_req, q_index)
    pos, req = _positions(num_seqlen_per_req, q_index)
    cs = cos_sin[pos]
    if qk_norm_policy == 2:
        q, k = rms_norm(q, q_norm_weight), rms_norm(k, k_norm_weight)
    q, k = rotary_neox(q, cs), rotary_neox(k, cs)
    if qk_norm_policy == 1:
        q, k = rms_norm(q, q_norm_weight), rms_norm(k, k_norm_weight)
    return q, k, v, pos, req


def _store(cache, rows, pos, req, num_seqlen_per_req, kv_indices):
    """paged write + zeroing of the tail of each request's last page (test_rope.py:104-117)."""
    blk = cache.shape[1]
    for t in range(rows.shape[0]):
        r, p = int(req[t]), int(pos[t])
        cb = int(kv_indices[r, p // blk])
        cache[cb, p % blk] = rows[t].to(cache.dtype)
        if p == int(num_seqlen_per_req[r]) - 1 and p % blk + 1 < blk:
            cache[cb, p % blk + 1:] = 0


def rope_norm_store_kv(kcache, vcache, qkv, cos_sin, num_seqlen_per_req, q_index, kv_indices,
                       q_norm_weight=None, k_norm_weight=None, qk_norm_policy=0):
    """bf16 path: updates kcache / vcache in place, returns q [rows, Hq, D] in qkv's dtype."""
    q, k, v, pos, req = rope_norm_qk(kcache, vcache, qkv, cos_sin, num_seqlen_per_req, q_index,
                                     q_norm_weight, k_norm_weight, qk_norm_policy)
    _store(kcache, k, pos, req, num_seqlen_per_req, kv_indices)
    _store(vcache, v, pos, req, num_seqlen_per_req, kv_indices)
    return q.to(qkv.dtype)


def rope_norm_store_kv_fp8(kcache, vcache, qkv, cos_sin, num_seqlen_per_req, q_index, kv_indices,
                           k_scale, v_scale, quant_policy, q_scale_inv=None, upper_max=448.0,
                           q_norm_weight=None, k_norm_weight=None, qk_norm_policy=0):
    """fp8 path: caches are e4m3, updated in place with x * (1 / scale); returns
    (q_fp8 [rows, Hq, D] e4m3, q_scale [rows, Hq] f32 or None)."""
    q, k, v, pos, req = rope_norm_qk(kcache, vcache, qkv, cos_sin, num_seqlen_per_req, q_index,
                                     q_norm_weight, k_norm_weight, qk_norm_policy)
    kq = (k * (1.0 / k_scale.float())).to(torch.float8_e4m3fn)
    vq = (v.float() * (1.0 / v_scale.float())).to(torch.float8_e4m3fn)
    _store(kcache, kq, pos, req, num_seqlen_per_req, kv_indices)
    _store(vcache, vq, pos, req, num_seqlen_per_req, kv_indices)
    if quant_policy == 1:
        q_scale = q.abs().amax(-1) / upper_max
        mult = torch.where(q_scale > 0, 1.0 / q_scale, torch.zeros_like(q_scale))
        return (q * mult[..., None]).to(torch.float8_e4m3fn), q_scale
    return (q * q_scale_inv.float()).to(torch.float8_e4m3fn), None
