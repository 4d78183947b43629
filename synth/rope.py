"""Synthetic inputs for rope_norm_store_kv[_fp8] (distributions and padding of reference
tests/test_rope.py:15-32,122-225)."""
import torch


def generate_cos_sin_cache(max_position, head_dim, base=10000.0):
    inv_freq = 1.0 / (base ** (torch.arange(0, head_dim, 2).float() / head_dim))
    t = torch.arange(max_position).float()
    freqs = torch.outer(t, inv_freq)
    return torch.cat([freqs.cos(), freqs.sin()], dim=-1)


def make_inputs(num_req, is_prefill, mtp, num_q_heads, num_kv_heads, qk_head_dim, v_head_dim=None,
                kv_block_size=64, max_num_kv_blocks=256, max_rope_position=2048, seed=0,
                len_range=(20, 200), pad_decode=True):
    """Returns a dict of CPU tensors; decode batches are padded to a multiple of 8 requests / rows
    like a CUDA-graph batch (padding requests have length 0 and q_index == padded row count)."""
    g = torch.Generator().manual_seed(seed)
    v_head_dim = v_head_dim or qk_head_dim
    hidden = num_q_heads * qk_head_dim + num_kv_heads * (qk_head_dim + v_head_dim)
    cos_sin = generate_cos_sin_cache(max_rope_position, qk_head_dim)
    kcache = torch.randn(max_num_kv_blocks, kv_block_size, num_kv_heads, qk_head_dim, generator=g).to(torch.bfloat16)
    vcache = torch.randn(max_num_kv_blocks, kv_block_size, num_kv_heads, v_head_dim, generator=g).to(torch.bfloat16)
    qw = torch.randn(qk_head_dim, generator=g)
    kw = torch.randn(qk_head_dim, generator=g)
    lens = torch.randint(len_range[0], len_range[1], (num_req,), generator=g)
    if is_prefill:
        q_len = torch.minimum((torch.rand(num_req, generator=g) * lens).long() + 1, lens)
        total = lens
    else:
        q_len = torch.full((num_req,), mtp + 1)
        total = lens + mtp + 1
    rows = int(q_len.sum())
    qkv = torch.randn(rows, hidden, generator=g).to(torch.bfloat16)
    q_index = torch.zeros(num_req + 1, dtype=torch.int32)
    q_index[1:] = torch.cumsum(q_len, 0)
    nblk = (total + kv_block_size - 1) // kv_block_size
    perm = torch.randperm(max_num_kv_blocks, generator=g)
    kv_idx = torch.full((num_req, int(nblk.max()) + 4), -1, dtype=torch.int32)
    off = 0
    for i in range(num_req):
        kv_idx[i, :int(nblk[i])] = perm[off:off + int(nblk[i])]
        off += int(nblk[i])
    num_seqlen = total.to(torch.int32)
    real_rows = None
    if not is_prefill and pad_decode:
        real_rows = rows
        pb, pr = (num_req + 7) // 8 * 8, (rows + 7) // 8 * 8
        qkv = torch.cat([qkv, torch.zeros(pr - rows, hidden, dtype=qkv.dtype)])
        num_seqlen = torch.cat([num_seqlen, torch.zeros(pb - num_req, dtype=torch.int32)])
        q_index = torch.cat([q_index, torch.full((pb - num_req,), pr, dtype=torch.int32)])
        kv_idx = torch.cat([kv_idx, torch.zeros(pb - num_req, kv_idx.shape[1], dtype=torch.int32)])
    return dict(qkv=qkv, num_seqlen=num_seqlen, q_index=q_index, kcache=kcache, vcache=vcache,
                kv_indices=kv_idx, q_norm_w=qw, k_norm_w=kw, cos_sin=cos_sin, real_rows=real_rows,
                num_req=num_req)
