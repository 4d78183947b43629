"""Synthetic inputs for FP8 paged decode attention (distributions of reference
benchmark/attention_decode/bench_attention_decode_fp8.py:135-186 and
tests/test_attention_decode_qpertoken_perhead_kvpertensor_fp8.py:84-150)."""
import math

import torch


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


def quant_paged_cache_pertoken(cache, block_size):
    """K cache bf16/f32 [blocks, bs + bs*4/D, Hkv, D] -> e4m3 with one scale per (token, head); the
    f32 scales are bit-cast into the extra rows: row t // 32 of head h holds tokens [32 r, 32 r + 32)
    (layout of reference tests/test_attention_decode_qkpertoken_perhead_vperhead_fp8.py:14-34)."""
    num_blocks, _, num_head_kv, head_dim = cache.shape
    scale = (cache[:, :block_size].float().abs().amax(-1) / 448).clamp_min(1e-8)  # [blocks, bs, Hkv]
    out = torch.empty(cache.shape, dtype=torch.float8_e4m3fn, device=cache.device)
    out[:, :block_size] = (cache[:, :block_size].float() / scale[..., None]).to(torch.float8_e4m3fn)
    rows = (scale.permute(0, 2, 1).contiguous().view(torch.float8_e4m3fn)
            .reshape(num_blocks, num_head_kv, -1, head_dim).permute(0, 2, 1, 3))
    out[:, block_size:] = rows
    return out


def make_decode_fp8_kpt_inputs(num_batch, num_seq_q, kv_lens_total, num_head_kv, num_head_q,
                               head_dim=128, block_size=64, seed=41, layout="NHD", device="cpu",
                               extra_blocks=8):
    """Seeded inputs for FP8 decode with q/k per-token-per-head and v per-head scales
    (distributions of reference tests/test_attention_decode_qkpertoken_perhead_vperhead_fp8.py:250-385).
    kvcache: e4m3 [blocks, 2, bs + 2, Hkv, D]; kcache / vcache / k_scale are views of it."""
    dev = torch.device(device)
    gen = torch.Generator(device=dev).manual_seed(seed)
    kv_lens_total = torch.as_tensor(kv_lens_total, dtype=torch.int32).cpu()
    nblocks = (kv_lens_total + block_size - 1) // block_size
    total_blocks = int(nblocks.sum())
    num_blocks = int(total_blocks * 1.2) + num_batch + extra_blocks
    srows = block_size * 4 // head_dim

    q = torch.randn((num_batch * num_seq_q, num_head_q, head_dim), generator=gen, device=dev)
    q = q / math.sqrt(head_dim)
    q_scale = q.abs().amax(-1).clamp_min(1e-6) / 10
    q8 = (q / q_scale[:, :, None]).to(torch.float8_e4m3fn)
    raw = torch.randn((num_blocks, 2, block_size + srows, num_head_kv, head_dim), generator=gen,
                      device=dev)
    kvcache = torch.empty(raw.shape, dtype=torch.float8_e4m3fn, device=dev)
    kvcache[:, 0] = quant_paged_cache_pertoken(raw[:, 0], block_size)
    vs = raw[:, 1, :block_size].abs().permute(2, 0, 1, 3).reshape(num_head_kv, -1).amax(-1) / 448
    kvcache[:, 1] = (raw[:, 1] / vs[None, None, :, None]).to(torch.float8_e4m3fn)
    v_scale = (vs * 0.1).float()

    perm = torch.randperm(num_blocks, generator=gen, device=dev)[:total_blocks].to(torch.int32).cpu()
    block_ids = torch.zeros((num_batch, int(nblocks.max())), dtype=torch.int32)
    cu = 0
    kv_u8 = kvcache.view(torch.uint8)
    for i in range(num_batch):
        nb = int(nblocks[i])
        block_ids[i, :nb] = perm[cu:cu + nb]
        cu += nb
        tail = int(kv_lens_total[i]) % block_size
        if tail:  # unused slots of the last block are zero (API contract)
            kv_u8[int(block_ids[i, nb - 1]), :, tail:block_size] = 0
    if layout == "HND":
        kvcache = kvcache.permute(0, 1, 3, 2, 4).contiguous().permute(0, 1, 3, 2, 4)
    kcache = kvcache[:, 0, :block_size]
    vcache = kvcache[:, 1, :block_size]
    k_scale = kvcache[:, 0, block_size:]
    return dict(q=q8, q_scale=q_scale.float(), kvcache=kvcache, kcache=kcache, vcache=vcache,
                k_scale=k_scale, v_scale=v_scale, block_ids=block_ids.to(dev),
                kv_lens_total=kv_lens_total.to(dev))


def make_decode_bf16_inputs(num_batch, num_seq_q, kv_lens_total, num_head_kv, num_head_q,
                            head_dim=128, block_size=64, seed=41, layout="NHD", device="cpu",
                            extra_blocks=8, q_std=1.5):
    """Seeded inputs for BF16 paged decode (shapes of reference tests/test_attention_decode_bf16.py:
    77-160; q is drawn wider than there so that the softmax is not close to uniform and an error in
    the scores is visible in the output). kv_lens_total includes the num_seq_q new tokens; unused
    slots of each request's last block are zero (API contract, reference hpc/attention.py:364)."""
    dev = torch.device(device)
    gen = torch.Generator(device=dev).manual_seed(seed)
    kv_lens_total = torch.as_tensor(kv_lens_total, dtype=torch.int32).cpu()
    nblocks = (kv_lens_total + block_size - 1) // block_size
    total_blocks = int(nblocks.sum())
    num_blocks = int(total_blocks * 1.2) + num_batch + extra_blocks

    q = (torch.randn((num_batch * num_seq_q, num_head_q, head_dim), generator=gen, device=dev)
         * q_std).to(torch.bfloat16)
    kvcache = torch.randn((num_blocks, 2, block_size, num_head_kv, head_dim), generator=gen,
                          device=dev).to(torch.bfloat16)
    perm = torch.randperm(num_blocks, generator=gen, device=dev)[:total_blocks].to(torch.int32).cpu()
    block_ids = torch.zeros((num_batch, int(nblocks.max())), dtype=torch.int32)
    cu = 0
    for i in range(num_batch):
        nb = int(nblocks[i])
        block_ids[i, :nb] = perm[cu:cu + nb]
        cu += nb
        tail = int(kv_lens_total[i]) % block_size
        if tail:
            kvcache[int(block_ids[i, nb - 1]), :, tail:] = 0
    if layout == "HND":
        kvcache = kvcache.permute(0, 1, 3, 2, 4).contiguous().permute(0, 1, 3, 2, 4)
    return dict(q=q, kvcache=kvcache, block_ids=block_ids.to(dev),
                kv_lens_total=kv_lens_total.to(dev))
