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
