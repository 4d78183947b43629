"""ORACLE — test infrastructure only (see oracle/__init__.py).

Restatement of the reference's in-test allreduce + RMSNorm reference
(/root/reference/tests/test_fuse_allreduce_rmsnorm_high_throughput.py:16-29, identical in
tests/test_fuse_allreduce_rmsnorm_low_latency.py): bf16 running sum over ranks, + residual (bf16),
RMS in fp32, normalised value rounded to bf16 before the bf16 multiply by gamma.
"""
import torch


def rmsnorm(x, w, eps):
    mean_square = x.float().pow(2).mean(-1, keepdim=True)
    return (x.float() * torch.rsqrt(mean_square + eps)).to(torch.bfloat16) * w.reshape(1, -1)


def allreduce_rmsnorm(input_list, residual, weight, eps):
    s = torch.zeros_like(input_list[0])
    for x in input_list:
        s += x
    out_residual = s + residual
    return out_residual, rmsnorm(out_residual, weight, eps)


def make_inputs(world_size, n, hidden, seed=10001):
    g = torch.Generator().manual_seed(seed)
    n_pad = (n + world_size - 1) // world_size * world_size
    xs = [torch.randn((n_pad, hidden), generator=g).to(torch.bfloat16) for _ in range(world_size)]
    residual = torch.randn((n_pad, hidden), generator=g).to(torch.bfloat16)
    weight = torch.randn((hidden,), generator=g).to(torch.bfloat16)
    return xs, residual, weight, n_pad


def token_slice(n_pad, world_size, rank):
    """Rows owned by `rank` in the high-throughput path (reference test :66-68)."""
    per = n_pad // world_size
    return per * rank, per * (rank + 1)
