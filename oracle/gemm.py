"""ORACLE — test infrastructure only. Route GEMM reference
(/root/reference/tests/test_gemm_bf16xfp32.py:28-38): fp32 matmul against the fp32 weight."""
import torch


def make_inputs(m, n, k, seed=10086, device="cpu"):
    g = torch.Generator(device=device).manual_seed(seed)
    x = torch.randn((m, k), generator=g, device=device).to(torch.bfloat16)
    w = torch.randn((n, k), generator=g, device=device)
    scale = 1 / 256
    w_high = w.to(torch.bfloat16)
    w_low = ((w - w_high.float()) / scale).to(torch.bfloat16)
    return x, w, w_high, w_low, scale


def gemm_fp32(x, w):
    return torch.matmul(x.float(), w.t())


def gemm_split_exact(x, w_high, w_low, scale):
    """What the kernel computes, in float64: X @ (w_high + scale * w_low)^T."""
    return x.double() @ (w_high.double() + scale * w_low.double()).t()
