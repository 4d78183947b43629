"""ORACLE — test infrastructure only (never imported by the product).

Stand-alone activation / quantisation references:
  * act_mul_and_quant: /root/reference/tests/test_act.py:20-28 (`_act_mul_and_quant`, the bf16-multiply
    form the reference kernel takes with use_bf16_mul=True) and the fp32-multiply form of
    /root/reference/src/activation/activation.cu:41-47 for use_bf16_mul=False.
  * scaled_fp8_quant: /root/reference/src/activation/activation.cu:461-500 (in * (1/scale) -> e4m3) ==
    /root/reference/benchmark/fused_moe/backends/base.py:64-67 up to the rounding of 1/scale.
Pinned by tests/golden/act_a.npz (the reference's own `_act_mul_and_quant` executed on CPU).
"""
import torch


def act_mul_and_quant(gate_up, scale, use_bf16_mul=True):
    def silu(x):
        return x / (1 + (-x).exp())

    gate, up = torch.chunk(gate_up.float(), 2, dim=-1)
    if use_bf16_mul:
        out = (silu(gate).to(torch.bfloat16) * up.to(torch.bfloat16)).to(torch.float32) * scale
    else:
        out = silu(gate) * up * scale
    return out.to(torch.float8_e4m3fn)


def scaled_fp8_quant(x, scale):
    inv = 1.0 / scale.float()
    return (x.float() * inv).clamp(-448.0, 448.0).to(torch.float8_e4m3fn)


def make_act_inputs(rows, half_cols, seed=41, device="cpu"):
    g = torch.Generator(device=device).manual_seed(seed)
    gate_up = torch.randn((rows, 2 * half_cols), generator=g, device=device).to(torch.bfloat16)
    scale = torch.rand((1,), generator=g, device=device) + 1.0
    return gate_up, scale
