"""GPU parity tests of the stand-alone activation / quantisation ops and the cp.async-style grouped
GEMM entry points (reference tests/test_act.py:42-59, tests/test_group_gemm_cp_async.py:55-111)."""
from pathlib import Path

import numpy as np
import pytest
import torch

from oracle import act as oact

pytestmark = pytest.mark.gpu
G = Path(__file__).resolve().parent / "golden"


def _fp8_mismatch(my, gt):
    """(#elements whose e4m3 code differs, max |difference| in e4m3 code steps)."""
    a = my.view(torch.uint8).cpu().to(torch.int32)
    b = gt.view(torch.uint8).cpu().to(torch.int32)
    diff = a != b
    # codes are sign-magnitude monotone: compare magnitudes of same-sign values
    step = ((a & 0x7F) - (b & 0x7F)).abs()[diff]
    return int(diff.sum()), int(step.max()) if step.numel() else 0


def test_act_mul_and_quant_golden(hpc):
    z = np.load(G / "act_a.npz")
    gate_up = torch.from_numpy(z["gate_up"]).to(torch.bfloat16).cuda()
    scale = torch.from_numpy(z["scale"]).cuda()
    out = hpc.act_mul_and_quant(gate_up, scale)
    gt = torch.from_numpy(z["gt"]).view(torch.float8_e4m3fn)
    bad, step = _fp8_mismatch(out, gt)
    # CUDA expf is within 2 ulp of the CPU's: a bf16 rounding of silu(gate) may flip, rarely
    assert bad <= 2 and step <= 1, (bad, step)


@pytest.mark.parametrize("rows,half_cols", [(1, 64), (77, 256), (4096, 4608), (300, 1000 * 8)])
@pytest.mark.parametrize("use_bf16_mul", [True, False])
@pytest.mark.parametrize("use_output", [True, False])
def test_act_mul_and_quant(hpc, rows, half_cols, use_bf16_mul, use_output):
    gate_up, scale = oact.make_act_inputs(rows, half_cols, seed=rows + half_cols)
    gt = oact.act_mul_and_quant(gate_up, scale, use_bf16_mul)
    gu, sc = gate_up.cuda(), scale.cuda()
    if use_output:
        out = torch.empty((rows, half_cols), dtype=torch.float8_e4m3fn, device="cuda")
        ret = hpc.act_mul_and_quant(gu, sc, use_bf16_mul, out)
        assert ret.data_ptr() == out.data_ptr()
    else:
        out = hpc.act_mul_and_quant(gu, sc, use_bf16_mul)
    assert out.shape == (rows, half_cols) and out.dtype == torch.float8_e4m3fn
    bad, step = _fp8_mismatch(out, gt)
    # bit-exact except where a <=2-ulp expf difference crosses a rounding boundary: <= 20 ppm of
    # the elements, never by more than one e4m3 code
    assert bad <= max(2, 20e-6 * out.numel()) and step <= 1, (bad, step, out.numel())


@pytest.mark.parametrize("dtype", [torch.float32, torch.float16, torch.bfloat16])
@pytest.mark.parametrize("shape", [(1,), (7,), (1000, 33), (4096, 4096)])
def test_scaled_fp8_quant(hpc, dtype, shape):
    g = torch.Generator().manual_seed(sum(shape))
    x = (torch.randn(shape, generator=g) * 3).to(dtype)
    scale = torch.tensor([0.37], dtype=torch.float32)
    gt = oact.scaled_fp8_quant(x, scale)
    out, sc = hpc.scaled_fp8_quant(x.cuda(), scale.cuda())
    assert sc.item() == pytest.approx(0.37)
    assert torch.equal(out.view(torch.uint8).cpu(), gt.view(torch.uint8))  # bit-exact
    pre = torch.empty(shape, dtype=torch.float8_e4m3fn, device="cuda")
    out2, _ = hpc.scaled_fp8_quant(x.cuda(), scale.cuda(), pre)
    assert out2.data_ptr() == pre.data_ptr() and torch.equal(out2.view(torch.uint8), out.view(torch.uint8))


def test_act_errors(hpc):
    with pytest.raises(RuntimeError):
        hpc.scaled_fp8_quant(torch.randn(8, device="cuda"), None)
    with pytest.raises(RuntimeError):
        hpc.act_mul_and_quant(torch.randn(4, 64, device="cuda"), torch.ones(1, device="cuda"))
    with pytest.raises(RuntimeError):
        hpc.act_mul_and_quant(torch.randn(4, 24, device="cuda").bfloat16(), torch.ones(1, device="cuda"))


# reference tests/test_group_gemm_cp_async.py:55-111 (ground truth there = hpc.group_gemm_fp8)
@pytest.mark.parametrize("shape", [(192, 256, 4096, 256), (192, 42, 4096, 192), (192, 8, 4096, 256),
                                   (16, 100, 512, 128)])
@pytest.mark.parametrize("scatter", [False, True])
@pytest.mark.parametrize("use_task_map", [False, True])
def test_group_gemm_cp_async(hpc, shape, scatter, use_task_map):
    num_group, actual_m, n, k = shape
    g = torch.Generator().manual_seed(10086)
    seqlens = torch.full((num_group,), actual_m, dtype=torch.int32)
    total = int(seqlens.sum())
    x_pool = torch.randn((total, k), generator=g).to(torch.float8_e4m3fn)
    w = torch.randn((num_group, n, k), generator=g).to(torch.float8_e4m3fn)
    scale = torch.rand((num_group,), generator=g) + 0.5
    row_indices = torch.randperm(total, generator=g).to(torch.int32)
    cu = torch.cat([torch.zeros(1, dtype=torch.int32), torch.cumsum(seqlens, 0).to(torch.int32)])
    tiles = (seqlens + 63) // 64
    cu_tiles = torch.cat([torch.zeros(1, dtype=torch.int32), torch.cumsum(tiles, 0).to(torch.int32)])
    x_compact = x_pool[row_indices.long()]
    # oracle: per-group fp32 matmul of the compact problem (reference tests/test_group_gemm_pertensor.py:20-44)
    gt = torch.empty((total, n), dtype=torch.float32)
    for i in range(num_group):
        a = x_compact[cu[i]:cu[i + 1]].float()
        gt[cu[i]:cu[i + 1]] = (a @ w[i].float().t()) * scale[i]
    dev = lambda t: t.cuda()  # noqa: E731
    if scatter:
        out = torch.ops.hpc.group_gemm_fp8_scatter_cp_async(
            dev(x_pool), dev(w), dev(scale), dev(row_indices), dev(seqlens), dev(cu), dev(tiles),
            dev(cu_tiles), use_task_map)
    else:
        out = torch.ops.hpc.group_gemm_fp8_cp_async(
            dev(x_compact), dev(w), dev(scale), dev(seqlens), dev(cu), dev(tiles), dev(cu_tiles),
            use_task_map)
    assert out.dtype == torch.bfloat16 and out.shape == (total, n)
    assert torch.allclose(out.float().cpu(), gt.bfloat16().float(), rtol=0.08, atol=1)  # reference :111
    # and tightly: bf16 rounding of an fp32-accumulated result
    assert torch.allclose(out.float().cpu(), gt, rtol=1e-2, atol=0.25)


def test_version_ops(hpc):
    assert torch.ops.hpc.version() == hpc.__version__
    assert "sm_100a" in torch.ops.hpc.built_json()
