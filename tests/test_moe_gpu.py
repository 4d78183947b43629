"""GPU parity tests: grouped FP8 GEMM and the FusedMoE pipeline vs the CPU oracle.
Tolerances are the reference's (tests/test_fuse_moe_blockwise.py:350, test_fuse_moe_pertensor.py:223,
test_group_gemm_blockwise.py:84, test_group_gemm_pertensor.py:77); routing integers are bit-exact."""
from pathlib import Path

import numpy as np
import pytest
import torch

from oracle import moe as om

pytestmark = pytest.mark.gpu
G = Path(__file__).resolve().parent / "golden"


def _cuda(d):
    return {k: (v.cuda() if isinstance(v, torch.Tensor) else v) for k, v in d.items()}


def _close(my, gt, rtol, atol, tag="", outlier_ppm=0.0, outlier_cap=0.0):
    """allclose at the reference tolerance. `outlier_ppm` > 0 tolerates that fraction of elements
    beyond it (each still within `outlier_cap`): the FP8 re-quantisation between the two GEMMs turns
    fp32 summation-order differences (tensor core vs CPU) into rare one-code flips."""
    my = my.float().cpu()
    gt = gt.float().cpu()
    assert torch.isfinite(my).all(), f"{tag}: non-finite"
    ok = torch.allclose(my, gt, rtol=rtol, atol=atol)
    if not ok:
        err = (my - gt).abs()
        i = int(err.argmax())
        bad = int((err > atol + rtol * gt.abs()).sum())
        if bad <= outlier_ppm * 1e-6 * my.numel() and float(err.max()) <= outlier_cap:
            return
        raise AssertionError(f"{tag}: max err {err.max():.4f} at {i} (my {my.flatten()[i]:.4f} "
                             f"gt {gt.flatten()[i]:.4f}); {bad}/{my.numel()} outside tolerance")


# ---------------------------------------------------------------------------------------------
def _blockwise_gemm_inputs(G_, rows_per_group, n, k, avg, seed=41):
    g = torch.Generator().manual_seed(seed)
    seqlens = torch.tensor(rows_per_group, dtype=torch.int32)
    cu = torch.zeros(G_ + 1, dtype=torch.int32)
    cu[1:] = torch.cumsum(seqlens, 0)
    m = int(cu[-1])
    x = (torch.randn((m, k), generator=g) / 10).to(torch.float8_e4m3fn)
    w = (torch.randn((G_, n, k), generator=g) / 10).to(torch.float8_e4m3fn)
    xs_rows = torch.randn((m, k // 128), generator=g)
    ws = torch.randn((G_, n // 128, (k // 128 + 3) // 4 * 4), generator=g)
    # transposed, tile-padded activation-scale layout
    from hpc.fuse_moe import _aligned_size
    tile = _aligned_size(avg)
    pads = [(r + tile - 1) // tile * tile for r in rows_per_group]
    m_pad = sum(pads) + tile
    xs_t = torch.zeros((k // 128, m_pad))
    col = 0
    for gi, r in enumerate(rows_per_group):
        s = int(cu[gi])
        xs_t[:, col:col + r] = xs_rows[s:s + r].t()
        col += pads[gi]
    return x, w, seqlens, cu, xs_rows, xs_t, ws


@pytest.mark.parametrize("rows", [[30] * 128, [0, 1, 127, 128, 129, 300, 0, 64], [260, 250] * 4])
@pytest.mark.parametrize("n,k", [(1024, 4096), (512, 512), (384, 256)])
def test_group_gemm_blockwise(hpc, rows, n, k):
    G_ = len(rows)
    avg = max(1, sum(rows) // G_)
    x, w, seqlens, cu, xs_rows, xs_t, ws = _blockwise_gemm_inputs(G_, rows, n, k, avg)
    my = hpc.group_gemm_blockwise_fp8(x.cuda(), w.cuda(), seqlens.cuda(), cu.cuda(), xs_t.cuda(),
                                      ws.cuda(), num_seq_per_group_avg=avg)
    gt = om.group_gemm_blockwise(x, w, seqlens, cu, xs_rows, ws)
    _close(my, gt, 0.01, 0.05, f"blockwise gemm rows={rows[:4]} n={n} k={k}")


def test_group_gemm_blockwise_reference_shape(hpc):
    """reference tests/test_group_gemm_blockwise.py:50-84: G=128, m=30 (pad 32), n=1024, k=4096,
    compared against that test's own (bf16-scaled) reference at its tolerance."""
    G_, actual_m, m_pad_g, n, k = 128, 30, 32, 1024, 4096
    g = torch.Generator().manual_seed(41)
    seqlens = torch.full((G_,), actual_m, dtype=torch.int32)
    cu = torch.zeros(G_ + 1, dtype=torch.int32)
    cu[1:] = torch.cumsum(seqlens, 0)
    x = (torch.randn((G_ * actual_m, k), generator=g) / 10).to(torch.float8_e4m3fn)
    w = (torch.randn((G_, n, k), generator=g) / 10).to(torch.float8_e4m3fn)
    xs_t = torch.randn((k // 128, m_pad_g * G_), generator=g)
    ws = torch.randn((G_, n // 128, k // 128), generator=g)
    my = hpc.group_gemm_blockwise_fp8(x.cuda(), w.cuda(), seqlens.cuda(), cu.cuda(), xs_t.cuda(),
                                      ws.cuda(), num_seq_per_group_avg=actual_m)
    gt = om.group_gemm_blockwise_standalone(x, w, seqlens, cu, xs_t, ws, m_pad_g)
    _close(my, gt, 0.08, 0.1, "reference-shape blockwise gemm")


@pytest.mark.parametrize("m_per", [8, 64, 200, 512])
def test_group_gemm_pertensor(hpc, m_per):
    """reference tests/test_group_gemm_pertensor.py:47-77 (G=8, n=4096, k=7168)."""
    G_, n, k = 8, 4096, 7168
    g = torch.Generator().manual_seed(41)
    seqlens = torch.full((G_,), m_per, dtype=torch.int32)
    cu = torch.zeros(G_ + 1, dtype=torch.int32)
    cu[1:] = torch.cumsum(seqlens, 0)
    x = torch.randn((G_ * m_per, k), generator=g).to(torch.float8_e4m3fn)
    w = torch.randn((G_, n, k), generator=g).to(torch.float8_e4m3fn)
    ys = torch.rand((G_,), generator=g) * 0.01
    my = hpc.group_gemm_pertensor_fp8(x.cuda(), w.cuda(), seqlens.cuda(), cu.cuda(), ys.cuda(),
                                      num_seq_per_group_avg=m_per)
    gt = om.group_gemm_pertensor(x, w, cu, ys)
    _close(my, gt, 0.08, 0.01, f"pertensor gemm m={m_per}")


def test_reformat_x_scale(hpc):
    """reference tests/test_group_gemm_blockwise.py:87-148."""
    G_, actual_m, m, k = 256, 30, 1280, 4096
    xscale = torch.rand((m * G_, k // 128))
    seqlens = torch.full((G_,), actual_m, dtype=torch.int32)
    cu = torch.arange(0, G_ + 1, dtype=torch.int32) * m
    out = torch.zeros((k // 128, m * G_), device="cuda")
    out = hpc.reformat_x_scale(xscale.cuda(), seqlens.cuda(), cu.cuda(), actual_m, out).cpu()
    tile = 32
    col = 0
    for gi in range(G_):
        ref = xscale[gi * m:gi * m + actual_m].t()
        assert torch.allclose(out[:, col:col + actual_m], ref, rtol=1e-5, atol=1e-5)
        col += (actual_m + tile - 1) // tile * tile


# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("num_tokens", [128, 1024, 4096])
@pytest.mark.parametrize("intermediate_size", [512, 256])
@pytest.mark.parametrize("rank_ep,size_ep", [(0, 1), (1, 4), (0, 8)])
@pytest.mark.parametrize("has_shared_output", [False, True])
def test_fuse_moe_blockwise(hpc, num_tokens, intermediate_size, rank_ep, size_ep, has_shared_output):
    """Grid of reference tests/test_fuse_moe_blockwise.py:265-272 (k=8, H=512, E=128)."""
    d = om.make_moe_blockwise_inputs(num_tokens, 8, 512, intermediate_size, 128, size_ep,
                                     has_shared_output, seed=41)
    c = _cuda(d)
    my = hpc.fuse_moe_blockwise_fp8(c["x"], c["x_scale"], c["gate_up_weight"],
                                    c["gate_up_weight_scale"], c["down_weight"],
                                    c["down_weight_scale"], c["topk_ids"], c["topk_scale"], rank_ep,
                                    128, c["shared_output"])
    gt = om.fuse_moe_blockwise(d["x"], d["x_scale"], d["gate_up_weight"], d["gate_up_weight_scale"],
                               d["down_weight"], d["down_weight_scale"], d["topk_ids"],
                               d["topk_scale"], rank_ep, d["shared_output"])
    _close(my, gt, 0.01, 0.01, f"moe blockwise T={num_tokens} I={intermediate_size} ep={rank_ep}/{size_ep}",
           outlier_ppm=5, outlier_cap=0.05)


def test_fuse_moe_blockwise_golden(hpc):
    from test_oracle_moe import load_blockwise

    for name in ("moe_blockwise_a.npz", "moe_blockwise_b.npz"):
        z, d, (T, K, H, I, E_total, size_ep, rank_ep) = load_blockwise(name)
        c = _cuda(d)
        my = hpc.fuse_moe_blockwise(c["x"], c["x_scale"], c["gate_up_weight"],
                                    c["gate_up_weight_scale"], c["down_weight"],
                                    c["down_weight_scale"], c["topk_ids"], c["topk_scale"], rank_ep,
                                    E_total, c["shared_output"])
        _close(my, torch.from_numpy(z["out"]), 0.01, 0.01, name)


def test_routing_bit_exact(hpc):
    """counts / cumsums / positions are integers: bit-exact vs the reference's gather order."""
    T, K, H, E_total, size_ep, rank_ep = 1000, 8, 256, 64, 2, 1
    g = torch.Generator().manual_seed(3)
    ids = torch.multinomial(torch.ones((T, E_total)), K, generator=g).to(torch.int32)
    x = torch.randn((T, H), generator=g).to(torch.float8_e4m3fn)
    E = E_total // size_ep
    out = hpc.count_and_gather(x.cuda(), ids.cuda(), E, rank_ep, 256, T * K // E_total)
    gathered, _, _, _, topk_pos, seqlens, cu_seqlens, tiles, cu_tiles = out
    y, _, pos, counts, cu = om.gather_expert_inputs(x, None, ids, E, rank_ep)
    assert torch.equal(topk_pos.cpu(), pos)
    assert torch.equal(seqlens.cpu(), counts) and torch.equal(cu_seqlens.cpu(), cu)
    n = int(cu[-1])
    assert torch.equal(gathered.cpu().view(torch.uint8)[:n], y.view(torch.uint8)[:n])


@pytest.mark.parametrize("T,H,I,E,use_bf16_mul", [(128, 512, 512, 128, True), (128, 4096, 192, 192, True),
                                                  (128, 4096, 192, 192, False), (777, 1024, 1536, 16, True)])
def test_fuse_moe_pertensor(hpc, T, H, I, E, use_bf16_mul):
    """reference tests/test_fuse_moe_pertensor.py:154-223 and tests/test_fuse_moe_cp_async.py:147-242."""
    g = torch.Generator().manual_seed(41)
    K = 8
    ids = torch.multinomial(torch.ones((T, E)), K, generator=g).to(torch.int32)
    ids, _ = torch.sort(ids, dim=1)
    ts = torch.rand((T, K), generator=g)
    x = torch.randn((T, H), generator=g).to(torch.float8_e4m3fn)
    guw = torch.randn((E, 2 * I, H), generator=g).to(torch.float8_e4m3fn)
    dw = torch.randn((E, H, I), generator=g).to(torch.float8_e4m3fn)
    gus = torch.rand((E,), generator=g) * 0.01
    ds = torch.rand((E,), generator=g) * 0.01
    acts = torch.rand((1,), generator=g) + 0.5
    my = hpc.fuse_moe(x.cuda(), guw.cuda(), dw.cuda(), gus.cuda(), ds.cuda(), acts.cuda(),
                      ids.cuda(), ts.cuda(), 0, E, use_bf16_mul=use_bf16_mul)
    gt = om.fuse_moe_pertensor(x, guw, dw, gus, ds, acts, ids, ts, 0, None, use_bf16_mul)
    _close(my, gt, 0.08, 0.1, f"moe pertensor T={T} H={H} I={I}")


def test_reduce_op(hpc):
    g = torch.Generator().manual_seed(1)
    x = torch.randn((64, 512), generator=g).to(torch.bfloat16)
    pos = torch.randint(-1, 64, (16, 4), generator=g).to(torch.int32)
    sc = torch.rand((16, 4), generator=g)
    sh = torch.randn((16, 512), generator=g).to(torch.bfloat16)
    my = hpc.reduce(x.cuda(), pos.cuda(), sc.cuda(), sh.cuda())
    gt = om.reduce(x, pos, sc, sh)
    _close(my, gt, 0.01, 0.01, "reduce")


def test_fuse_moe_is_deterministic(hpc):
    d = _cuda(om.make_moe_blockwise_inputs(512, 8, 512, 256, 128, 1, False, seed=9))
    a = hpc.fuse_moe_blockwise_fp8(d["x"], d["x_scale"], d["gate_up_weight"], d["gate_up_weight_scale"],
                                   d["down_weight"], d["down_weight_scale"], d["topk_ids"],
                                   d["topk_scale"], 0, 128)
    b = hpc.fuse_moe_blockwise_fp8(d["x"], d["x_scale"], d["gate_up_weight"], d["gate_up_weight_scale"],
                                   d["down_weight"], d["down_weight_scale"], d["topk_ids"],
                                   d["topk_scale"], 0, 128)
    assert torch.equal(a, b)
