"""MoE oracles reproduce the reference's own in-test references (golden fixtures). CPU only."""
from pathlib import Path

import numpy as np
import torch

from oracle import moe as om

G = Path(__file__).resolve().parent / "golden"


def _f8(a):
    return torch.from_numpy(a).view(torch.float8_e4m3fn)


def load_blockwise(name):
    z = np.load(G / name)
    T, K, H, I, E_total, size_ep, rank_ep, shared = map(int, z["meta"])
    d = dict(x=_f8(z["x"]), x_scale=torch.from_numpy(z["x_scale"]), gate_up_weight=_f8(z["guw"]),
             gate_up_weight_scale=torch.from_numpy(z["guws"]), down_weight=_f8(z["dw"]),
             down_weight_scale=torch.from_numpy(z["dws"]), topk_ids=torch.from_numpy(z["topk_ids"]),
             topk_scale=torch.from_numpy(z["topk_scale"]),
             shared_output=torch.from_numpy(z["shared"]).to(torch.bfloat16) if shared else None)
    return z, d, (T, K, H, I, E_total, size_ep, rank_ep)


def test_blockwise_oracle_matches_reference_function():
    for name in ("moe_blockwise_a.npz", "moe_blockwise_b.npz"):
        z, d, (T, K, H, I, E_total, size_ep, rank_ep) = load_blockwise(name)
        E = E_total // size_ep
        _, _, pos, counts, cu = om.gather_expert_inputs(d["x"], d["x_scale"], d["topk_ids"], E, rank_ep)
        assert np.array_equal(pos.numpy(), z["topk_pos"])
        assert np.array_equal(counts.numpy(), z["counts"]) and np.array_equal(cu.numpy(), z["cu"])
        y = om.fuse_moe_blockwise(d["x"], d["x_scale"], d["gate_up_weight"],
                                  d["gate_up_weight_scale"], d["down_weight"],
                                  d["down_weight_scale"], d["topk_ids"], d["topk_scale"], rank_ep,
                                  d["shared_output"])
        ref = torch.from_numpy(z["out"])
        assert torch.allclose(y.float(), ref, rtol=0.01, atol=0.01), (name, (y.float() - ref).abs().max())


def test_pertensor_oracle_matches_reference_function():
    z = np.load(G / "moe_pertensor_a.npz")
    T, K, H, I, E_total, size_ep, rank_ep = map(int, z["meta"])
    y = om.fuse_moe_pertensor(_f8(z["x"]), _f8(z["guw"]), _f8(z["dw"]), torch.from_numpy(z["gus"]),
                              torch.from_numpy(z["ds"]), torch.from_numpy(z["acts"]),
                              torch.from_numpy(z["topk_ids"]), torch.from_numpy(z["topk_scale"]),
                              rank_ep)
    ref = torch.from_numpy(z["out"])
    assert torch.allclose(y.float(), ref, rtol=0.08, atol=0.1), (y.float() - ref).abs().max()


def test_group_gemm_oracles_match_reference_functions():
    """Stand-alone grouped GEMM oracles vs the reference's naive functions
    (tests/test_group_gemm_pertensor.py:20-44, tests/test_group_gemm_blockwise.py:20-47)."""
    z = np.load(G / "group_gemm_a.npz")
    x, w = _f8(z["x"]), _f8(z["w"])
    seqlens, cu = torch.from_numpy(z["seqlens"]), torch.from_numpy(z["cu"])
    per = int(z["per"][0])
    Gn = w.shape[0]
    # per-tensor: the reference passes `scale` as scale_a and scale_b of _scaled_mm -> y * scale^2
    s2 = float(z["scale"]) ** 2
    cu_full = torch.cat([cu, torch.tensor([Gn * per], dtype=torch.int32)])
    y = om.group_gemm_pertensor(x, w, cu_full, torch.full((Gn,), s2))
    y_bw = om.group_gemm_blockwise_standalone(x, w, seqlens, cu, torch.from_numpy(z["xscale"]),
                                              torch.from_numpy(z["wscale"]), per)
    ref_pt, ref_bw = torch.from_numpy(z["y_pertensor"]), torch.from_numpy(z["y_blockwise"])
    for g in range(Gn):
        s, c = int(cu[g]), int(seqlens[g])
        assert torch.allclose(y[s:s + c].float(), ref_pt[s:s + c], rtol=1e-2, atol=1e-2), g
        # the reference function multiplies in bf16; executed on the CPU (as the fixture was) that
        # GEMM accumulates in bf16 blocks, i.e. it is noisier than the oracle, which accumulates the
        # same bf16-rounded operands in fp32. rtol is the reference test's (test_group_gemm_blockwise.py:84),
        # atol covers the fixture's own accumulation noise (outputs reach |82|, bf16 ulp there = 0.5)
        assert torch.allclose(y_bw[s:s + c].float(), ref_bw[s:s + c], rtol=0.08, atol=0.3), g
        assert (ref_pt[s + c:s + per] == 0).all() and (ref_bw[s + c:s + per] == 0).all()
