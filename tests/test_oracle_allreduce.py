"""Allreduce + RMSNorm oracle reproduces the reference's own in-test reference (golden fixture)."""
from pathlib import Path

import numpy as np
import torch

from oracle import allreduce as oar

G = Path(__file__).resolve().parent / "golden"


def test_allreduce_oracle_matches_reference_function():
    z = np.load(G / "allreduce_w4.npz")
    world, n, hidden, seed = map(int, z["meta"])
    xs, residual, weight, n_pad = oar.make_inputs(world, n, hidden, seed)
    res, out = oar.allreduce_rmsnorm(xs, residual, weight, 1e-6)
    assert n_pad == 16 and res.dtype == torch.bfloat16 and out.dtype == torch.bfloat16
    assert np.array_equal(res.float().numpy(), z["out_residual"])  # bit-exact
    assert np.array_equal(out.float().numpy(), z["out"])
    # slices of the high-throughput path cover the padded token range exactly once
    spans = [oar.token_slice(n_pad, world, r) for r in range(world)]
    assert spans[0][0] == 0 and spans[-1][1] == n_pad
    assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
