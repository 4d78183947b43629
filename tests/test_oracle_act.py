"""Activation oracles reproduce the reference's own in-test reference (golden fixture). CPU only."""
from pathlib import Path

import numpy as np
import torch

from oracle import act as oact

G = Path(__file__).resolve().parent / "golden"


def test_act_mul_and_quant_oracle_matches_reference_function():
    z = np.load(G / "act_a.npz")
    gate_up = torch.from_numpy(z["gate_up"]).to(torch.bfloat16)
    scale = torch.from_numpy(z["scale"])
    out = oact.act_mul_and_quant(gate_up, scale, use_bf16_mul=True)
    assert np.array_equal(out.view(torch.uint8).numpy(), z["gt"])  # bit-exact


def test_scaled_fp8_quant_oracle_known_answers():
    x = torch.tensor([0.0, 1.0, -2.0, 448.0, 1000.0, -1000.0, 0.0625, 3.3], dtype=torch.float32)
    q = oact.scaled_fp8_quant(x, torch.tensor([2.0])).float()
    assert q.tolist() == [0.0, 0.5, -1.0, 224.0, 448.0, -448.0, 0.03125, 1.625]
