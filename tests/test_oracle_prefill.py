"""Prefill oracle reproduces the reference's in-test references (golden fixtures). CPU only."""
from pathlib import Path

import numpy as np
import torch

from oracle import prefill as op

G = Path(__file__).resolve().parent / "golden"


def load(name):
    z = np.load(G / name)
    B, seq, Hq, Hkv, kpt, layout = map(int, z["meta"])
    kv = torch.from_numpy(z["kv"]).view(torch.float8_e4m3fn)
    d = dict(q=torch.from_numpy(z["q"]).view(torch.float8_e4m3fn), kcache=kv[:, 0], vcache=kv[:, 1],
             qscale=torch.from_numpy(z["qscale"]), kscale=torch.from_numpy(z["kscale"]),
             vscale=torch.from_numpy(z["vscale"]), block_ids=torch.from_numpy(z["block_ids"]),
             block_mask=torch.from_numpy(z["mask"]) if z["mask"].size else None,
             cu_seqlens_q=torch.arange(0, B + 1, dtype=torch.int32) * seq,
             seqlens_kv=torch.full((B,), seq, dtype=torch.int32), max_q=seq)
    return z, d, bool(kpt), layout


def test_prefill_oracle_matches_reference_functions():
    for name in ("prefill_kvpt.npz", "prefill_kpertoken.npz"):
        z, d, kpt, _ = load(name)
        y = op.blocksparse_prefill(d["q"], d["kcache"], d["vcache"], d["qscale"], d["kscale"],
                                   d["vscale"], d["cu_seqlens_q"], d["seqlens_kv"], d["block_ids"],
                                   d["block_mask"], kpt)
        ref = torch.from_numpy(z["out"])
        assert torch.equal(y.float(), ref), (name, (y.float() - ref).abs().max())


def test_mask_always_keeps_the_causal_diagonal():
    m = op.generate_block_sparse_mask(2, 4, 8, 8, 0.9, True, torch.Generator().manual_seed(1))
    for r in range(8):
        assert bool(m[:, :, r, r].all())
        assert not bool(m[:, :, r, r + 1:].any())
