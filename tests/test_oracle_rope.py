"""The RoPE / QK-norm / KV-store oracle reproduces the reference's own `rope_norm_ref`
(tests/golden/rope_*.npz made by tests/golden/make_golden.py from /root/reference/tests/test_rope.py).
CPU only."""
from pathlib import Path

import numpy as np
import torch

from oracle import rope as orp

G = Path(__file__).resolve().parent / "golden"


def _bf16(a):
    return torch.from_numpy(a.copy()).view(torch.bfloat16)


def load_rope(name):
    z = np.load(G / name)
    d = dict(qkv=_bf16(z["qkv"]), num_seqlen=torch.from_numpy(z["num_seqlen"]),
             q_index=torch.from_numpy(z["q_index"]), kv_indices=torch.from_numpy(z["kv_indices"]),
             kcache=_bf16(z["kcache_in"]), vcache=_bf16(z["vcache_in"]),
             q_norm_w=torch.from_numpy(z["q_norm_w"]), k_norm_w=torch.from_numpy(z["k_norm_w"]),
             cos_sin=torch.from_numpy(z["cos_sin"]))
    out = dict(q=_bf16(z["out_q"]), kcache=_bf16(z["kcache_out"]), vcache=_bf16(z["vcache_out"]))
    num_req, is_prefill, mtp, hq, hkv, policy = map(int, z["meta"])
    return d, out, (num_req, bool(is_prefill), mtp, hq, hkv, policy)


def test_rope_oracle_matches_reference_function():
    for name in ("rope_prefill_p2.npz", "rope_decode_p1.npz"):
        d, out, (_, _, _, _, _, policy) = load_rope(name)
        kc, vc = d["kcache"].clone(), d["vcache"].clone()
        q = orp.rope_norm_store_kv(kc, vc, d["qkv"], d["cos_sin"], d["num_seqlen"], d["q_index"],
                                   d["kv_indices"], d["q_norm_w"], d["k_norm_w"], policy)
        assert torch.equal(q, out["q"]), name
        assert torch.equal(kc, out["kcache"]) and torch.equal(vc, out["vcache"]), name


def test_rope_fp8_oracle_dequantises_to_the_bf16_oracle():
    """fp8 rules (dynamic q scale = amax / 448, static k / v scales): dequantised values agree with
    the bf16 path within the reference's fp8 tolerance (tests/test_rope.py:367 atol 0.5)."""
    d, _, (_, _, _, _, _, policy) = load_rope("rope_decode_p1.npz")
    kc, vc = d["kcache"].clone(), d["vcache"].clone()
    ref_q = orp.rope_norm_store_kv(kc, vc, d["qkv"], d["cos_sin"], d["num_seqlen"], d["q_index"],
                                   d["kv_indices"], d["q_norm_w"], d["k_norm_w"], policy)
    k8, v8 = d["kcache"].to(torch.float8_e4m3fn), d["vcache"].to(torch.float8_e4m3fn)
    ks, vs = torch.tensor([0.1]), torch.tensor([0.1])
    q8, qs = orp.rope_norm_store_kv_fp8(k8, v8, d["qkv"], d["cos_sin"], d["num_seqlen"], d["q_index"],
                                        d["kv_indices"], ks, vs, 1, None, 448.0, d["q_norm_w"],
                                        d["k_norm_w"], policy)
    assert torch.allclose(q8.float() * qs[..., None], ref_q.float(), atol=0.5)
    blk = kc.shape[1]
    r, p = 0, int(d["num_seqlen"][0]) - 1
    cb = int(d["kv_indices"][r, p // blk])
    assert torch.allclose(k8[cb, p % blk].float() * 0.1, kc[cb, p % blk].float(), atol=0.1, rtol=0.07)
