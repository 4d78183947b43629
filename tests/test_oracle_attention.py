"""The attention oracles reproduce the reference's own in-test references (golden fixtures made by
tests/golden/make_golden.py from /root/reference/tests). CPU only."""
from pathlib import Path

import numpy as np
import torch

from oracle import attention as oa

G = Path(__file__).resolve().parent / "golden"


def _load_fp8(name):
    z = np.load(G / name)
    B, sq, hkv, hq, D, bs = map(int, z["meta"])
    q = torch.from_numpy(z["q"]).view(torch.float8_e4m3fn)
    kv = torch.from_numpy(z["kvcache"]).view(torch.float8_e4m3fn)
    return z, q, kv, (B, sq, hkv, hq, D, bs)


def test_decode_fp8_oracle_matches_reference_function():
    for name in ("decode_fp8_b2_nhd.npz", "decode_fp8_b5_hnd.npz"):
        z, q, kv, (B, sq, hkv, hq, D, bs) = _load_fp8(name)
        out = oa.decode_fp8_kvpertensor(
            q, kv[:, 0], kv[:, 1], torch.from_numpy(z["block_ids"]),
            torch.from_numpy(z["kv_lens_total"]), torch.from_numpy(z["q_scale"]),
            torch.from_numpy(z["k_scale"]), torch.from_numpy(z["v_scale"]), sq)
        ref = torch.from_numpy(z["out"])
        assert torch.equal(out.float(), ref), (name, (out.float() - ref).abs().max())


def load_kpt(name):
    """golden of the k-per-token decode variant -> (z, dict of tensors as the API takes them)."""
    z = np.load(G / name)
    B, sq, hkv, hq, D, bs = map(int, z["meta"])
    kv = torch.from_numpy(z["kvcache"]).view(torch.float8_e4m3fn)  # [blocks, 2, bs + 2, Hkv, D]
    d = dict(q=torch.from_numpy(z["q"]).view(torch.float8_e4m3fn), kvcache=kv,
             kcache=kv[:, 0, :bs], vcache=kv[:, 1, :bs], k_scale=kv[:, 0, bs:],
             block_ids=torch.from_numpy(z["block_ids"]), kv_lens_total=torch.from_numpy(z["kv_lens_total"]),
             q_scale=torch.from_numpy(z["q_scale"]), v_scale=torch.from_numpy(z["v_scale"]))
    return z, d, (B, sq, hkv, hq, D, bs), int(z["layout"][0])


def test_decode_fp8_kpertoken_oracle_matches_reference_function():
    for name in ("decode_fp8_kpt_b3_nhd.npz", "decode_fp8_kpt_b4_hnd.npz"):
        z, d, (B, sq, hkv, hq, D, bs), _ = load_kpt(name)
        out = oa.decode_fp8_kpertoken(d["q"], d["kcache"], d["vcache"], d["block_ids"],
                                      d["kv_lens_total"], d["q_scale"], d["k_scale"], d["v_scale"], sq)
        ref = torch.from_numpy(z["out"])
        assert torch.equal(out.float(), ref), (name, (out.float() - ref).abs().max())


def test_decode_bf16_config_c1_cpu_plumbing():
    """BASELINE config 0: bs=2 h=4 d=64 seq<=128 bf16 decode through the torch CPU reference path."""
    z = np.load(G / "decode_bf16_c1.npz")
    q = torch.from_numpy(z["q"]).to(torch.bfloat16)
    kv = torch.from_numpy(z["kvcache"]).to(torch.bfloat16)
    out = oa.decode_bf16(q, kv[:, 0], kv[:, 1], torch.from_numpy(z["block_ids"]),
                         torch.from_numpy(z["kv_lens_total"]), 1)
    assert torch.allclose(out.float(), torch.from_numpy(z["out"]), atol=0.016)


def load_bf16_decode(name):
    z = np.load(G / name)
    B, sq, hkv, hq, D, bs = map(int, z["meta"])
    kv = torch.from_numpy(z["kvcache"]).view(torch.bfloat16)
    if int(z["layout"][0]) == 1:
        kv = kv.permute(0, 1, 3, 2, 4).contiguous().permute(0, 1, 3, 2, 4)
    d = dict(q=torch.from_numpy(z["q"]).view(torch.bfloat16), kvcache=kv,
             block_ids=torch.from_numpy(z["block_ids"]),
             kv_lens_total=torch.from_numpy(z["kv_lens_total"]))
    return z, d, (B, sq, hkv, hq, D, bs)


def test_decode_bf16_dim128_oracle_matches_reference_goldens():
    """Head dim 128, page sizes 16 and 64, MTP: bit-equal to the reference's own test function."""
    for name in ("decode_bf16_b3_bs16_nhd.npz", "decode_bf16_b4_bs64_hnd.npz"):
        z, d, (B, sq, hkv, hq, D, bs) = load_bf16_decode(name)
        out = oa.decode_bf16(d["q"], d["kvcache"][:, 0], d["kvcache"][:, 1], d["block_ids"],
                             d["kv_lens_total"], sq)
        ref = torch.from_numpy(z["out"])
        assert torch.equal(out.float(), ref), (name, (out.float() - ref).abs().max())


def test_input_builder_respects_zero_tail_contract():
    d = oa.make_decode_fp8_inputs(3, 2, [5, 64, 70], 2, 8, seed=3)
    kv = d["kvcache"].view(torch.uint8)
    for i, L in enumerate([5, 64, 70]):
        nb = (L + 63) // 64
        last = int(d["block_ids"][i, nb - 1])
        tail = L % 64
        if tail:
            assert int(kv[last, :, tail:].sum()) == 0
