"""A/B of the rotated bin walk of the fp8 decode kernel (csrc/decode_common.cuh) at the C2 shape,
token-major (NHD) cache. HPC_B200_DECODE_ROTATE / HPC_B200_KV_PROMO are read at every launch, so one
process times all variants on the same box. GPU box only.

    python tools/decode_rotate_ab.py
"""
import json
import os
import sys
from pathlib import Path

REPO = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(REPO))
sys.path.insert(0, str(REPO / "hpc-ops_b200"))
sys.path.insert(0, str(REPO / "tools"))
import torch  # noqa: E402

import hpc  # noqa: E402
from bench_extras import peaks, time_eager  # noqa: E402


VARIANTS = (("front_to_back_promo128", "0", None, "0"), ("rotated_promo256", "1", None, "0"),
            ("rotated_promo256_evict_normal", "1", None, "1"),
            ("rotated_promo256_evict_last", "1", None, "2"),
            ("rotated_promo128", "1", "2", "0"), ("front_to_back_promo256", "0", "3", "0"))


def main():
    dev = "cuda"
    B, ctx, hq, hkv, D, bs = 64, 8192, 32, 8, 128, 64
    nb = ctx // bs
    pk = peaks()
    g = torch.Generator(device=dev).manual_seed(0)
    kv = torch.randn(B * nb, 2, bs, hkv, D, device=dev, generator=g).to(torch.float8_e4m3fn)
    q = torch.randn(B, hq, D, device=dev, generator=g).to(torch.float8_e4m3fn)
    ids = torch.randperm(B * nb, device=dev).to(torch.int32).view(B, nb)
    y = torch.empty(B, hq, D, device=dev, dtype=torch.bfloat16)
    qs = torch.ones(B, hq, device=dev)
    one = torch.ones(1, device=dev)
    out = {}
    for lens_name in ("equal", "ragged"):
        if lens_name == "equal":
            lens = torch.full((B,), ctx, dtype=torch.int32, device=dev)
        else:
            lens = torch.randint(1024, ctx + 1, (B,), generator=g, device=dev, dtype=torch.int32)
        tm = hpc.get_attention_decode_task_workspace(B, ctx, hkv, 64)
        hpc.assign_attention_decode_task(lens, tm, hkv, 1, True, 64)
        fn = lambda: hpc.attention_decode_fp8(q, kv[:, 0], kv[:, 1], ids, lens, qs, one, one,  # noqa: E731
                                              new_kv_included=True, task_map=tm, output=y)
        byts = 2 * int(lens.sum()) * hkv * D
        ref = None
        for rep in range(2):
            for name, rot, promo, pol in VARIANTS:
                os.environ["HPC_B200_DECODE_ROTATE"] = rot
                os.environ["HPC_B200_KV_POLICY"] = pol
                if promo is None:
                    os.environ.pop("HPC_B200_KV_PROMO", None)
                else:
                    os.environ["HPC_B200_KV_PROMO"] = promo
                ms = time_eager(fn, 100)
                yy = y.clone()
                if ref is None:
                    ref = yy
                same = bool(torch.equal(ref, yy))
                out.setdefault(lens_name, {}).setdefault(name, []).append(
                    {"ms": round(ms, 5), "gbs": round(byts / ms / 1e6, 1),
                     "frac_hbm": round(byts / ms / 1e6 / pk["hbm_gbs"], 4), "bit_equal_to_first": same})
    os.environ.pop("HPC_B200_DECODE_ROTATE", None)
    os.environ.pop("HPC_B200_KV_PROMO", None)
    os.environ.pop("HPC_B200_KV_POLICY", None)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
