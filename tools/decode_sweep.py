"""Time the decode attention kernel alone (CUDA events) over a few variants. GPU box only.

    python tools/decode_sweep.py [--layout NHD|HND] [--batch 64] [--seq 8192] [--iters 200]
"""
import argparse
import json
import os
import sys
from pathlib import Path

REPO = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(REPO))
sys.path.insert(0, str(REPO / "hpc-ops_b200"))
import torch  # noqa: E402

import hpc  # noqa: E402
from hpc import _ffi  # noqa: E402
from hpc import attention as hatt  # noqa: E402
from synth import decode as oa  # noqa: E402


def time_partial(d, B, Sq, Hkv, S, mpl, iters):
    kc, vc = d["kvcache"][:, 0], d["kvcache"][:, 1]
    tm = hpc.get_attention_decode_task_workspace(B, S, Hkv, mpl)
    hpc.assign_attention_decode_task(d["kv_lens_total"], tm, Hkv, Sq, True, mpl)
    out = torch.empty((B * Sq, d["q"].shape[1], 128), dtype=torch.bfloat16, device="cuda")
    y, args, keep = hatt._decode_fp8_prepare(d["q"], kc, vc, d["block_ids"], d["kv_lens_total"],
                                             d["q_scale"], d["k_scale"], d["v_scale"], Sq - 1, True,
                                             1, True, tm, None, out)
    for _ in range(5):
        _ffi.check(_ffi.lib.hpc_attention_decode_fp8_partial_async(*args))
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        _ffi.lib.hpc_attention_decode_fp8_partial_async(*args)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--seq", type=int, default=8192)
    ap.add_argument("--hkv", type=int, default=8)
    ap.add_argument("--hq", type=int, default=32)
    ap.add_argument("--sq", type=int, default=1)
    ap.add_argument("--iters", type=int, default=200)
    ap.add_argument("--mpl", type=int, default=64)
    a = ap.parse_args()
    res = []
    for layout in ("NHD", "HND"):
        d = oa.make_decode_fp8_inputs(a.batch, a.sq, [a.seq] * a.batch, a.hkv, a.hq, seed=41,
                                      layout=layout, device="cuda")
        bytes_ = a.batch * a.hkv * a.seq * 256
        for promo in ("", "0", "2", "3"):
            if promo:
                os.environ["HPC_B200_KV_PROMO"] = promo
            else:
                os.environ.pop("HPC_B200_KV_PROMO", None)
            ms = time_partial(d, a.batch, a.sq, a.hkv, a.seq, a.mpl, a.iters)
            r = dict(layout=layout, promo=promo or "default", ms=ms, gbs=bytes_ / ms / 1e6)
            res.append(r)
            print(json.dumps(r))
        del d
        torch.cuda.empty_cache()
