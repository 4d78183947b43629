"""C2 decode step / kernel timing only (no CPU baseline, no extras): used for on-box A/B of library builds.
    python tools/decode_ab.py"""
import json
import sys
from pathlib import Path

REPO = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(REPO))
sys.path.insert(0, str(REPO / "hpc-ops_b200"))
import torch  # noqa: E402

import hpc  # noqa: E402
from hpc import _ffi, attention as hatt  # noqa: E402
from synth.decode import make_decode_fp8_inputs  # noqa: E402

B, Sq, Hkv, Hq, S, MPL = 64, 1, 8, 32, 8192, 64
d = make_decode_fp8_inputs(B, Sq, [S] * B, Hkv, Hq, seed=41, device="cuda")
kc, vc = d["kvcache"][:, 0], d["kvcache"][:, 1]
tm = hpc.get_attention_decode_task_workspace(B, S, Hkv, MPL)
out = torch.empty((B, Hq, 128), dtype=torch.bfloat16, device="cuda")


def step():
    hpc.assign_attention_decode_task(d["kv_lens_total"], tm, Hkv, Sq, True, MPL)
    hpc.attention_decode_fp8(d["q"], kc, vc, d["block_ids"], d["kv_lens_total"], d["q_scale"], d["k_scale"],
                             d["v_scale"], mtp=0, new_kv_included=True, task_map=tm, output=out)


def timed(fn, n):
    for _ in range(20):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


step()
y, args, keep = hatt._decode_fp8_prepare(d["q"], kc, vc, d["block_ids"], d["kv_lens_total"], d["q_scale"],
                                         d["k_scale"], d["v_scale"], 0, True, 1, True, tm, None, out)
res = {"step_us": timed(step, 1000),
       "kernel_us": timed(lambda: _ffi.lib.hpc_attention_decode_fp8_partial_async(*args), 300)}
res["step_us_2"] = timed(step, 1000)
print(json.dumps(res))
