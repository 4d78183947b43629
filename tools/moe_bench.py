"""Time the FusedMoE blockwise pipeline at BASELINE config C3 (and its kernels). GPU box only.

    python tools/moe_bench.py [--tokens 4096] [--experts 128] [--topk 8] [--hidden 4096] [--inter 14336]
"""
import argparse
import json
import sys
from pathlib import Path

REPO = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(REPO))
sys.path.insert(0, str(REPO / "hpc-ops_b200"))
import torch  # noqa: E402

import hpc  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--tokens", type=int, default=4096)
    ap.add_argument("--experts", type=int, default=128)
    ap.add_argument("--topk", type=int, default=8)
    ap.add_argument("--hidden", type=int, default=4096)
    ap.add_argument("--inter", type=int, default=14336)
    ap.add_argument("--iters", type=int, default=10)
    a = ap.parse_args()
    dev = "cuda"
    g = torch.Generator(device=dev).manual_seed(41)
    T, E, K, H, I = a.tokens, a.experts, a.topk, a.hidden, a.inter
    ids = torch.multinomial(torch.ones((T, E), device=dev), K, generator=g).to(torch.int32)
    ids, _ = torch.sort(ids, dim=1)
    ts = torch.rand((T, K), generator=g, device=dev)
    ts = ts / ts.sum(1, keepdim=True)
    x = (torch.randn((T, H), generator=g, device=dev) / 100).to(torch.float8_e4m3fn)
    xs = torch.rand((T, H // 128), generator=g, device=dev) + 0.5

    def w8(shape):
        w = torch.empty(shape, dtype=torch.float8_e4m3fn, device=dev)
        for e in range(shape[0]):
            w[e] = torch.randn(shape[1:], generator=g, device=dev).to(torch.float8_e4m3fn)
        return w

    guw = w8((E, 2 * I, H))
    dw = w8((E, H, I))
    guws = torch.rand((E, 2 * I // 128, (H // 128 + 3) // 4 * 4), generator=g, device=dev) * 0.02
    dws = torch.rand((E, H // 128, (I // 128 + 3) // 4 * 4), generator=g, device=dev) * 0.02

    def run():
        return hpc.fuse_moe_blockwise_fp8(x, xs, guw, guws, dw, dws, ids, ts, 0, E)

    y = run()
    torch.cuda.synchronize()
    import os
    if not os.environ.get("HPC_B200_MOE_DEBUG"):
        assert torch.isfinite(y.float()).all()
    for _ in range(2):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.iters):
        run()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.iters
    flops = 2 * T * K * (2 * I * H + H * I)
    wbytes = E * (2 * I * H + H * I)
    out = {"ms": ms, "tok_per_s": T / ms * 1e3, "tflops": flops / ms / 1e9,
           "frac_fp8_4500": flops / ms / 1e9 / 4500, "weight_gbs": wbytes / ms / 1e6, "cfg": vars(a)}
    import os
    if int(os.environ.get("HPC_B200_MOE_DEBUG", "0")) & 8:
        # per-CTA counters of the MMA threads (last Gate-Up and Down launches)
        import ctypes
        import numpy as np
        buf = np.zeros((2, 256, 16), dtype=np.int64)
        hpc._ffi.lib.hpc_group_gemm_debug_counters.restype = ctypes.c_int
        hpc._ffi.lib.hpc_group_gemm_debug_counters.argtypes = [ctypes.c_void_p]
        hpc._ffi.check(hpc._ffi.lib.hpc_group_gemm_debug_counters(buf.ctypes.data), "debug counters")
        for name, b in (("gate_up", buf[0]), ("down", buf[1])):
            b = b[b[:, 1] > 0].astype(np.float64)
            kb = b[:, 1]
            out[name] = {
                "cycles_per_kblock": float((b[:, 0] / kb).mean()),
                "sm_clock_mhz": float((b[:, 0] / b[:, 2]).mean() * 1e3),
                "busy_ms": float(b[:, 2].mean() / 1e6), "tiles_per_cta": float(b[:, 3].mean()),
                # cycles per K block each role spends blocked
                "mma_wait_full": float((b[:, 4] / kb).mean()), "mma_wait_acc_free": float((b[:, 5] / kb).mean()),
                "mma_wait_tileq": float((b[:, 6] / kb).mean()),
                "prod_wait_stage": float((b[:, 9] / kb).mean()), "prod_wait_tileq": float((b[:, 10] / kb).mean()),
                "epi_total": float((b[:, 11] / kb).mean()), "epi_wait_xs": float((b[:, 12] / kb).mean()),
                "epi_wait_acc_ready": float((b[:, 13] / kb).mean()), "epi_wait_tileq": float((b[:, 14] / kb).mean()),
                "epi_tile_epilogue": float((b[:, 15] / kb).mean()),
                "epi_tile_epilogue_cycles_per_tile": float((b[:, 15] / b[:, 3]).mean())}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
