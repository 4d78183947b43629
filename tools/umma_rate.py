"""UMMA issue-rate probe (diagnostics, GPU box only): cycles per K block (4 x M x N x 32 fp8 UMMAs)
for cta_group::1 / ::2, with and without the per-K-block ready/drained handshake of the grouped
GEMM. 512 cycles per K block is the nominal rate at N=256 (8192 MAC/clk/SM).

    python tools/umma_rate.py
"""
import json
import sys
from pathlib import Path

REPO = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(REPO / "hpc-ops_b200"))
import torch  # noqa: E402

from hpc import _ffi  # noqa: E402


def main():
    iters = 4000
    out = torch.zeros(256, dtype=torch.int64, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    for pair in (0, 1):
        for n in (256, 128, 64):
            for hs in (0, 1):
                out.zero_()
                _ffi.check(_ffi.lib.hpc_selftest_umma_rate(pair, n, 10, hs, out.data_ptr(), st), "rate")
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                _ffi.check(_ffi.lib.hpc_selftest_umma_rate(pair, n, iters, hs, out.data_ptr(), st), "rate")
                e1.record()
                torch.cuda.synchronize()
                cyc = out[out > 0].double()
                ms = e0.elapsed_time(e1)
                m = 256 if pair else 128
                ctas = int((out > 0).sum()) * (2 if pair else 1)
                flops = 2.0 * 128 * n * 128 * iters * ctas
                print(json.dumps({"cta_group": pair + 1, "M": m, "N": n, "handshake": hs,
                                  "cycles_per_kblock_mean": float(cyc.mean()) / iters,
                                  "cycles_per_kblock_max": float(cyc.max()) / iters,
                                  "ms": ms, "tflops": flops / ms / 1e9,
                                  "nominal_cycles": 128 * n * 4 / 256}))


if __name__ == "__main__":
    main()
