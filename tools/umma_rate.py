"""UMMA issue-rate probe (diagnostics, GPU box only): cycles per K block (4 x M x N x 32 fp8 UMMAs)
for cta_group::1 / ::2 with the grouped GEMM's pipeline protocol added piece by piece (flags, see
csrc/selftest.cu). 512 cycles per K block is the nominal rate at N=256 (8192 MAC/clk/SM).

    python tools/umma_rate.py
"""
import json
import sys
from pathlib import Path

REPO = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(REPO / "hpc-ops_b200"))
import torch  # noqa: E402

from hpc import _ffi  # noqa: E402

FLAG_NAMES = {1: "handshake", 2: "stage-protocol", 4: "8-consumer-warps", 8: "random-data", 16: "drain"}


def main():
    iters = 8000
    out = torch.zeros(1024, dtype=torch.int64, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    cases = [(0, 256, f) for f in (0, 8, 1, 1 | 8, 1 | 2 | 8, 1 | 4 | 8, 1 | 2 | 4 | 8, 1 | 2 | 4 | 8 | 16,
                                   1 | 4 | 16, 1 | 2 | 4 | 16)]
    cases += [(1, 256, f) for f in (0, 8, 1 | 8, 1 | 2 | 4 | 8, 1 | 2 | 4 | 8 | 16)]
    cases += [(0, 128, 8), (1, 128, 8)]
    for pair, n, flags in cases:
        out.zero_()
        _ffi.check(_ffi.lib.hpc_selftest_umma_rate(pair, n, 10, flags, out.data_ptr(), st), "rate")
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _ffi.check(_ffi.lib.hpc_selftest_umma_rate(pair, n, iters, flags, out.data_ptr(), st), "rate")
        e1.record()
        torch.cuda.synchronize()
        v = out.view(-1, 2)
        v = v[v[:, 0] > 0].double()
        ms = e0.elapsed_time(e1)
        ctas = v.shape[0] * (2 if pair else 1)
        flops = 2.0 * 128 * n * 128 * iters * ctas
        print(json.dumps({"cta_group": pair + 1, "N": n,
                          "features": [name for b, name in FLAG_NAMES.items() if flags & b],
                          "cycles_per_kblock": float(v[:, 0].mean()) / iters,
                          "sm_clock_mhz": float((v[:, 0] / v[:, 1]).mean()) * 1e3,
                          "ms": ms, "tflops": flops / ms / 1e9, "nominal_cycles": 128 * n * 4 / 256}))


if __name__ == "__main__":
    main()
