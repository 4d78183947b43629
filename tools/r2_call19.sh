#!/bin/bash
# round-2 call 19 (1 GPU): rotated bin walk of the fp8 decode kernel: parity, then A/B timing
mkdir -p gpurun_out
( timeout 500 python -m pytest tests/test_decode_gpu.py -q -m gpu -x 2>&1 | tail -15 ) > gpurun_out/r2_decode_rotate_pytest.log 2>&1
tail -6 gpurun_out/r2_decode_rotate_pytest.log
( timeout 200 python tools/decode_rotate_ab.py ) > gpurun_out/r2_decode_rotate_ab.log 2>&1
tail -2 gpurun_out/r2_decode_rotate_ab.log | cut -c1-2500
