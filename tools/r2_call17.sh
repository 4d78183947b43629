#!/bin/bash
# round-2 call 17 (1 GPU): bf16 decode bring-up: descriptor self tests, parity grid, timing
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_decode_bf16_gpu.py -q -m gpu -x 2>&1 | tail -25 ) > gpurun_out/r2_decode_bf16_pytest.log 2>&1
tail -12 gpurun_out/r2_decode_bf16_pytest.log
( timeout 200 python -m pytest tests/test_decode_gpu.py -q -m gpu -x -k "golden or edge or without_task_map" 2>&1 | tail -3 )
( timeout 150 python tools/decode_bf16_bench.py ) > gpurun_out/r2_decode_bf16_bench.log 2>&1
tail -2 gpurun_out/r2_decode_bf16_bench.log | cut -c1-600
