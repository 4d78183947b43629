#!/bin/bash
# round-2 call 9 (1 GPU): rope tests + bench, final MoE number, full GPU suite timing
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_rope_gpu.py -q -m gpu 2>&1 | tail -12 ) > gpurun_out/r2_rope_pytest.log 2>&1
tail -8 gpurun_out/r2_rope_pytest.log
( timeout 100 python tools/moe_bench.py ) > gpurun_out/r2_moe_final.log 2>&1
tail -1 gpurun_out/r2_moe_final.log | cut -c1-200
( timeout 1500 python -m pytest tests -q -m gpu -x --deselect tests/test_rope_gpu.py 2>&1 | tail -5 ) > gpurun_out/r2_full_pytest.log 2>&1
tail -4 gpurun_out/r2_full_pytest.log
