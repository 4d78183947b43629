#!/bin/bash
# round-2 call 11 (1 GPU): on-box A/B of three library builds (is the 14 ms MoE a code or a box effect?), rope tests
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,power.limit,temperature.gpu,clocks_event_reasons.active --format=csv
bash tools/r2_ab.sh current 468f43b 89b00ab
( timeout 600 python -m pytest tests/test_rope_gpu.py -q -m gpu 2>&1 | tail -5 ) > gpurun_out/r2_rope_pytest.log 2>&1
tail -3 gpurun_out/r2_rope_pytest.log
( timeout 100 python tools/rope_bench.py ) > gpurun_out/r2_rope_bench.log 2>&1
tail -1 gpurun_out/r2_rope_bench.log | cut -c1-600
