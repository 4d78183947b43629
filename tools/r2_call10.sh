#!/bin/bash
# round-2 call 10 (1 GPU): rope mtp=1 failure detail, MoE re-measure, prefill with single-lane polling
mkdir -p gpurun_out
( timeout 300 python -m pytest tests/test_rope_gpu.py -q -m gpu -x -k "test_rope_norm_store_kv and False-1-7-0-8-1-128" 2>&1 | grep -E "^E |assert|Error" | head -20 ) > gpurun_out/r2_rope_fail.log 2>&1
cat gpurun_out/r2_rope_fail.log | cut -c1-300
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw,temperature.gpu,clocks_event_reasons.active --format=csv
( timeout 100 python tools/moe_bench.py ) > gpurun_out/r2_moe_final.log 2>&1
tail -1 gpurun_out/r2_moe_final.log | cut -c1-200
( HPC_B200_MOE_DEBUG=8 timeout 200 python tools/moe_bench.py ) > gpurun_out/r2_moe_attr8.log 2>&1
tail -1 gpurun_out/r2_moe_attr8.log | python -c "
import sys,json
d=json.loads(sys.stdin.read())
print('dbg8 ms %.2f'%d['ms'])
for k in ('gate_up','down'):
    print(' ',k,{a:round(b,1) for a,b in d[k].items()})
"
( HPC_B200_PDL=0 timeout 100 python tools/moe_bench.py ) > gpurun_out/r2_moe_nopdl.log 2>&1
tail -1 gpurun_out/r2_moe_nopdl.log | cut -c1-200
( timeout 100 python tools/prefill_bench.py; timeout 100 python tools/prefill_bench.py --kpt 0 ) > gpurun_out/r2_prefill_poll.log 2>&1
cut -c1-200 gpurun_out/r2_prefill_poll.log
( timeout 100 python tools/rope_bench.py ) > gpurun_out/r2_rope_bench.log 2>&1
tail -1 gpurun_out/r2_rope_bench.log | cut -c1-600
( timeout 400 python -m pytest tests/test_prefill_gpu.py tests/test_moe_gpu.py -q -m gpu -x 2>&1 | tail -3 ) > gpurun_out/r2_poll_pytest.log 2>&1
tail -2 gpurun_out/r2_poll_pytest.log
