#!/bin/bash
# round-2 call 8 (1 GPU): grouped GEMM: scales on the stage barrier + L2 prefetch of weight tiles
mkdir -p gpurun_out
( timeout 400 python -m pytest tests/test_moe_gpu.py tests/test_act_gpu.py tests/test_baseline_shapes_gpu.py -q -m gpu -x 2>&1 | tail -6 ) > gpurun_out/r2_moe_pytest.log 2>&1
( timeout 100 python tools/moe_bench.py ) > gpurun_out/r2_moe_pf.log 2>&1
( HPC_B200_MOE_DEBUG=16 timeout 100 python tools/moe_bench.py ) > gpurun_out/r2_moe_nopf.log 2>&1
( HPC_B200_MOE_DEBUG=8 timeout 200 python tools/moe_bench.py ) > gpurun_out/r2_moe_attr8.log 2>&1
tail -3 gpurun_out/r2_moe_pytest.log; tail -1 gpurun_out/r2_moe_pf.log | cut -c1-200; tail -1 gpurun_out/r2_moe_nopf.log | cut -c1-200
tail -1 gpurun_out/r2_moe_attr8.log | python -c "
import sys,json
d=json.loads(sys.stdin.read())
print('dbg8 ms %.2f'%d['ms'])
for k in ('gate_up','down'):
    print(' ',k,{a:round(b,1) for a,b in d[k].items()})
"
