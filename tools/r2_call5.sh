#!/bin/bash
# round-2 call 5 (1 GPU): wait attribution of the grouped GEMM; kpt decode + dense prefill tests
mkdir -p gpurun_out
for dbg in 8 12 15; do
( HPC_B200_MOE_DEBUG=$dbg timeout 200 python tools/moe_bench.py ) > gpurun_out/r2_moe_attr$dbg.log 2>&1
done
( timeout 900 python -m pytest tests/test_decode_gpu.py tests/test_prefill_gpu.py -q -m gpu 2>&1 | tail -8 ) > gpurun_out/r2_decode_pytest.log 2>&1
for dbg in 8 12 15; do tail -1 gpurun_out/r2_moe_attr$dbg.log | python -c "
import sys,json
d=json.loads(sys.stdin.read())
print('dbg$dbg ms %.2f'%d['ms'])
for k in ('gate_up','down'):
    print(' ',k,{a:round(b,1) for a,b in d[k].items()})
"; done
tail -6 gpurun_out/r2_decode_pytest.log
bash tools/run_reference_tests.sh gpurun_out/r2_reference_tests_b.txt 400 test_fuse_moe_blockwise.py test_act.py 2>&1 | cut -c1-600
