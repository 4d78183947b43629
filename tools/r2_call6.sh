#!/bin/bash
# round-2 call 6 (1 GPU): warp-uniform issue in grouped GEMM, prefill and decode kernels
mkdir -p gpurun_out
( timeout 400 python -m pytest tests/test_moe_gpu.py tests/test_act_gpu.py tests/test_baseline_shapes_gpu.py -q -m gpu -x 2>&1 | tail -6 ) > gpurun_out/r2_moe_pytest.log 2>&1
( timeout 100 python tools/moe_bench.py ) > gpurun_out/r2_moe_uniform.log 2>&1
for dbg in 8 12 15; do
( HPC_B200_MOE_DEBUG=$dbg timeout 200 python tools/moe_bench.py ) > gpurun_out/r2_moe_attr$dbg.log 2>&1
done
tail -3 gpurun_out/r2_moe_pytest.log; tail -1 gpurun_out/r2_moe_uniform.log | cut -c1-300
for dbg in 8 12 15; do tail -1 gpurun_out/r2_moe_attr$dbg.log | python -c "
import sys,json
d=json.loads(sys.stdin.read())
print('dbg$dbg ms %.2f'%d['ms'])
for k in ('gate_up','down'):
    print(' ',k,{a:round(b,1) for a,b in d[k].items()})
"; done
( timeout 100 python tools/prefill_bench.py; timeout 100 python tools/prefill_bench.py --kpt 0 ) > gpurun_out/r2_prefill_uniform.log 2>&1
cut -c1-200 gpurun_out/r2_prefill_uniform.log
( timeout 900 python -m pytest tests/test_prefill_gpu.py tests/test_decode_gpu.py -q -m gpu -x 2>&1 | tail -5 ) > gpurun_out/r2_attn_pytest.log 2>&1
tail -3 gpurun_out/r2_attn_pytest.log
( timeout 200 python bench.py --no-extra --steps 1000 ) > gpurun_out/r2_bench_uniform.json 2> gpurun_out/r2_bench_uniform.err
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r2_bench_uniform.json').read().strip().splitlines()[-1])
    print({k:d[k] for k in ('value','ms_per_step')}, d['roofline']['frac'], d['roofline']['kernel_ms'], d['e2e']['ms_per_step'])
except Exception as e:
    print('bench parse failed', e)
PY
tail -3 gpurun_out/r2_bench_uniform.err
