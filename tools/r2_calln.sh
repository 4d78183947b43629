#!/bin/bash
# multi-GPU check: tools/r2_calln.sh N [skip-pytest]: allreduce tests at W=N, then bench.py under torchrun at N
N=${1:-2}; mkdir -p gpurun_out
if [ -z "$2" ]; then
( HPC_B200_TEST_WORLDS=$N timeout 600 python -m pytest tests/test_allreduce_gpu.py -q -m gpu 2>&1 | tail -6 ) > gpurun_out/r2_ar_pytest$N.log 2>&1
tail -4 gpurun_out/r2_ar_pytest$N.log
fi
( timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2953$N bench.py --gpus $N --steps 200 --warmup 5 ) > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n$N.err
python - $N <<'PY'
import json, sys
N = sys.argv[1]
try:
    d=json.loads(open(f'gpurun_out/r2_bench_n{N}.json').read().strip().splitlines()[-1])
    print({k:d[k] for k in ('value','n_gpus','ms_per_step')})
    ar=d.get('extra',{}).get('allreduce_c5',{})
    for e in ar.get('ht',[]): print('HT', {k:e.get(k) for k in ('tokens','hidden','us','protocol','link_gbs','frac_nvlink_770','nccl_allreduce_only_us','multicast')}, (e.get('parity') or {}).get('max_abs_err'))
    for e in ar.get('ll',[]): print('LL', {k:e.get(k) for k in ('tokens','us','protocol')}, (e.get('parity') or {}).get('max_abs_err'))
    if 'error' in ar: print('AR ERROR', ar['error'])
except Exception as e:
    print('bench parse failed', e)
PY
tail -4 gpurun_out/r2_bench_n$N.err | cut -c1-300
