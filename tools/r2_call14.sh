#!/bin/bash
# round-2 call 14 (4 GPUs): allreduce tests at W=4, bench.py under torchrun at N=4
mkdir -p gpurun_out
( HPC_B200_TEST_WORLDS=4 timeout 900 python -m pytest tests/test_allreduce_gpu.py -q -m gpu 2>&1 | tail -6 ) > gpurun_out/r2_ar_pytest4.log 2>&1
tail -4 gpurun_out/r2_ar_pytest4.log
( timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 4 --steps 200 --warmup 5 ) > gpurun_out/r2_bench_n4.json 2> gpurun_out/r2_bench_n4.err
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r2_bench_n4.json').read().strip().splitlines()[-1])
    print({k:d[k] for k in ('value','n_gpus','ms_per_step')})
    ar=d.get('extra',{}).get('allreduce_c5',{})
    for e in ar.get('ht',[]): print('HT', {k:e.get(k) for k in ('tokens','hidden','us','protocol','link_gbs','frac_nvlink_770','nccl_allreduce_only_us','multicast')}, (e.get('parity') or {}).get('max_abs_err'))
    for e in ar.get('ll',[]): print('LL', {k:e.get(k) for k in ('tokens','us','protocol')}, (e.get('parity') or {}).get('max_abs_err'))
    if 'error' in ar: print('AR ERROR', ar['error'])
except Exception as e:
    print('bench parse failed', e)
PY
tail -4 gpurun_out/r2_bench_n4.err | cut -c1-300
