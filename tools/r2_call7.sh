#!/bin/bash
# round-2 call 7 (2 GPUs): bench.py under torchrun at N=2 (decode replicas + sharded allreduce extra), allreduce tests
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_allreduce_gpu.py -q -m gpu -k "high_throughput or c5_shape or one_shot or varying" 2>&1 | tail -5 ) > gpurun_out/r2_ar_pytest2.log 2>&1
( timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 2 --steps 200 --warmup 5 ) > gpurun_out/r2_bench_n2.json 2> gpurun_out/r2_bench_n2.err
( timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29522 bench.py --impl reference --gpus 2 --steps 1 --warmup 0 ) > gpurun_out/r2_bench_n2_ref.json 2> gpurun_out/r2_bench_n2_ref.err
tail -3 gpurun_out/r2_ar_pytest2.log
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r2_bench_n2.json').read().strip().splitlines()[-1])
    print({k:d[k] for k in ('value','n_gpus','ms_per_step','timed_regions')})
    ex=d.get('extra',{})
    print({k:(v if k!='allreduce_c5' else '...') for k,v in ex.items()})
    ar=ex.get('allreduce_c5',{})
    for e in ar.get('ht',[]): print('HT', json.dumps(e)[:700])
    for e in ar.get('ll',[]): print('LL', {k:e.get(k) for k in ('tokens','us','protocol')}, (e.get('parity') or {}).get('max_abs_err'))
    if 'error' in ar: print('AR ERROR', ar['error'])
except Exception as e:
    print('bench parse failed', e)
PY
tail -4 gpurun_out/r2_bench_n2.err | cut -c1-400; cut -c1-600 gpurun_out/r2_bench_n2_ref.json
