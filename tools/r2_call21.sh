#!/bin/bash
# round-2 call 21 (1 GPU): HEAD build: quick decode parity, bench line, smoke
mkdir -p gpurun_out
( timeout 200 python -m pytest tests/test_decode_gpu.py -q -m gpu -x -k "golden or edge or c2 or without_task_map or paging" 2>&1 | tail -3 )
( timeout 400 python bench.py ) > gpurun_out/r2_bench_n1_final.json 2> gpurun_out/r2_bench_n1_final.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench_n1_final.json').read().strip().splitlines()[-1])
print({k:d.get(k) for k in ('value','ms_per_step','gpu_launches')}, d['roofline'], d['e2e']['value'], d['clocks'])
for k,v in d.get('extra',{}).items():
    print(k, json.dumps(v)[:300])
PY
tail -2 gpurun_out/r2_bench_n1_final.err | cut -c1-300
( timeout 100 python -c "import __graft_entry__ as g; g.smoke()" ) 2>&1 | tail -1
