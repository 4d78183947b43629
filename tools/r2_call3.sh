#!/bin/bash
# round-2 call 3 (1 GPU): UMMA probe with protocol features, PDL decode A/B, new bench.py line, BASELINE-shape tests
mkdir -p gpurun_out
( timeout 200 python tools/umma_rate.py ) > gpurun_out/r2_umma_rate2.log 2>&1
( timeout 600 python -m pytest tests/test_baseline_shapes_gpu.py tests/test_decode_gpu.py -q -m gpu -x 2>&1 | tail -8 ) > gpurun_out/r2_c3_pytest.log 2>&1
( HPC_B200_PDL=0 timeout 200 python bench.py --no-extra --steps 500 ) > gpurun_out/r2_bench_nopdl.json 2> gpurun_out/r2_bench_nopdl.err
( timeout 600 python bench.py ) > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err
cat gpurun_out/r2_umma_rate2.log | cut -c1-330
tail -5 gpurun_out/r2_c3_pytest.log
cut -c1-900 gpurun_out/r2_bench_nopdl.json; tail -3 gpurun_out/r2_bench_nopdl.err
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r2_bench.json').read().strip().splitlines()[-1])
    print({k:d[k] for k in ('value','ms_per_step','timed_regions')}, d['roofline']['frac'], d['e2e']['value'])
    for k,v in d.get('extra',{}).items():
        print(k, json.dumps(v)[:1500])
except Exception as e:
    print('bench parse failed', e)
PY
tail -5 gpurun_out/r2_bench.err
