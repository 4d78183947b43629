#!/bin/bash
# round-2 call 12 (1 GPU): validation of the current build (per-warp arrivals, rope rewrite, counters), bench line, small-M MoE, ncu pass
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -5 ) > gpurun_out/r2_full_pytest.log 2>&1
tail -3 gpurun_out/r2_full_pytest.log
( timeout 100 python tools/prefill_bench.py; timeout 100 python tools/prefill_bench.py --kpt 0 ) > gpurun_out/r2_prefill_warparrive.log 2>&1
cut -c1-120 gpurun_out/r2_prefill_warparrive.log
( timeout 100 python tools/moe_bench.py ) > gpurun_out/r2_moe_final.log 2>&1
tail -1 gpurun_out/r2_moe_final.log | cut -c1-160
( timeout 100 python tools/moe_small_bench.py ) > gpurun_out/r2_moe_small.log 2>&1
tail -1 gpurun_out/r2_moe_small.log | cut -c1-500
( timeout 100 python tools/rope_bench.py ) > gpurun_out/r2_rope_bench.log 2>&1
tail -1 gpurun_out/r2_rope_bench.log | cut -c1-500
( timeout 600 python bench.py ) > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r2_bench_n1.json').read().strip().splitlines()[-1])
    print({k:d[k] for k in ('value','ms_per_step')}, 'roof', round(d['roofline']['frac'],4), 'e2e', d['e2e']['value'], d['clocks'])
    for k,v in d.get('extra',{}).items():
        if isinstance(v, dict):
            print(k, {a:(round(b,4) if isinstance(b,float) else b) for a,b in v.items() if a in ('ms','tok_per_s','tflops','frac_fp8_nominal_4500','frac_fp8_2x_measured_bf16','frac_hbm','error','wall_s','clocks')})
except Exception as e:
    print('bench parse failed', e)
PY
tail -3 gpurun_out/r2_bench_n1.err | cut -c1-300
bash tools/r2_ncu.sh 2>&1 | tail -10
