#!/bin/bash
# Full verification pass on a B200 box (what the round's last gpurun calls ran):
#   gpurun --timeout 1500 -- 'bash tools/gpu_check.sh'
# Outputs land in gpurun_out/ (scratch); copy what should be judged into profiles/.
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -12 ) > gpurun_out/check_pytest.log 2>&1
( timeout 300 python bench.py --impl reference ) > gpurun_out/check_bench_ref.json 2> gpurun_out/check_bench_ref.err
( timeout 300 python bench.py ) > gpurun_out/check_bench.json 2> gpurun_out/check_bench.err
( timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" ) > gpurun_out/check_smoke.log 2>&1
( timeout 200 python tools/gemm_bench.py --out gpurun_out/route_gemm_bench.json ) > gpurun_out/check_gemm.log 2>&1
( timeout 100 python tools/prefill_bench.py ; timeout 100 python tools/prefill_bench.py --kpt 0 ) > gpurun_out/check_prefill.log 2>&1
( timeout 200 python tools/moe_bench.py ) > gpurun_out/check_moe.log 2>&1
# launch list of the bench step (kernel shares; times under ncu are not bench values)
( timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'decode|assign' -c 90 --csv \
    --log-file gpurun_out/decode_launches.csv python bench.py --steps 30 --warmup 3 ) > gpurun_out/check_ncu.log 2>&1
tail -3 gpurun_out/check_pytest.log; cut -c1-300 gpurun_out/check_bench.json; tail -1 gpurun_out/check_smoke.log
tail -2 gpurun_out/check_gemm.log | cut -c1-200; cut -c1-100 gpurun_out/check_prefill.log; tail -1 gpurun_out/check_moe.log | cut -c1-160
