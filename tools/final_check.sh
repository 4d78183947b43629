mkdir -p gpurun_out
( timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -8 ) > gpurun_out/final_pytest.log 2>&1
( timeout 300 python bench.py --impl reference ) > gpurun_out/final_bench_ref.json 2> gpurun_out/final_bench_ref.err
( timeout 300 python bench.py ) > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err
( timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" ) > gpurun_out/final_smoke.log 2>&1
( timeout 200 python tools/gemm_bench.py --out gpurun_out/route_gemm_bench.json ) > gpurun_out/final_gemm.log 2>&1
( timeout 200 python tools/moe_bench.py ) > gpurun_out/final_moe.log 2>&1
( timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'decode|assign' -c 120 --csv --log-file gpurun_out/r1_decode_launches.csv python bench.py --steps 30 --warmup 3 ) > gpurun_out/final_ncu.log 2>&1
tail -3 gpurun_out/final_pytest.log; cat gpurun_out/final_bench.json | cut -c1-400; tail -1 gpurun_out/final_smoke.log; tail -3 gpurun_out/final_gemm.log; tail -4 gpurun_out/final_moe.log
