mkdir -p gpurun_out
( timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -8 ) > gpurun_out/f3_pytest.log 2>&1
( timeout 300 python bench.py ) > gpurun_out/f3_bench.json 2> gpurun_out/f3_bench.err
( timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" ) > gpurun_out/f3_smoke.log 2>&1
( timeout 200 python tools/gemm_bench.py --out gpurun_out/route_gemm_bench.json ) > gpurun_out/f3_gemm.log 2>&1
( timeout 100 python tools/prefill_bench.py ; timeout 100 python tools/prefill_bench.py --kpt 0 ) > gpurun_out/f3_prefill.log 2>&1
( timeout 200 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16 -s 5 -c 1 -f -o gpurun_out/r1_route_gemm python tools/gemm_bench.py --only-m 4096 ) > gpurun_out/f3_ncu_gemm.log 2>&1
( timeout 300 ncu --set full --clock-control none --import-source on -k regex:prefill -s 2 -c 1 -f -o gpurun_out/prefill_r1e python tools/prefill_bench.py --kpt 0 ) > gpurun_out/f3_ncu_prefill.log 2>&1
( timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'decode|assign' -c 90 --csv --log-file gpurun_out/r1_decode_launches.csv python bench.py --steps 30 --warmup 3 ) > gpurun_out/f3_ncu_dec.log 2>&1
tail -3 gpurun_out/f3_pytest.log; cut -c1-200 gpurun_out/f3_bench.json; tail -1 gpurun_out/f3_smoke.log; cut -c1-160 gpurun_out/f3_gemm.log; cut -c1-100 gpurun_out/f3_prefill.log; tail -2 gpurun_out/f3_ncu_gemm.log; tail -2 gpurun_out/f3_ncu_prefill.log
