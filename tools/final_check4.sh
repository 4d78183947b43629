mkdir -p gpurun_out
( timeout 300 python -m pytest tests/test_act_gpu.py -q -m gpu 2>&1 | tail -25 ) > gpurun_out/f4_act.log 2>&1
( timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -12 ) > gpurun_out/f4_pytest.log 2>&1
( timeout 300 python bench.py ) > gpurun_out/f4_bench.json 2> gpurun_out/f4_bench.err
( timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" ) > gpurun_out/f4_smoke.log 2>&1
tail -25 gpurun_out/f4_act.log; tail -4 gpurun_out/f4_pytest.log; cut -c1-220 gpurun_out/f4_bench.json; tail -3 gpurun_out/f4_bench.err; tail -1 gpurun_out/f4_smoke.log
