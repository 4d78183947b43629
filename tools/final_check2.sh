mkdir -p gpurun_out
( timeout 300 python -m pytest tests/test_gemm_gpu.py -q -m gpu -x 2>&1 | tail -5 ) > gpurun_out/c2_gemm_pytest.log 2>&1
( timeout 200 python tools/gemm_bench.py --out gpurun_out/route_gemm_bench.json ) > gpurun_out/c2_gemm.log 2>&1
( timeout 100 python tools/prefill_bench.py ; timeout 100 python tools/prefill_bench.py --kpt 0 ) > gpurun_out/c2_prefill.log 2>&1
( timeout 200 python tools/moe_bench.py ) > gpurun_out/c2_moe.log 2>&1
( timeout 200 python bench.py --steps 500 ) > gpurun_out/c2_bench.json 2> gpurun_out/c2_bench.err
tail -3 gpurun_out/c2_gemm_pytest.log; cut -c1-330 gpurun_out/c2_gemm.log; cut -c1-100 gpurun_out/c2_prefill.log; tail -1 gpurun_out/c2_moe.log | cut -c1-120; python -c "
import json;d=json.load(open('gpurun_out/c2_bench.json'));print(d['ms_per_step'],d['roofline']['kernel_ms'],d['roofline']['frac'],d['e2e']['ms_per_step'])"
