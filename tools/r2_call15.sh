#!/bin/bash
# round-2 call 15 (1 GPU): rope d128 kernel tests + bench, allreduce W=1, smoke
mkdir -p gpurun_out
( timeout 400 python -m pytest tests/test_rope_gpu.py tests/test_allreduce_gpu.py -q -m gpu -x 2>&1 | tail -4 ) > gpurun_out/r2_rope_pytest.log 2>&1
tail -3 gpurun_out/r2_rope_pytest.log
( timeout 100 python tools/rope_bench.py ) > gpurun_out/r2_rope_bench.log 2>&1
tail -1 gpurun_out/r2_rope_bench.log | cut -c1-500
( timeout 120 python -c "import __graft_entry__ as g; g.smoke()" ) 2>&1 | tail -1
