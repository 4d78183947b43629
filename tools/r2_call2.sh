#!/bin/bash
# round-2 call 2 (2 GPUs): UMMA issue-rate probe, reworked allreduce tests + bench at W=1,2
mkdir -p gpurun_out
( timeout 120 python tools/umma_rate.py ) > gpurun_out/r2_umma_rate.log 2>&1
( timeout 900 python -m pytest tests/test_allreduce_gpu.py -q -m gpu -x 2>&1 | tail -8 ) > gpurun_out/r2_ar_pytest.log 2>&1
( timeout 200 python tools/allreduce_bench.py ) > gpurun_out/r2_ar_bench_w1.log 2>&1
( timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/allreduce_bench.py ) > gpurun_out/r2_ar_bench_w2.log 2>&1
( HPC_B200_AR_P2P_MAX_WORLD=0 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/allreduce_bench.py ) > gpurun_out/r2_ar_bench_w2_nvls.log 2>&1
cat gpurun_out/r2_umma_rate.log | cut -c1-260; tail -5 gpurun_out/r2_ar_pytest.log
grep -h '"path"' gpurun_out/r2_ar_bench_w1.log gpurun_out/r2_ar_bench_w2.log gpurun_out/r2_ar_bench_w2_nvls.log | cut -c1-230
tail -3 gpurun_out/r2_ar_bench_w2.log | cut -c1-300
