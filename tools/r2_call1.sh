#!/bin/bash
# round-2 diagnostics: where does the grouped GEMM's time go; do the two unvalidated variants work
mkdir -p gpurun_out
for dbg in 0 1 3 4 5 7; do
  ( HPC_B200_MOE_DEBUG=$dbg timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'group_gemm|moe_' --csv \
      --log-file gpurun_out/r2_moe_dbg$dbg.csv python tools/moe_bench.py --iters 1 ) > gpurun_out/r2_moe_dbg$dbg.log 2>&1
done
( timeout 100 python tools/moe_bench.py ) > gpurun_out/r2_moe_base.log 2>&1
( HPC_B200_MOE_CLUSTER=1 timeout 300 python -m pytest tests/test_moe_gpu.py -x -q -m gpu 2>&1 | tail -5 ) > gpurun_out/r2_moe_cluster_pytest.log 2>&1
( HPC_B200_MOE_CLUSTER=1 timeout 100 python tools/moe_bench.py ) > gpurun_out/r2_moe_cluster_bench.log 2>&1
( HPC_B200_PREFILL_WG2=1 timeout 300 python -m pytest tests/test_prefill_gpu.py -x -q -m gpu 2>&1 | tail -5 ) > gpurun_out/r2_prefill_wg2_pytest.log 2>&1
( HPC_B200_PREFILL_WG2=1 timeout 100 python tools/prefill_bench.py; HPC_B200_PREFILL_WG2=1 timeout 100 python tools/prefill_bench.py --kpt 0 ) > gpurun_out/r2_prefill_wg2_bench.log 2>&1
( timeout 100 python tools/prefill_bench.py; timeout 100 python tools/prefill_bench.py --kpt 0 ) > gpurun_out/r2_prefill_base.log 2>&1
for f in gpurun_out/r2_moe_dbg*.csv; do echo $f; grep -E "group_gemm|moe_" $f | awk -F'","' '{print $5, $NF}' | tail -6; done
tail -2 gpurun_out/r2_moe_base.log gpurun_out/r2_moe_cluster_pytest.log gpurun_out/r2_moe_cluster_bench.log gpurun_out/r2_prefill_wg2_pytest.log gpurun_out/r2_prefill_wg2_bench.log gpurun_out/r2_prefill_base.log | cut -c1-300
