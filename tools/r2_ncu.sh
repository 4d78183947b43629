#!/bin/bash
# round-2 profiling pass (1 GPU): launch list of the bench step + one `ncu --set full` capture per kernel family.
# Numbers printed under ncu are never bench values; summaries are written into profiles/ from the .ncu-rep files.
mkdir -p gpurun_out
NCU="ncu --clock-control none"
( timeout 300 $NCU --metrics gpu__time_duration.sum -k regex:'decode|assign' -s 30 -c 90 --csv --log-file gpurun_out/r2_decode_launches.csv python bench.py --no-extra --steps 30 --warmup 3 ) > gpurun_out/r2_ncu_launches.log 2>&1
( timeout 300 $NCU --set full -k regex:decode_attn_fp8 -s 4 -c 1 -o gpurun_out/r2_prof_decode -f python bench.py --no-extra --steps 3 --warmup 3 ) > gpurun_out/r2_ncu_decode.log 2>&1
( timeout 300 $NCU --set full -k regex:decode_combine -s 4 -c 1 -o gpurun_out/r2_prof_combine -f python bench.py --no-extra --steps 3 --warmup 3 ) > gpurun_out/r2_ncu_combine.log 2>&1
( timeout 400 $NCU --set full -k regex:group_gemm_fp8 -s 4 -c 2 -o gpurun_out/r2_prof_moe -f python tools/moe_bench.py --iters 1 ) > gpurun_out/r2_ncu_moe.log 2>&1
( timeout 300 $NCU --set full -k regex:prefill_blocksparse -s 2 -c 1 -o gpurun_out/r2_prof_prefill -f python tools/prefill_bench.py --kpt 0 --iters 1 ) > gpurun_out/r2_ncu_prefill.log 2>&1
( timeout 300 $NCU --set full -k regex:ar_rmsnorm_ht -s 10 -c 1 -o gpurun_out/r2_prof_ar_w1 -f python tools/allreduce_bench.py --iters 5 ) > gpurun_out/r2_ncu_ar.log 2>&1
( timeout 200 $NCU --set full -k regex:rope_norm -s 3 -c 1 -o gpurun_out/r2_prof_rope -f python tools/rope_bench.py --iters 3 ) > gpurun_out/r2_ncu_rope.log 2>&1
ls -la gpurun_out/*.ncu-rep
