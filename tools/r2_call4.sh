#!/bin/bash
# round-2 call 4 (1 GPU): reworked grouped GEMM (smem scale rings, commit order, PDL), kpt decode, reference tests
mkdir -p gpurun_out
( timeout 400 python -m pytest tests/test_moe_gpu.py tests/test_act_gpu.py -q -m gpu -x 2>&1 | tail -6 ) > gpurun_out/r2_moe_pytest.log 2>&1
( HPC_B200_MOE_DEBUG=8 timeout 200 python tools/moe_bench.py ) > gpurun_out/r2_moe_dbg8.log 2>&1
( HPC_B200_MOE_DEBUG=12 timeout 200 python tools/moe_bench.py ) > gpurun_out/r2_moe_dbg12.log 2>&1
( HPC_B200_MOE_DEBUG=15 timeout 200 python tools/moe_bench.py ) > gpurun_out/r2_moe_dbg15.log 2>&1
( timeout 400 python -m pytest tests/test_baseline_shapes_gpu.py -q -m gpu 2>&1 | tail -6 ) > gpurun_out/r2_c3_pytest.log 2>&1
( timeout 900 python -m pytest tests/test_decode_gpu.py -q -m gpu -x 2>&1 | tail -6 ) > gpurun_out/r2_decode_pytest.log 2>&1
tail -3 gpurun_out/r2_moe_pytest.log; tail -1 gpurun_out/r2_moe_dbg8.log | cut -c1-900; tail -1 gpurun_out/r2_moe_dbg12.log | cut -c1-900;  tail -1 gpurun_out/r2_moe_dbg15.log | cut -c1-900
tail -4 gpurun_out/r2_c3_pytest.log; tail -4 gpurun_out/r2_decode_pytest.log
bash tools/run_reference_tests.sh gpurun_out/r2_reference_tests.txt 240 test_group_gemm_blockwise.py test_group_gemm_pertensor.py test_fuse_moe_blockwise.py test_fuse_moe_pertensor.py test_fuse_moe_cp_async.py test_group_gemm_cp_async.py test_gemm_bf16xfp32.py test_act.py test_version.py test_attention_decode_qpertoken_perhead_kvpertensor_fp8.py test_attention_decode_qkpertoken_perhead_vperhead_fp8.py test_attention_blocksparse_qpertoken_perhead_kvpertensor_fp8.py test_attention_blocksparse_qkpertoken_perhead_vperhead_fp8.py 2>&1 | cut -c1-400
