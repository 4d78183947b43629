#!/bin/bash
# GPU box: run the reference's own tests (staged by tools/stage_reference_tests.sh) against this
# package, file by file, and record pass / fail / skip per file.
#   bash tools/run_reference_tests.sh [out-file] [per-file-timeout-s] [files...]
out=${1:-gpurun_out/r2_reference_tests.txt}; tmo=${2:-600}; shift 2 2>/dev/null
cd "$(dirname "$0")/.."
mkdir -p "$(dirname "$out")"
files=${@:-"test_attention_decode_qpertoken_perhead_kvpertensor_fp8.py test_attention_decode_qkpertoken_perhead_vperhead_fp8.py test_attention_blocksparse_qpertoken_perhead_kvpertensor_fp8.py test_attention_blocksparse_qkpertoken_perhead_vperhead_fp8.py test_fuse_moe_blockwise.py test_fuse_moe_pertensor.py test_fuse_moe_cp_async.py test_group_gemm_blockwise.py test_group_gemm_pertensor.py test_group_gemm_cp_async.py test_gemm_bf16xfp32.py test_act.py test_version.py test_fuse_allreduce_rmsnorm_high_throughput.py test_fuse_allreduce_rmsnorm_low_latency.py"}
: > "$out"
for f in $files; do
  if [ ! -f baseline/_ref/tests/$f ]; then echo "$f: not staged" >> "$out"; continue; fi
  r=$( cd baseline/_ref/tests && timeout $tmo python -m pytest $f -q --no-header -p no:cacheprovider 2>&1 | tail -3 | tr '\n' ' ' )
  echo "$f: $r" >> "$out"
done
cat "$out"
