#!/bin/bash
# on-box A/B of prebuilt libraries (hpc/_C.so.<tag>): MoE C3 and prefill C4 per library, interleaved twice
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
H=hpc-ops_b200/hpc
cp $H/_C.so $H/_C.so.current
: > gpurun_out/r2_ab.log
for rep in 1 2; do
for tag in "$@"; do
  cp $H/_C.so.$tag $H/_C.so
  m=$(timeout 100 python tools/moe_bench.py 2>/dev/null | tail -1 | python -c "import sys,json; print('%.2f'%json.loads(sys.stdin.read())['ms'])" 2>/dev/null)
  p0=$(timeout 100 python tools/prefill_bench.py --kpt 0 2>/dev/null | tail -1 | python -c "import sys,json; print('%.3f'%json.loads(sys.stdin.read())['ms'])" 2>/dev/null)
  p1=$(timeout 100 python tools/prefill_bench.py --kpt 1 2>/dev/null | tail -1 | python -c "import sys,json; print('%.3f'%json.loads(sys.stdin.read())['ms'])" 2>/dev/null)
  echo "rep $rep lib $tag : moe_c3 ${m} ms  prefill_c4 kvpt ${p0} ms kpt ${p1} ms" | tee -a gpurun_out/r2_ab.log
done
done
cp $H/_C.so.current $H/_C.so
