#!/bin/bash
# round-2 call 13 (1 GPU): decode A/B (per-warp vs per-thread softmax arrivals, direct bf16 write), decode tests, sanitizer smoke
mkdir -p gpurun_out
H=hpc-ops_b200/hpc
cp $H/_C.so $H/_C.so.current
: > gpurun_out/r2_decode_ab.log
for rep in 1 2; do
for tag in current perthread 89b00ab; do
  cp $H/_C.so.$tag $H/_C.so
  r=$(timeout 100 python tools/decode_ab.py 2>/dev/null | tail -1)
  echo "rep $rep lib $tag : $r" | tee -a gpurun_out/r2_decode_ab.log
done
done
cp $H/_C.so.current $H/_C.so
( timeout 600 python -m pytest tests/test_decode_gpu.py -q -m gpu -x 2>&1 | tail -3 ) > gpurun_out/r2_decode_pytest.log 2>&1
tail -2 gpurun_out/r2_decode_pytest.log
for tool in memcheck synccheck racecheck; do
  ( timeout 600 compute-sanitizer --tool $tool python tools/sanitize_smoke.py 2>&1 | tail -12 ) > gpurun_out/r2_sanitizer_$tool.log 2>&1
  echo "== $tool"; tail -6 gpurun_out/r2_sanitizer_$tool.log | cut -c1-300
done
