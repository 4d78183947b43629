#!/bin/bash
# round-2 call 23 (1 GPU): one `ncu --set full` capture of the decode kernel as shipped (rotated walk)
mkdir -p gpurun_out
( timeout 240 ncu --set full --clock-control none --import-source on -k regex:decode_attn_fp8 -s 4 -c 1 -o gpurun_out/r2_prof_decode_rot -f python bench.py --no-extra --steps 3 --warmup 3 ) > gpurun_out/r2_ncu_decode_rot.log 2>&1
tail -3 gpurun_out/r2_ncu_decode_rot.log | cut -c1-300
ls -la gpurun_out/*.ncu-rep | tail -2
