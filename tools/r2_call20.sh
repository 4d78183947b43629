#!/bin/bash
# round-2 call 20 (1 GPU): rotated walk: L2 policy A/B, DRAM traffic with and without the rotation
mkdir -p gpurun_out
( timeout 200 python tools/decode_rotate_ab.py ) > gpurun_out/r2_decode_rotate_ab2.log 2>&1
tail -1 gpurun_out/r2_decode_rotate_ab2.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
for k,v in d.items():
    for n,r in v.items(): print(k, n, [x['ms'] for x in r])
"
M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct
( timeout 200 ncu --clock-control none --metrics $M -k regex:decode_attn_fp8 -s 3 -c 2 --csv --log-file gpurun_out/r2_decode_rot1_dram.csv python bench.py --no-extra --steps 3 --warmup 3 ) > /dev/null 2>&1
( HPC_B200_DECODE_ROTATE=0 timeout 200 ncu --clock-control none --metrics $M -k regex:decode_attn_fp8 -s 3 -c 2 --csv --log-file gpurun_out/r2_decode_rot0_dram.csv python bench.py --no-extra --steps 3 --warmup 3 ) > /dev/null 2>&1
grep -h "dram__bytes_read\|gpu__time\|hit_rate" gpurun_out/r2_decode_rot1_dram.csv | cut -d, -f5,13- | head -8
echo ---
grep -h "dram__bytes_read\|gpu__time\|hit_rate" gpurun_out/r2_decode_rot0_dram.csv | cut -d, -f5,13- | head -8
