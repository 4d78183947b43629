#!/bin/bash
mkdir -p gpurun_out
( timeout 200 python tools/decode_layout_probe.py ) > gpurun_out/r2_decode_layout_probe.log 2>&1
tail -2 gpurun_out/r2_decode_layout_probe.log | cut -c1-900
