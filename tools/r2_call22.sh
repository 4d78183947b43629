#!/bin/bash
# round-2 call 22 (1 GPU): the whole GPU suite on the HEAD build
mkdir -p gpurun_out
( timeout 640 python -m pytest tests -q -m gpu -x 2>&1 | tail -12 ) > gpurun_out/r2_gpu_suite_final.log 2>&1
tail -6 gpurun_out/r2_gpu_suite_final.log
