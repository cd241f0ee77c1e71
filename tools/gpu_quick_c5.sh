#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/c5_62.log
for w in 3 0 6; do echo "== FB_PLOIDY_WAVE=$w (0 = all at once)" >> gpurun_out/c5_62.log; FB_PLOIDY_WAVE=$w FB_HOST_PROF=1 timeout 900 python tools/scale_run.py c5 62 2>&1 | grep "pass 2\|host ms" | tail -6 >> gpurun_out/c5_62.log; done
cat gpurun_out/c5_62.log
