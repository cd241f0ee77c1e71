#!/bin/bash
mkdir -p gpurun_out
export FB_REQUIRE_GPU=1
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -8 > gpurun_out/all_tests.log
: > gpurun_out/waves.log
for w in 0 3 2; do echo "FB_PLOIDY_WAVE=$w" >> gpurun_out/waves.log; FB_PLOIDY_WAVE=$w FB_HOST_PROF=1 timeout 900 python tools/scale_run.py c5 500 2>&1 | grep "pass 2\|host ms" | tail -2 >> gpurun_out/waves.log; done
timeout 300 python tools/run_once.py 3 2>&1 | tail -1 >> gpurun_out/waves.log
cat gpurun_out/all_tests.log gpurun_out/waves.log
