#!/bin/bash
mkdir -p gpurun_out
export FB_REQUIRE_GPU=1
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -6 > gpurun_out/all_tests.log
: > gpurun_out/uni.log
timeout 300 python tools/run_once.py 4 2>&1 | tail -3 >> gpurun_out/uni.log
FB_BEAM_PROF=1 timeout 300 python tools/run_once.py 1 2>&1 | grep prof | tail -2 >> gpurun_out/uni.log
timeout 300 python tools/c3_probe.py 100000 50000 4 2>&1 | grep "^rep" | cut -c1-60 >> gpurun_out/uni.log
FB_BEAM_PROF=1 timeout 300 python tools/c3_probe.py 100000 50000 4 2>&1 | grep prof | head -1 >> gpurun_out/uni.log
timeout 600 python tools/scale_run.py c5 100 2>&1 | grep "pass 2" >> gpurun_out/uni.log
cat gpurun_out/all_tests.log gpurun_out/uni.log
