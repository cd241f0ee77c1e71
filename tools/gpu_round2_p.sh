#!/bin/bash
mkdir -p gpurun_out
export FB_REQUIRE_GPU=1
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -4 > gpurun_out/ab3.log
FB_HOST_PROF=1 timeout 300 python tools/run_once.py 4 2>&1 | tail -6 >> gpurun_out/ab3.log
timeout 600 python tools/share_one.py 8 3 3 2>&1 | tail -1 >> gpurun_out/ab3.log
timeout 600 python tools/share_one.py 1 0 2 2>&1 | tail -1 >> gpurun_out/ab3.log
cat gpurun_out/ab3.log
