#!/bin/bash
mkdir -p gpurun_out
export FB_REQUIRE_GPU=1
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_multi.py tests/test_gpu_baseline_sizes.py -x -q -m gpu 2>&1 | tail -3 > gpurun_out/ab3.log
FB_HOST_PROF=1 timeout 600 python tools/share_one.py 1 0 2 2>&1 | tail -7 >> gpurun_out/ab3.log
timeout 600 python tools/share_one.py 8 3 3 2>&1 | tail -1 >> gpurun_out/ab3.log
cat gpurun_out/ab3.log
