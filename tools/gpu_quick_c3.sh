#!/bin/bash
mkdir -p gpurun_out
export FB_REQUIRE_GPU=1
timeout 900 python -m pytest tests/test_gpu_beam_wide.py -x -q 2>&1 | tail -3 > gpurun_out/wide_tests.log
timeout 300 python tools/c3_probe.py 100000 50000 4 2>&1 | grep "^rep" | cut -c1-60 > gpurun_out/c3_quick.log
FB_BEAM_PROF=1 timeout 300 python tools/c3_probe.py 100000 50000 4 2>&1 | grep prof | head -1 >> gpurun_out/c3_quick.log
cat gpurun_out/wide_tests.log gpurun_out/c3_quick.log
