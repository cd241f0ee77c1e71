#!/bin/bash
mkdir -p gpurun_out
export FB_REQUIRE_GPU=1
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -12 > gpurun_out/all_tests.log
FB_BEAM_PROF=1 timeout 300 python tools/c3_probe.py 100000 50000 4 2>&1 | grep -v "^{" | head -3 > gpurun_out/c3_prof.log
FB_BEAM_PROF=1 timeout 300 python tools/run_once.py 2>&1 | tail -5 > gpurun_out/c1_prof.log
cat gpurun_out/all_tests.log gpurun_out/c3_prof.log gpurun_out/c1_prof.log
