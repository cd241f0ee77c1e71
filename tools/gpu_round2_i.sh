#!/bin/bash
mkdir -p gpurun_out
export FB_REQUIRE_GPU=1
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -6 > gpurun_out/all_tests.log
timeout 900 python tools/final_hapq_run.py 0.25 > gpurun_out/final_hapq.log 2>&1
bash tools/gpu_round2_prof.sh > gpurun_out/prof_script.log 2>&1
cat gpurun_out/all_tests.log gpurun_out/final_hapq.log; tail -20 gpurun_out/prof_script.log | cut -c1-400
