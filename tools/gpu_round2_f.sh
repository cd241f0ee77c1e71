#!/bin/bash
mkdir -p gpurun_out
FB_BEAM_PROF=1 timeout 300 python tools/run_once.py 3 2>&1 | tail -8 > gpurun_out/c1_prof.log
timeout 600 python tools/scale_run.py c5 100 2>&1 | tail -6 > gpurun_out/c5_100.log
cat gpurun_out/c1_prof.log gpurun_out/c5_100.log
