#!/bin/bash
mkdir -p gpurun_out
FB_HOST_PROF=1 timeout 900 python tools/scale_run.py c5 500 2>&1 | tail -14 > gpurun_out/c5_500.log
cat gpurun_out/c5_500.log
