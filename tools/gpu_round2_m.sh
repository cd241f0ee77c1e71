#!/bin/bash
mkdir -p gpurun_out
export FB_REQUIRE_GPU=1
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_multi.py tests/test_gpu_baseline_sizes.py tests/test_config0_real_data.py -x -q -m gpu 2>&1 | tail -5 > gpurun_out/some_tests.log
FB_HOST_PROF=1 timeout 900 python tools/scale_run.py c5 500 2>&1 | grep "pass 2\|host ms" | tail -2 > gpurun_out/waves.log
timeout 300 python tools/run_once.py 3 2>&1 | tail -1 >> gpurun_out/waves.log
cat gpurun_out/some_tests.log gpurun_out/waves.log
