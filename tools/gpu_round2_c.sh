#!/bin/bash
mkdir -p gpurun_out
export FB_REQUIRE_GPU=1
timeout 900 python -m pytest tests/test_gpu_beam_wide.py -x -q 2>&1 | tail -8 > gpurun_out/wide_tests.log
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_golden.py tests/test_config0_real_data.py -x -q -m gpu 2>&1 | tail -8 > gpurun_out/parity_tests.log
FB_BEAM_PROF=1 timeout 300 python tools/c3_probe.py 100000 50000 4 2>&1 | grep -v "^{" > gpurun_out/c3_prof.log
cat gpurun_out/wide_tests.log gpurun_out/parity_tests.log gpurun_out/c3_prof.log
