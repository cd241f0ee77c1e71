#!/bin/bash
mkdir -p gpurun_out
export FB_REQUIRE_GPU=1
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_bench_helpers.py tests/test_golden.py -x -q -m gpu 2>&1 | tail -5 > gpurun_out/sweep_tests.log
for v in 1 0; do echo "FB_SWEEP_PAIR=$v" >> gpurun_out/sweep_ab.log; for p in 4 2; do FB_SWEEP_PAIR=$v timeout 300 python tools/c3_once.py 100000 50000 $p 5 2>&1 | tail -1 >> gpurun_out/sweep_ab.log; done; done
cat gpurun_out/sweep_tests.log gpurun_out/sweep_ab.log
