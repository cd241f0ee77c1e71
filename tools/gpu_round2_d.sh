#!/bin/bash
mkdir -p gpurun_out
export FB_REQUIRE_GPU=1
timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_beam_wide.py -x -q 2>&1 | tail -12 > gpurun_out/new_tests.log
FB_BEAM_PROF=1 timeout 300 python tools/c3_probe.py 100000 50000 4 2>&1 | grep -v "^{" | head -3 > gpurun_out/c3_prof.log
timeout 1500 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
cat gpurun_out/new_tests.log gpurun_out/c3_prof.log; tail -5 gpurun_out/bench_n1.err; cat gpurun_out/bench_n1.json
