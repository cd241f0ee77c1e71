#!/bin/bash
# every GPU test + the N=1 bench line (both arms)
mkdir -p gpurun_out
export FB_REQUIRE_GPU=1
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -5 > gpurun_out/all_tests.log
timeout 1200 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
timeout 900 python bench.py --impl reference > gpurun_out/bench_n1_ref.json 2> gpurun_out/bench_n1_ref.err
cat gpurun_out/all_tests.log; tail -3 gpurun_out/bench_n1.err; cat gpurun_out/bench_n1.json | cut -c1-3000; cat gpurun_out/bench_n1_ref.json | cut -c1-800
