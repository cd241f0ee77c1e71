#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_n1_ref.json 2>> gpurun_out/bench_n1.err
tail -3 gpurun_out/bench_n1.err; cat gpurun_out/bench_n1.json | cut -c1-3000; echo; cat gpurun_out/bench_n1_ref.json | cut -c1-600
