#!/bin/bash
# first GPU pass of round 2: box facts, the new wide-beam tests, the old suite, a configs[2] timing probe
mkdir -p gpurun_out
{ nproc; free -g | head -2; lscpu | grep -i "model name\|socket\|thread" ; nvidia-smi --query-gpu=name,memory.total --format=csv; } > gpurun_out/box.txt 2>&1
export FB_REQUIRE_GPU=1
timeout 1500 python -m pytest tests/test_gpu_beam_wide.py -x -q 2>&1 | tail -15 > gpurun_out/wide_tests.log
timeout 600 python tools/c3_probe.py > gpurun_out/c3_probe.log 2>&1
timeout 1200 python -m pytest tests -x -q -m gpu --deselect tests/test_gpu_beam_wide.py 2>&1 | tail -8 > gpurun_out/gpu_tests.log
tail -5 gpurun_out/wide_tests.log gpurun_out/gpu_tests.log; cat gpurun_out/c3_probe.log
