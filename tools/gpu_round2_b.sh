#!/bin/bash
mkdir -p gpurun_out
export FB_REQUIRE_GPU=1
timeout 600 python -m pytest tests/test_gpu_beam_wide.py -x -q -k golden 2>&1 | tail -5 > gpurun_out/golden_test.log
for g in 148 74 37; do
  echo "== grid $g" >> gpurun_out/c3_prof.log
  FB_BEAM_WIDE_GRID=$g FB_BEAM_PROF=1 timeout 600 python tools/c3_probe.py 100000 50000 4 2>&1 | grep -v "^{" >> gpurun_out/c3_prof.log
done
cat gpurun_out/golden_test.log gpurun_out/c3_prof.log
