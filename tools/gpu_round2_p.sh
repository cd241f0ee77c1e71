#!/bin/bash
mkdir -p gpurun_out
export FB_REQUIRE_GPU=1
rm -f gpurun_out/c3_prof.log
timeout 900 python -m pytest tests/test_gpu_beam_wide.py -x -q -m gpu 2>&1 | tail -3 > gpurun_out/p_tests.log
for lib in libfloria_b200.so; do
  FB_LIB=$PWD/floria_b200/$lib FB_BEAM_PROF=1 timeout 300 python tools/c3_probe.py 100000 50000 4 2>&1 | grep "k_beam_wide prof" | head -2 | cut -c1-420 >> gpurun_out/c3_prof.log
  FB_LIB=$PWD/floria_b200/$lib timeout 300 python tools/c3_probe.py 100000 50000 4 2>&1 | grep "^rep" | cut -c1-30 >> gpurun_out/c3_prof.log
done
cat gpurun_out/p_tests.log gpurun_out/c3_prof.log
