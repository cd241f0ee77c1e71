#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/c3_prof.log
for lib in libfloria_b200_alt.so; do
  FB_LIB=$PWD/floria_b200/$lib FB_BEAM_PROF=1 timeout 300 python tools/c3_probe.py 100000 50000 4 2>&1 | grep "k_beam_wide prof" | head -2 | cut -c1-700 >> gpurun_out/c3_prof.log
done
cat gpurun_out/c3_prof.log
