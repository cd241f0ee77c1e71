#!/bin/bash
mkdir -p gpurun_out
for lib in libfloria_b200.so libfloria_b200_alt.so; do
  echo "== $lib" >> gpurun_out/ab2.log
  FB_LIB=$PWD/floria_b200/$lib timeout 300 python tools/run_once.py 8 2>&1 | tail -6 >> gpurun_out/ab2.log
done
FB_HOST_PROF=1 timeout 300 python tools/run_once.py 3 2>&1 | tail -8 >> gpurun_out/ab2.log
cat gpurun_out/ab2.log
