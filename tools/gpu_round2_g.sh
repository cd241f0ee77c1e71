#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/ab.log
for lib in floria_b200/libfloria_b200.so floria_b200/libfloria_b200_v1.so; do
  echo "=== $lib" >> gpurun_out/ab.log
  FB_LIB=$PWD/$lib timeout 300 python tools/run_once.py 4 2>&1 | tail -3 >> gpurun_out/ab.log
  FB_LIB=$PWD/$lib timeout 300 python tools/c3_probe.py 100000 50000 4 2>&1 | grep "^rep" | cut -c1-60 >> gpurun_out/ab.log
  FB_LIB=$PWD/$lib timeout 600 python tools/scale_run.py c5 100 2>&1 | grep "pass 2" >> gpurun_out/ab.log
done
cat gpurun_out/ab.log
