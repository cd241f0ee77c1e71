#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/p_tests.log
for lib in libfloria_b200.so libfloria_b200_alt.so libfloria_b200.so libfloria_b200_alt.so; do
  echo "== $lib" >> gpurun_out/p_tests.log
  FB_LIB=$PWD/floria_b200/$lib timeout 300 python tools/c3_probe.py 100000 50000 4 2>&1 | grep "^rep 1" | cut -c1-30 >> gpurun_out/p_tests.log
  FB_LIB=$PWD/floria_b200/$lib timeout 300 python tools/run_once.py 5 2>&1 | tail -2 >> gpurun_out/p_tests.log
done
cat gpurun_out/p_tests.log
