#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/wave.log
for cfg in "3 1" "4 2" "4 1" "2 2" "5 1"; do
  set -- $cfg
  echo "== FB_PLOIDY_WAVE=$1 FB_PLOIDY_STEP=$2" >> gpurun_out/wave.log
  FB_PLOIDY_WAVE=$1 FB_PLOIDY_STEP=$2 timeout 600 python tools/share_one.py 8 3 2 2>&1 | tail -1 | cut -c1-120 >> gpurun_out/wave.log
  FB_PLOIDY_WAVE=$1 FB_PLOIDY_STEP=$2 timeout 600 python tools/share_one.py 1 0 2 2>&1 | tail -1 | cut -c1-120 >> gpurun_out/wave.log
done
cat gpurun_out/wave.log
