#!/bin/bash
mkdir -p gpurun_out
export FB_REQUIRE_GPU=1
rm -f gpurun_out/rot.log
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_multi.py tests/test_gpu_baseline_sizes.py -x -q -m gpu 2>&1 | tail -3 >> gpurun_out/rot.log
for r in 0 1; do
  echo "== FB_BEAM_ROT=$r" >> gpurun_out/rot.log
  FB_BEAM_ROT=$r FB_HOST_PROF=1 timeout 600 python tools/share_one.py 8 3 2 2>&1 | grep "run_beam\|world" | tail -5 | cut -c1-150 >> gpurun_out/rot.log
  FB_BEAM_ROT=$r timeout 600 python tools/share_one.py 1 0 2 2>&1 | tail -1 | cut -c1-150 >> gpurun_out/rot.log
done
cat gpurun_out/rot.log
