#!/bin/bash
mkdir -p gpurun_out
export FB_REQUIRE_GPU=1
timeout 600 python -m pytest tests/test_gpu_final_parts.py tests/test_twin_breakers.py tests/test_c_abi.py -x -q -m gpu 2>&1 | tail -4 > gpurun_out/new_tests.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_beam -c 1 -f -o gpurun_out/src_beam python tools/run_once.py 1 > gpurun_out/src_beam.log 2>&1
python tools/ncu_source_top.py gpurun_out/src_beam.ncu-rep gpurun_out/src_beam 80
rm -f gpurun_out/src_beam.ncu-rep
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_beam_wide -c 1 -f -o gpurun_out/src_wide python tools/c3_probe.py 8000 50000 4 > gpurun_out/src_wide.log 2>&1
python tools/ncu_source_top.py gpurun_out/src_wide.ncu-rep gpurun_out/src_wide 80
rm -f gpurun_out/src_wide.ncu-rep
cat gpurun_out/new_tests.log; head -30 gpurun_out/src_beam.cuda.txt | cut -c1-250
