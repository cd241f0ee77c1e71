#!/bin/bash
mkdir -p gpurun_out
export FB_REQUIRE_GPU=1
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -3 > gpurun_out/all_tests.log
timeout 900 ncu --section SourceCounters --clock-control none --import-source on -k regex:k_beam_wide -c 1 -f -o gpurun_out/prof_kbw python tools/c3_probe.py 3000 50000 4 > gpurun_out/prof_kbw.log 2>&1
python tools/ncu_sass_dump.py gpurun_out/prof_kbw.ncu-rep gpurun_out/kbw_rows.csv
rm -f gpurun_out/prof_kbw.ncu-rep
xz -9 -f gpurun_out/kbw_rows.csv
cat gpurun_out/all_tests.log; ls -la gpurun_out/kbw_rows.csv.xz; tail -2 gpurun_out/prof_kbw.log | cut -c1-200
