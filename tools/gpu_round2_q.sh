#!/bin/bash
# source-level stall samples of k_beam<256> on configs[1]
mkdir -p gpurun_out
timeout 900 ncu --section SourceCounters --section WarpStateStats --section SchedulerStats --clock-control none --import-source on -k regex:k_beam -s 1 -c 1 -f -o gpurun_out/prof_kbeam python tools/run_once.py 2 > gpurun_out/prof_kbeam.log 2>&1
python tools/ncu_source_top.py gpurun_out/prof_kbeam.ncu-rep gpurun_out/kbeam_src 400
rm -f gpurun_out/prof_kbeam.ncu-rep
tail -3 gpurun_out/prof_kbeam.log; head -30 gpurun_out/kbeam_src.sass.txt | cut -c1-250
