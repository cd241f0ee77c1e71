#!/bin/bash
# quick GPU visit: parity tests + configs[2] sweep/hist timing (+ optional ncu capture with NCU=1)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
python tools/c3_once.py 100000 50000 4 5 2>&1 | tee gpurun_out/c3_p4.log
python tools/c3_once.py 100000 50000 2 5 2>&1 | tee gpurun_out/c3_p2.log
python tools/prof_phase.py 2>&1 | tee gpurun_out/phase.log
if [ -n "$NCU" ]; then
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_sweep|k_hist$' -c 2 -f -o gpurun_out/prof_c3 python tools/c3_once.py 100000 50000 4 1 > gpurun_out/prof_c3.log 2>&1
ncu -i gpurun_out/prof_c3.ncu-rep --page raw --csv > gpurun_out/prof_c3.raw.csv 2>/dev/null
ncu -i gpurun_out/prof_c3.ncu-rep --page source --csv > gpurun_out/prof_c3.source.csv 2>/dev/null
rm -f gpurun_out/prof_c3.ncu-rep
fi
