#!/bin/bash
# one GPU-box visit: parity tests, bench line, CPU arm, ncu launch list + full captures (outputs under gpurun_out/)
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 600 gpurun_out/bench.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2>> gpurun_out/bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-c3 > gpurun_out/bench_under_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_sweep|k_hist$' -c 2 -f -o gpurun_out/prof_c3 python tools/c3_once.py 100000 50000 4 1 > gpurun_out/prof_c3.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:k_beam -c 1 -f -o gpurun_out/prof_beam python tools/run_once.py 1 > gpurun_out/prof_beam.log 2>&1
# the .ncu-rep files embed the whole module (40+ MB each): export the pages that are read later and drop the reports
for r in prof_c3 prof_beam; do
  ncu -i gpurun_out/$r.ncu-rep --page raw --csv > gpurun_out/$r.raw.csv 2>/dev/null
  rm -f gpurun_out/$r.ncu-rep
done
python tools/c3_once.py 100000 50000 4 5 > gpurun_out/c3_p4.log 2>&1
python tools/c3_once.py 100000 50000 2 5 > gpurun_out/c3_p2.log 2>&1
cat gpurun_out/c3_p4.log gpurun_out/c3_p2.log
cat gpurun_out/bench.json
