#!/bin/bash
# round-2 profiles: ncu launch list of the bench command (configs[2] leg), `ncu --set full` captures of the dominant kernel
# (k_beam_wide) and of the two HBM-bound kernels on the 100k x 50k block; summaries exported on the box (the reports
# embed the module and exceed the transfer limit)
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_configs2.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e --no-secondary > gpurun_out/bench_under_ncu.log 2>&1
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:k_beam_wide -c 1 -f -o gpurun_out/prof_beam_wide python tools/c3_probe.py 100000 50000 4 > gpurun_out/prof_beam_wide.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_sweep|k_hist$' -c 2 -f -o gpurun_out/prof_c3 python tools/c3_once.py 100000 50000 4 1 > gpurun_out/prof_c3.log 2>&1
for r in prof_beam_wide prof_c3; do
  ncu -i gpurun_out/$r.ncu-rep --page raw --csv > gpurun_out/$r.raw.csv 2>/dev/null
  rm -f gpurun_out/$r.ncu-rep
done
python tools/ncu_export.py gpurun_out/r02_ncu_full_summaries.json k_beam_wide=gpurun_out/prof_beam_wide.raw.csv c3=gpurun_out/prof_c3.raw.csv > /dev/null
tail -3 gpurun_out/prof_beam_wide.log; head -c 1500 gpurun_out/r02_ncu_full_summaries.json; tail -5 gpurun_out/r02_launches_configs2.csv | cut -c1-200
