#!/bin/bash
mkdir -p gpurun_out
export FB_REQUIRE_GPU=1
rm -f gpurun_out/pipe_sms.log
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -5 > gpurun_out/all_tests.log
for n in 8 16 32; do
  FB_PIPE_SMS=$n timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu --no-secondary > gpurun_out/bp.json 2> gpurun_out/bp.err
  python - <<PY >> gpurun_out/pipe_sms.log
import json
d = json.load(open("gpurun_out/bp.json"))
print("FB_PIPE_SMS=$n resident", round(d["ms_per_step"],1), "e2e", round(d["e2e"]["ms_per_step"],1), d["e2e"]["same_result"])
PY
done
FB_PIPELINE_UPLOAD=0 timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu --no-secondary > gpurun_out/bp.json 2> gpurun_out/bp.err
python - <<PY >> gpurun_out/pipe_sms.log
import json
d = json.load(open("gpurun_out/bp.json"))
print("no pipeline: resident", round(d["ms_per_step"],1), "e2e", round(d["e2e"]["ms_per_step"],1), d["e2e"]["same_result"])
PY
cat gpurun_out/all_tests.log gpurun_out/pipe_sms.log
