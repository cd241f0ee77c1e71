#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e > gpurun_out/bq.json 2> gpurun_out/bq.err
tail -3 gpurun_out/bq.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/bq.json"))
print("c3", d["ms_per_step"], "| c1", d["configs1"]["ms_per_step"], "| shard500", d["shard500"]["ms_per_step"], d["shard500"]["e2e"]["ms_per_step"], d["shard500"].get("gpu_launches"), "| shard_blocks", d["shard_blocks"])
PY
for t in 0 1; do echo "FB_SWEEP_TMA=$t"; FB_SWEEP_TMA=$t timeout 300 python tools/c3_once.py 100000 50000 4 5 2>&1 | tail -1; done
