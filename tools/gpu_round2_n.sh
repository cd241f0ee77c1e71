#!/bin/bash
mkdir -p gpurun_out
export FB_REQUIRE_GPU=1
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -6 > gpurun_out/all_tests.log
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err
cat gpurun_out/all_tests.log; python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_quick.json"))
print("c3", d["ms_per_step"], "configs1", d["configs1"]["ms_per_step"], d["configs1"]["e2e_ms_per_step"], "shard500", d["shard500"]["ms_per_step"], d["shard500"]["e2e"]["ms_per_step"])
PY
