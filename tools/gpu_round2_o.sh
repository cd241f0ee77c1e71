#!/bin/bash
mkdir -p gpurun_out
export FB_REQUIRE_GPU=1
timeout 900 python -m pytest tests/test_gpu_beam_wide.py tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -6 > gpurun_out/pipe_tests.log
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu --no-secondary > gpurun_out/bench_pipe.json 2> gpurun_out/bench_pipe.err
cat gpurun_out/pipe_tests.log; tail -3 gpurun_out/bench_pipe.err; python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_pipe.json"))
print("c3 resident", d["ms_per_step"], "e2e", d["e2e"])
PY
