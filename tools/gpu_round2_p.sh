#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/h3.log
for lib in libfloria_b200.so libfloria_b200_alt.so; do
  echo "== $lib" >> gpurun_out/h3.log
  FB_LIB=$PWD/floria_b200/$lib timeout 300 python tools/c3_once.py 100000 50000 4 7 2>&1 | tail -1 >> gpurun_out/h3.log
  FB_LIB=$PWD/floria_b200/$lib timeout 300 python tools/run_once.py 4 2>&1 | tail -2 >> gpurun_out/h3.log
done
FB_LIB=$PWD/floria_b200/libfloria_b200_alt.so FB_REQUIRE_GPU=1 timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -2 >> gpurun_out/h3.log
python - <<'PY' >> gpurun_out/h3.log
import __graft_entry__ as g
g.smoke()
print("smoke ok")
PY
cat gpurun_out/h3.log
