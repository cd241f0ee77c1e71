#!/bin/bash
# beam-focused GPU visit: beam/phase parity tests + phase profile
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_golden.py -m gpu -x -q -k "beam or phase or golden" 2>&1 | tail -3
FB_BEAM_PROF=1 python tools/run_once.py 2 2>&1 | grep "k_beam prof" | tail -1
python tools/prof_phase.py 2>&1 | tail -3
