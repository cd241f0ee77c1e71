#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python tools/share_balance.py 8 2>&1 | tail -10 > gpurun_out/balance.log
cat gpurun_out/balance.log
