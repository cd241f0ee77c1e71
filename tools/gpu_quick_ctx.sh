#!/bin/bash
mkdir -p gpurun_out
(timeout 600 python tools/share_ctx.py 8 0; timeout 900 python tools/share_ctx.py 1 0) 2>&1 | grep "^world" > gpurun_out/share_ctx.log
cat gpurun_out/share_ctx.log
