#!/bin/bash
mkdir -p gpurun_out
for cfg in "1 128256 4096 14336" "2 128256 4096 14336" "1 128256 4096 14336"; do
  timeout 300 python tools/dbg_step2.py $cfg 2>&1 | tail -1
done
timeout 600 python tools/step_timeline.py llama-3.1-8b 8 2048 > gpurun_out/r2c_step_tl_8b_ctx2048.txt 2>&1; cat gpurun_out/r2c_step_tl_8b_ctx2048.txt | tail -30
