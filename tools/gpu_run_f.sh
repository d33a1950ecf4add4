#!/bin/bash
for i in 1 2 3; do timeout 300 python tools/dbg_step2.py 1 128256 4096 14336 2>&1 | grep -v Traceback | tail -1; done
timeout 300 python tools/dbg_step2.py 2 128256 4096 14336 2>&1 | grep -v Traceback | tail -1
timeout 600 python tools/step_timeline.py llama-3.1-8b 8 2048 > gpurun_out/r2c_step_tl_8b_ctx2048.txt 2>&1; cat gpurun_out/r2c_step_tl_8b_ctx2048.txt | tail -32
