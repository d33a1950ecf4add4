#!/bin/bash
mkdir -p gpurun_out
for p in tiny-bigvocab tiny-deep; do
  timeout 300 python tools/dbg_step.py $p 2>&1 | tail -4
done
timeout 600 compute-sanitizer --tool memcheck --print-limit 5 python tools/dbg_step.py tiny-bigvocab 2>&1 | grep -v "^$" | head -60 > gpurun_out/r2d_sanitizer.txt; head -50 gpurun_out/r2d_sanitizer.txt
