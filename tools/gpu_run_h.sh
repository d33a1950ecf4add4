#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ps_k_step -s 3 -c 1 -f -o gpurun_out/r2h_step python tools/prof_step.py llama-3.1-8b 4 2048 6 > gpurun_out/r2h_ncu.log 2>&1
tail -5 gpurun_out/r2h_ncu.log; ls -la gpurun_out/r2h_step.ncu-rep
