#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/tp_timeline.py 2 4 2048 p2p > gpurun_out/tp2_timeline.txt 2>&1; cat gpurun_out/tp2_timeline.txt | cut -c1-200 | head -40
