#!/bin/bash
mkdir -p gpurun_out
for k in 2 3 4 5; do
timeout 300 python tools/timeline.py llama-3.1-8b 8 2048 --opt=cta_trace=$k --cta > gpurun_out/p_cta_$k.txt 2>&1
echo "== kind $k"; grep -A30 "per-CTA" gpurun_out/p_cta_$k.txt | (head -4; tail -6)
done
