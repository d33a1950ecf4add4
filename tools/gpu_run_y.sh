#!/bin/bash
# final round-2 artefacts: bench line, launch list + ncu --set full of the decode kernels, timeline, 1B / Qwen2 A-B
mkdir -p gpurun_out
timeout 1500 python bench.py > gpurun_out/r2y_bench_8b.json 2> gpurun_out/r2y_bench_8b.err; tail -2 gpurun_out/r2y_bench_8b.err; head -c 600 gpurun_out/r2y_bench_8b.json; echo
bash tools/gpu_profile.sh
timeout 600 python tools/timeline.py llama-3.1-8b 8 2048 > gpurun_out/r2y_timeline_8b_ctx2048.txt 2>&1; grep -A10 "per-kernel-kind" gpurun_out/r2y_timeline_8b_ctx2048.txt | head -11
