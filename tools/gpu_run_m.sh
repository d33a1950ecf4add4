#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_decode.py -x -q 2>&1 | tail -4 | tee gpurun_out/r2m_pytest.log
timeout 900 python tools/ab_decode.py llama-3.1-8b 8 2048 --timeline "rw_ksplit=0" "rw_ksplit=8" "rw_ksplit=4" "rw_ksplit=2" "rw_ksplit=8,rw_kb=1" "rw_ksplit=8,rw_kb=2" "rw_ksplit=8,rw_kb=8" "rw_ksplit=0" > gpurun_out/r2m_ab.txt 2>&1; cat gpurun_out/r2m_ab.txt | tail -20
timeout 600 python tools/ab_decode.py llama-3.2-1b 16 256 --timeline "rw_ksplit=0" "rw_ksplit=8" > gpurun_out/r2m_ab_1b.txt 2>&1; cat gpurun_out/r2m_ab_1b.txt | tail -4
