#!/bin/bash
# 1-GPU box: smoke, gpu tests, bench N=1, reference arm, launch list.
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3 | tee gpurun_out/smoke_r1c.log
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -25 | tee gpurun_out/pytest_gpu_r1c.log
timeout 600 python bench.py --steps 50 --warmup 5 | tee gpurun_out/bench_n1_r1c.json
