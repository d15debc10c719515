#!/bin/bash
# 1-GPU box: smoke, all gpu tests, MLP micro-bench, bench N=1.
mkdir -p gpurun_out
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -8 | tee gpurun_out/pytest_gpu_r1d.log
timeout 300 python tools/bench_mlp.py 2>&1 | tail -4
timeout 600 python bench.py --steps 50 --warmup 5 | tee gpurun_out/bench_n1_r1d.json
