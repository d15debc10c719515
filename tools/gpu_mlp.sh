#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_mlp.py -q -s 2>&1 | tail -30 | tee gpurun_out/pytest_mlp.log
timeout 300 python tools/bench_mlp.py 2>&1 | tail -12 | tee gpurun_out/bench_mlp.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mlp_tail_kernel -s 3 -c 1 -f -o gpurun_out/prof_mlp_tail_kernel python tools/bench_mlp.py > gpurun_out/ncu_mlp_tail_kernel.log 2>&1
