#!/bin/bash
mkdir -p gpurun_out
timeout 120 python -m pytest tests/test_gpu_mlp.py -q -s -x 2>&1 | tail -30 | tee gpurun_out/pytest_mlp.log
timeout 120 python tools/bench_mlp.py 2>&1 | tail -12 | tee gpurun_out/bench_mlp.log
ARMNET_GEMM_1CTA=1 timeout 120 python tools/bench_mlp.py 2>&1 | tail -12 | head -6
