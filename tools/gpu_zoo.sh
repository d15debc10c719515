#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -6 | tee gpurun_out/pytest_gpu_r1f.log
timeout 300 python tools/bench_zoo.py 2>&1 | tail -1 | tee gpurun_out/bench_zoo.json
