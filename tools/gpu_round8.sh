#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -4 | tee gpurun_out/pytest_gpu_r1h.log
timeout 300 python tools/bench_train.py --optimizer fused 2>&1 | tail -1 | tee gpurun_out/train_n1_fused.json
timeout 300 python bench.py --steps 50 --warmup 5 > gpurun_out/bench_n1_r1j.json; python tools/show_bench.py gpurun_out/bench_n1_r1j.json
