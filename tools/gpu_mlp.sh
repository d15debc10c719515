#!/bin/bash
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_mlp.py -q -x 2>&1 | tail -4 | tee gpurun_out/pytest_mlp.log
timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/bench_n1_r1k.json; python tools/show_bench.py gpurun_out/bench_n1_r1k.json 2>/dev/null | head -1
