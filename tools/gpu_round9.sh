#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_train.py tests/test_gpu_mlp.py -q -s 2>&1 | grep -E "tcgen05 3xTF32|passed|failed|Error|assert" | tee gpurun_out/pytest_train.log
timeout 300 python tools/bench_train.py --optimizer fused 2>&1 | tail -1 | tee gpurun_out/train_n1_fused.json
