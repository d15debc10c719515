#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_train.py -q -x 2>&1 | tail -15
for o in torch fused; do timeout 300 python tools/bench_train.py --optimizer $o 2>&1 | tail -1 | tee gpurun_out/train_n1_$o.json; done
