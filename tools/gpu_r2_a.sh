#!/bin/bash
# round 2, call A: tcgen05 building-block check + small-shape micro-benchmark (modes 4/5) + the new host-side tests
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2a_smi.txt 2>&1
timeout 60 tools/ubench/tmem_logits_check > gpurun_out/r2a_tmem_logits_check.txt 2>&1; echo "check rc=$?" >> gpurun_out/r2a_tmem_logits_check.txt
timeout 120 tools/ubench/tcgen05_small > gpurun_out/r2a_tcgen05_small.txt 2>&1; echo "ubench rc=$?" >> gpurun_out/r2a_tcgen05_small.txt
cat gpurun_out/r2a_tmem_logits_check.txt gpurun_out/r2a_tcgen05_small.txt
timeout 900 python -m pytest tests/test_gpu_train.py tests/test_gpu_mlp.py -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/r2a_pytest.txt
