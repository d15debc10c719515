#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -5 | tee gpurun_out/pytest_gpu_r1g.log
echo "--- train.py (FlatAdam default), synthetic data, 1 GPU"
timeout 600 python train.py --model armnet --nfield 39 --nfeat 100000 --dataset synthetic --synthetic_rows 65536 --epoch 2 --patience 2 --report_freq 100 --exp_name flat1 --log_dir gpurun_out/log/ 2>&1 | grep -E "train\s|val\s|test\s|Total" | tail -5
echo "--- same with --torch_adam"
timeout 600 python train.py --model armnet --nfield 39 --nfeat 100000 --dataset synthetic --synthetic_rows 65536 --epoch 2 --patience 2 --report_freq 100 --torch_adam --exp_name torch1 --log_dir gpurun_out/log/ 2>&1 | grep -E "train\s|val\s|test\s|Total" | tail -5
