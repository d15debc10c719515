#!/bin/bash
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 tools/bench_train.py --optimizer fused 2>/dev/null | tail -1 | tee gpurun_out/train_n2_fused.json
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29523 bench.py --gpus 2 --steps 50 --warmup 5 2>/dev/null | tail -1 > gpurun_out/bench_n2_r1j.json; python tools/show_bench.py gpurun_out/bench_n2_r1j.json
echo "--- train.py 2 GPUs (FlatAdam, device-resident synthetic data)"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29524 train.py --model armnet --nfield 39 --nfeat 100000 --dataset synthetic --synthetic_rows 65536 --epoch 2 --patience 2 --report_freq 100 --exp_name dp2 --log_dir gpurun_out/log/ 2>&1 | grep -E "train\s|val\s|test\s|Total" | tail -5
