#!/bin/bash
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 tools/bench_train.py --optimizer fused 2>/dev/null | tail -1 | tee gpurun_out/train_n2_fused.json
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29522 tools/bench_train.py --optimizer torch 2>/dev/null | tail -1 | tee gpurun_out/train_n2_torch.json
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29523 bench.py --gpus 2 --steps 50 --warmup 5 2>/dev/null | tail -1 | tee gpurun_out/bench_n2_r1d.json
