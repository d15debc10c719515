#!/bin/bash
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | wc -l
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 50 --warmup 5 2>gpurun_out/bench_n8.err | tail -1 > gpurun_out/bench_n8.json; python tools/show_bench.py gpurun_out/bench_n8.json | head -1
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29542 bench.py --impl reference --gpus 8 --steps 2 --warmup 1 2>/dev/null | tail -1 | cut -c1-160
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29543 tools/bench_train.py --optimizer fused 2>/dev/null | tail -1 | tee gpurun_out/train_n8_fused.json | cut -c1-400
