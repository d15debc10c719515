#!/bin/bash
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 4 --steps 50 --warmup 5 2>gpurun_out/bench_n4.err | tail -1 | tee gpurun_out/bench_n4.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('N=4 value %.2fM e2e %.2fM ms %.4f' % (d['value']/1e6, d['e2e']['value']/1e6, d['ms_per_step']))"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29532 bench.py --impl reference --gpus 4 --steps 2 --warmup 1 2>/dev/null | tail -1 | cut -c1-200
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29533 tools/bench_train.py --optimizer fused 2>/dev/null | tail -1 | tee gpurun_out/train_n4_fused.json
tail -3 gpurun_out/bench_n4.err
