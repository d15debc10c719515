#!/bin/bash
# 2-GPU box: tests, bench N=1/N=2 (torchrun), train.py smoke on synthetic data, ncu captures of the fused kernel.
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -4 | tee gpurun_out/pytest_gpu_r1b.log
timeout 600 python bench.py --steps 50 --warmup 5 | tee gpurun_out/bench_n1.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 50 --warmup 5 2>gpurun_out/bench_n2.err | tee gpurun_out/bench_n2.json
timeout 300 python bench.py --impl reference --steps 4 --warmup 1 | tee gpurun_out/bench_ref_r1b.json
echo "--- train.py 1 GPU"
timeout 600 python train.py --model armnet --nfield 39 --nfeat 100000 --dataset synthetic --synthetic_rows 65536 --epoch 2 --patience 2 --report_freq 8 --exp_name smoke1 --log_dir gpurun_out/log/ 2>&1 | grep -E "train|val|test|best|Total" | tail -8
echo "--- train.py 2 GPUs"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 train.py --model armnet --nfield 39 --nfeat 100000 --dataset synthetic --synthetic_rows 65536 --epoch 2 --patience 2 --report_freq 8 --exp_name smoke2 --log_dir gpurun_out/log/ 2>&1 | grep -E "train\s|val\s|test\s|best|Total" | tail -8
timeout 900 ncu --set full --clock-control none --import-source on -k regex:armnet_fwd_kernel -s 3 -c 1 -o gpurun_out/prof_fwd_v31 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_v31.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_v31.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
ls gpurun_out | head -40
