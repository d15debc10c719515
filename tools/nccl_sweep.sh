#!/bin/bash
# all-reduce of the config-4 gradient bucket (72.8 MB) at N GPUs under different NCCL settings: tools/nccl_sweep.sh 8
N=${1:-8}
run() { echo "== $1"; env $1 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 tools/bench_train.py --steps 20 --warmup 5 2>/dev/null | tail -2; }
run "NCCL_DEBUG=WARN"
run "NCCL_ALGO=NVLS"
run "NCCL_ALGO=NVLSTree"
run "NCCL_ALGO=Ring"
run "NCCL_ALGO=Tree"
run "NCCL_NVLS_ENABLE=0"
