#!/bin/bash
mkdir -p gpurun_out
N=${1:-2}
timeout -s KILL 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 50 --warmup 5 > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n$N.err
python tools/show_bench.py gpurun_out/r2_bench_n$N.json 2>/dev/null | head -1; tail -3 gpurun_out/r2_bench_n$N.err
python -c "
import json;d=json.load(open('gpurun_out/r2_bench_n$N.json'));print(json.dumps(d.get('train_c4'))[:800])"
