#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 1500 python -m pytest tests/test_gpu_train.py tests/test_gpu_parity.py -m gpu -x -q -k "train or backward or trajectory or flat_adam" 2>&1 | tail -4
timeout -s KILL 600 python tools/bench_train.py --steps 20 --warmup 5 > gpurun_out/r2q_train_c4.json 2> gpurun_out/r2q_train.err; cat gpurun_out/r2q_train_c4.json; tail -3 gpurun_out/r2q_train.err
bash tools/gpu_train_prof.sh 2>&1 | tail -24
