#!/bin/bash
mkdir -p gpurun_out
export ARMNET_B200_LIB=$PWD/armnet_b200/tuning/libc20.so
timeout -s KILL 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "tensor_memory" 2>&1 | tail -2
timeout -s KILL 600 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --no-train-leg --no-eager-leg > gpurun_out/r2o_bench_c20.json 2> gpurun_out/r2o.err
python tools/show_bench.py gpurun_out/r2o_bench_c20.json 2>/dev/null | head -1; tail -2 gpurun_out/r2o.err
