#!/bin/bash
# quick A/B on the GPU box: tmem parity tests, trace of the tuned build, bench (C2a) of the default library
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -x -q -m gpu  2>&1 | tail -3
ARMNET_B200_LIB=$PWD/armnet_b200/tuning/libtrace.so python tools/trace_tmem.py --regime init 2>&1 | tail -26
python bench.py --no-train-leg --no-eager-leg --no-cpu-baseline --steps 50 --warmup 5 > gpurun_out/q_bench.json 2> gpurun_out/q_bench.err
python tools/show_bench.py gpurun_out/q_bench.json
