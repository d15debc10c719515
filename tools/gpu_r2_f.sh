#!/bin/bash
# round 2, call F: 16 consumer warps x 96 registers (tuning/libc16.so) against the 12 x 128 default, same box
mkdir -p gpurun_out
for lib in default c16; do
  if [ $lib = c16 ]; then export ARMNET_B200_LIB=$PWD/armnet_b200/tuning/libc16.so; fi
  timeout -s KILL 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "tensor_memory" 2>&1 | tail -2
  timeout -s KILL 600 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --no-train-leg --no-eager-leg > gpurun_out/r2f_bench_$lib.json 2> gpurun_out/r2f_bench_$lib.err
  python tools/show_bench.py gpurun_out/r2f_bench_$lib.json | head -1; tail -2 gpurun_out/r2f_bench_$lib.err
done
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:armnet_fwd_tmem --launch-skip 2 -c 1 \
     -o gpurun_out/r2f_tmem_c16_init -f python tools/prof_hot.py --regime init > gpurun_out/r2f_ncu.log 2>&1
