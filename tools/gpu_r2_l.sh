#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 400 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "tensor_memory" 2>&1 | tail -3
for lib in default u5; do
  if [ $lib = u5 ]; then export ARMNET_B200_LIB=$PWD/armnet_b200/tuning/libu5.so; fi
  timeout -s KILL 600 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --no-train-leg --no-eager-leg > gpurun_out/r2l_bench_$lib.json 2> gpurun_out/r2l_bench_$lib.err
  python tools/show_bench.py gpurun_out/r2l_bench_$lib.json | head -1; tail -2 gpurun_out/r2l_bench_$lib.err
done
unset ARMNET_B200_LIB
for regime in init trained; do
  timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:armnet_fwd_tmem --launch-skip 2 -c 1 \
     -o gpurun_out/r2l_tmem_${regime} -f python tools/prof_hot.py --regime $regime > gpurun_out/r2l_ncu_${regime}.log 2>&1
done
