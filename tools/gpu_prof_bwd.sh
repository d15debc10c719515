#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k armnet_fwd_kernel -s 5 -c 1 -f -o gpurun_out/prof_bwd python tools/bench_train.py --steps 2 --warmup 2 > gpurun_out/ncu_bwd.log 2>&1
ls -la gpurun_out/prof_bwd.ncu-rep; tail -3 gpurun_out/ncu_bwd.log
