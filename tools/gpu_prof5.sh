#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:armnet_fwd_kernel -s 3 -c 1 -f -o gpurun_out/prof_fwd_v6 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_v6.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:mlp_gemm -s 3 -c 1 -f -o gpurun_out/prof_gemm_v6 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_gemm_v6.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_v6.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
ls -la gpurun_out | tail -5
