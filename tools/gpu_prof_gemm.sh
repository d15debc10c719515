#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mlp_gemm_tf32x3_pair -s 3 -c 1 -f -o gpurun_out/prof_gemm_pair python tools/bench_mlp.py > gpurun_out/ncu_gemm_pair.log 2>&1
# splits=4 launch is the 3rd kernel in bench order (splits 1,2,4...): capture the (32,1,4) grid explicitly
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mlp_gemm_tf32x3_pair -s 120 -c 1 -f -o gpurun_out/prof_gemm_pair4 python tools/bench_mlp.py > gpurun_out/ncu_gemm_pair4.log 2>&1
ls -la gpurun_out | grep prof_gemm_pair
