#!/bin/bash
# 1-GPU box: all gpu tests, bench N=1, full ncu captures of the three top kernels.
mkdir -p gpurun_out
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -8 | tee gpurun_out/pytest_gpu_r1d.log
timeout 300 python tools/bench_mlp.py 2>&1 | tail -12
timeout 600 python bench.py --steps 50 --warmup 5 | tee gpurun_out/bench_n1_r1d.json
for k in mlp_tail_kernel; do
timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 3 -c 1 -f -o gpurun_out/prof_$k python tools/bench_mlp.py > gpurun_out/ncu_$k.log 2>&1
done
ls -la gpurun_out | tail -8
