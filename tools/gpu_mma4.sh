#!/bin/bash
# Tensor-core fused kernel, nemb-16 instance (config 4 shape): parity (new cases, c4 fixture, training tests), bench c4.
mkdir -p gpurun_out
ARMNET_MMA=1 timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_train.py -x -q -m gpu -k "nemb16 or c4 or train or backward" > gpurun_out/pytest_mma4.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_mma4.log
tail -5 gpurun_out/pytest_mma4.log
ARMNET_MMA=0 timeout 300 python bench.py --workload c4 --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench_c4_fp32.json 2> gpurun_out/bench_c4.err
ARMNET_MMA=1 timeout 300 python bench.py --workload c4 --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench_c4_mma_w16.json 2>> gpurun_out/bench_c4.err
ARMNET_MMA=1 ARMNET_MMA_WARPS=12 timeout 300 python bench.py --workload c4 --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench_c4_mma_w12.json 2>> gpurun_out/bench_c4.err
for f in bench_c4_fp32 bench_c4_mma_w16 bench_c4_mma_w12; do echo $f; python tools/show_bench.py gpurun_out/$f.json 2>/dev/null | head -1; done
tail -3 gpurun_out/bench_c4.err
