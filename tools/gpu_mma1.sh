#!/bin/bash
# Tensor-core fused kernel: parity tests, then bench.py with the MMA kernel (16 and 12 warps) and with the FP32 kernel.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/pytest_mma.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_mma.log
tail -15 gpurun_out/pytest_mma.log
ARMNET_MMA=1 timeout 300 python bench.py --steps 50 --warmup 5 > gpurun_out/bench_n1_mma.json 2> gpurun_out/bench_n1_mma.err
ARMNET_MMA=1 ARMNET_FORCE_NW=12 timeout 300 python bench.py --steps 50 --warmup 5 > gpurun_out/bench_n1_mma_nw12.json 2>/dev/null
timeout 300 python bench.py --steps 50 --warmup 5 > gpurun_out/bench_n1_nomma.json 2>/dev/null
for f in bench_n1_mma bench_n1_mma_nw12 bench_n1_nomma; do echo $f; python tools/show_bench.py gpurun_out/$f.json; done
tail -3 gpurun_out/bench_n1_mma.err
