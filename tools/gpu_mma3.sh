#!/bin/bash
# Tensor-core fused kernel, third pass (plain / debug instances, 16 vs 12 warps): parity, then bench on the same box.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "tensor_core" > gpurun_out/pytest_mma3.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_mma3.log
ARMNET_MMA=1 timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "fused_stages or plain_call or module_forward or batchnorm_epilogue or full_size" > gpurun_out/pytest_mma3_golden.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_mma3_golden.log
tail -5 gpurun_out/pytest_mma3.log; tail -3 gpurun_out/pytest_mma3_golden.log
ARMNET_MMA=1 timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/bench_n1_mma3_w16.json 2> gpurun_out/bench_n1_mma3.err
ARMNET_MMA=1 ARMNET_MMA_WARPS=12 timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/bench_n1_mma3_w12.json 2>/dev/null
timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/bench_n1_mma3_fp32.json 2>/dev/null
for f in bench_n1_mma3_w16 bench_n1_mma3_w12 bench_n1_mma3_fp32; do echo $f; python tools/show_bench.py gpurun_out/$f.json | head -1; done
