#!/bin/bash
# round 2, call B: first run of armnet_fwd_tmem_kernel -- parity tests under a hard timeout, then the headline bench
mkdir -p gpurun_out
timeout -s KILL 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "tensor_memory" 2>&1 | tail -40 > gpurun_out/r2b_tmem_tests.txt
cat gpurun_out/r2b_tmem_tests.txt
if grep -q "passed" gpurun_out/r2b_tmem_tests.txt && ! grep -q "failed" gpurun_out/r2b_tmem_tests.txt; then
  timeout -s KILL 600 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/r2b_bench_tmem.json 2> gpurun_out/r2b_bench_tmem.err
  tail -c 3000 gpurun_out/r2b_bench_tmem.json; tail -5 gpurun_out/r2b_bench_tmem.err
  ARMNET_TMEM=0 timeout -s KILL 600 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/r2b_bench_fp32.json 2> gpurun_out/r2b_bench_fp32.err
  tail -c 3000 gpurun_out/r2b_bench_fp32.json
fi
