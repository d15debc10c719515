#!/bin/bash
# round 2, call C: headline bench with the tensor-memory kernel (default) and with it forced off, same box
mkdir -p gpurun_out
timeout -s KILL 600 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/r2c_bench_tmem.json 2> gpurun_out/r2c_bench_tmem.err
python tools/show_bench.py gpurun_out/r2c_bench_tmem.json; tail -3 gpurun_out/r2c_bench_tmem.err
ARMNET_TMEM=0 timeout -s KILL 600 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/r2c_bench_fp32.json 2> gpurun_out/r2c_bench_fp32.err
python tools/show_bench.py gpurun_out/r2c_bench_fp32.json
timeout -s KILL 600 python bench.py --workload c2b --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/r2c_bench_c2b_tmem.json 2>/dev/null
python tools/show_bench.py gpurun_out/r2c_bench_c2b_tmem.json
