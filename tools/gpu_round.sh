#!/bin/bash
# One gpurun call: GPU tests, bench, ncu launch list, ncu full capture of the fused kernel. Outputs under gpurun_out/.
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 50 --warmup 5 2> gpurun_out/bench.err | tee gpurun_out/bench.json
tail -3 gpurun_out/bench.err
timeout 300 python bench.py --impl reference --steps 4 --warmup 1 2>> gpurun_out/bench.err | tee gpurun_out/bench_ref.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:armnet_fwd_kernel -s 3 -c 2 \
    -o gpurun_out/prof_fwd python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out
