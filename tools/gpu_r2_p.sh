#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2p_pytest_gpu.log; tail -6 gpurun_out/r2p_pytest_gpu.log
timeout -s KILL 900 python bench.py --no-cpu-baseline --no-train-leg --no-eager-leg > gpurun_out/r2p_bench.json 2> gpurun_out/r2p_bench.err; python tools/show_bench.py gpurun_out/r2p_bench.json | head -1; tail -3 gpurun_out/r2p_bench.err
for wl in c1 c2b c3 c4; do
timeout -s KILL 600 python bench.py --workload $wl --no-cpu-baseline --no-train-leg --no-eager-leg > gpurun_out/r2p_bench_$wl.json 2>/dev/null; echo $wl; python tools/show_bench.py gpurun_out/r2p_bench_$wl.json | head -1
done
