#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 400 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "tensor_memory" 2>&1 | tail -3
timeout -s KILL 600 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --no-train-leg --no-eager-leg > gpurun_out/r2m_bench.json 2> gpurun_out/r2m_bench.err
python tools/show_bench.py gpurun_out/r2m_bench.json | head -1; tail -2 gpurun_out/r2m_bench.err
timeout -s KILL 600 python bench.py --workload c2b --steps 50 --warmup 5 --no-cpu-baseline --no-train-leg --no-eager-leg > gpurun_out/r2m_bench_c2b.json 2>/dev/null
python tools/show_bench.py gpurun_out/r2m_bench_c2b.json | head -1
