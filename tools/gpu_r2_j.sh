#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 400 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "tensor_memory" 2>&1 | tail -15
for rows in 2 1; do
ARMNET_TMEM_ROWS=$rows timeout -s KILL 600 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --no-train-leg --no-eager-leg > gpurun_out/r2j_bench_rows$rows.json 2> gpurun_out/r2j_bench_rows$rows.err
python tools/show_bench.py gpurun_out/r2j_bench_rows$rows.json | head -1; tail -2 gpurun_out/r2j_bench_rows$rows.err
done
timeout -s KILL 600 python bench.py --workload c2b --steps 50 --warmup 5 --no-cpu-baseline --no-train-leg --no-eager-leg > gpurun_out/r2j_bench_c2b.json 2>/dev/null
python tools/show_bench.py gpurun_out/r2j_bench_c2b.json | head -1
for regime in init trained; do
  timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:armnet_fwd_tmem --launch-skip 2 -c 1 \
     -o gpurun_out/r2j_tmem_${regime} -f python tools/prof_hot.py --regime $regime > gpurun_out/r2j_ncu_${regime}.log 2>&1
done
