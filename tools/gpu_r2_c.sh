#!/bin/bash
# round 2, call C: headline bench with the tensor-memory kernel (default) and with it forced off, same box; ncu captures
mkdir -p gpurun_out
timeout -s KILL 600 python bench.py --steps 50 --warmup 5 > gpurun_out/r2c_bench_tmem.json 2> gpurun_out/r2c_bench_tmem.err
python tools/show_bench.py gpurun_out/r2c_bench_tmem.json; tail -3 gpurun_out/r2c_bench_tmem.err
ARMNET_TMEM=0 timeout -s KILL 600 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --no-train-leg --no-eager-leg > gpurun_out/r2c_bench_fp32.json 2> gpurun_out/r2c_bench_fp32.err
python tools/show_bench.py gpurun_out/r2c_bench_fp32.json
timeout -s KILL 600 python bench.py --workload c2b --steps 50 --warmup 5 --no-cpu-baseline --no-train-leg --no-eager-leg > gpurun_out/r2c_bench_c2b_tmem.json 2>/dev/null
python tools/show_bench.py gpurun_out/r2c_bench_c2b_tmem.json
for regime in init trained; do
  timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:armnet_fwd_tmem --launch-skip 2 -c 1 \
     -o gpurun_out/r2c_tmem_${regime} -f python tools/prof_hot.py --regime $regime > gpurun_out/r2c_ncu_${regime}.log 2>&1
  tail -2 gpurun_out/r2c_ncu_${regime}.log
done
ls -la gpurun_out/*.ncu-rep
