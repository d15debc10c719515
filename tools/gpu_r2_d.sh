#!/bin/bash
# round 2, call D: tensor-memory kernel with specialised producer warps -- parity, bench, ncu
mkdir -p gpurun_out
timeout -s KILL 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "tensor_memory" 2>&1 | tail -5 > gpurun_out/r2h_tmem_tests.txt
cat gpurun_out/r2h_tmem_tests.txt
grep -q "passed" gpurun_out/r2h_tmem_tests.txt && ! grep -q "failed\|error" gpurun_out/r2h_tmem_tests.txt || exit 1
timeout -s KILL 600 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --no-train-leg --no-eager-leg > gpurun_out/r2h_bench_tmem.json 2> gpurun_out/r2h_bench_tmem.err
python tools/show_bench.py gpurun_out/r2h_bench_tmem.json; tail -3 gpurun_out/r2h_bench_tmem.err
timeout -s KILL 600 python bench.py --workload c2b --steps 50 --warmup 5 --no-cpu-baseline --no-train-leg --no-eager-leg > gpurun_out/r2h_bench_c2b_tmem.json 2>/dev/null
python tools/show_bench.py gpurun_out/r2h_bench_c2b_tmem.json
for regime in init trained; do
  timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:armnet_fwd_tmem --launch-skip 2 -c 1 \
     -o gpurun_out/r2h_tmem_${regime} -f python tools/prof_hot.py --regime $regime > gpurun_out/r2h_ncu_${regime}.log 2>&1
  tail -1 gpurun_out/r2h_ncu_${regime}.log
done
