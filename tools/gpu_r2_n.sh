#!/bin/bash
# round 2, call N: the whole GPU suite, smoke, full bench (both arms), config-1 Frappe training, ncu evidence
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2n_smi.txt 2>&1
timeout -s KILL 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2n_pytest_gpu.log; tail -5 gpurun_out/r2n_pytest_gpu.log
timeout -s KILL 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2n_smoke.log 2>&1; tail -3 gpurun_out/r2n_smoke.log
timeout -s KILL 900 python bench.py > gpurun_out/r2n_bench.json 2> gpurun_out/r2n_bench.err; python tools/show_bench.py gpurun_out/r2n_bench.json; tail -3 gpurun_out/r2n_bench.err
timeout -s KILL 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2n_bench_reference.json 2> gpurun_out/r2n_bench_reference.err; cat gpurun_out/r2n_bench_reference.json | cut -c1-600
timeout -s KILL 900 python train.py --model armnet_1h --nemb 10 --h 10 --alpha 1.7 --lr 0.001 --data_dir data/ --dataset frappe --log_dir gpurun_out/frappe_log/ --exp_name frappe_c1 > gpurun_out/r2n_frappe_train.log 2>&1; grep -E "Final|best valid|Total" gpurun_out/r2n_frappe_train.log | tail -4
for regime in init trained; do
  timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:armnet_fwd_tmem --launch-skip 2 -c 1 \
     -o gpurun_out/r2n_tmem_${regime} -f python tools/prof_hot.py --regime $regime > gpurun_out/r2n_ncu_${regime}.log 2>&1
done
timeout -s KILL 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2n_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-train-leg --no-eager-leg > gpurun_out/r2n_launches_bench.log 2>&1
