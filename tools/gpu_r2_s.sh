#!/bin/bash
# round 2, second half (startup / hidden-layer kernel work), single-GPU evidence run: whole GPU suite, smoke, bench (both arms, all legs), training step, ncu
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2s_smi.txt 2>&1
timeout -s KILL 1500 python -m pytest tests -m gpu -q 2>&1 | tail -8 > gpurun_out/r2s_pytest_gpu.log; tail -4 gpurun_out/r2s_pytest_gpu.log
timeout -s KILL 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2s_smoke.log 2>&1; tail -2 gpurun_out/r2s_smoke.log
timeout -s KILL 900 python bench.py > gpurun_out/r2s_bench.json 2> gpurun_out/r2s_bench.err; python tools/show_bench.py gpurun_out/r2s_bench.json 2>/dev/null | head -1; tail -3 gpurun_out/r2s_bench.err
timeout -s KILL 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2s_bench_reference.json 2> gpurun_out/r2s_bench_reference.err
for wl in c1 c2b c3 c4; do
timeout -s KILL 600 python bench.py --workload $wl --no-cpu-baseline --no-train-leg --no-eager-leg > gpurun_out/r2s_bench_$wl.json 2>/dev/null; echo $wl; python tools/show_bench.py gpurun_out/r2s_bench_$wl.json 2>/dev/null | head -1
done
bash tools/gpu_train_prof.sh 2>&1 | tail -24 > gpurun_out/r2s_train_step_launches.txt; head -3 gpurun_out/r2s_train_step_launches.txt
cp gpurun_out/train_launches.csv gpurun_out/r2s_train_launches.csv
for regime in init trained; do
  timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:armnet_fwd_tmem --launch-skip 2 -c 1 \
     -o gpurun_out/r2s_tmem_${regime} -f python tools/prof_hot.py --regime $regime > gpurun_out/r2s_ncu_${regime}.log 2>&1
done
timeout -s KILL 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2s_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-train-leg --no-eager-leg > gpurun_out/r2s_launches_bench.log 2>&1
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:mlp_hidden --launch-skip 2 -c 1 \
   -o gpurun_out/r2s_hidden_tc -f python tools/prof_hidden.py > gpurun_out/r2s_ncu_hidden.log 2>&1
timeout -s KILL 300 python tools/bench_mlp.py > gpurun_out/r2s_bench_mlp.txt 2>&1
timeout -s KILL 300 python tools/sweep_scorer.py > gpurun_out/r2s_sweep_scorer.txt 2>&1
