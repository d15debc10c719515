timeout 300 python -m pytest tests/test_gpu_mlp.py -x -q -m gpu 2>&1 | tail -2
timeout 200 python tools/bench_mlp.py 2>&1 | tail -3
for i in 1 2; do
python bench.py --no-train-leg --no-eager-leg --no-cpu-baseline --steps 50 --warmup 5 > gpurun_out/q_tc.json 2>/dev/null; python tools/show_bench.py gpurun_out/q_tc.json 2>/dev/null| head -1
python bench.py --no-train-leg --no-eager-leg --no-cpu-baseline --steps 50 --warmup 5 --no-hidden-tc > gpurun_out/q_notc.json 2>/dev/null; python tools/show_bench.py gpurun_out/q_notc.json 2>/dev/null | head -1
done
