#!/bin/bash
# Final validation of the committed state (nemb-16 tensor-core instance on by default): smoke, whole GPU suite,
# bench.py c2a + c4, config-4 training step.
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke_v7b.log 2>&1; tail -1 gpurun_out/smoke_v7b.log
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu_v7b.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_v7b.log
tail -4 gpurun_out/pytest_gpu_v7b.log
timeout 300 python bench.py --steps 50 --warmup 5 > gpurun_out/bench_n1_v7b.json 2> gpurun_out/bench_n1_v7b.err
timeout 300 python bench.py --workload c4 --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench_c4_v7b.json 2>/dev/null
for f in bench_n1_v7b bench_c4_v7b; do echo $f; python tools/show_bench.py gpurun_out/$f.json 2>/dev/null | head -1; done
timeout 300 python tools/bench_train.py --optimizer fused 2>&1 | tail -1 | tee gpurun_out/train_n1_fused_v7b.json | cut -c1-300
