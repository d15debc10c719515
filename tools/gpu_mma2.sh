#!/bin/bash
# Tensor-core fused kernel, second pass: parity (both operand splits), bench, ncu full capture + launch list.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/pytest_mma2.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_mma2.log
ARMNET_MMA=1 ARMNET_MMA_SPLIT=rna timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "tensor_core or fused_stages" > gpurun_out/pytest_mma2_rna.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_mma2_rna.log
tail -12 gpurun_out/pytest_mma2.log; tail -4 gpurun_out/pytest_mma2_rna.log
ARMNET_MMA=1 timeout 300 python bench.py --steps 50 --warmup 5 > gpurun_out/bench_n1_mma2.json 2> gpurun_out/bench_n1_mma2.err
ARMNET_MMA=1 ARMNET_MMA_SPLIT=rna timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/bench_n1_mma2_rna.json 2>/dev/null
for f in bench_n1_mma2 bench_n1_mma2_rna; do echo $f; python tools/show_bench.py gpurun_out/$f.json | head -1; done
ARMNET_MMA=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:armnet_fwd_mma -s 3 -c 1 -f -o gpurun_out/prof_fwd_mma python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_mma.log 2>&1
ARMNET_MMA=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_mma.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
ls -la gpurun_out | tail -4
