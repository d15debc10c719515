#!/bin/bash
# Final validation of the committed state: smoke(), the whole GPU suite, bench.py (default path), launch list.
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke_v7.log 2>&1; tail -2 gpurun_out/smoke_v7.log
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu_v7.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_v7.log
tail -6 gpurun_out/pytest_gpu_v7.log
timeout 300 python bench.py --steps 50 --warmup 5 > gpurun_out/bench_n1_v7.json 2> gpurun_out/bench_n1_v7.err
python tools/show_bench.py gpurun_out/bench_n1_v7.json | head -1
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_v7.json 2>/dev/null; cut -c1-200 gpurun_out/bench_ref_v7.json
