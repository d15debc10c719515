#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -5 | tee gpurun_out/pytest_gpu_r1e.log
for wl in c1 c2b c3 c4; do
  timeout 300 python bench.py --workload $wl --steps 30 --warmup 5 --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/bench_$wl.json
  python -c "
import json; d=json.load(open('gpurun_out/bench_$wl.json'))
print('$wl', 'value %.2fM' % (d['value']/1e6), 'ms %.4f' % d['ms_per_step'], 'frac %.3f' % d['roofline']['frac'], 'e2e %.2fM' % (d['e2e']['value']/1e6), 'sync %.2fM' % (d['e2e']['per_call_sync']['value']/1e6), 'trained-like %.2fM' % (d['trained_like']['value']/1e6))"
done
