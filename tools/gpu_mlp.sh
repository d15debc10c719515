#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_mlp.py -q -s 2>&1 | tail -30 | tee gpurun_out/pytest_mlp.log
timeout 300 python tools/bench_mlp.py 2>&1 | tail -12 | tee gpurun_out/bench_mlp.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"mlp_|sgemm|gemm" -c 400 --csv --log-file gpurun_out/mlp_launches.csv python tools/bench_mlp.py > /dev/null 2>&1
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open('gpurun_out/mlp_launches.csv')) if len(r) > 5]
hdr = rows[0]; ki = hdr.index('Kernel Name'); vi = hdr.index('Metric Value'); gi = hdr.index('Grid Size')
d = collections.defaultdict(list)
for r in rows[1:]:
    try: d[(r[ki][:60], r[gi])].append(float(r[vi].replace(',', '')))
    except ValueError: pass
for k, v in d.items(): print(k, 'n=%d' % len(v), 'median %.1f us' % (sorted(v)[len(v)//2] / 1000 if max(v) > 1000 else sorted(v)[len(v)//2]))
PY
