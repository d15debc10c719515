#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/train_launches.csv python tools/bench_train.py --steps 2 --warmup 2 > /dev/null 2>&1
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open('gpurun_out/train_launches.csv')) if len(r) > 5]
hdr = rows[0]; ki = hdr.index('Kernel Name'); vi = hdr.index('Metric Value'); ii = hdr.index('ID')
recs = []
for r in rows[1:]:
    try: recs.append((int(r[ii]), r[ki], float(r[vi].replace(',', '')) / 1000))
    except ValueError: pass
# last step = launches after the last clamp_adam-but-one
idx = [i for i, r in enumerate(recs) if 'clamp_adam' in r[1]]
last = recs[idx[-2] + 1: idx[-1] + 1]
tot = sum(r[2] for r in last)
print('launches in one step: %d, total %.1f us' % (len(last), tot))
for _, k, t in last:
    if t > 15: print('%8.1f us  %s' % (t, k[:110]))
PY
