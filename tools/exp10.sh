#!/bin/bash
# occupancy sensitivity of the fused kernel: throughput vs warps per CTA
for nw in 4 6 8 10 12; do
  echo -n "NW=$nw "; ARMNET_FORCE_NW=$nw timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value']/1e6, d['ms_per_step'], d['trained_like']['value']/1e6)"
done
