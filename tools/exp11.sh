#!/bin/bash
# warps per CTA vs registers per thread for the fused kernel: 12 x 168 (shipped), 14 x 144, 16 x 128
P='import json,sys; d=json.loads(sys.stdin.read()); print(d["value"]/1e6, d["ms_per_step"], "trained", d["trained_like"]["value"]/1e6)'
echo -n "w12 "; timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "$P"
for w in 14 16; do echo -n "w$w "; ARMNET_B200_LIB=$PWD/armnet_b200/tuning/libarmnet_w$w.so timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "$P"; done
