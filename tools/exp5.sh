P='import json,sys; d=json.loads(sys.stdin.read()); print(d["ms_per_step"])'
for sk in 0 1 2 4 3 5 6 7; do echo -n "skip=$sk  ms="; ARMNET_DEBUG_SKIP=$sk python bench.py --steps 30 --warmup 3 --no-cpu-baseline | python -c "$P"; done
