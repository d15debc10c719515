P='import json,sys; d=json.loads(sys.stdin.read()); print(d["value"], d["ms_per_step"], "trained", d["trained_like"]["value"])'
for w in c2a f20 f10; do for nw in 4 8 12; do echo "== $w NW=$nw"; ARMNET_FORCE_NW=$nw python bench.py --steps 30 --warmup 3 --no-cpu-baseline --workload $w | python -c "$P"; done; done
