P='import json,sys; d=json.loads(sys.stdin.read()); print(d["value"], d["ms_per_step"], "trained", d["trained_like"]["value"], "e2e", d["e2e"]["value"])'
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
for w in c2a c4 c3 c2b c1; do echo "== $w"; python bench.py --steps 30 --warmup 3 --no-cpu-baseline --workload $w | python -c "$P"; done
echo "== c2a dynamic"; ARMNET_DYNAMIC=1 python bench.py --steps 30 --warmup 3 --no-cpu-baseline | python -c "$P"
