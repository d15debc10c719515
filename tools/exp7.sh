P='import json,sys; d=json.loads(sys.stdin.read()); print(d["value"], d["ms_per_step"], "trained", d["trained_like"]["value"], "e2e", d["e2e"]["value"])'
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
for dyn in 0 1; do
if [ $dyn = 1 ]; then export ARMNET_DYNAMIC=1; else unset ARMNET_DYNAMIC; fi
for w in c2a c3 c2b c1; do echo "== dynamic=$dyn $w"; python bench.py --steps 30 --warmup 3 --no-cpu-baseline --workload $w | python -c "$P"; done
echo -n "== dynamic=$dyn c2a skip7 ms "; ARMNET_DEBUG_SKIP=7 python bench.py --steps 30 --warmup 3 --no-cpu-baseline | python -c 'import json,sys; d=json.loads(sys.stdin.read()); print(d["ms_per_step"])'
for lk in 1 2 4; do echo -n "== dynamic=$dyn c2a look=$lk ms "; ARMNET_FORCE_LOOK=$lk python bench.py --steps 30 --warmup 3 --no-cpu-baseline | python -c 'import json,sys; d=json.loads(sys.stdin.read()); print(d["ms_per_step"])'; done
done
