timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -5
P='import json,sys; d=json.loads(sys.stdin.read()); print(d["value"], d["ms_per_step"], "trained", d["trained_like"]["value"], "e2e", d["e2e"]["value"])'
echo "== 12 warps (default lib)"; python bench.py --steps 30 --warmup 3 --no-cpu-baseline | python -c "$P"
for nw in 4 8; do echo "== default lib NW=$nw"; ARMNET_FORCE_NW=$nw python bench.py --steps 30 --warmup 3 --no-cpu-baseline | python -c "$P"; done
export ARMNET_B200_LIB=$PWD/build/libarmnet_b200_w16.so
echo "== 16 warps lib"; python bench.py --steps 30 --warmup 3 --no-cpu-baseline | python -c "$P"
for w in c2b c4 c3 c1; do echo "== w16 workload $w"; python bench.py --steps 30 --warmup 3 --no-cpu-baseline --workload $w | python -c "$P"; done
unset ARMNET_B200_LIB
for w in c2b c4 c3 c1; do echo "== w12 workload $w"; python bench.py --steps 30 --warmup 3 --no-cpu-baseline --workload $w | python -c "$P"; done
