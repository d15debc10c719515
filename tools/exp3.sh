P='import json,sys; d=json.loads(sys.stdin.read()); print(d["value"], d["ms_per_step"], "trained", d["trained_like"]["value"], "e2e", d["e2e"]["value"])'
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1
for lib in "" $PWD/build/libarmnet_b200_w16.so; do
export ARMNET_B200_LIB=$lib
for dyn in 0 1; do
if [ $dyn = 1 ]; then export ARMNET_DYNAMIC=1; else unset ARMNET_DYNAMIC; fi
for w in c2a c3; do echo "== lib=${lib##*/} dynamic=$dyn $w"; python bench.py --steps 30 --warmup 3 --no-cpu-baseline --workload $w | python -c "$P"; done
done; done
