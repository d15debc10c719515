P='import json,sys; d=json.loads(sys.stdin.read()); print(d["ms_per_step"])'
export ARMNET_DEBUG_SKIP=7
echo -n "skip7 lockstep  ms="; python bench.py --steps 30 --warmup 3 --no-cpu-baseline | python -c "$P"
echo -n "skip7 dynamic  ms="; ARMNET_DYNAMIC=1 python bench.py --steps 30 --warmup 3 --no-cpu-baseline | python -c "$P"
echo -n "skip7 no-tma-store  ms="; ARMNET_NO_TMA_STORE=1 python bench.py --steps 30 --warmup 3 --no-cpu-baseline | python -c "$P"
echo -n "skip7 no-tma-gather  ms="; ARMNET_NO_TMA_GATHER=1 python bench.py --steps 30 --warmup 3 --no-cpu-baseline | python -c "$P"
echo -n "skip7 no-flush  ms="; python bench.py --steps 30 --warmup 3 --no-cpu-baseline --no-flush | python -c "$P"
echo -n "skip7 NW=4  ms="; ARMNET_FORCE_NW=4 python bench.py --steps 30 --warmup 3 --no-cpu-baseline | python -c "$P"
echo -n "skip7 NW=8  ms="; ARMNET_FORCE_NW=8 python bench.py --steps 30 --warmup 3 --no-cpu-baseline | python -c "$P"
echo -n "skip7 dynamic no-flush  ms="; ARMNET_DYNAMIC=1 python bench.py --steps 30 --warmup 3 --no-cpu-baseline --no-flush | python -c "$P"
