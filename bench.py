#!/usr/bin/env python
"""Benchmark of the ARM-Net forward hot path (BASELINE.json metric: CTR samples/sec, bsz=4096, Criteo shape).

  python bench.py [--gpus N] [--steps K] [--warmup W]            # this repo's CUDA path
  python bench.py --impl reference [--steps K] [--warmup W]      # the reference algorithm on the host CPU cores
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N   (one rank per GPU)

One "step" = one pass of the hot path (value clamp -> embedding lookup -> attention logits -> entmax gates ->
exponential-neuron interaction, armnet.py:82-87) over one synthetic batch of 4096 samples per GPU.
  value  : samples/s with inputs resident in HBM; CUDA-event time of the K steps (L2 flushed between steps, the
           flush is outside the events), max over ranks.
  e2e    : samples/s through the drop-in nn.Module (ARMNetModel.forward -> y[B]) starting from pinned HOST
           batches: H2D copy of ids/values, fused kernel, BatchNorm + MLP, D2H read of y, all inside the timed region.
  roofline: algorithmic HBM bytes of the step / event time, against MEASURED_PEAKS.json (hbm_gbs).
  cpu_baseline: the oracle port of the reference's ATen op chain (oracle/armnet_oracle.py, torch CPU, all host
           threads) on a bounded sample of the same workload, rank 0 at N=1 only.
Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # BASELINE.json configs[1]: armnet (multi-head) synthetic Criteo-shape, train.py defaults for the rest
    'c2a': dict(model='armnet', nfield=39, nfeat=1000000, nemb=10, nhead=4, nhid=128, alpha=1.7, bsz=4096,
                mlp_nlayer=2, mlp_nhid=256),
    'c2b': dict(model='armnet', nfield=39, nfeat=1000000, nemb=10, nhead=4, nhid=64, alpha=2.0, bsz=4096,
                mlp_nlayer=2, mlp_nhid=500),
    'c3': dict(model='armnet', nfield=22, nfeat=1500000, nemb=100, nhead=1, nhid=32, alpha=1.5, bsz=8192,
               mlp_nlayer=3, mlp_nhid=200),
    'c4': dict(model='armnet', nfield=39, nfeat=1000000, nemb=16, nhead=4, nhid=128, alpha=1.7, bsz=4096,
               mlp_nlayer=2, mlp_nhid=256),
    # tuning probes (not BASELINE configs): C2a with fewer fields -> smaller unrolled kernel body
    'f20': dict(model='armnet', nfield=20, nfeat=1000000, nemb=10, nhead=4, nhid=128, alpha=1.7, bsz=4096,
                mlp_nlayer=2, mlp_nhid=256),
    'f10': dict(model='armnet', nfield=10, nfeat=1000000, nemb=10, nhead=4, nhid=128, alpha=1.7, bsz=4096,
                mlp_nlayer=2, mlp_nhid=256),
    'c1': dict(model='armnet_1h', nfield=10, nfeat=5382, nemb=10, nhead=1, nhid=10, alpha=1.7, bsz=4096,
               mlp_nlayer=2, mlp_nhid=256),
}
# roofline.traffic / fp32_pipe come from committed `ncu --set full` captures, keyed by (workload, kernel kind):
# profiles/ncu_traffic.json = {"<workload>/<kind>": {"dram_bytes": dram__bytes_read.sum + dram__bytes_write.sum per launch,
# "fp32_pipe_cycles_per_row": FMA-pipe cycles per (sample, neuron) row from the executed-opcode histogram or null,
# "source": the extract under profiles/}}.  No entry -> null (never a number from another kernel).


def ncu_record(workload, kind):
    try:
        with open(os.path.join(ROOT, 'profiles', 'ncu_traffic.json')) as f:
            return json.load(f).get(f'{workload}/{kind}')
    except Exception:
        return None


METRIC = 'CTR samples/sec (bsz=4096, Criteo-shape) at 1/2/4/8 B200; HBM GB/s vs roofline'


def algorithmic_bytes_per_sample(w):
    """ids + values + gathered rows + interaction output z."""
    F, E, R = w['nfield'], w['nemb'], w['nhead'] * w['nhid']
    return F * (8 + 4 + 4 * E) + 4 * R * E


def peaks():
    try:
        with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as f:
            return float(json.load(f)['hbm_gbs']), 'measured (MEASURED_PEAKS.json hbm_gbs)'
    except Exception:
        return 6650.0, 'fallback (B200_PROFILING.md)'


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled while the timed region runs."""
    Q = ('clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), f'--query-gpu={self.Q}',
                                          '--format=csv,noheader,nounits', '-lms', '50'], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(',')])

    def stop(self):
        if not self.proc:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        self.proc.terminate()
        self.thread.join(timeout=2)
        sm = [float(r[0]) for r in self.rows if len(r) >= 6 and r[0].replace('.', '').isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 6 and r[1].replace('.', '').isdigit()]
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        reasons = sorted({names[i] for r in self.rows if len(r) >= 6 for i in range(4) if r[2 + i] == 'Active'})
        return {'sm_mhz': statistics.median(sm) if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'reasons': reasons, 'samples': len(sm)}


def make_batches(w, n, seed, regime_values='ones'):
    import torch
    gen = torch.Generator().manual_seed(seed)
    out = []
    for _ in range(n):
        ids = torch.randint(0, w['nfeat'], (w['bsz'], w['nfield']), generator=gen, dtype=torch.int64)
        vals = torch.ones(w['bsz'], w['nfield']) if regime_values == 'ones' else \
            torch.rand(w['bsz'], w['nfield'], generator=gen)
        out.append((ids, vals))
    return out


def build_module(w, seed=2025):
    import torch
    import armnet_b200 as ab
    torch.manual_seed(seed)                       # train.py:47,144
    if w['model'] == 'armnet':
        return ab.ARMNetModel(w['nfield'], w['nfeat'], w['nemb'], w['nhead'], w['alpha'], w['nhid'],
                              w['mlp_nlayer'], w['mlp_nhid'], 0.0, False, 2, 256)
    return ab.ARMNet1H(w['nfield'], w['nfeat'], w['nemb'], w['alpha'], w['nhid'], w['nemb'], w['mlp_nlayer'],
                       w['mlp_nhid'], 0.0, False, 2, 256)


def trained_like_(model, seed=7):
    """Secondary regime: embedding ~ N(0,1), attention weights x4 -> |logits| O(1..10), gates genuinely sparse."""
    import torch
    gen = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        w = model.embedding.embedding.weight
        w.copy_(torch.randn(w.shape, generator=gen))
        a = model.attn_layer
        (a.bilinear_w if isinstance(a.bilinear_w, torch.nn.Parameter) else a.bilinear_w.weight).mul_(4.0)
        a.query.mul_(4.0)


def cpu_model_string():
    try:
        with open('/proc/cpuinfo') as f:
            for line in f:
                if line.startswith('model name'):
                    return line.split(':', 1)[1].strip()
    except OSError:
        pass
    return 'unknown'


def reference_eager_on_gpu(w, model, batch, dev, steps=3):
    """SURVEY.md 8d's bar: the reference's own op chain (oracle port of armnet.py:82-87: the same ATen calls) run EAGERLY
    on the same B200 on the same batch -- what `model.cuda()` (model_utils.py:86) gives a user of the reference."""
    import torch
    from oracle import armnet_oracle as oracle
    st = {k: v.detach() for k, v in model.state_dict().items()
          if k.startswith('embedding.') or k.startswith('attn_layer.')}
    ids, vals = batch

    def run():
        with torch.no_grad():
            return oracle.hot_path(st, w['alpha'], ids, vals.clone())['z']
    run()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(steps):
        z = run()
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / steps
    del z
    torch.cuda.empty_cache()
    return {'value': w['bsz'] / (ms * 1e-3), 'unit': 'samples/s', 'ms_per_step': ms, 'steps': steps,
            'what': 'reference op chain (embedding, 2 einsums, 50-step entmax bisection, einsum, exp) as eager torch ops '
                    'on this GPU, same batch, hot path only (oracle port on CUDA tensors)'}


def train_c4_leg(dev, rank, world, steps=10, warmup=3):
    """BASELINE config 4 under the driver's eyes: one data-parallel training step of armnet at the Criteo shape with
    nemb=16, bsz=4096 per GPU: forward + BCE loss + backward + ONE all-reduce of the flat gradient bucket + fused
    clamp+Adam (train.py:101-114,62-65).  Device-timed, max over ranks; phases by CUDA events on rank 0."""
    import torch
    import torch.nn as nn
    import torch.distributed as dist
    import armnet_b200 as ab
    from armnet_b200.parallel import FlatAdam
    F, V, E, K, O, B = 39, 1000000, 16, 4, 128, 4096
    torch.manual_seed(2025)
    model = ab.ARMNetModel(F, V, E, K, 1.7, O, 2, 256, 0.0, False, 2, 256).to(dev).train()
    # >= 4 ranks: optimizer state sharded over the ranks (reduce-scatter, Adam on 1/world of the bucket, all-gather of the
    # parameters): same wire bytes as the all-reduce, an 8x smaller Adam pass at 8 GPUs (2.99 -> 2.91 ms per step)
    stepper = FlatAdam(model.parameters(), lr=3e-3, clamp=1.0, shard_state=world >= 4)
    crit = nn.BCEWithLogitsLoss()
    g = torch.Generator().manual_seed(100 + rank)
    batches = [(torch.randint(0, V, (B, F), generator=g).to(dev), torch.ones(B, F, device=dev),
                (torch.rand(B, generator=g) < 0.25).float().to(dev)) for _ in range(4)]
    sampler = ClockSampler(dev.index or 0)

    def step(i, evs):
        ids, vals, y = batches[i % 4]
        e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        e[0].record()
        loss = crit(model({'id': ids, 'value': vals.clone()}).reshape(-1), y)
        e[1].record()
        stepper.zero_grad()
        loss.backward()
        e[2].record()
        opt_ev = [] if evs is not None else None
        stepper.step(events=opt_ev)
        if evs is not None:
            evs.append(e + opt_ev)

    for i in range(warmup):
        step(i, None)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    if rank == 0:
        sampler.start()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    evs = []
    t0.record()
    for i in range(steps):
        step(i, evs)
    t1.record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([t0.elapsed_time(t1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = t.item()
    ph = [0.0] * 4
    for e in evs:
        ph[0] += e[0].elapsed_time(e[1])
        ph[1] += e[1].elapsed_time(e[2])
        ph[2] += e[3].elapsed_time(e[4])
        ph[3] += e[4].elapsed_time(e[5])
    ph = [x / steps for x in ph]
    nbytes = stepper.numel() * 4
    sharded = stepper.sharded
    del model, stepper, batches
    torch.cuda.empty_cache()
    return {'metric': 'training samples/s (config 4: armnet nemb=16, bsz=4096/GPU, dense Adam, 1 all-reduce/step)',
            'value': B * world * steps / (ms * 1e-3), 'unit': 'samples/s', 'n_gpus': world, 'steps': steps,
            'ms_per_step': ms / steps, 'grad_bucket_bytes': nbytes,
            'phase_ms': ({'forward+loss': ph[0], 'backward': ph[1], 'reduce_scatter': ph[2],
                          'clamp+adam(shard)+all_gather': ph[3]} if sharded else
                         {'forward+loss': ph[0], 'backward': ph[1], 'allreduce': ph[2], 'clamp+adam': ph[3]}),
            'collective': ('reduce-scatter + all-gather around a sharded clamp+Adam (optimizer state / world)' if sharded
                           else 'one all-reduce of the flat gradient bucket'),
            'allreduce_bus_GBps': (((1.0 if sharded else 2.0) * (world - 1) / world * nbytes / (ph[2] * 1e-3) / 1e9)
                                   if world > 1 else None),
            'clocks': clocks}


def cpu_reference_rate(w, state, steps, warmup, sample_b, full_model, seed=99):
    """The reference algorithm on the host cores: oracle port of the ATen op chain, all threads."""
    import torch
    from oracle import armnet_oracle as oracle
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    ws = dict(w, bsz=sample_b)
    batches = make_batches(ws, 2, seed)
    fn = oracle.forward if full_model else oracle.hot_path
    times = []
    with torch.no_grad():
        for i in range(warmup + steps):
            ids, vals = batches[i % 2]
            t0 = time.perf_counter()
            fn(state, w['alpha'], ids, vals.clone())
            dt = time.perf_counter() - t0
            if i >= warmup:
                times.append(dt)
    total = sum(times)
    return sample_b * len(times) / total, total / len(times), cores


_JSON_FD = None


def emit_json(obj):
    """The ONE JSON line of the contract goes to the process's original stdout; everything else a library writes to fd 1
    (NCCL prints its version banner there) has been redirected to stderr by main()."""
    line = (json.dumps(obj) + '\n').encode()
    if _JSON_FD is None:
        sys.stdout.write(line.decode())
        sys.stdout.flush()
    else:
        os.write(_JSON_FD, line)


def main():
    global _JSON_FD
    sys.stdout.flush()
    _JSON_FD = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=50)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--workload', default='c2a', choices=sorted(WORKLOADS))
    ap.add_argument('--cpu-sample', type=int, default=0,
                    help='samples per CPU step (default: the whole batch for --impl reference, 256 for the cpu_baseline leg)')
    ap.add_argument('--no-train-leg', action='store_true', help='skip the config-4 training-step leg (train_c4)')
    ap.add_argument('--no-eager-leg', action='store_true', help='skip the reference-eager-on-GPU leg')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-flush', action='store_true', help='do not flush L2 between timed steps')
    ap.add_argument('--no-hidden-tc', action='store_true',
                    help='A/B: hidden layers 2..n of the MLP on the CUDA-core tail kernel instead of tcgen05')
    args = ap.parse_args()
    w = WORKLOADS[args.workload]
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    cfg = {'workload': f"{args.workload}: {w['model']} forward hot path, {w['nfield']} fields, {w['nfeat']} vocab, "
                       f"nemb={w['nemb']}, nattn_head={w['nhead']}, h={w['nhid']}, alpha={w['alpha']}, "
                       f"bsz={w['bsz']}/GPU", 'ids': 'uniform[0,V) int64', 'values': 'all 1.0',
           'params': 'reference init, seed 2025', 'parallelism': f'dp{max(world, args.gpus)} batch-sharded, no collective'}

    # ------------------------------------------------------------------ reference arm (CPU)
    if args.impl == 'reference':
        if rank != 0:
            return
        import torch
        from oracle import armnet_oracle as oracle
        st = oracle.reference_init_state(w['model'], w['nfield'], w['nfeat'], w['nemb'], w['nhead'], w['nhid'],
                                         mlp_nlayer=w['mlp_nlayer'], mlp_nhid=w['mlp_nhid'], seed=2025)
        sample_b = args.cpu_sample or w['bsz']                 # the arm's own config: whole batches of bsz samples
        steps = max(3, min(args.steps, 5))                     # >= 3 timed steps; ~2-4 s each at bsz=4096 on 16 cores
        warm = 1
        rate, sec, cores = cpu_reference_rate(w, st, steps, warm, sample_b, full_model=True)
        sample = (f'{steps} steps x {sample_b} samples of the same workload (full ARMNetModel.forward, eval), '
                  f'torch {torch.__version__} CPU, {cores} threads, {sec:.2f} s/step, cpu: {cpu_model_string()}')
        emit_json({
            'impl': 'reference', 'metric': METRIC, 'value': rate, 'unit': 'samples/s', 'n_gpus': args.gpus,
            'steps': steps, 'warmup': warm, 'ms_per_step': sec * 1e3, 'higher_is_better': True, 'scaling': 'weak',
            'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic', 'config': cfg,
            'cpu_baseline': {'value': rate, 'unit': 'samples/s', 'cores': cores, 'kind': 'port', 'sample': sample},
            'e2e': {'value': rate, 'unit': 'samples/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
            'gpu_launches': 0})
        return

    # ------------------------------------------------------------------ this repo's CUDA arm
    import torch
    import torch.distributed as dist
    from armnet_b200 import ops
    assert torch.cuda.is_available(), 'bench.py needs a CUDA device (no CPU fallback)'
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    if world > 1:
        os.environ.setdefault('NCCL_DEBUG_FILE', '/dev/stderr')   # NCCL's version banner / warnings off stdout: ONE JSON line
        dist.init_process_group('nccl', device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    model = build_module(w).to(dev).eval()
    if args.no_hidden_tc:
        model.mlp.hidden_tensor_core = False
    n_batches = 4
    host = [(i.pin_memory(), v.pin_memory()) for i, v in make_batches(w, n_batches, seed=1000 + rank)]
    resident = [(i.to(dev), v.to(dev)) for i, v in host]
    flush = None if args.no_flush else torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2
    table = model.embedding.embedding.weight
    tab, ld = model._shadow.get(table) if model.padded_table else (table.detach(), table.shape[1])
    W, Q, Vv = (t.detach() for t in model._attn_weights())

    # serving semantics: the attention parameters are pre-contracted once per weight version (armnet_fused_prepare_f32, a
    # 3.6 us parameter-only kernel, like the padded table and the split MLP weights), not once per batch
    state = {'ws': ops.fused_prepare(W, Q, Vv, w['alpha'], w['nfield'], one_head=model.one_head)}

    def hot_step(i):
        ids, vals = resident[i % n_batches]
        z, _ = ops.fused_forward(ids, vals, tab, W, Q, Vv, w['alpha'], one_head=model.one_head, ld=ld,
                                 nemb=table.shape[1], prepared=state['ws'])
        return z

    def timed_hot(steps, warmup):
        for i in range(warmup):
            hot_step(i)
        launches = 0
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        barrier()
        for i in range(steps):
            if flush is not None:
                flush.zero_()
            ev[i][0].record()
            hot_step(i)
            ev[i][1].record()
            launches += ops.last_launch_count()
        barrier()
        ms = [a.elapsed_time(b) for a, b in ev]
        return sum(ms), ms, launches

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    total_ms, per_step, launches = timed_hot(args.steps, max(args.warmup, 3))

    # end to end from pinned HOST batches.  (a) the serving API: BatchScorer pipelines H2D(i+1) / forward(i) / D2H(i) and
    # replays one CUDA graph per batch; every step's copies are inside the timed region.  (b) the plain module call with
    # a blocking read of y per batch (what train.py's eval loop does), reported next to it.
    from armnet_b200 import BatchScorer
    scorer = BatchScorer(model, w['bsz'], w['nfield'], depth=6, compute_streams=3)

    def e2e_pipelined(steps):
        pending, acc = [], 0.0
        for i in range(steps):
            if len(pending) == scorer.depth:
                acc += float(scorer.result(pending.pop(0))[0])      # the step's result is read on the host
            pending.append(scorer.submit(*host[i % n_batches]))
        for t in pending:
            acc += float(scorer.result(t)[0])
        return acc

    def e2e_step(i):
        ids_h, vals_h = host[i % n_batches]
        x = {'id': ids_h.to(dev, non_blocking=True), 'value': vals_h.to(dev, non_blocking=True)}
        with torch.no_grad():
            y = model(x)
        return y.cpu()

    e2e_pipelined(max(args.warmup, 3))
    barrier()
    t0 = time.perf_counter()
    e2e_pipelined(args.steps)
    barrier()
    e2e_s = time.perf_counter() - t0
    for i in range(max(args.warmup, 3)):
        e2e_step(i)
    barrier()
    t0 = time.perf_counter()
    for i in range(args.steps):
        e2e_step(i)
    barrier()
    e2e_sync_s = time.perf_counter() - t0
    clocks = sampler.stop() if rank == 0 else None

    # the MLP kernels of the e2e step, timed alone (CUDA events, warm L2: in the pipeline their input was just written)
    mlp_info = None
    with torch.no_grad():
        mlp_fast = model.mlp._fast_ok(torch.empty(2, model.mlp.ninput, device=dev))
    if rank == 0 and mlp_fast:
        with torch.no_grad():
            x_mlp = model.interaction({'id': resident[0][0], 'value': resident[0][1]}, fold_bn=True).reshape(w['bsz'], -1)
            w_hi, w_lo, packed, (ac, hsplits, hout) = model.mlp._prepared()

            def ev_time(fn, n=20):
                for _ in range(3):
                    fn()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                torch.cuda.synchronize()
                a.record()
                for _ in range(n):
                    fn()
                b.record()
                torch.cuda.synchronize()
                return a.elapsed_time(b) / n * 1e-3

            t_gemm = ev_time(lambda: ops.mlp_first_linear(x_mlp, w_hi, w_lo))
            part = ops.mlp_first_linear(x_mlp, w_hi, w_lo)
            t_tail = ev_time(lambda: ops.mlp_tail(part, packed, model.mlp.nlayers - 1, model.mlp.noutput, w['bsz']))
            hidden_tc = (model.mlp.hidden_tensor_core and model.mlp.nlayers >= 2 and model.mlp.nhid <= 256
                         and model.mlp.noutput <= 4 and w['bsz'] >= model.mlp.hidden_tensor_core_min_batch)
            t_hidden = ev_time(lambda: ops.mlp_hidden_tc(part, w['bsz'], ac, hsplits, hout)) if hidden_tc else None
        flops = 3 * 2.0 * w['bsz'] * model.mlp.ninput * model.mlp.nhid         # three TF32 MMAs per fp32 product
        try:
            with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as f:
                tf32_peak = float(json.load(f)['bf16_tflops']) / 2.0           # TF32 runs at half the BF16 rate
            src = 'MEASURED_PEAKS.json bf16_tflops / 2'
        except Exception:
            tf32_peak, src = 1125.0, 'nominal 2250 / 2'
        mlp_info = {'gemm_kernel': 'mlp_gemm_tf32x3_pair_kernel (first Linear, 3xTF32 on tcgen05, cta_group::2)',
                    'gemm_us': t_gemm * 1e6, 'gemm_tf32_tflops': flops / t_gemm / 1e12,
                    'gemm_fp32_equivalent_tflops': flops / 3 / t_gemm / 1e12,
                    'roofline': {'bound': 'tensor', 'achieved': flops / t_gemm / 1e12, 'peak': tf32_peak,
                                 'unit': 'TFLOP/s', 'frac': flops / t_gemm / 1e12 / tf32_peak, 'peak_source': src},
                    'tail_kernel_us': t_tail * 1e6,
                    'hidden_tc_kernel_us': t_hidden * 1e6 if t_hidden is not None else None,
                    'after_first_linear': ('mlp_hidden_tc_kernel (hidden layers 2..n + output Linear on tcgen05, 128 samples '
                                           'per CTA)' if hidden_tc else 'mlp_tail_kernel (CUDA cores, 32 samples per CTA)')}

    eager = None
    if rank == 0 and world == 1 and not args.no_eager_leg:
        try:
            eager = reference_eager_on_gpu(w, model, resident[0], dev)
        except Exception as ex:                                  # e.g. out of memory on a shared box: report, keep going
            eager = {'unavailable': repr(ex)[:200]}
    trained_like_(model)
    tab, ld = model._shadow.get(table) if model.padded_table else (table.detach(), table.shape[1])
    W, Q, Vv = (t.detach() for t in model._attn_weights())
    state['ws'] = ops.fused_prepare(W, Q, Vv, w['alpha'], w['nfield'], one_head=model.one_head)
    tl_ms, _, _ = timed_hot(max(args.steps // 2, 5), 3)
    tl_steps = max(args.steps // 2, 5)

    t = torch.tensor([total_ms, e2e_s, tl_ms, e2e_sync_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms, e2e_s, tl_ms, e2e_sync_s = t.tolist()
    n = world
    samples = w['bsz'] * args.steps * n
    value = samples / (total_ms * 1e-3)
    ms_per_step = total_ms / args.steps
    peak, peak_src = peaks()
    abytes = algorithmic_bytes_per_sample(w) * w['bsz']
    achieved = abytes / (ms_per_step * 1e-3) / 1e9        # per GPU: one launch processes one batch
    fp32_pipe = None
    kind = ops.fused_fwd_kernel_kind(w['nfield'], w['nemb'], w['nhead'], w['nhid'], w['alpha'])
    kernel_name = {1: 'armnet_fwd_kernel', 2: 'armnet_fwd_mma_kernel (E x F products as 3xTF32 warp MMAs)',
                   3: 'armnet_fwd_tmem_kernel (attention logits by tcgen05.mma into tensor memory)'}[kind]
    rec = ncu_record(args.workload, kind)
    if rec and rec.get('fp32_pipe_cycles_per_row') and clocks and clocks.get('sm_mhz'):
        # warp-level FP32-pipe cycles the kernel needs per second / what 148 SMs x 4 sub-partitions offer at the sampled clock
        rows_per_s = value / n * w['nhead'] * w['nhid']
        need = rows_per_s / 32.0 * rec['fp32_pipe_cycles_per_row']   # one warp instruction serves 32 row-threads
        have = 148 * 4 * clocks['sm_mhz'] * 1e6
        fp32_pipe = {'frac': need / have, 'what': 'FMA-pipe busy fraction implied by the measured rate and the executed-'
                     'opcode histogram of ' + rec.get('source', 'the committed ncu capture')}
    out = {
        'metric': METRIC, 'value': value, 'unit': 'samples/s', 'n_gpus': n, 'steps': args.steps,
        'warmup': max(args.warmup, 3), 'ms_per_step': ms_per_step, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': dict(cfg, l2='flushed between timed steps (256 MiB memset outside the events)'
                       if flush is not None else 'not flushed'),
        'roofline': {'bound': 'hbm', 'achieved': achieved, 'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak,
                     'traffic': rec.get('dram_bytes') if rec else None,
                     'traffic_source': rec.get('source') if rec else None, 'peak_source': peak_src,
                     'algorithmic_bytes_per_launch': abytes,
                     'kernel': kernel_name + ' (attention parameters pre-contracted once per weight version by the '
                               'prepare kernels, ~4 us, outside the per-batch step)',
                     'note': 'path is instruction-issue/MUFU bound (entmax), not HBM bound; see DESIGN.md',
                     'fp32_pipe': fp32_pipe},
        'e2e': {'value': w['bsz'] * args.steps * n / e2e_s, 'unit': 'samples/s',
                'h2d_bytes_per_step': w['bsz'] * w['nfield'] * 12, 'd2h_bytes_per_step': w['bsz'] * 4,
                'ms_per_step': e2e_s / args.steps * 1e3,
                'what': 'armnet_b200.BatchScorer over pinned host batches: H2D ids+values, full ARMNetModel forward '
                        '(fused kernel, tcgen05 GEMM of the first Linear, tcgen05 hidden-layer + output kernel) as one CUDA graph per batch, D2H y; copies of '
                        'batch i+1 overlap the forward of batch i; 6 slots on 3 compute streams (the GEMM / tail of one batch overlaps the '
                        'fused kernel of the next)',
                'per_call_sync': {'value': w['bsz'] * args.steps * n / e2e_sync_s, 'ms_per_step': e2e_sync_s / args.steps * 1e3,
                                  'what': 'y = model(batch_from_pinned_host).cpu() per batch, no overlap'}},
        'gpu_launches': launches,
        'trained_like': {'value': w['bsz'] * tl_steps * n / (tl_ms * 1e-3), 'unit': 'samples/s',
                         'what': 'same shapes, embedding~N(0,1), attention weights x4 (sparse gates)'},
        'step_ms_min_med_max': [min(per_step), statistics.median(per_step), max(per_step)],
        'mlp': mlp_info,
        'clocks': clocks,
        'reference_gpu_eager': eager,
    }
    if not args.no_train_leg:
        del scorer, model
        torch.cuda.empty_cache()
        try:
            out['train_c4'] = train_c4_leg(dev, rank, world)
        except Exception as ex:
            out['train_c4'] = {'unavailable': repr(ex)[:300]}
    if rank == 0:
        if n == 1 and not args.no_cpu_baseline:
            st = {k: v.detach().cpu() for k, v in build_module(w).state_dict().items()}
            sample_b = args.cpu_sample or 1024
            rate, sec, cores = cpu_reference_rate(w, st, 3, 1, sample_b, full_model=False)
            out['cpu_baseline'] = {'value': rate, 'unit': 'samples/s', 'cores': cores, 'kind': 'port',
                                   'sample': f'3 steps x {sample_b} samples of the same workload (hot path '
                                             f'only), oracle port on torch CPU, {cores} threads, {sec:.2f} s/step, '
                                             f'cpu: {cpu_model_string()}'}
        emit_json(out)
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
