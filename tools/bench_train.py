"""BASELINE config 4: data-parallel training step of armnet at the Criteo shape (nemb=16, bsz=4096 per GPU):
forward + BCE loss + backward + ONE gradient all-reduce (flat bucket) + clamp + dense Adam (train.py:109-114,62-65).
Prints samples/s (whole job, max over ranks) and a per-phase CUDA-event breakdown on rank 0.
  python tools/bench_train.py                         # 1 GPU
  python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/bench_train.py
"""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn as nn
import torch.distributed as dist

ap = argparse.ArgumentParser()
ap.add_argument('--steps', type=int, default=20)
ap.add_argument('--warmup', type=int, default=5)
ap.add_argument('--nemb', type=int, default=16)
ap.add_argument('--optimizer', default='fused', choices=['torch', 'fused'])
ap.add_argument('--shard', type=int, default=0, help='FlatAdam(shard_state=...): reduce-scatter / sharded Adam / all-gather')
args = ap.parse_args()
rank, world = int(os.environ.get('RANK', '0')), int(os.environ.get('WORLD_SIZE', '1'))
local = int(os.environ.get('LOCAL_RANK', '0'))
torch.cuda.set_device(local)
dev = torch.device('cuda', local)
if world > 1:
    dist.init_process_group('nccl', device_id=dev)
import armnet_b200 as ab
from armnet_b200.parallel import GradAllReducer

F, V, E, K, O, B = 39, 1000000, args.nemb, 4, 128, 4096
torch.manual_seed(2025)
model = ab.ARMNetModel(F, V, E, K, 1.7, O, 2, 256, 0.0, False, 2, 256).to(dev).train()
if args.optimizer == 'fused':
    from armnet_b200.parallel import FlatAdam
    stepper = FlatAdam(model.parameters(), lr=3e-3, clamp=1.0, shard_state=bool(args.shard))
else:
    opt = torch.optim.Adam(model.parameters(), lr=3e-3)
    reducer = GradAllReducer(model.parameters(), clamp=1.0)
crit = nn.BCEWithLogitsLoss()
g = torch.Generator().manual_seed(100 + rank)
batches = [(torch.randint(0, V, (B, F), generator=g).to(dev), torch.ones(B, F, device=dev),
            (torch.rand(B, generator=g) < 0.25).float().to(dev)) for _ in range(4)]
names = ['forward+loss', 'backward', 'allreduce+clamp+adam']
acc = [0.0] * 3


def step(i, timed):
    ids, vals, y = batches[i % 4]
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    ev[0].record()
    out = model({'id': ids, 'value': vals.clone()})
    loss = crit(out.reshape(-1), y)
    ev[1].record()
    if args.optimizer == 'fused':
        stepper.zero_grad()
    else:
        opt.zero_grad(set_to_none=True)
    loss.backward()
    ev[2].record()
    if args.optimizer == 'fused':
        stepper.step()
    else:
        reducer.step()
        opt.step()
    ev[3].record()
    return ev


for i in range(args.warmup):
    step(i, False)
if world > 1:
    dist.barrier()
torch.cuda.synchronize()
t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
t0.record()
evs = [step(i, True) for i in range(args.steps)]
t1.record()
if world > 1:
    dist.barrier()
torch.cuda.synchronize()
ms = t0.elapsed_time(t1)
t = torch.tensor([ms], dtype=torch.float64, device=dev)
if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
ms = t.item()
for ev in evs:
    for j in range(3):
        acc[j] += ev[j].elapsed_time(ev[j + 1])
checksum = float(sum(p.detach().double().abs().sum() for p in model.parameters()))   # equal trajectories -> equal checksums
if rank == 0:
    nparam = sum(p.numel() for p in model.parameters())
    print('param |.| checksum after %d steps: %.10e (shard_state=%d)' % (args.steps + args.warmup, checksum, args.shard))
    print(json.dumps({'metric': 'training samples/s (config 4: armnet nemb=%d, bsz=4096/GPU, dense Adam, 1 all-reduce/step)' % E,
                      'value': B * world * args.steps / (ms * 1e-3), 'n_gpus': world, 'ms_per_step': ms / args.steps,
                      'optimizer': args.optimizer, 'params': nparam, 'grad_bucket_MB': nparam * 4 / 1e6,
                      'phase_ms': {n: a / args.steps for n, a in zip(names, acc)}}))
if world > 1:
    dist.destroy_process_group()
