"""Data-parallel plumbing for the hot path: one process per GPU, batch-sharded, parameters replicated.

The forward path needs NO collective (every sample depends only on its own ids/values and on replicated parameters,
armnet.py:82-87).  Training adds exactly one all-reduce per step over a single flat fp32 gradient bucket, after which
the gradient is averaged and clamped to [-1, 1] -- the reference clamps every parameter gradient with a hook
(train.py:64-65); clamping AFTER the all-reduce keeps N ranks x (B/N) samples equivalent to the reference's single
process on the concatenated batch.  Works with any torch.distributed backend (NCCL on the B200 box, gloo in CPU tests).
"""
from typing import Dict, Iterable, List, Tuple

import torch
import torch.distributed as dist


def shard_bounds(n_rows: int, rank: int, world: int) -> Tuple[int, int]:
    """Rows [lo, hi) of a global batch owned by `rank`: contiguous, sizes differ by at most one, empty allowed."""
    base, rem = divmod(n_rows, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_batch(batch: Dict[str, torch.Tensor], rank: int, world: int) -> Dict[str, torch.Tensor]:
    """Slice every [B, ...] tensor of a {'id','value','y'} batch (data_loader.py:52-55) to this rank's rows."""
    n = next(iter(batch.values())).shape[0]
    lo, hi = shard_bounds(n, rank, world)
    return {k: v[lo:hi] for k, v in batch.items()}


class GradAllReducer:
    """One flat fp32 bucket for all parameter gradients; step(): all-reduce(sum) -> /world -> clamp -> scatter back."""

    def __init__(self, params: Iterable[torch.nn.Parameter], clamp: float = 1.0, group=None):
        self.params: List[torch.nn.Parameter] = [p for p in params if p.requires_grad]
        self.clamp = clamp
        self.group = group
        n = sum(p.numel() for p in self.params)
        p0 = self.params[0]
        self.bucket = torch.zeros(n, dtype=torch.float32, device=p0.device)
        self.views = []
        off = 0
        for p in self.params:
            self.views.append(self.bucket[off:off + p.numel()].view_as(p))
            off += p.numel()

    def numel(self) -> int:
        return self.bucket.numel()

    @torch.no_grad()
    def step(self, weight: float = 1.0) -> None:
        """`weight` = this rank's share of the global batch times world size (1.0 for equal shards), so the result
        is the gradient of the mean loss over the concatenated batch."""
        world = dist.get_world_size(self.group) if dist.is_initialized() else 1
        for p, v in zip(self.params, self.views):
            if p.grad is None:
                v.zero_()
            else:
                v.copy_(p.grad)
        if weight != 1.0:
            self.bucket.mul_(weight)
        if world > 1:
            dist.all_reduce(self.bucket, op=dist.ReduceOp.SUM, group=self.group)
            self.bucket.div_(world)
        if self.clamp is not None:
            self.bucket.clamp_(-self.clamp, self.clamp)      # train.py:65 semantics at global-batch level
        for p, v in zip(self.params, self.views):
            p.grad = v                                        # optimizer reads straight from the bucket


class FlatAdam:
    """Dense Adam for data-parallel training on ONE flat parameter buffer (train.py:62-65 semantics).

    The parameters are re-pointed to views of a flat fp32 buffer and their .grad to views of a flat gradient bucket, so
    autograd accumulates straight into the bucket and step() is: [one NCCL all-reduce(sum) of the bucket] -> ONE kernel
    (armnet_clamp_adam_f32) that averages, clamps to [-clamp, clamp] and applies the Adam update in a single pass over
    (p, g, m, v).  Equivalent to GradAllReducer.step() + torch.optim.Adam.step().

    shard_state=True (world > 1): the optimizer state is sharded over the ranks -- reduce-scatter of the bucket, the
    clamp+Adam kernel on this rank's 1/world slice of (p, g, m, v), all-gather of the updated parameters.  Same bytes on
    the wire as the all-reduce (reduce-scatter + all-gather IS the ring all-reduce), same update element by element, but
    the Adam pass and the m / v buffers shrink by the world size."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, clamp=1.0, group=None, shard_state=False):
        self.params = [p for p in params if p.requires_grad]
        self.lr, self.betas, self.eps, self.clamp, self.group = lr, betas, eps, clamp, group
        dev = self.params[0].device
        if dev.type != 'cuda':
            raise RuntimeError('FlatAdam runs on CUDA parameters only (armnet_b200 has no CPU path)')
        world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.sharded = bool(shard_state) and world > 1     # any true value; callers pass `world >= 4` (measured break-even)
        sizes = [(p.numel() + 3) // 4 * 4 for p in self.params]           # every view starts 16-byte aligned
        n = sum(sizes)
        if self.sharded:                                                  # equal, 16-byte aligned shards
            n = (n + 4 * world - 1) // (4 * world) * (4 * world)
        self.flat_p = torch.zeros(n, dtype=torch.float32, device=dev)
        self.flat_g = torch.zeros(n, dtype=torch.float32, device=dev)
        n_state = n // world if self.sharded else n
        self.exp_avg = torch.zeros(n_state, dtype=torch.float32, device=dev)
        self.exp_avg_sq = torch.zeros(n_state, dtype=torch.float32, device=dev)
        self.g_shard = torch.zeros(n_state, dtype=torch.float32, device=dev) if self.sharded else None
        self.t = 0
        off = 0
        with torch.no_grad():
            for p, sz in zip(self.params, sizes):
                view = self.flat_p[off:off + p.numel()].view_as(p)
                view.copy_(p)
                p.data = view
                p.grad = self.flat_g[off:off + p.numel()].view_as(p)
                p._armnet_flat_grad = True      # kernels may accumulate straight into p.grad (ops._FusedInteractionFn)
                off += sz

    def numel(self):
        return self.flat_p.numel()

    def zero_grad(self):
        self.flat_g.zero_()

    @torch.no_grad()
    def step(self, weight=1.0, events=None):
        """`weight` as in GradAllReducer.step (this rank's share of the global batch times world size).
        `events`: optional list; three CUDA events (start, after the all-reduce, after clamp+Adam) are appended to it
        (bench.py's train_c4 leg: exposed all-reduce time and bus bandwidth)."""
        from . import ops
        world = dist.get_world_size(self.group) if dist.is_initialized() else 1
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)] if events is not None else None
        if ev:
            ev[0].record()
        self.t += 1
        if self.sharded:
            rank = dist.get_rank(self.group)
            S = self.g_shard.numel()
            dist.reduce_scatter_tensor(self.g_shard, self.flat_g, op=dist.ReduceOp.SUM, group=self.group)
            if ev:
                ev[1].record()
            p_shard = self.flat_p[rank * S:(rank + 1) * S]
            ops.clamp_adam(p_shard, self.g_shard, self.exp_avg, self.exp_avg_sq, weight / world, self.clamp, self.lr,
                           self.betas[0], self.betas[1], self.eps, self.t)
            dist.all_gather_into_tensor(self.flat_p, p_shard, group=self.group)
        else:
            if world > 1:
                dist.all_reduce(self.flat_g, op=dist.ReduceOp.SUM, group=self.group)
            if ev:
                ev[1].record()
            ops.clamp_adam(self.flat_p, self.flat_g, self.exp_avg, self.exp_avg_sq, weight / world, self.clamp, self.lr,
                           self.betas[0], self.betas[1], self.eps, self.t)
        # the kernel writes through raw pointers: tell torch, so every cache keyed on (data_ptr, _version) -- padded
        # embedding table, pre-contracted attention workspace, folded arm_bn, split MLP weights -- is rebuilt
        torch.autograd.graph.increment_version(self.params)
        if ev:
            ev[2].record()
            events.extend(ev)


@torch.no_grad()
def broadcast_buffers(model: torch.nn.Module, src: int = 0, group=None) -> None:
    """BatchNorm running statistics are per-replica in DP (the reference has no SyncBN); make them rank `src`'s
    before evaluation / checkpointing."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    for b in model.buffers():
        dist.broadcast(b, src=src, group=group)


def max_over_ranks(value: float, device, group=None) -> float:
    """Timing rule of bench.py: a multi-GPU number is the MAX over ranks."""
    t = torch.tensor([value], dtype=torch.float64, device=device)
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())
