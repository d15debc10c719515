"""N > 1 host logic on CPU: world_size 2, gloo backend. The forward path has no collective, so what
is tested is batch sharding, max-over-ranks timing and the single flat-bucket gradient all-reduce: two ranks on half
batches must reproduce the single-process gradient of the concatenated batch (clamped after the reduction)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _toy_model():
    torch.manual_seed(5)
    return torch.nn.Sequential(torch.nn.Linear(6, 8), torch.nn.ReLU(), torch.nn.Linear(8, 1))


def _toy_batch(n=10):
    g = torch.Generator().manual_seed(9)
    return {'value': torch.randn(n, 6, generator=g) * 3, 'y': (torch.rand(n, generator=g) < 0.3).float()}


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from armnet_b200.parallel import GradAllReducer, broadcast_buffers, max_over_ranks, shard_batch, shard_bounds
    model = _toy_model()
    full = _toy_batch()
    mine = shard_batch(full, rank, world)
    lo, hi = shard_bounds(10, rank, world)
    assert mine['y'].shape[0] == hi - lo
    loss = torch.nn.BCEWithLogitsLoss(reduction='mean')(model(mine['value']).squeeze(1), mine['y'])
    loss.backward()
    red = GradAllReducer(model.parameters(), clamp=1.0)
    red.step(weight=(hi - lo) * world / 10)      # unequal shards are weighted by their share of the batch
    flat = torch.cat([p.grad.reshape(-1) for p in model.parameters()])
    t = max_over_ranks(1.0 + rank, torch.device('cpu'))
    bn = torch.nn.BatchNorm1d(3)
    bn.running_mean.fill_(float(rank + 1))
    broadcast_buffers(bn, src=0)
    if rank == 0:
        ret.put((flat.clone(), t, bn.running_mean.clone(), red.numel()))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_bounds_cover_and_are_ragged_safe():
    from armnet_b200.parallel import shard_bounds
    for n in (0, 1, 7, 8, 4096, 4097):
        for world in (1, 2, 3, 8):
            spans = [shard_bounds(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


@pytest.mark.timeout(120)
def test_two_rank_gradient_equals_single_process():
    ctx = mp.get_context('spawn')
    ret = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, ret)) for r in range(2)]
    for p in procs:
        p.start()
    flat, t, rm, numel = ret.get(timeout=90)
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
    # single process on the concatenated batch, clamp like train.py:65
    model = _toy_model()
    full = _toy_batch()
    loss = torch.nn.BCEWithLogitsLoss(reduction='mean')(model(full['value']).squeeze(1), full['y'])
    loss.backward()
    ref = torch.cat([p.grad.reshape(-1) for p in model.parameters()]).clamp(-1, 1)
    assert numel == ref.numel()
    torch.testing.assert_close(flat, ref, rtol=1e-5, atol=1e-6)
    assert t == 2.0                                   # max over ranks
    assert torch.equal(rm, torch.ones(3))             # rank 0's buffers everywhere
