#!/usr/bin/env python
"""Training / evaluation driver with the reference's command line (train.py:15-50 of nusdbsystem/ARM-Net), running the
armnet / armnet_1h rows of create_model on the B200-native path.

    python train.py --model armnet_1h --h 10 --alpha 1.7 --lr 0.001 --data_dir /path/to/data/ --dataset frappe
    python train.py --model armnet --nfield 39 --nfeat 1000000 --dataset synthetic --eval_freq 50
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 train.py ...   (data parallel)

Same flags, defaults, loss (BCEWithLogits, mean), optimiser (dense Adam), gradient clamp to [-1, 1], early stopping on
validation AUC and log line format as the reference. Differences, all on the host side: the per-batch AUC is computed
on the GPU (rank statistic) instead of sklearn on the CPU, batches can be synthetic (no dataset download offline), and
under torchrun every rank takes a contiguous shard of each batch with ONE gradient all-reduce per step
(armnet_b200/parallel.py). run.sh's stale aliases --nlayer/--mlp_hid/--dnn_hid are accepted.
"""
import argparse
import logging
import os
import sys
import time

import numpy as np
import torch
from torch import nn, optim

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def get_args(argv=None):
    p = argparse.ArgumentParser(description='ARM-Net on armnet_b200')
    p.add_argument('--exp_name', default='test', type=str, help='exp name for log & checkpoint')
    # model config (train.py:18-33)
    p.add_argument('--model', default='armnet', type=str, help='armnet | armnet_1h')
    p.add_argument('--nfeat', type=int, default=5500, help='the number of features')
    p.add_argument('--nfield', type=int, default=10, help='the number of fields')
    p.add_argument('--nemb', type=int, default=10, help='embedding size')
    p.add_argument('--k', type=int, default=3, help='(unused by armnet; kept for CLI compatibility)')
    p.add_argument('--h', type=int, default=128, help='exponential neurons per head')
    p.add_argument('--mlp_nlayer', '--nlayer', type=int, default=2, help='the number of mlp layers')
    p.add_argument('--mlp_nhid', '--mlp_hid', type=int, default=256, help='mlp hidden units')
    p.add_argument('--dropout', default=0.0, type=float, help='dropout rate')
    p.add_argument('--nattn_head', type=int, default=4, help='the number of attention heads')
    p.add_argument('--ensemble', action='store_true', default=False, help='to ensemble with DNNs')
    p.add_argument('--dnn_nlayer', type=int, default=2, help='the number of mlp layers')
    p.add_argument('--dnn_nhid', '--dnn_hid', type=int, default=256, help='mlp hidden units')
    p.add_argument('--alpha', default=1.7, type=float, help='entmax alpha to control sparsity')
    # optimizer (train.py:35-39)
    p.add_argument('--epoch', type=int, default=100, help='number of maximum epochs')
    p.add_argument('--patience', type=int, default=1, help='number of epochs for stopping training')
    p.add_argument('--batch_size', type=int, default=4096, help='batch size (global, split across ranks)')
    p.add_argument('--lr', default=0.003, type=float, help='learning rate, default 3e-3')
    p.add_argument('--eval_freq', type=int, default=10000, help='max number of batches to train per epoch')
    # dataset (train.py:41-43)
    p.add_argument('--dataset', type=str, default='frappe', help="dataset folder under data_dir, or 'synthetic'")
    p.add_argument('--data_dir', type=str, default='./data/', help='path to dataset')
    p.add_argument('--workers', default=4, type=int, help='(unused: batches are sliced from in-memory tensors)')
    # log (train.py:45-48)
    p.add_argument('--log_dir', type=str, default='./log/', help='path to logs')
    p.add_argument('--report_freq', type=int, default=30, help='report frequency')
    p.add_argument('--seed', type=int, default=2025, help='seed for reproducibility')
    p.add_argument('--repeat', type=int, default=1, help='number of repeats with seeds [seed, seed+repeat)')
    p.add_argument('--synthetic_rows', type=int, default=200000, help='rows of the synthetic training split')
    p.add_argument('--torch_adam', action='store_true',
                   help='torch.optim.Adam + separate gradient bucket (default: FlatAdam, one fused clamp+Adam kernel over '
                        'a flat parameter / gradient bucket, armnet_b200/parallel.py)')
    p.add_argument('--host_data', action='store_true',
                   help='keep the splits in host memory and copy every batch (default: whole splits resident in HBM, '
                        'batches cut on the device, armnet_b200/data.py)')
    return p.parse_args(argv)


# ------------------------------------------------------------------------------------------------ data

def load_libsvm(path, nfield):
    """`label id:val id:val ...` per line -> (id [N,F] int32, value [N,F] f32, y [N] f32) through the native parser and
    its binary cache (armnet_b200/data.py); malformed lines are skipped like data_loader.py:37-44."""
    from armnet_b200.data import load_split
    ids, vals, y = load_split(path, nfield)
    return (torch.from_numpy(np.array(ids)), torch.from_numpy(np.array(vals)), torch.from_numpy(np.array(y)))


def synthetic_split(n, nfield, nfeat, seed):
    """Seeded CTR-shaped data with a learnable signal: ids uniform, values 1, label from a hidden linear model."""
    g = torch.Generator().manual_seed(seed)
    ids = torch.randint(0, nfeat, (n, nfield), generator=g, dtype=torch.int64)
    w = torch.randn(min(nfeat, 1 << 20), generator=torch.Generator().manual_seed(1234))
    score = w[ids % w.numel()].sum(1) / nfield ** 0.5
    y = (torch.rand(n, generator=g) < torch.sigmoid(2.0 * score - 1.0)).float()
    return ids, torch.ones(n, nfield), y


def load_data(args):
    if args.dataset == 'synthetic':
        n = args.synthetic_rows
        return [synthetic_split(m, args.nfield, args.nfeat, args.seed + i)
                for i, m in enumerate((n, max(n // 8, 1), max(n // 8, 1)))]
    import glob
    d = os.path.join(args.data_dir, args.dataset)
    files = []
    for pat in ('tr*libsvm', 'va*libsvm', 'te*libsvm'):           # data_loader.py:58-61
        hits = sorted(glob.glob(os.path.join(d, pat)))
        if not hits:    # only the binary cache was shipped (tools/make_frappe_cache.py): <name>.libsvm.armnet_bin/
            hits = [h[:-len('.armnet_bin')] for h in sorted(glob.glob(os.path.join(d, pat + '.armnet_bin')))]
        if not hits:
            raise FileNotFoundError(f'no {pat} under {d}')
        files.append(hits[0])
    return [load_libsvm(f, args.nfield) for f in files]


def batches(split, bsz, shuffle, gen):
    ids, vals, y = split
    n = y.shape[0]
    order = torch.randperm(n, generator=gen) if shuffle else torch.arange(n)
    for i in range(0, n, bsz):
        sel = order[i:i + bsz]
        yield {'id': ids[sel], 'value': vals[sel], 'y': y[sel]}


# ------------------------------------------------------------------------------------------------ metrics

def auc_on_device(logits, target):
    """ROC-AUC of one batch by the Mann-Whitney rank statistic (ties averaged), as a 0-dim DEVICE tensor: no host
    synchronisation. 0 when one class is missing, like utils/utils.py:104-106 (sklearn raises, the reference returns 0)."""
    pos = target > 0.5
    n_pos = pos.sum().double()
    n_neg = target.numel() - n_pos
    vals, inv, counts = torch.unique(logits.float(), sorted=True, return_inverse=True, return_counts=True)
    ends = torch.cumsum(counts, 0).double()
    avg_rank = ends - (counts.double() - 1) / 2
    r_pos = (avg_rank[inv] * pos.double()).sum()
    denom = n_pos * n_neg
    auc = (r_pos - n_pos * (n_pos + 1) / 2) / torch.clamp(denom, min=1.0)
    return torch.where(denom > 0, auc, torch.zeros_like(auc))


class Meter:
    """Running (last, weighted sum, count). Values may be 0-dim device tensors: they are accumulated on the device and
    only read back (one synchronisation) when .val / .avg are formatted -- at report lines and at the end of a split,
    not once per batch like the reference's loss.item() / sklearn AUC (train.py:118-122)."""

    def __init__(self):
        self.val = self.sum = self.cnt = 0.0

    def update(self, v, n=1):
        self.val = v
        self.sum = self.sum + v * n
        self.cnt += n

    @property
    def avg(self):
        return float(self.sum) / max(self.cnt, 1)


# ------------------------------------------------------------------------------------------------ loop

def run(epoch, model, split, args, plogger, dev, rank, world, optimizer=None, reducer=None, namespace='train'):
    from armnet_b200.parallel import shard_batch
    model.train() if optimizer else model.eval()
    crit = nn.BCEWithLogitsLoss(reduction='mean')
    time_avg, loss_avg, auc_avg = Meter(), Meter(), Meter()
    stamp = time.time()
    gen = torch.Generator().manual_seed(args.seed * 1000 + epoch)       # same shuffle on every rank
    from armnet_b200.data import DeviceSplit
    if isinstance(split, DeviceSplit):      # batches are cut (and sharded) on the device
        it = split.batches(args.batch_size, optimizer is not None, gen, rank, world if optimizer else 1)
        n_rows = len(split)
    else:
        it = batches(split, args.batch_size, optimizer is not None, gen)
        n_rows = split[2].shape[0]
    for bi, batch in enumerate(it):
        if isinstance(split, DeviceSplit):
            full_n, mine = batch['global_rows'], batch
        else:
            full_n = batch['y'].shape[0]
            mine = shard_batch(batch, rank, world) if (world > 1 and optimizer) else batch
        x = {'id': mine['id'].to(dev, non_blocking=True), 'value': mine['value'].to(dev, non_blocking=True)}
        target = mine['y'].to(dev, non_blocking=True)
        if optimizer:
            y = model(x)
            loss = crit(y.reshape(-1), target)
            if reducer is None:                                        # FlatAdam: bucket zero -> backward -> one
                optimizer.zero_grad()                                  # all-reduce -> one fused clamp+Adam kernel
                loss.backward()
                optimizer.step(weight=target.numel() * world / full_n)
            else:
                optimizer.zero_grad(set_to_none=True)
                loss.backward()
                reducer.step(weight=target.numel() * world / full_n)   # one all-reduce, then clamp (train.py:65)
                optimizer.step()
        else:
            with torch.no_grad():
                y = model(x)
                loss = crit(y.reshape(-1), target)
        loss_avg.update(loss.detach().double(), target.numel())       # device-side accumulation, no per-batch sync
        auc_avg.update(auc_on_device(y.detach().reshape(-1), target), target.numel())
        time_avg.update(time.time() - stamp)
        stamp = time.time()
        if bi % args.report_freq == 0 and rank == 0:
            plogger.info(f'Epoch [{epoch:3d}/{args.epoch}][{bi:3d}/{(n_rows + args.batch_size - 1) // args.batch_size}]\t{time_avg.val:.3f} ({time_avg.avg:.3f}) '
                         f'AUC {float(auc_avg.val):4f} ({auc_avg.avg:4f}) Loss {float(loss_avg.val):8.4f} ({loss_avg.avg:8.4f})')
        if bi >= args.eval_freq:
            break
    if hasattr(model, 'check_ids'):
        model.check_ids()            # the reference's IndexError for ids outside [0, nfeat), once per split (lazy flag)
    if rank == 0:
        plogger.info(f'{namespace}\tTime {time_avg.sum:10.1f}s AUC {auc_avg.avg:8.4f} Loss {loss_avg.avg:8.4f}')
    return auc_avg.avg


def main(args, data, rank, world, dev):
    import armnet_b200 as ab
    from armnet_b200.parallel import GradAllReducer, broadcast_buffers
    os.makedirs(os.path.join(args.log_dir, args.exp_name), exist_ok=True)
    plogger = logging.getLogger(f'{args.exp_name}_{rank}')
    plogger.setLevel(logging.INFO)
    plogger.handlers = [logging.StreamHandler(sys.stdout)]
    if rank == 0:
        plogger.handlers.append(logging.FileHandler(os.path.join(args.log_dir, args.exp_name, 'stdout.log')))
    model = ab.create_model(args, plogger).to(dev)
    if not args.host_data:
        from armnet_b200.data import DeviceSplit
        data = [DeviceSplit(*split, device=dev) for split in data]
        plogger.info(f'splits resident on {dev}: {sum(d.nbytes() for d in data) / 1e6:.1f} MB')
    plogger.info(vars(args))
    if args.torch_adam:
        optimizer = optim.Adam(model.parameters(), lr=args.lr)          # dense Adam over every parameter (train.py:62)
        reducer = GradAllReducer(model.parameters(), clamp=1.0)
    else:
        from armnet_b200.parallel import FlatAdam
        optimizer, reducer = FlatAdam(model.parameters(), lr=args.lr, clamp=1.0, shard_state=world >= 4), None   # train.py:62-65 in one kernel; optimizer state sharded from 4 ranks on
    best_valid, best_test, patience = 0.0, 0.0, 0
    start = time.time()
    for epoch in range(args.epoch):
        plogger.info(f'Epoch [{epoch:3d}/{args.epoch:3d}]')
        run(epoch, model, data[0], args, plogger, dev, rank, world, optimizer, reducer)
        broadcast_buffers(model)                                        # per-replica BN statistics -> rank 0's
        valid_auc = run(epoch, model, data[1], args, plogger, dev, rank, world, namespace='val')
        test_auc = run(epoch, model, data[2], args, plogger, dev, rank, world, namespace='test')
        if valid_auc >= best_valid:
            patience, best_valid, best_test = 0, valid_auc, test_auc
            plogger.info(f'best valid auc: valid {valid_auc:.4f}, test {test_auc:.4f}')
        else:
            patience += 1
            plogger.info(f'valid {valid_auc:.4f}, test {test_auc:.4f}')
            plogger.info(f'Early stopped, {patience}-th best auc at epoch {epoch - 1}')
        if patience >= args.patience:
            plogger.info(f'Final best valid auc {best_valid:.4f}, with test auc {best_test:.4f}')
            break
    plogger.info(f'Total running time: {time.time() - start:.1f}s')
    return best_valid, best_test


if __name__ == '__main__':
    args = get_args()
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        sys.exit('train.py: armnet_b200 needs a CUDA device (no CPU fallback)')
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group('nccl', device_id=dev)
    data = load_data(args)
    base = args.exp_name
    for args.seed in range(args.seed, args.seed + args.repeat):
        torch.manual_seed(args.seed)                                    # utils/utils.py:124-131 seed_everything
        torch.cuda.manual_seed_all(args.seed)
        args.exp_name = f'{base}_{args.seed}'
        main(args, data, rank, world, dev)
    if world > 1:
        dist.destroy_process_group()
