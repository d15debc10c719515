#!/usr/bin/env python
"""Golden TRAINING trajectories from the UNMODIFIED reference on CPU (dev container only, needs /root/reference):
    python tests/golden/make_train_traj.py
Re-enacts train.py's optimisation step by step on seeded synthetic batches --
    BCEWithLogitsLoss (train.py:60), optim.Adam(lr) (train.py:62), p.register_hook(clamp(-1, 1)) (train.py:64-65),
    y = model(batch); loss.backward(); optimizer.step() (train.py:109-114) --
and stores the initial state_dict, the batches, the per-step losses and the final parameters / BatchNorm buffers in
tests/golden/traj_<case>.npz.  tests/test_gpu_train.py replays the same steps with the CUDA path (fused forward and
backward kernels, train-mode arm_bn kernels, fused clamp+Adam) and compares.
"""
import os
import sys

import numpy as np
import torch

REF = os.environ.get('ARMNET_REFERENCE', '/root/reference')
sys.path.insert(0, REF)
from models.armnet import ARMNetModel as RefMH          # noqa: E402
from models.armnet_1h import ARMNetModel as Ref1H       # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
CASES = {
    # BASELINE config 1 shape (armnet_1h, Frappe: 10 fields, nemb 10, h 10), run.sh lr 1e-3 scaled up to move the loss
    'c1_1h': dict(model='armnet_1h', nfield=10, nfeat=600, nemb=10, nhead=1, alpha=1.7, nhid=10, d_k=10, mlp_nlayer=2,
                  mlp_nhid=32, B=256, steps=12, lr=1e-2),
    # BASELINE config 4 shape (armnet, 39 fields, nemb 16), train.py default lr 3e-3
    'c4_mh': dict(model='armnet', nfield=39, nfeat=800, nemb=16, nhead=4, alpha=1.7, nhid=16, d_k=16, mlp_nlayer=2,
                  mlp_nhid=32, B=128, steps=10, lr=3e-3),
}


def build(c):
    if c['model'] == 'armnet':
        return RefMH(c['nfield'], c['nfeat'], c['nemb'], c['nhead'], c['alpha'], c['nhid'], c['mlp_nlayer'],
                     c['mlp_nhid'], 0.0, False, 2, 16)
    return Ref1H(c['nfield'], c['nfeat'], c['nemb'], c['alpha'], c['nhid'], c['d_k'], c['mlp_nlayer'], c['mlp_nhid'],
                 0.0, False, 2, 16)


def run(name, c):
    torch.manual_seed(2025)
    model = build(c)
    gen = torch.Generator().manual_seed(7)
    with torch.no_grad():      # embeddings large enough that the gates are not uniform and gradients are not tiny
        model.embedding.embedding.weight.normal_(0.0, 0.3, generator=gen)
    state0 = {k: v.detach().clone() for k, v in model.state_dict().items()}
    ids = torch.randint(0, c['nfeat'], (c['steps'], c['B'], c['nfield']), generator=gen)
    vals = torch.rand(c['steps'], c['B'], c['nfield'], generator=gen) * 1.2
    # labels from a fixed random linear teacher on the ids, so the loss can actually go down
    teacher = torch.randn(c['nfeat'], generator=gen)
    target = (teacher[ids].sum(-1) + 0.5 * torch.randn(c['steps'], c['B'], generator=gen) > 0.8).float()
    crit = torch.nn.BCEWithLogitsLoss(reduction='mean')
    opt = torch.optim.Adam(model.parameters(), lr=c['lr'])
    for p in model.parameters():
        p.register_hook(lambda g: g.clamp(-1.0, 1.0))
    model.train()
    losses = []
    for t in range(c['steps']):
        y = model({'id': ids[t], 'value': vals[t].clone()})
        loss = crit(y, target[t])
        opt.zero_grad()
        loss.backward()
        opt.step()
        losses.append(float(loss.detach()))
    blob = {'ids': ids.numpy(), 'values': vals.numpy(), 'target': target.numpy(), 'losses': np.array(losses, np.float64)}
    for k, v in c.items():
        blob['cfg/' + k] = np.array(v)
    for k, v in state0.items():
        blob['state0/' + k] = v.numpy()
    for k, v in model.state_dict().items():
        blob['final/' + k] = v.detach().numpy()
    path = os.path.join(OUT, f'traj_{name}.npz')
    np.savez_compressed(path, **blob)
    print(f'{name}: {os.path.getsize(path) / 1024:.0f} KiB, losses {losses[0]:.5f} -> {losses[-1]:.5f}')


if __name__ == '__main__':
    torch.set_num_threads(8)
    for n, c in CASES.items():
        run(n, c)
