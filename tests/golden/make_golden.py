#!/usr/bin/env python
"""Generate golden input/output vectors by running the UNMODIFIED reference on CPU.

Run in the dev container only (needs /root/reference):
    python tests/golden/make_golden.py
Writes tests/golden/<case>.npz.  Each file holds the seeded inputs, the reference
state_dict, and the per-stage tensors captured with forward hooks from the reference
modules themselves (models/armnet.py, models/armnet_1h.py, models/layers.py,
utils/entmax.py in /root/reference):
    e  embedding output          layers.py:20-21
    g  attention logits          armnet.py:33-34 / armnet_1h.py:30-32  (input of .sparsemax)
    p  sparse gates              entmax.py:29-68                        (output of .sparsemax)
    w  gates * values            armnet.py:36 / armnet_1h.py:34
    s  pre-exp interaction sums  armnet.py:87 (same einsum re-run on the captured e, w)
    z  exp(s)                    armnet.py:86 (input of arm_bn)
    y  model output (eval mode)  armnet.py:101
plus, for cases with train=True, a train-mode forward/backward: y_train, loss, dz and
every parameter gradient (BCEWithLogits, mean reduction, train.py:60,110).
The GPU box has no /root/reference, so these fixtures are what the parity tests read.
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

# name -> dict(model, nfield, nfeat, nemb, nhead, alpha, nhid, d_k, mlp_nlayer, mlp_nhid,
#              ensemble, B, values, emb_std, [attn_scale, val_scale], train)
CASES = {
    # BASELINE config 1 shape (armnet_1h on Frappe), reference init, all-ones values
    'c1_1h_frappe': dict(model='armnet_1h', nfield=10, nfeat=5382, nemb=10, nhead=1, alpha=1.7, nhid=10,
                         d_k=10, mlp_nlayer=2, mlp_nhid=16, ensemble=False, B=64, values='ones',
                         emb_std=None, train=True),
    # BASELINE config 2 shape (C2a), reference init
    'c2a_init': dict(model='armnet', nfield=39, nfeat=1000, nemb=10, nhead=4, alpha=1.7, nhid=128,
                     d_k=10, mlp_nlayer=2, mlp_nhid=16, ensemble=False, B=4, values='ones',
                     emb_std=None, train=False),
    # C2a shape, trained-like embedding scale (sparse gates), values exercising the clamp
    'c2a_sparse': dict(model='armnet', nfield=39, nfeat=1000, nemb=10, nhead=4, alpha=1.7, nhid=128,
                       d_k=10, mlp_nlayer=2, mlp_nhid=16, ensemble=False, B=4, values='uniform',
                       emb_std=1.0, attn_scale=4.0, val_scale=2.0, train=False),
    # C2b: run.sh Criteo line, alpha=2 (sparsemax), h=64
    'c2b_alpha2': dict(model='armnet', nfield=39, nfeat=1000, nemb=10, nhead=4, alpha=2.0, nhid=64,
                       d_k=10, mlp_nlayer=2, mlp_nhid=16, ensemble=False, B=6, values='uniform',
                       emb_std=0.7, attn_scale=3.0, val_scale=2.0, train=True),
    # C3 shape: Avazu-like, nemb=100, alpha=1.5
    'c3_wide': dict(model='armnet', nfield=22, nfeat=2000, nemb=100, nhead=1, alpha=1.5, nhid=32,
                    d_k=100, mlp_nlayer=3, mlp_nhid=16, ensemble=False, B=16, values='uniform',
                    emb_std=0.3, attn_scale=12.0, val_scale=2.0, train=True),
    # C4 shape: nemb=16, training
    'c4_e16': dict(model='armnet', nfield=39, nfeat=1000, nemb=16, nhead=4, alpha=1.7, nhid=32,
                   d_k=16, mlp_nlayer=2, mlp_nhid=16, ensemble=False, B=12, values='uniform',
                   emb_std=0.5, attn_scale=4.0, val_scale=2.0, train=True),
    # softmax branch (alpha == 1.)
    'softmax_1h': dict(model='armnet_1h', nfield=10, nfeat=500, nemb=10, nhead=1, alpha=1.0, nhid=24,
                       d_k=10, mlp_nlayer=1, mlp_nhid=16, ensemble=False, B=16, values='uniform',
                       emb_std=1.0, train=True),
    # alpha > 2 (run.sh MovieLens line: alpha 2.5, 3 fields), ensemble branch on
    'alpha25_ens': dict(model='armnet', nfield=3, nfeat=300, nemb=10, nhead=1, alpha=2.5, nhid=8,
                        d_k=10, mlp_nlayer=2, mlp_nhid=16, ensemble=True, B=32, values='ones',
                        emb_std=1.0, train=True),
    # one-head with d_k != nemb, ensemble, odd sizes, B == 1 (0-dim output)
    'odd_1h_b1': dict(model='armnet_1h', nfield=7, nfeat=97, nemb=6, nhead=1, alpha=1.3, nhid=5,
                      d_k=9, mlp_nlayer=1, mlp_nhid=8, ensemble=True, B=1, values='uniform',
                      emb_std=1.0, train=False),
    # Diabetes130 shape (43 fields), many heads of one neuron (run.sh: --h 1 --nattn_head 32)
    'diab_43f': dict(model='armnet', nfield=43, nfeat=369, nemb=10, nhead=32, alpha=1.7, nhid=1,
                     d_k=10, mlp_nlayer=1, mlp_nhid=16, ensemble=False, B=9, values='uniform',
                     emb_std=1.5, attn_scale=2.0, train=True),
}


def build(cfg):
    if cfg['model'] == 'armnet':
        return RefMH(cfg['nfield'], cfg['nfeat'], cfg['nemb'], cfg['nhead'], cfg['alpha'], cfg['nhid'],
                     cfg['mlp_nlayer'], cfg['mlp_nhid'], 0.0, cfg['ensemble'], 2, 16)
    return Ref1H(cfg['nfield'], cfg['nfeat'], cfg['nemb'], cfg['alpha'], cfg['nhid'], cfg['d_k'],
                 cfg['mlp_nlayer'], cfg['mlp_nhid'], 0.0, cfg['ensemble'], 2, 16)


def make_inputs(cfg, gen):
    B, F, V = cfg['B'], cfg['nfield'], cfg['nfeat']
    ids = torch.randint(0, V, (B, F), generator=gen, dtype=torch.int64)
    if cfg['values'] == 'ones':
        vals = torch.ones(B, F)
    else:  # U(-0.1, 1.3): some below 1e-3 (and negative), some above 1 -> clamp both ways
        vals = torch.rand(B, F, generator=gen) * 1.4 - 0.1
    y = (torch.rand(B, generator=gen) < 0.25).float()
    return ids, vals, y


def run_case(name, cfg):
    torch.manual_seed(2025)                      # train.py:47 default seed
    model = build(cfg)
    gen = torch.Generator().manual_seed(2025)
    if cfg['emb_std'] is not None:               # "trained-like" scale: gates become sparse
        with torch.no_grad():
            model.embedding.embedding.weight.normal_(0.0, cfg['emb_std'], generator=gen)
            # non-trivial BN statistics so eval-mode BN is exercised
            model.arm_bn.running_mean.uniform_(0.5, 1.5, generator=gen)
            model.arm_bn.running_var.uniform_(0.5, 2.0, generator=gen)
            model.arm_bn.weight.uniform_(0.5, 1.5, generator=gen)
            model.arm_bn.bias.uniform_(-0.5, 0.5, generator=gen)
            # trained attention weights are far from Xavier init: scale them so |g| reaches O(1..10)
            a = cfg.get('attn_scale', 1.0)
            q_par = model.attn_layer.query
            w_par = (model.attn_layer.bilinear_w if cfg['model'] == 'armnet'
                     else model.attn_layer.bilinear_w.weight)
            q_par.mul_(a)
            w_par.mul_(a)
            model.attn_layer.values.mul_(cfg.get('val_scale', 1.0))
    ids, vals, target = make_inputs(cfg, gen)

    cap = {}
    def grab(**names):                            # hook that records tensors and returns None
        def hook(m, i, o):
            for key, src in names.items():
                cap[key] = i[0] if src == 'in' else o
        return hook

    hooks = [
        model.embedding.register_forward_hook(grab(e='out')),
        model.attn_layer.sparsemax.register_forward_hook(grab(g='in', p='out')),
        model.attn_layer.register_forward_hook(grab(w='out')),
        model.arm_bn.register_forward_hook(grab(z='in')),
    ]
    out = {}
    state = {k: v.detach().clone() for k, v in model.state_dict().items()}

    model.eval()
    with torch.no_grad():
        v_in = vals.clone()
        y = model({'id': ids, 'value': v_in})
    out['values_after'] = v_in.numpy()           # armnet.py:82 clamps the caller's tensor in place
    for k in ('e', 'g', 'p', 'w', 'z'):
        out[k] = cap[k].detach().numpy().copy()
    e, w = cap['e'], cap['w']
    if cfg['model'] == 'armnet':
        out['s'] = torch.einsum('bfe,bkof->bkoe', e, w).numpy()
    else:
        out['s'] = torch.einsum('bfe,bof->boe', e, w).numpy()
    out['y'] = y.numpy()

    if cfg['train']:
        model.train()
        model.zero_grad()
        zs = []
        def keep_z(m, i):
            i[0].retain_grad()
            zs.append(i[0])

        h = model.arm_bn.register_forward_pre_hook(keep_z)
        y_t = model({'id': ids, 'value': vals.clone()})
        loss = torch.nn.BCEWithLogitsLoss(reduction='mean')(y_t, target)
        loss.backward()
        h.remove()
        out['y_train'] = y_t.detach().numpy()
        out['loss'] = loss.detach().numpy()
        out['dz'] = zs[0].grad.numpy()
        for n, p_ in model.named_parameters():
            out['grad/' + n] = p_.grad.numpy()
        for k, v in model.state_dict().items():   # BN running stats after one train step
            if 'running' in k or 'num_batches' in k:
                out['after/' + k] = v.numpy().copy()
    for hk in hooks:
        hk.remove()

    blob = {'ids': ids.numpy(), 'values': vals.numpy(), 'target': target.numpy()}
    for k, v in cfg.items():
        if v is not None:
            blob['cfg/' + k] = np.array(v)
    for k, v in state.items():
        blob['state/' + k] = v.numpy()
    for k, v in out.items():
        blob['out/' + k] = v
    path = os.path.join(OUT, name + '.npz')
    np.savez_compressed(path, **blob)
    print(f'{name}: {os.path.getsize(path) / 1024:.0f} KiB  y[:3]={np.atleast_1d(out["y"])[:3]}')


if __name__ == '__main__':
    torch.set_num_threads(8)
    for name, cfg in CASES.items():
        run_case(name, cfg)
