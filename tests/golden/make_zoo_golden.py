#!/usr/bin/env python
"""Golden vectors for the config-5 zoo models (afm, dcn, dcn+, cin, xdfm, afn) from the UNMODIFIED reference on CPU
(dev container only, needs /root/reference):   python tests/golden/make_zoo_golden.py
Writes tests/golden/zoo/<model>.npz: seeded inputs, the reference state_dict (BatchNorm statistics and a few
constant-initialised parameters randomised so nothing is trivially zero), the eval-mode output y, the embedding /
linear-term stage outputs, and a train-mode loss with every parameter gradient."""
import os
import sys

import numpy as np
import torch

REF = os.environ.get('ARMNET_REFERENCE', '/root/reference')
sys.path.insert(0, REF)
from models.afm import AFMModel                     # noqa: E402
from models.dcn import CrossNetModel, DCNModel      # noqa: E402
from models.xdfm import CINModel, xDeepFMModel      # noqa: E402
from models.afn import AFNModel                     # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'zoo')
F_, V_, E_, B_ = 13, 400, 10, 48
CASES = {
    'afm': lambda: AFMModel(V_, E_, 8, 0.0),
    'dcn': lambda: CrossNetModel(F_, V_, E_, 3),
    'dcn+': lambda: DCNModel(F_, V_, E_, 3, 2, 16, 0.0),
    'cin': lambda: CINModel(F_, V_, E_, 2, 6),
    'xdfm': lambda: xDeepFMModel(F_, V_, E_, 2, 6, 2, 16, 0.0),
    'afn': lambda: AFNModel(F_, V_, E_, 12, 2, 16, 0.0, False, 2, 16),
    'afn_ens': lambda: AFNModel(F_, V_, E_, 12, 2, 16, 0.0, True, 2, 16),
}


def run(name, make):
    torch.manual_seed(2025)
    model = make()
    gen = torch.Generator().manual_seed(11)
    with torch.no_grad():
        for n, p in model.named_parameters():
            if p.abs().max() == 0 or n.endswith('ensemble_layer.weight'):      # zero / constant inits -> random
                p.copy_(torch.randn(p.shape, generator=gen) * 0.1)
        for n, b in model.named_buffers():
            if n.endswith('running_mean'):
                b.copy_(torch.randn(b.shape, generator=gen) * 0.2)
            elif n.endswith('running_var'):
                b.copy_(torch.rand(b.shape, generator=gen) + 0.5)
        model.embedding.embedding.weight.mul_(3.0)
    state = {k: v.detach().clone() for k, v in model.state_dict().items()}
    ids = torch.randint(0, V_, (B_, F_), generator=gen)
    vals = torch.rand(B_, F_, generator=gen) * 1.3 - 0.1 if name.startswith('afn') else torch.rand(B_, F_, generator=gen)
    target = (torch.rand(B_, generator=gen) < 0.3).float()
    out = {}
    model.eval()
    with torch.no_grad():
        v = vals.clone()
        out['y'] = model({'id': ids, 'value': v}).numpy()
        out['values_after'] = v.numpy()
        out['table_after'] = model.embedding.embedding.weight.detach().numpy().copy()   # afn rewrites it in place
        out['e'] = model.embedding({'id': ids, 'value': v}).numpy()
        if hasattr(model, 'linear'):
            out['lin'] = model.linear({'id': ids, 'value': v}).numpy()
    model.load_state_dict(state)
    model.train()
    y = model({'id': ids, 'value': vals.clone()})
    loss = torch.nn.BCEWithLogitsLoss()(y, target)
    loss.backward()
    out['loss'] = loss.detach().numpy()
    for n, p in model.named_parameters():
        out['grad/' + n] = p.grad.numpy()
    blob = {'ids': ids.numpy(), 'values': vals.numpy(), 'target': target.numpy()}
    blob.update({'state/' + k: v.numpy() for k, v in state.items()})
    blob.update({'out/' + k: v for k, v in out.items()})
    os.makedirs(OUT, exist_ok=True)
    path = os.path.join(OUT, name + '.npz')
    np.savez_compressed(path, **blob)
    print(f'{name}: {os.path.getsize(path) / 1024:.0f} KiB, y[:3] = {out["y"][:3]}')


if __name__ == '__main__':
    torch.set_num_threads(8)
    for n, mk in CASES.items():
        run(n, mk)
