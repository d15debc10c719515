"""Shared helpers of the zoo tests: fixture loading and model construction (mirrors tests/golden/make_zoo_golden.py)."""
import os

import numpy as np
import torch

ZOO_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'zoo')
ZOO_CASES = sorted(f[:-4] for f in os.listdir(ZOO_DIR) if f.endswith('.npz'))
F_, V_, E_ = 13, 400, 10


def build(name):
    from armnet_b200 import zoo
    return {
        'afm': lambda: zoo.AFMModel(V_, E_, 8, 0.0),
        'dcn': lambda: zoo.CrossNetModel(F_, V_, E_, 3),
        'dcn+': lambda: zoo.DCNModel(F_, V_, E_, 3, 2, 16, 0.0),
        'cin': lambda: zoo.CINModel(F_, V_, E_, 2, 6),
        'xdfm': lambda: zoo.xDeepFMModel(F_, V_, E_, 2, 6, 2, 16, 0.0),
        'afn': lambda: zoo.AFNModel(F_, V_, E_, 12, 2, 16, 0.0, False, 2, 16),
        'afn_ens': lambda: zoo.AFNModel(F_, V_, E_, 12, 2, 16, 0.0, True, 2, 16),
    }[name]()


def load(name):
    z = np.load(os.path.join(ZOO_DIR, name + '.npz'))
    state = {k[6:]: torch.from_numpy(z[k]) for k in z.files if k.startswith('state/')}
    out = {k[4:]: torch.from_numpy(z[k]) for k in z.files if k.startswith('out/')}
    return torch.from_numpy(z['ids']), torch.from_numpy(z['values']), torch.from_numpy(z['target']), state, out


def check_against_fixture(name, device, tol_y=2e-5, tol_g=2e-4):
    """Eval forward + train-mode loss / gradients of the drop-in zoo model against the reference's fixture."""
    ids, vals, target, state, out = load(name)
    model = build(name)
    model.load_state_dict(state)                       # same parameter names and shapes as the reference
    model = model.to(device).eval()
    v = vals.clone().to(device)
    with torch.no_grad():
        y = model({'id': ids.to(device), 'value': v})
        e = model.embedding({'id': ids.to(device), 'value': v})
    scale = max(out['y'].abs().max().item(), 1.0)
    assert (y.cpu() - out['y']).abs().max().item() <= tol_y * scale
    assert torch.equal(e.cpu(), out['e'])             # the shared gather kernel: bit-exact
    assert torch.equal(v.cpu(), out['values_after'])  # afn clamps the caller's values in place (afn.py:52)
    assert torch.equal(model.embedding.embedding.weight.detach().cpu(), out['table_after'])   # afn.py:74-77
    if 'lin' in out:
        with torch.no_grad():
            lin = model.linear({'id': ids.to(device), 'value': v})
        assert (lin.cpu() - out['lin']).abs().max().item() <= 1e-6 * max(out['lin'].abs().max().item(), 1.0)
    model.load_state_dict(state)
    model.train()
    yt = model({'id': ids.to(device), 'value': vals.clone().to(device)})
    loss = torch.nn.BCEWithLogitsLoss()(yt, target.to(device))
    loss.backward()
    assert abs(loss.item() - out['loss'].item()) <= 2e-5
    gmax = max(out['grad/' + n].abs().max().item() for n, _ in model.named_parameters())
    for n, p in model.named_parameters():
        ref = out['grad/' + n]
        err = (p.grad.cpu() - ref).abs().max().item()
        assert err <= tol_g * ref.abs().max().item() + 1e-5 * gmax, (n, err)
