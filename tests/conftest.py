import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, 'tests', 'golden')
GOLDEN_CASES = sorted(f[:-4] for f in os.listdir(GOLDEN_DIR) if f.endswith('.npz') and not f.startswith('traj_'))


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with -m gpu)')


class Golden:
    """One fixture written by tests/golden/make_golden.py (outputs of the unmodified reference)."""

    def __init__(self, name):
        self.name = name
        d = np.load(os.path.join(GOLDEN_DIR, name + '.npz'))
        self.cfg = {k[4:]: d[k].item() for k in d.files if k.startswith('cfg/')}
        self.state = {k[6:]: torch.from_numpy(d[k]) for k in d.files if k.startswith('state/')}
        self.out = {k[4:]: torch.from_numpy(np.asarray(d[k])) for k in d.files if k.startswith('out/')}
        self.ids = torch.from_numpy(d['ids'])
        self.values = torch.from_numpy(d['values'])
        self.target = torch.from_numpy(d['target'])

    @property
    def one_head(self):
        return self.cfg['model'] == 'armnet_1h'


@pytest.fixture(params=GOLDEN_CASES)
def golden(request):
    return Golden(request.param)


def load_golden(name):
    return Golden(name)


def norm_rel(a, b):
    """max|a-b| / max|b|  (the norm-relative parity metric)."""
    a = torch.as_tensor(a).double()
    b = torch.as_tensor(b).double()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()
