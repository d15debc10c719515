"""train.py's on-device AUC (Mann-Whitney rank statistic, no host sync) against sklearn.roc_auc_score, the metric the
reference computes per batch (utils/utils.py:85-106): ties, single-class batches (reference: 0), weighted averaging."""
import importlib.util
import os

import numpy as np
import pytest
import torch
from sklearn.metrics import roc_auc_score

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _train_module():
    spec = importlib.util.spec_from_file_location('armnet_train_cli', os.path.join(ROOT, 'train.py'))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


@pytest.mark.parametrize('n,ties', [(4096, False), (513, True), (7, True), (2, False)])
def test_auc_on_device_matches_sklearn(n, ties):
    tr = _train_module()
    g = torch.Generator().manual_seed(n)
    logits = torch.randn(n, generator=g)
    if ties:
        logits = (logits * 2).round() / 2                      # many exact ties
    target = (torch.rand(n, generator=g) < 0.4).float()
    target[0], target[1] = 1.0, 0.0                            # both classes present
    got = tr.auc_on_device(logits, target)
    assert got.dim() == 0
    assert abs(float(got) - roc_auc_score(target.numpy(), logits.numpy())) < 1e-9


def test_auc_single_class_is_zero_like_the_reference():
    tr = _train_module()
    logits = torch.randn(16)
    assert float(tr.auc_on_device(logits, torch.ones(16))) == 0.0
    assert float(tr.auc_on_device(logits, torch.zeros(16))) == 0.0


def test_meter_accumulates_tensors_without_reading_them():
    tr = _train_module()
    m = tr.Meter()
    m.update(torch.tensor(0.5, dtype=torch.float64), 10)
    m.update(torch.tensor(1.0, dtype=torch.float64), 30)
    assert isinstance(m.sum, torch.Tensor)
    assert abs(m.avg - 0.875) < 1e-12
