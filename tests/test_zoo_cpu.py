"""CPU check of the zoo model BODIES (dense torch ops) and their state_dict compatibility against the reference's
fixtures. The two lookups are CUDA kernels with no CPU path; here -- in the test only -- they are substituted by their
two-line torch definitions (layers.py:20-21, :36-37) so the bodies can be validated without a GPU. The product path is
unchanged: without the substitution the models raise on CPU tensors."""
import pytest
import torch

from zoo_common import ZOO_CASES, build, check_against_fixture, load


@pytest.fixture
def torch_lookups(monkeypatch):
    from armnet_b200 import ops

    def embed_gather(ids, values, table, clamp=None, **kw):
        if clamp is not None:
            values.clamp_(*clamp)
        return table[ids] * values.unsqueeze(2)

    def linear_gather(ids, values, weight, bias=None, err_flag=None):
        y = (weight.reshape(-1)[ids] * values).sum(1)
        return y + bias if bias is not None else y

    monkeypatch.setattr(ops, 'embed_gather', embed_gather)
    monkeypatch.setattr(ops, 'linear_gather', linear_gather)


@pytest.mark.parametrize('name', ZOO_CASES)
def test_zoo_bodies_match_reference_on_cpu(name, torch_lookups):
    check_against_fixture(name, torch.device('cpu'))


def test_zoo_models_have_no_cpu_path():
    ids, vals, _, state, _ = load('afm')
    model = build('afm')
    with pytest.raises(RuntimeError, match='CUDA tensors only'):
        model({'id': ids, 'value': vals})


def test_create_model_routes_zoo_models():
    import argparse
    import logging
    import armnet_b200 as ab
    args = argparse.Namespace(nfield=13, nfeat=400, nemb=10, nattn_head=4, alpha=1.7, h=8, k=3, mlp_nlayer=2, mlp_nhid=16,
                              dropout=0.0, ensemble=False, dnn_nlayer=2, dnn_nhid=16)
    for name, cls in [('afm', ab.zoo.AFMModel), ('dcn', ab.zoo.CrossNetModel), ('dcn+', ab.zoo.DCNModel),
                      ('cin', ab.zoo.CINModel), ('xdfm', ab.zoo.xDeepFMModel), ('afn', ab.zoo.AFNModel)]:
        args.model = name
        assert isinstance(ab.create_model(args, logging.getLogger('t')), cls)
