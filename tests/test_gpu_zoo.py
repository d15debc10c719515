"""BASELINE config 5: afm / dcn / dcn+ / cin / xdfm / afn on the shared CUDA gather kernels, against fixtures generated
from the unmodified reference (tests/golden/make_zoo_golden.py)."""
import pytest
import torch

from zoo_common import ZOO_CASES, check_against_fixture

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('name', ZOO_CASES)
def test_zoo_model_matches_reference(name):
    check_against_fixture(name, torch.device('cuda:0'))


def test_linear_gather_int32_and_bad_ids():
    from armnet_b200 import ops
    d = torch.device('cuda:0')
    g = torch.Generator().manual_seed(0)
    w = torch.randn(1000, 1, generator=g).to(d)
    ids = torch.randint(0, 1000, (257, 39), generator=g).to(d)
    vals = torch.rand(257, 39, generator=g).to(d)
    bias = torch.tensor([0.25], device=d)
    ref = (w[ids].squeeze(2) * vals).sum(1) + bias
    assert torch.allclose(ops.linear_gather(ids, vals, w, bias), ref, rtol=1e-6, atol=1e-6)
    assert torch.allclose(ops.linear_gather(ids.int(), vals, w, bias), ref, rtol=1e-6, atol=1e-6)
    flag = ops.new_error_flag(d)
    ids[3, 5] = 1000
    ops.linear_gather(ids, vals, w, bias, err_flag=flag)
    with pytest.raises(IndexError):
        ops.raise_if_bad_ids(flag)
