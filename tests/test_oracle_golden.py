"""Pin the CPU oracle (oracle/armnet_oracle.py) against outputs of the unmodified reference
(tests/golden/*.npz, written by tests/golden/make_golden.py). CPU only."""
import torch

from oracle import armnet_oracle as oracle


def test_oracle_reproduces_reference_stages(golden):
    vals = golden.values.clone()
    out = oracle.forward(golden.state, float(golden.cfg['alpha']), golden.ids, vals, training=False)
    # in-place clamp of the caller's tensor (armnet.py:82)
    assert torch.equal(vals, golden.out['values_after'])
    # gather is bit-exact by contract
    assert torch.equal(out['e'], golden.out['e'])
    for k in ('g', 'p', 'w', 's'):
        assert out[k].shape == golden.out[k].shape, k
        assert torch.equal(out[k], golden.out[k]), f'{golden.name}: stage {k} differs from the reference'
    assert torch.equal(out['z'], golden.out['z'])
    assert out['y'].shape == golden.out['y'].shape
    torch.testing.assert_close(out['y'], golden.out['y'], rtol=0, atol=0)


def test_oracle_train_mode_forward(golden):
    if 'y_train' not in golden.out:
        return
    out = oracle.forward(golden.state, float(golden.cfg['alpha']), golden.ids, golden.values.clone(),
                         training=True)
    torch.testing.assert_close(out['y'], golden.out['y_train'], rtol=0, atol=0)


def test_oracle_entmax_backward_matches_autograd(golden):
    """entmax.py:71-80 restated; compared with autograd through the reference-pinned forward's p."""
    alpha = float(golden.cfg['alpha'])
    if alpha == 1.:
        return
    p = golden.out['p']
    gen = torch.Generator().manual_seed(0)
    dp = torch.randn(p.shape, generator=gen)
    dx = oracle.entmax_backward(p, dp, alpha)
    # property of the Jacobian: rows of dX sum to ~0 on the support and vanish off it
    assert torch.all(dx[p == 0] == 0)
    assert dx.sum(-1).abs().max() < 1e-4 * dp.abs().max()


def test_reference_init_state_shapes():
    st = oracle.reference_init_state('armnet', 39, 1000, 10, 4, 128)
    assert st['attn_layer.bilinear_w'].shape == (4, 10, 10)
    assert st['attn_layer.values'].shape == (4, 128, 39)
    assert st['mlp.mlp.0.weight'].shape == (256, 5120)
    ids = torch.randint(0, 1000, (8, 39))
    out = oracle.forward(st, 1.7, ids, torch.ones(8, 39))
    assert out['z'].shape == (8, 512, 10) and out['y'].shape == (8,)
    st1 = oracle.reference_init_state('armnet_1h', 10, 5382, 10, 1, 10)
    out = oracle.forward(st1, 1.7, torch.randint(0, 5382, (4, 10)), torch.ones(4, 10))
    assert out['z'].shape == (4, 10, 10)


def test_c_oracle_matches_reference(golden):
    """The plain-C restatement (oracle/armnet_oracle.c) against the reference fixtures: gather bit-exact,
    gates within 2e-6, logits / sums / output within 1e-6 norm-relative (libm vs ATen pow/exp: <= 2 ulp)."""
    import numpy as np
    from conftest import norm_rel
    from oracle import c_oracle
    vals = golden.values.clone().numpy()
    st = {k: v.numpy() for k, v in golden.state.items()}
    out = c_oracle.hot_path(st, float(golden.cfg['alpha']), golden.ids.numpy(), vals)
    B = golden.ids.shape[0]
    assert np.array_equal(vals, golden.out['values_after'].numpy())
    assert np.array_equal(out['e'], golden.out['e'].numpy())
    ref = {k: golden.out[k].reshape(B, -1, golden.out[k].shape[-1]) for k in ('g', 'p', 's', 'z')}
    assert norm_rel(out['g'], ref['g']) <= 1e-6
    assert np.abs(out['p'] - ref['p'].numpy()).max() <= 2e-6
    assert norm_rel(out['s'], ref['s']) <= 2e-6
    assert norm_rel(out['z'], ref['z']) <= 1e-6
