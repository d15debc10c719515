"""Parity of the CUDA path (through the C ABI) against the reference-generated golden fixtures and the oracle.
Tolerances (SURVEY.md 8d / BASELINE.json north_star):
  gather e                      bit-exact
  logits g, sums s, output z    max|a-b| / max|b| <= 1e-5
  gates p                       abs <= 2e-6   (entries at the support boundary are ill-conditioned)
"""
import pytest
import torch

from conftest import load_golden, norm_rel, GOLDEN_CASES

pytestmark = pytest.mark.gpu

TOL_NORM = 1e-5
TOL_P = 2e-6


def dev():
    return torch.device('cuda:0')


def hot_params(g):
    st = g.state
    if g.one_head:
        return st['attn_layer.bilinear_w.weight'], st['attn_layer.query'], st['attn_layer.values']
    return st['attn_layer.bilinear_w'], st['attn_layer.query'], st['attn_layer.values']


def build_model(g, device):
    import armnet_b200 as ab
    c = g.cfg
    if c['model'] == 'armnet':
        m = ab.ARMNetModel(c['nfield'], c['nfeat'], c['nemb'], c['nhead'], c['alpha'], c['nhid'], c['mlp_nlayer'],
                           c['mlp_nhid'], 0.0, bool(c['ensemble']), 2, 16)
    else:
        m = ab.ARMNet1H(c['nfield'], c['nfeat'], c['nemb'], c['alpha'], c['nhid'], c['d_k'], c['mlp_nlayer'],
                        c['mlp_nhid'], 0.0, bool(c['ensemble']), 2, 16)
    m.load_state_dict(g.state)
    return m.to(device)


def test_gather_bit_exact(golden):
    from armnet_b200 import ops
    d = dev()
    vals = golden.values.clone().to(d)
    e = ops.embed_gather(golden.ids.to(d), vals, golden.state['embedding.embedding.weight'].to(d),
                         clamp=(0.001, 1.0))
    assert torch.equal(e.cpu(), golden.out['e'])
    assert torch.equal(vals.cpu(), golden.out['values_after'])      # in-place clamp, armnet.py:82
    # int32 ids are accepted too
    e32 = ops.embed_gather(golden.ids.to(d).int(), vals, golden.state['embedding.embedding.weight'].to(d))
    assert torch.equal(e32.cpu(), golden.out['e'])


@pytest.mark.parametrize('solver', ['auto', 'bisect'])
@pytest.mark.parametrize('padded', [False, True])
def test_fused_stages_match_reference(golden, solver, padded):
    from armnet_b200 import ops
    d = dev()
    c = golden.cfg
    W, Q, Vv = (t.to(d) for t in hot_params(golden))
    table = golden.state['embedding.embedding.weight'].to(d)
    E = table.shape[1]
    ld = E
    if padded:
        ld = (E + 3) // 4 * 4 + 4
        tp = torch.zeros(table.shape[0], ld, device=d)
        tp[:, :E] = table
        table = tp
    vals = golden.values.clone().to(d)
    z, ex = ops.fused_forward(golden.ids.to(d), vals, table, W, Q, Vv, float(c['alpha']),
                              one_head=golden.one_head, ld=ld, nemb=E,
                              solver=ops.SOLVER_BISECT if solver == 'bisect' else ops.SOLVER_AUTO,
                              want_tau=True, want_p=True, want_g=True, want_s=True)
    torch.cuda.synchronize()
    assert torch.equal(vals.cpu(), golden.out['values_after'])
    B = golden.ids.shape[0]
    ref = {k: golden.out[k].reshape(B, -1, golden.out[k].shape[-1]) for k in ('g', 'p', 's', 'z')}
    assert z.shape == ref['z'].shape
    assert norm_rel(ex['g'].cpu(), ref['g']) <= TOL_NORM
    assert (ex['p'].cpu() - ref['p']).abs().max().item() <= TOL_P
    assert norm_rel(ex['s'].cpu(), ref['s']) <= TOL_NORM
    assert norm_rel(z.cpu(), ref['z']) <= TOL_NORM
    # gates are a distribution; (tau, sum) rebuilds them
    assert (ex['p'].sum(-1) - 1).abs().max().item() < 1e-5
    assert torch.isfinite(ex['tau']).all()


def test_fused_plain_call_equals_debug_call(golden):
    """The optional outputs must not change z."""
    from armnet_b200 import ops
    d = dev()
    W, Q, Vv = (t.to(d) for t in hot_params(golden))
    table = golden.state['embedding.embedding.weight'].to(d)
    a = float(golden.cfg['alpha'])
    z0, _ = ops.fused_forward(golden.ids.to(d), golden.values.clone().to(d), table, W, Q, Vv, a,
                              one_head=golden.one_head)
    z1, _ = ops.fused_forward(golden.ids.to(d), golden.values.clone().to(d), table, W, Q, Vv, a,
                              one_head=golden.one_head, want_p=True, want_g=True, want_s=True, want_tau=True)
    assert torch.equal(z0, z1)


def test_module_forward_matches_reference(golden):
    d = dev()
    m = build_model(golden, d).eval()
    for padded, fuse_bn in ((True, True), (False, True), (True, False)):
        m.padded_table = padded
        m.fuse_bn = fuse_bn
        vals = golden.values.clone().to(d)
        with torch.no_grad():
            y = m({'id': golden.ids.to(d), 'value': vals})
        assert y.shape == golden.out['y'].shape          # 0-dim when B == 1 (armnet.py:101)
        assert torch.equal(vals.cpu(), golden.out['values_after'])
        err = (y.cpu() - golden.out['y']).abs().max().item()
        assert err <= 2e-5 * max(1.0, golden.out['y'].abs().max().item()), err


def test_fused_batchnorm_epilogue(golden):
    """Eval-mode arm_bn folded into the kernel (armnet.py:89) against the oracle's batch_norm of the reference z."""
    from armnet_b200 import ops
    from oracle import armnet_oracle as oracle
    d = dev()
    W, Q, Vv = (t.to(d) for t in hot_params(golden))
    table = golden.state['embedding.embedding.weight'].to(d)
    st = golden.state
    scale = st['arm_bn.weight'] * torch.rsqrt(st['arm_bn.running_var'] + 1e-5)
    post = (st['arm_bn.running_mean'].to(d), scale.to(d), st['arm_bn.bias'].to(d))
    zb, _ = ops.fused_forward(golden.ids.to(d), golden.values.clone().to(d), table, W, Q, Vv,
                              float(golden.cfg['alpha']), one_head=golden.one_head, post=post)
    ref = oracle.batchnorm1d(golden.out['z'], st, 'arm_bn.', training=False)
    # BN divides z - mean (z ~ 1) by sqrt(var): a 1-ulp difference in z is amplified by scale, so the bound is
    # relative to scale * |z| rather than to the (possibly tiny) normalised value
    bound = 2e-6 * (scale.abs().max() * golden.out['z'].abs().max()).item() + 1e-5 * ref.abs().max().item()
    assert (zb.cpu() - ref).abs().max().item() <= bound


def test_entmax_op_forward_backward(golden):
    from armnet_b200 import ops
    from oracle import armnet_oracle as oracle
    d = dev()
    a = float(golden.cfg['alpha'])
    g = golden.out['g'].to(d).requires_grad_(True)
    for solver in (ops.SOLVER_AUTO, ops.SOLVER_BISECT):
        p = ops.entmax(g, alpha=a, dim=-1, solver=solver)
        assert (p.detach().cpu() - golden.out['p']).abs().max().item() <= TOL_P
    gen = torch.Generator().manual_seed(1)
    dp = torch.randn(golden.out['p'].shape, generator=gen)
    p.backward(dp.to(d))
    if a == 1.:
        pr = golden.out['p']
        ref = pr * (dp - (dp * pr).sum(-1, keepdim=True))
    else:
        ref = oracle.entmax_backward(golden.out['p'], dp, a)
    assert norm_rel(g.grad.cpu(), ref) <= 2e-5


def test_train_step_matches_reference_autograd(golden):
    if 'y_train' not in golden.out:
        pytest.skip('fixture has no train-mode pass')
    d = dev()
    m = build_model(golden, d).train()
    y = m({'id': golden.ids.to(d), 'value': golden.values.clone().to(d)})
    loss = torch.nn.BCEWithLogitsLoss(reduction='mean')(y, golden.target.to(d))
    loss.backward()
    assert abs(loss.item() - golden.out['loss'].item()) <= 1e-5
    assert norm_rel(y.detach().cpu(), golden.out['y_train']) <= 5e-5
    gmax = max(golden.out['grad/' + n].abs().max().item() for n, _ in m.named_parameters())
    for n, p in m.named_parameters():
        ref = golden.out['grad/' + n]
        # floor relative to the largest gradient: biases that feed a train-mode BatchNorm have a true gradient of 0
        # and hold only fp32 noise (1e-9..1e-7) in both implementations
        err = (p.grad.cpu() - ref).abs().max().item()
        assert err <= 2e-4 * ref.abs().max().item() + 1e-5 * gmax, (n, err)


def test_fused_backward_matches_unfused_autograd(golden):
    """The fused backward kernel (+ cuBLAS finishing products) against the unfused autograd stages, same device."""
    from armnet_b200 import ops
    c = golden.cfg
    if not ops.fused_bwd_supported(c['nfield'], c['nemb']):
        pytest.skip('no backward instance for this shape')
    d = dev()
    gen = torch.Generator().manual_seed(4)
    grads = []
    for fused in (True, False):
        m = build_model(golden, d).train()
        m.fused_backward = fused
        z = m._interaction_autograd({'id': golden.ids.to(d), 'value': golden.values.clone().to(d)})
        if fused:
            dz = torch.randn(z.shape, generator=gen).to(d)
        (z * dz).sum().backward()
        W = m.attn_layer.bilinear_w.weight if golden.one_head else m.attn_layer.bilinear_w
        grads.append([m.embedding.embedding.weight.grad, W.grad, m.attn_layer.query.grad, m.attn_layer.values.grad])
    for a, b, n in zip(grads[0], grads[1], ('table', 'bilinear_w', 'query', 'values')):
        assert norm_rel(a.cpu(), b.cpu()) <= 5e-5, n


# ---------------------------------------------------------------- edge cases and full-size properties

def _criteo_state(seed=2025, nfeat=100000, nemb=10, nhid=128, nhead=4, nfield=39):
    from oracle import armnet_oracle as oracle
    return oracle.reference_init_state('armnet', nfield, nfeat, nemb, nhead, nhid, mlp_nhid=32, seed=seed)


def test_empty_batch_and_ragged_tiles():
    from armnet_b200 import ops
    from oracle import armnet_oracle as oracle
    d = dev()
    st = _criteo_state(nhid=16)
    W, Q, Vv = (st[k].to(d) for k in ('attn_layer.bilinear_w', 'attn_layer.query', 'attn_layer.values'))
    table = st['embedding.embedding.weight'].to(d)
    z, _ = ops.fused_forward(torch.zeros(0, 39, dtype=torch.int64, device=d), torch.zeros(0, 39, device=d), table,
                             W, Q, Vv, 1.7)
    assert z.shape == (0, 64, 10)
    gen = torch.Generator().manual_seed(3)
    for B in (1, 2, 7, 149, 300):            # R = 64: tiles hold 8 samples, so these end in ragged tiles
        ids = torch.randint(0, 100000, (B, 39), generator=gen)
        vals = torch.rand(B, 39, generator=gen)
        ref = oracle.hot_path(st, 1.7, ids, vals.clone())
        z, _ = ops.fused_forward(ids.to(d), vals.to(d), table, W, Q, Vv, 1.7)
        assert norm_rel(z.cpu(), ref['z']) <= TOL_NORM, B


def test_out_of_range_id_raises_index_error():
    import armnet_b200 as ab
    d = dev()
    torch.manual_seed(0)
    m = ab.ARMNetModel(5, 50, 10, 2, 1.7, 8, 1, 8, 0.0, False, 1, 8).to(d).eval()
    m.validate_ids = True
    ids = torch.randint(0, 50, (4, 5), device=d)
    with torch.no_grad():
        m({'id': ids, 'value': torch.ones(4, 5, device=d)})
        ids[2, 3] = 50
        with pytest.raises(IndexError):
            m({'id': ids, 'value': torch.ones(4, 5, device=d)})
        ids[2, 3] = -1
        with pytest.raises(IndexError):
            m({'id': ids, 'value': torch.ones(4, 5, device=d)})


@pytest.mark.parametrize('K,O,alpha', [(8, 128, 1.7), (4, 300, 2.0), (16, 128, 1.5)])
def test_large_neuron_counts_run_in_row_chunks(K, O, alpha):
    """--h x --nattn_head beyond what one CTA's shared memory holds (the reference accepts any size, train.py:24,28):
    the forward runs in passes over chunks of K*O rows; same function, in-place clamp applied once, eval-mode arm_bn
    epilogue per row."""
    from armnet_b200 import ops
    from oracle import armnet_oracle as oracle
    d = dev()
    torch.manual_seed(K * O)
    B, F, E, V = 67, 39, 10, 4000
    R = K * O
    ids = torch.randint(0, V, (B, F))
    values = torch.rand(B, F) * 1.2 - 0.1
    table = torch.randn(V, E) * 0.6
    W, Q, Vv = torch.randn(K, E, E) * 0.5, torch.randn(K, O, E) * 0.5, torch.randn(K, O, F)
    st = {'embedding.embedding.weight': table, 'attn_layer.bilinear_w': W, 'attn_layer.query': Q,
          'attn_layer.values': Vv}
    ref = oracle.hot_path(st, alpha, ids, values.clone())
    vals = values.clone().to(d)
    post = (torch.randn(R, device=d) * 0.1 + 1, torch.rand(R, device=d) + 0.5, torch.randn(R, device=d))
    z, _ = ops.fused_forward(ids.to(d), vals, table.to(d), W.to(d), Q.to(d), Vv.to(d), alpha)
    assert ops.last_launch_count() >= 3                        # prepare + at least two row chunks
    zb, _ = ops.fused_forward(ids.to(d), values.clone().to(d), table.to(d), W.to(d), Q.to(d), Vv.to(d), alpha, post=post)
    torch.cuda.synchronize()
    assert torch.equal(vals.cpu(), values.clamp(0.001, 1.0))
    assert norm_rel(z.cpu(), ref['z']) <= TOL_NORM
    zb_ref = (z - post[0][None, :, None]) * post[1][None, :, None] + post[2][None, :, None]
    assert torch.allclose(zb, zb_ref, rtol=1e-6, atol=1e-6)
    with pytest.raises(RuntimeError):                          # validation outputs need all rows in one pass
        ops.fused_forward(ids.to(d), values.clone().to(d), table.to(d), W.to(d), Q.to(d), Vv.to(d), alpha, want_p=True)


def test_out_of_range_id_lazy_check_and_scorer():
    """Default validate_ids = 'lazy': no per-batch synchronisation, the kernels record the bad id in a device flag;
    model.check_ids() and BatchScorer.result() raise the reference's IndexError (layers.py:20).  Rows with a bad id are
    computed as zero embedding rows, the rest of the batch is unaffected."""
    import armnet_b200 as ab
    d = dev()
    torch.manual_seed(0)
    m = ab.ARMNetModel(39, 500, 10, 2, 1.7, 64, 1, 8, 0.0, False, 1, 8).to(d).eval()
    assert m.validate_ids == 'lazy'
    ids = torch.randint(0, 500, (64, 39), device=d)
    vals = torch.ones(64, 39, device=d)
    with torch.no_grad():
        y0 = m({'id': ids, 'value': vals.clone()})
        m.check_ids()                                  # clean
        bad = ids.clone()
        bad[5, 7] = 500
        y1 = m({'id': bad, 'value': vals.clone()})     # does not raise here ...
        with pytest.raises(IndexError):
            m.check_ids()                              # ... but here
        m.check_ids()                                  # flag was reset
        keep = torch.arange(64, device=d) != 5
        assert torch.equal(y0[keep], y1[keep])
    scorer = ab.BatchScorer(m, 64, 39, depth=2, compute_streams=1)
    t = scorer.submit(ids.cpu().pin_memory(), vals.cpu().pin_memory())
    assert torch.allclose(scorer.result(t).to(d), y0, rtol=1e-6, atol=1e-6)
    t = scorer.submit(bad.cpu().pin_memory(), vals.cpu().pin_memory())
    with pytest.raises(IndexError):
        scorer.result(t)
    t = scorer.submit(ids.cpu().pin_memory(), vals.cpu().pin_memory())
    assert torch.allclose(scorer.result(t).to(d), y0, rtol=1e-6, atol=1e-6)


def test_bad_arguments_fail_loudly():
    from armnet_b200 import ops
    d = dev()
    t = torch.zeros(10, 10, device=d)
    with pytest.raises(RuntimeError):        # CPU tensor: no fallback
        ops.embed_gather(torch.zeros(2, 3, dtype=torch.int64), torch.ones(2, 3), torch.zeros(10, 10))
    # a strided view of values is accepted like the reference does, and still clamped in place (armnet.py:82)
    v_nc = (torch.ones(3, 2, device=d) * 2).t()
    e_nc = ops.embed_gather(torch.zeros(2, 3, dtype=torch.int64, device=d), v_nc, t, clamp=(0.001, 1.0))
    assert e_nc.shape == (2, 3, 10) and torch.equal(v_nc, torch.ones(2, 3, device=d))
    with pytest.raises(TypeError):           # integer values are not
        ops.embed_gather(torch.zeros(2, 3, dtype=torch.int64, device=d), torch.ones(2, 3, device=d).long(), t)
    with pytest.raises(RuntimeError):        # alpha < 1 -> ARMNET_ERR_SHAPE
        ops.entmax_forward(torch.zeros(4, 8, device=d), 0.5)
    with pytest.raises(RuntimeError):        # 65 fields -> ARMNET_ERR_UNSUPPORTED
        ops.entmax_forward(torch.zeros(4, 65, device=d), 1.5)


def _full_size_case(F, E, K, O, V, B, alpha, regime, tuning, expect_kind, ld_pad=None, n_pick=48):
    """One BASELINE configuration at its full size: the oracle on a sample of rows spread over the batch, plus
    size-independent properties on the whole batch -- finite / positive output, in-place clamp, permutation equivariance
    over samples, determinism, and independence of the table's row pitch (the TMA-gather path)."""
    from armnet_b200 import ops
    from oracle import armnet_oracle as oracle
    d = dev()
    st = oracle.reference_init_state('armnet', F, V, E, K, O, mlp_nhid=32, seed=2025)
    st = {k: v for k, v in st.items() if k.startswith('embedding.') or k.startswith('attn_layer.')}
    gen = torch.Generator().manual_seed(11)
    if regime == 'trained':
        st['embedding.embedding.weight'].normal_(0, 1.0, generator=gen)
        st['attn_layer.bilinear_w'] *= 4
        st['attn_layer.query'] *= 4
    ids = torch.randint(0, V, (B, F), generator=gen)
    ids[0, 0], ids[-1, -1] = V - 1, V - 1                       # the last table row (byte offsets beyond 2^31 at C3)
    vals = torch.rand(B, F, generator=gen) * 1.2
    W, Q, Vv = (st[k].to(d) for k in ('attn_layer.bilinear_w', 'attn_layer.query', 'attn_layer.values'))
    table = st['embedding.embedding.weight'].to(d)
    ld = E if ld_pad is None else ld_pad
    if ld != E:
        tab = torch.zeros(V, ld, device=d)
        tab[:, :E] = table
    else:
        tab = table
    with ops.tuning(**tuning):
        assert ops.fused_fwd_kernel_kind(F, E, K, O, alpha) == expect_kind
        v_d = vals.clone().to(d)
        z, _ = ops.fused_forward(ids.to(d), v_d, tab, W, Q, Vv, alpha, ld=ld, nemb=E)
        assert z.shape == (B, K * O, E)
        assert torch.isfinite(z).all() and (z > 0).all()
        assert torch.equal(v_d.cpu(), vals.clamp(0.001, 1.0))
        # oracle on samples spread over the batch (first and last included)
        pick = torch.cat([torch.arange(0, B, max(B // n_pick, 1))[:n_pick], torch.tensor([B - 1])])
        ref = oracle.hot_path(st, alpha, ids[pick], vals[pick].clone())
        assert norm_rel(z[pick.to(d)].cpu(), ref['z']) <= TOL_NORM
        # permutation equivariance over samples (each row only depends on its own sample; tile / warp neighbours change)
        perm = torch.randperm(B, generator=gen)
        z2, _ = ops.fused_forward(ids[perm].to(d), vals[perm].clone().to(d), tab, W, Q, Vv, alpha, ld=ld, nemb=E)
        assert norm_rel(z2.cpu(), z[perm.to(d)].cpu()) <= 2e-6
        # determinism
        z3, _ = ops.fused_forward(ids.to(d), vals.clone().to(d), tab, W, Q, Vv, alpha, ld=ld, nemb=E)
        assert torch.equal(z3, z)
    return z, ids, vals, table, (W, Q, Vv)


@pytest.mark.parametrize('regime', ['init', 'trained'])
@pytest.mark.parametrize('kernel', ['fp32', 'tmem'])
def test_full_size_criteo_batch_properties(regime, kernel):
    """BASELINE config 2 at full size (B=4096, 39 fields, 1M vocab, nemb 10, 4 heads x 128 neurons), through
    armnet_fwd_kernel and through armnet_fwd_tmem_kernel (the default on the 16-byte-pitch table the module keeps)."""
    from armnet_b200 import ops
    d = dev()
    tuning = dict(tmem=0, mma=0) if kernel == 'fp32' else dict(tmem=1)
    z, ids, vals, table, (W, Q, Vv) = _full_size_case(39, 10, 4, 128, 1000000, 4096, 1.7, regime, tuning,
                                                      1 if kernel == 'fp32' else 3, ld_pad=12)
    # the 40-byte-row table (no TMA gather; always armnet_fwd_kernel) gives the same function
    with ops.tuning(**tuning):
        z4, _ = ops.fused_forward(ids.to(d), vals.clone().to(d), table, W, Q, Vv, 1.7)
    if kernel == 'fp32':
        assert torch.equal(z4, z)                       # same kernel, same bits whatever the gather path
    else:
        assert norm_rel(z4.cpu(), z.cpu()) <= 2e-6      # two kernels, one function


def test_full_size_avazu_shape_c3():
    """BASELINE config 3 at full size: 22 fields, 1.5M vocab, nemb 100 (rows split over 4 lanes), h 32, alpha 1.5,
    B=8192; the 600 MB table puts row offsets beyond 2^31 bytes."""
    _full_size_case(22, 100, 1, 32, 1500000, 8192, 1.5, 'trained', dict(tmem=-1, mma=-1), 1, n_pick=32)


@pytest.mark.parametrize('regime', ['init', 'trained'])
def test_full_size_config4_shape_nemb16(regime):
    """BASELINE config 4's forward at full size: Criteo shape with nemb 16, K*O = 512, B=4096 -- the default-on
    armnet_fwd_mma_kernel instance (the two E x F products as 3xTF32 warp MMAs)."""
    _full_size_case(39, 16, 4, 128, 1000000, 4096, 1.7, regime, dict(tmem=-1, mma=-1), 2)


@pytest.mark.parametrize('alpha', [1.0, 1.3, 1.5, 1.7, 2.0])
@pytest.mark.parametrize('F,scale', [(39, 0.05), (39, 3.0), (40, 1.0), (33, 1.0), (36, 8.0)])
def test_tensor_core_kernel_matches_fp32_kernel_and_oracle(alpha, F, scale):
    """armnet_fwd_mma_kernel (the two E x F products as 3xTF32 warp MMAs; ARMNET_MMA=1) against
    armnet_fwd_kernel (same products on the FP32 pipe; ARMNET_MMA=0) and against the oracle, for every Newton-type
    solver mode, padded field blocks (F % 8 != 0), dense (scale 0.05) to very sparse (scale 8) gates, clamped values
    and both gather paths.  nemb 10: one MMA step + 2 leftover lanes on the FP32 pipe."""
    _check_tensor_core_case(alpha, F, scale, 10)


@pytest.mark.parametrize('alpha', [1.0, 1.5, 1.7, 2.0])
@pytest.mark.parametrize('F,scale', [(39, 0.05), (39, 3.0), (34, 1.0)])
def test_tensor_core_kernel_nemb16(alpha, F, scale):
    """Same for the nemb-16 instance (config 4 shape): two MMA steps, no leftover lanes."""
    _check_tensor_core_case(alpha, F, scale, 16)


def _oracle_with_own_error(st, alpha, ids, values):
    """The oracle in fp32 (the parity target) and in fp64, plus the reference's OWN fp32 error per stage
    (max |fp32 - fp64|): where that exceeds north_star's tolerance no fp32 implementation can hold the tolerance against
    the fp32 oracle, so a stage is accepted at  err(cuda vs fp64) <= max(tolerance, 2 x own error)."""
    from oracle import armnet_oracle as oracle
    ref = oracle.hot_path(st, alpha, ids, values.clone())
    st64 = {k: v.double() for k, v in st.items()}
    ref64 = oracle.hot_path(st64, alpha, ids, values.double().clone())
    own = {k: (ref[k].double() - ref64[k]).abs().max().item() for k in ('g', 'p', 's', 'z')
           if torch.isfinite(ref64[k]).all()}
    return ref, ref64, own


def _stage_bounds(ref64, own, alpha):
    """Absolute error bounds per stage: north_star's tolerance (norm-relative 1e-5 on g / s / z, 2e-6 on gates) or twice
    the reference's own fp32-vs-fp64 error, whichever is larger.  Gates additionally get the floor of ANY fp32
    evaluation: a gate is a function of differences X_f - tau of the scaled logits X = (alpha-1) g (entmax.py:42-61),
    so it cannot be more accurate than a few units in the last place of the largest |X| (4 ulp = 4.8e-7 max|X|; at
    alpha = 2 the gate IS such a difference)."""
    b = {}
    for k, tol in (('g', TOL_NORM), ('s', TOL_NORM), ('z', TOL_NORM)):
        if k in own:
            b[k] = max(tol * ref64[k].abs().max().item(), 2.0 * own[k])
    x_max = abs(alpha - 1.0) * ref64['g'].abs().max().item() if alpha != 1.0 else ref64['g'].abs().max().item()
    b['p'] = max(TOL_P, 2.0 * own['p'], 4 * 1.1920929e-07 * x_max)
    return b


def _check_tensor_core_case(alpha, F, scale, E):
    from armnet_b200 import ops
    d = dev()
    torch.manual_seed(1234 + F)
    B, V, K, O, D = 37, 5000, 2, 96, 10
    with ops.tuning(tmem=0, mma=0):
        assert ops.fused_fwd_kernel_kind(F, E, K, O, alpha) == 1           # forced off
    with ops.tuning(tmem=0, mma=1):
        assert ops.fused_fwd_kernel_kind(F, E, K, O, alpha) == 2
        assert ops.fused_fwd_kernel_kind(F, E, K, O, 2.5) == 1            # alpha > 2: the literal bisection
        assert ops.fused_fwd_kernel_kind(F, E, K, O + 1, alpha) == 1      # K*O % 64 != 0
        if alpha > 1.0:   # alpha == 1 is softmax whatever the solver
            assert ops.fused_fwd_kernel_kind(F, E, K, O, alpha, ops.SOLVER_BISECT) == 1
    ids = torch.randint(0, V, (B, F))
    values = torch.rand(B, F) * 1.2 - 0.1                              # both clamp bounds are hit
    table = torch.randn(V, E) * scale
    W = torch.randn(K, E, D) * 0.5
    Q = torch.randn(K, O, D) * 0.5
    Vv = torch.randn(K, O, F)
    st = {'embedding.embedding.weight': table, 'attn_layer.bilinear_w': W, 'attn_layer.query': Q,
          'attn_layer.values': Vv}
    ref, ref64, own = _oracle_with_own_error(st, alpha, ids, values)
    bound = _stage_bounds(ref64, own, alpha)
    outs = {}
    for kind in ('mma', 'fp32'):
        with ops.tuning(tmem=0, mma=0 if kind == 'fp32' else 1):
            for padded in (False, True):
                tab = table.to(d)
                ld = E
                if padded:
                    ld = E + 4 - E % 4 if E % 4 else E + 4          # 12 for nemb 10, 20 for nemb 16: 16-byte aligned rows
                    tab = torch.zeros(V, ld, device=d)
                    tab[:, :E] = table.to(d)
                vals = values.clone().to(d)
                z, ex = ops.fused_forward(ids.to(d), vals, tab, W.to(d), Q.to(d), Vv.to(d), alpha, ld=ld, nemb=E,
                                          want_tau=True, want_p=True, want_g=True, want_s=True)
                z_plain, _ = ops.fused_forward(ids.to(d), values.clone().to(d), tab, W.to(d), Q.to(d), Vv.to(d), alpha,
                                               ld=ld, nemb=E)
                torch.cuda.synchronize()
                assert torch.equal(z, z_plain)
                assert torch.equal(vals.cpu(), values.clamp(0.001, 1.0))
                outs[kind, padded] = (z.cpu(), {k: v.cpu() for k, v in ex.items()})
    assert torch.equal(outs['mma', False][0], outs['mma', True][0])    # gather path does not change the result
    R = K * O
    for key, (z, ex) in outs.items():
        # every stage against the fp64 evaluation of the reference, within max(north_star tolerance, 2 x the reference's
        # own fp32 error at this input) -- no hand-tuned widening
        for k, shape in (('g', (B, R, F)), ('p', (B, R, F)), ('s', (B, R, E))):
            err = (ex[k].double() - ref64[k].reshape(shape)).abs().max().item()
            assert err <= bound[k], (key, k, err, bound[k], own[k])
        if 'z' in bound:
            # z = exp(s): an error ds in s is an error z ds in z, so where |s| is large the bound on z follows s's
            z64 = ref64['z'].reshape(B, R, E)
            ok = (z.double() - z64).abs() <= torch.clamp(z64.abs() * (1.5 * bound['s']), min=bound['z'])
            assert ok.all(), (key, 'z', (z.double() - z64).abs().max().item(), bound['z'], bound['s'], own['z'])
        assert (ex['p'].sum(-1) - 1).abs().max().item() < 1e-5, key
    zm, zf = outs['mma', True][0], outs['fp32', True][0]
    assert torch.equal(torch.isfinite(zm), torch.isfinite(zf))


TMEM_CASES = [
    # alpha, F, E, K, O, table scale, B
    (1.7, 39, 10, 4, 128, 0.05, 37),     # C2a shape, dense gates (closed-form start), odd batch tail
    (1.7, 39, 10, 4, 128, 1.0, 64),      # C2a shape, sparse gates (Hoelder pre-solve + q-norm Newton)
    (1.7, 39, 10, 4, 128, 3.0, 33),
    (1.7, 40, 10, 2, 128, 1.0, 40),      # even field count, one item per tile
    (1.9, 39, 10, 6, 128, 2.0, 9),       # three items per tile (K*O = 768: the tensor-memory limit)
    (1.3, 39, 10, 4, 64, 1.0, 21),       # q > 2: no Hoelder bound, bound-started q-norm Newton
    (1.7, 39, 10, 3, 128, 1.0, 19),      # K*O = 384: not a multiple of 256 -> one row per thread only
    (1.5, 39, 10, 4, 64, 1.0, 21),       # alpha = 1.5: square mode, no MUFU
    (2.0, 39, 10, 4, 64, 2.0, 21),       # C2b: sparsemax (Michelot)
    (1.0, 40, 10, 2, 128, 1.0, 21),      # softmax
    (1.7, 39, 8, 4, 64, 1.0, 21),        # nemb 8: two 16-byte chunks per row
    (1.7, 40, 5, 2, 128, 1.0, 21),       # nemb 5: padded to 8 floats per row
    (1.7, 39, 10, 4, 64, 8.0, 21),       # very sparse, large logits
    (1.7, 39, 16, 4, 128, 0.05, 37),     # config-4 shape (nemb 16): wide packed layout [e | l] / [M | M | m], six MMA steps
    (1.7, 39, 16, 4, 128, 1.0, 33),
    (1.7, 40, 13, 2, 128, 1.0, 21),      # nemb 13 on the wide instance (padded table rows)
    (2.0, 39, 16, 2, 128, 2.0, 9),
]


@pytest.mark.parametrize('alpha,F,E,K,O,scale,B', TMEM_CASES)
def test_tensor_memory_kernel_matches_oracle(alpha, F, E, K, O, scale, B):
    """armnet_fwd_tmem_kernel (logits by tcgen05.mma into tensor memory, entmax with the MUFU-free pre-solve, kind 3)
    against the oracle in fp64 within max(north_star tolerance, 2 x the reference's own fp32 error) and against
    armnet_fwd_kernel (kind 1) on the same inputs; values hit both clamp bounds and are clamped in place."""
    from armnet_b200 import ops
    d = dev()
    torch.manual_seed(99 + F + E)
    V, D, R = 3000, E, K * O
    with ops.tuning(tmem=1):
        assert ops.fused_fwd_kernel_kind(F, E, K, O, alpha) == 3
    ids = torch.randint(0, V, (B, F))
    values = torch.rand(B, F) * 1.2 - 0.1
    table = torch.randn(V, E) * scale
    W = torch.randn(K, E, D) * 0.5
    Q = torch.randn(K, O, D) * 0.5
    Vv = torch.randn(K, O, F)
    st = {'embedding.embedding.weight': table, 'attn_layer.bilinear_w': W, 'attn_layer.query': Q,
          'attn_layer.values': Vv}
    ref, ref64, own = _oracle_with_own_error(st, alpha, ids, values)
    ld = (E + 3) // 4 * 4 if E % 4 else E
    tab = torch.zeros(V, ld, device=d)
    tab[:, :E] = table.to(d)
    out = {}
    # both thread mappings of the kernel: one row per thread (logits in registers) and two rows per thread (logits
    # streamed from tensor memory; needs K*O % 256 == 0, the default there), and armnet_fwd_kernel
    kinds = [('tmem1', dict(tmem=1, tmem_rows=1, mma=0)), ('fp32', dict(tmem=0, mma=0))]
    if R % 256 == 0 and E <= 10:   # the two-rows mapping exists for the narrow packed layout only
        kinds.insert(0, ('tmem2', dict(tmem=1, tmem_rows=2, mma=0)))
    for kind, tune in kinds:
        with ops.tuning(**tune):
            vals = values.clone().to(d)
            z, _ = ops.fused_forward(ids.to(d).int() if kind == 'tmem1' else ids.to(d), vals, tab, W.to(d), Q.to(d),
                                     Vv.to(d), alpha, ld=ld, nemb=E)
            torch.cuda.synchronize()
            assert torch.equal(vals.cpu(), values.clamp(0.001, 1.0))
            out[kind] = z.cpu()
    z64 = ref64['z'].reshape(B, R, E)
    s64 = ref64['s'].reshape(B, R, E)
    fin = torch.isfinite(z64) & (z64.abs() < 1e30)
    zf = out['fp32']
    for kind in [k for k, _ in kinds if k != 'fp32']:
        zt = out[kind]
        assert zt.shape == (B, R, E)
        assert torch.equal(torch.isfinite(zt) & (zt.abs() < 1e30), fin), kind
        if 'z' in own and fin.all():
            bound = max(TOL_NORM * z64.abs().max().item(), 2.0 * own['z'])
            err = (zt.double() - z64).abs().max().item()
            assert err <= bound, (kind, err, bound, own['z'])
        # log domain (s = log z): the tolerance north_star states for the interaction logits, element by element where
        # the reference itself is that accurate
        s_err = (torch.log(zt.double().clamp_min(1e-300)) - s64).abs()[fin].max().item()
        assert s_err <= max(TOL_NORM * s64.abs().max().item(), 2.0 * own['s']) + 3e-7 * s64.abs().max().item(), (kind, s_err)
        # the kernels agree far inside the tolerance
        assert ((zt - zf)[fin].abs() <= 2e-5 * zf[fin].abs().clamp_min(1e-30) *
                max(1.0, ref['s'].abs().max().item())).all(), kind


def test_tensor_memory_kernel_long_batch_ring_wraparound():
    """B = 10 001 at the C2b shape: 34 tiles per CTA, so the raw-row ring (32 samples), the B-tile ring (4 tiles) and
    both D slots wrap many times; rows sampled from the start / middle / end of the batch against the oracle, the whole
    output against armnet_fwd_kernel, determinism, and the eval-mode arm_bn epilogue."""
    from armnet_b200 import ops
    from oracle import armnet_oracle as oracle
    d = dev()
    torch.manual_seed(5)
    B, F, E, K, O, V, alpha = 10001, 39, 10, 4, 64, 200000, 1.7
    R = K * O
    ids = torch.randint(0, V, (B, F))
    values = torch.rand(B, F)
    table = torch.randn(V, E) * 0.7
    W, Q, Vv = torch.randn(K, E, E) * 0.5, torch.randn(K, O, E) * 0.5, torch.randn(K, O, F)
    tab = torch.zeros(V, 12, device=d)
    tab[:, :E] = table.to(d)
    post = (torch.randn(R, device=d) * 0.1 + 1, torch.rand(R, device=d) + 0.5, torch.randn(R, device=d))
    with ops.tuning(tmem=1, tmem_rows=1):
        z1r, _ = ops.fused_forward(ids.to(d), values.clone().to(d), tab, W.to(d), Q.to(d), Vv.to(d), alpha, ld=12, nemb=E)
    with ops.tuning(tmem=1):
        z1, _ = ops.fused_forward(ids.to(d), values.clone().to(d), tab, W.to(d), Q.to(d), Vv.to(d), alpha, ld=12, nemb=E)
        z2, _ = ops.fused_forward(ids.to(d), values.clone().to(d), tab, W.to(d), Q.to(d), Vv.to(d), alpha, ld=12, nemb=E)
        zb, _ = ops.fused_forward(ids.to(d), values.clone().to(d), tab, W.to(d), Q.to(d), Vv.to(d), alpha, ld=12, nemb=E,
                                  post=post)
    with ops.tuning(tmem=0, mma=0):
        zf, _ = ops.fused_forward(ids.to(d), values.clone().to(d), tab, W.to(d), Q.to(d), Vv.to(d), alpha, ld=12, nemb=E)
    torch.cuda.synchronize()
    assert torch.equal(z1, z2)                                                     # deterministic
    assert norm_rel(z1r.cpu(), z1.cpu()) <= 2e-6                                   # both thread mappings, one function
    assert ((z1 - zf).abs() <= 2e-5 * zf.abs() * max(1.0, zf.log().abs().max().item())).all()
    zb_ref = (z1 - post[0][None, :, None]) * post[1][None, :, None] + post[2][None, :, None]
    assert torch.allclose(zb, zb_ref, rtol=1e-6, atol=1e-6)
    rows = torch.cat([torch.arange(0, 48), torch.arange(5000, 5048), torch.arange(B - 33, B)])
    st = {'embedding.embedding.weight': table, 'attn_layer.bilinear_w': W, 'attn_layer.query': Q,
          'attn_layer.values': Vv}
    ref = oracle.hot_path(st, alpha, ids[rows], values[rows].clone())
    assert norm_rel(z1[rows.to(d)].cpu(), ref['z']) <= TOL_NORM

