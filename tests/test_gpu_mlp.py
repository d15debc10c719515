"""The trailing MLP in eval mode (layers.py:68-88): tcgen05 3xTF32 GEMM for the first Linear + one tail kernel,
through the C ABI, against fp64 / stock torch fp32 evaluations of the same layers.
Tolerance: max|a-b| / max|b| <= 1e-5 for the GEMM (cuBLAS fp32 SGEMM itself sits at ~4e-6 from fp64 at K=5120)."""
import pytest
import torch
import torch.nn as nn

pytestmark = pytest.mark.gpu

TOL_GEMM = 1e-5


def dev():
    return torch.device('cuda:0')


def nrel(a, b):
    return ((a.double() - b.double()).abs().max() / b.double().abs().max()).item()


@pytest.mark.parametrize('B,K,N,splits', [(4096, 5120, 256, None), (4096, 5120, 256, 4), (512, 1280, 256, 1), (300, 132, 48, 3),
                                           (130, 5120, 500, 2), (1, 64, 8, 1), (257, 36, 250, None)])
def test_first_linear_3xtf32_matches_fp64(B, K, N, splits):
    from armnet_b200 import ops
    g = torch.Generator().manual_seed(B * 7 + K + N)
    x = (torch.rand(B, K, generator=g) * 2 + 0.1).to(dev())          # like arm_bn(exp(.)): positive, O(1)
    w = (torch.randn(N, K, generator=g) * K ** -0.5).to(dev())
    hi, lo = ops.mlp_split_weight(w)
    assert torch.equal((hi.double() + lo.double()).float(), w)        # exact split
    assert torch.all((hi.view(torch.int32) & 0x1FFF) == 0)           # hi is a TF32 value
    part = ops.mlp_first_linear(x, hi, lo, splits=splits)
    y = ops.partials_to_dense(part, B)
    ref = x.double() @ w.double().t()
    err = nrel(y, ref)
    old = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    err_sgemm = nrel(x @ w.t(), ref)
    torch.backends.cuda.matmul.allow_tf32 = old
    print(f'B={B} K={K} N={N} splits={part.shape[0]}: 3xTF32 err {err:.2e}, cuBLAS fp32 err {err_sgemm:.2e}')
    assert err <= TOL_GEMM


@pytest.mark.parametrize('ninput,nlayers,nhid,noutput,B', [(5120, 2, 256, 1, 4096), (640, 3, 200, 1, 513),
                                                            (132, 1, 48, 2, 77), (2560, 2, 500, 1, 300)])
def test_mlp_eval_fast_path_matches_stock_modules(ninput, nlayers, nhid, noutput, B):
    from armnet_b200.layers import MLP
    from armnet_b200 import ops
    torch.manual_seed(ninput + nhid)
    m = MLP(ninput, nlayers, nhid, 0.1, noutput=noutput)
    with torch.no_grad():
        for mod in m.mlp:
            if isinstance(mod, nn.BatchNorm1d):                        # non-trivial running statistics / affine
                mod.running_mean.normal_(0, 0.3)
                mod.running_var.uniform_(0.5, 2.0)
                mod.weight.uniform_(0.5, 1.5)
                mod.bias.normal_(0, 0.2)
    m = m.to(dev()).eval()
    x = torch.rand(B, ninput, device=dev()) * 2
    old = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    with torch.no_grad():
        y_fast = m(x)
        n_launch = ops.last_launch_count()
        m.tensor_core = False
        y_ref = m(x)
        y64 = m.double()(x.double())
    torch.backends.cuda.matmul.allow_tf32 = old
    assert n_launch == 1 and y_fast.shape == y_ref.shape == (B, noutput)
    e_fast, e_ref = nrel(y_fast, y64), nrel(y_ref, y64)
    print(f'fast path err vs fp64 {e_fast:.2e}; stock fp32 modules {e_ref:.2e}')
    assert e_fast <= 1e-5
    # the cache follows parameter updates
    m = m.float()
    m.tensor_core = True
    with torch.no_grad():
        m.mlp[0].weight.mul_(0.5)
        y2 = m(x)
        m.tensor_core = False
        y2_ref = m(x)
    assert nrel(y2, y2_ref) <= 1e-5


@pytest.mark.parametrize('ninput,nlayers,nhid,noutput,B', [(5120, 2, 256, 1, 4096), (640, 3, 200, 1, 2049),
                                                            (1280, 2, 252, 3, 1100), (5120, 3, 256, 1, 4100),
                                                            (64, 2, 8, 4, 513)])
def test_mlp_hidden_layers_on_tensor_cores(ninput, nlayers, nhid, noutput, B):
    """Hidden layers 2..n + output Linear on tcgen05 (armnet_mlp_hidden_tc_f32, one launch per hidden layer) against the
    fp64 evaluation of the stock modules and against the CUDA-core tail kernel; ragged batches, widths below 256."""
    from armnet_b200.layers import MLP
    torch.manual_seed(ninput + nhid + B)
    m = MLP(ninput, nlayers, nhid, 0.1, noutput=noutput)
    with torch.no_grad():
        for mod in m.mlp:
            if isinstance(mod, nn.BatchNorm1d):
                mod.running_mean.normal_(0, 0.3)
                mod.running_var.uniform_(0.5, 2.0)
                mod.weight.uniform_(0.5, 1.5)
                mod.bias.normal_(0, 0.2)
    m = m.to(dev()).eval()
    x = torch.rand(B, ninput, device=dev()) * 2
    with torch.no_grad():
        assert m.hidden_tensor_core and B >= m.hidden_tensor_core_min_batch
        y_tc = m(x)
        m.hidden_tensor_core = False
        y_tail = m(x)
        y64 = m.double()(x.double())
    e_tc, e_tail = nrel(y_tc, y64), nrel(y_tail, y64)
    print(f'hidden layers on tcgen05: err vs fp64 {e_tc:.2e}; CUDA-core tail {e_tail:.2e}')
    assert y_tc.shape == (B, noutput) and e_tc <= 1e-5 and e_tail <= 1e-5


def test_mlp_train_mode_uses_stock_path():
    from armnet_b200.layers import MLP
    m = MLP(64, 2, 32, 0.0).to(dev()).train()
    x = torch.rand(16, 64, device=dev())
    y = m(x)
    y.sum().backward()
    assert m.mlp[0].weight.grad is not None


@pytest.mark.parametrize('use_graph,depth,streams', [(True, 2, 1), (False, 2, 1), (True, 4, 2), (True, 6, 3)])
def test_batch_scorer_matches_module_forward(use_graph, depth, streams):
    """serving.BatchScorer (pipelined copies, CUDA-graph replay) returns exactly what model(x) returns."""
    import armnet_b200 as ab
    torch.manual_seed(3)
    F, V, B = 39, 5000, 256
    model = ab.ARMNetModel(F, V, 10, 4, 1.7, 32, 2, 64, 0.0, False, 2, 32).to(dev()).eval()
    g = torch.Generator().manual_seed(11)
    batches = [(torch.randint(0, V, (B, F), generator=g).pin_memory(),
                (torch.rand(B, F, generator=g) * 1.3).pin_memory()) for _ in range(13)]
    want = []
    with torch.no_grad():
        for ids, vals in batches:
            want.append(model({'id': ids.to(dev()), 'value': vals.clone().to(dev())}).cpu())
    scorer = ab.BatchScorer(model, B, F, depth=depth, use_graph=use_graph, compute_streams=streams)
    got = list(scorer.score(batches))
    assert len(got) == len(want)
    for a, b in zip(got, want):
        assert torch.equal(a, b)
    assert torch.equal(batches[0][1], batches[0][1].clone())     # host values are not clamped in place
    with pytest.raises(ValueError):
        scorer.submit(batches[0][0][:10], batches[0][1][:10])
    # parameters updated in place after capture: the scorer re-captures and follows them
    with torch.no_grad():
        model.attn_layer.values.mul_(1.3)
        model.mlp.mlp[0].weight.mul_(0.9)
        want2 = model({'id': batches[1][0].to(dev()), 'value': batches[1][1].clone().to(dev())}).cpu()
    got2 = scorer.result(scorer.submit(*batches[1])).clone()
    assert torch.equal(got2, want2) and not torch.equal(got2, want[1])


def test_full_size_eval_forward_matches_oracle():
    """BASELINE config 2 at full size (39 fields, 1M vocabulary, nemb 10, 4 x 128 neurons, bsz 4096, mlp 2 x 256): the whole
    eval forward -- fused kernel + arm_bn epilogue, tcgen05 GEMM, tail kernel -- through ARMNetModel.forward and through
    BatchScorer, against the CPU oracle of the reference's forward on a sample of rows (samples are independent in eval
    mode, armnet.py:82-101)."""
    import armnet_b200 as ab
    from oracle import armnet_oracle as oracle
    F, V, E, K, O, B = 39, 1000000, 10, 4, 128, 4096
    st = oracle.reference_init_state('armnet', F, V, E, K, O, mlp_nlayer=2, mlp_nhid=256, seed=2025)
    g = torch.Generator().manual_seed(5)
    st['embedding.embedding.weight'].normal_(0, 0.4, generator=g)        # trained-like: non-uniform gates
    st['arm_bn.running_mean'].normal_(1.0, 0.2, generator=g)
    st['arm_bn.running_var'].uniform_(0.5, 1.5, generator=g)
    st['mlp.mlp.1.running_mean'].normal_(0, 0.2, generator=g)
    st['mlp.mlp.1.running_var'].uniform_(0.5, 1.5, generator=g)
    ids = torch.randint(0, V, (B, F), generator=g)
    vals = torch.rand(B, F, generator=g) * 1.2
    rows = torch.cat([torch.arange(0, 64), torch.arange(2000, 2064), torch.arange(B - 64, B)])
    ref = oracle.forward(st, 1.7, ids[rows], vals[rows].clone())['y']
    model = ab.ARMNetModel(F, V, E, K, 1.7, O, 2, 256, 0.0, False, 2, 256)
    model.load_state_dict(st)
    model = model.to(dev()).eval()
    with torch.no_grad():
        y = model({'id': ids.to(dev()), 'value': vals.clone().to(dev())}).cpu()
    scale = max(ref.abs().max().item(), 1.0)
    err = (y[rows] - ref).abs().max().item() / scale
    scorer = ab.BatchScorer(model, B, F, depth=4, compute_streams=2)
    t = scorer.submit(ids.pin_memory(), vals.pin_memory())
    y2 = scorer.result(t).clone()
    print(f'full-size eval forward: max |y - oracle| / scale = {err:.2e}')
    assert err <= 2e-5
    assert torch.equal(y2, y)                                             # same kernels, same results, graph or not


def test_prepared_attention_workspace_equals_per_call_preparation():
    """armnet_fused_prepare_f32 + armnet_fused_fwd_prepared_f32 == armnet_fused_fwd_f32, bit for bit, and the module's
    cache follows parameter updates."""
    import armnet_b200 as ab
    from armnet_b200 import ops
    torch.manual_seed(1)
    F, V, B = 39, 3000, 200
    m = ab.ARMNetModel(F, V, 10, 4, 1.7, 32, 2, 32, 0.0, False, 2, 32).to(dev()).eval()
    ids = torch.randint(0, V, (B, F), device=dev())
    vals = torch.rand(B, F, device=dev())
    W, Q, Vv = (t.detach() for t in m._attn_weights())
    tab = m.embedding.embedding.weight.detach()
    z0, _ = ops.fused_forward(ids, vals.clone(), tab, W, Q, Vv, 1.7)
    ws = ops.fused_prepare(W, Q, Vv, 1.7, F)
    z1, _ = ops.fused_forward(ids, vals.clone(), tab, W, Q, Vv, 1.7, prepared=ws)
    assert torch.equal(z0, z1) and ops.last_launch_count() == 1
    with torch.no_grad():
        y0 = m({'id': ids, 'value': vals.clone()})
        m.attn_layer.query.mul_(1.5)                       # in-place update -> version bump -> cache rebuilt
        y1 = m({'id': ids, 'value': vals.clone()})
        m.cache_attention = False
        y2 = m({'id': ids, 'value': vals.clone()})
    assert not torch.equal(y0, y1) and torch.equal(y1, y2)


def test_mlp_shapes_outside_the_tail_kernel_limits_use_the_stock_path():
    """armnet_mlp_tail_f32 handles at most 64 outputs and a non-empty batch (csrc/mlp.cu); zoo.DCNModel builds
    MLP(..., noutput=mlp_hid) with mlp_hid = 256: such modules, and empty eval batches, must not raise."""
    from armnet_b200.layers import MLP
    torch.manual_seed(0)
    m = MLP(128, 2, 128, 0.0, noutput=128).cuda().eval()
    x = torch.randn(64, 128, device='cuda')
    with torch.no_grad():
        assert not m._fast_ok(x)
        y = m(x)
        assert torch.allclose(y, m.mlp(x))
        m1 = MLP(128, 2, 128, 0.0, noutput=1).cuda().eval()
        assert m1._fast_ok(x)
        assert m1(torch.empty(0, 128, device='cuda')).shape == (0, 1)
