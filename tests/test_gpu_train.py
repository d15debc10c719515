"""Training-side kernels against stock torch: fused clamp+Adam over the flat bucket (train.py:62-65 semantics)."""
import pytest
import torch
import torch.nn as nn

pytestmark = pytest.mark.gpu


def dev():
    return torch.device('cuda:0')


def test_flat_adam_matches_torch_adam_with_clamp_hooks():
    from armnet_b200.parallel import FlatAdam
    torch.manual_seed(5)

    def make():
        torch.manual_seed(5)
        return nn.Sequential(nn.Linear(37, 19), nn.ReLU(), nn.Linear(19, 3)).to(dev())

    ref, ours = make(), make()
    opt = torch.optim.Adam(ref.parameters(), lr=3e-3)                     # train.py:62
    for p in ref.parameters():
        p.register_hook(lambda g: g.clamp(-1.0, 1.0))                     # train.py:64-65
    fa = FlatAdam(ours.parameters(), lr=3e-3, clamp=1.0)
    assert fa.numel() % 4 == 0
    g = torch.Generator().manual_seed(1)
    for step in range(6):
        x = (torch.randn(64, 37, generator=g) * 30).to(dev())            # large inputs: the clamp is active
        y = torch.randn(64, 3, generator=g).to(dev())
        opt.zero_grad()
        ((ref(x) - y) ** 2).sum().backward()
        opt.step()
        fa.zero_grad()
        ((ours(x) - y) ** 2).sum().backward()
        fa.step()
        for a, b in zip(ref.parameters(), ours.parameters()):
            assert torch.allclose(a, b, rtol=1e-5, atol=1e-6), (step, (a - b).abs().max().item())


def test_flat_adam_trains_the_drop_in_model():
    """FlatAdam re-points parameters to views of one buffer: the fused forward / backward must keep working on them."""
    import armnet_b200 as ab
    from armnet_b200.parallel import FlatAdam
    torch.manual_seed(0)
    F, V, B = 10, 500, 128
    model = ab.ARMNetModel(F, V, 16, 2, 1.7, 8, 1, 16, 0.0, False, 1, 8).to(dev()).train()
    fa = FlatAdam(model.parameters(), lr=1e-2, clamp=1.0)
    ids = torch.randint(0, V, (B, F), device=dev())
    target = (torch.rand(B, device=dev()) < 0.3).float()
    crit = nn.BCEWithLogitsLoss()
    losses = []
    for _ in range(30):
        fa.zero_grad()
        loss = crit(model({'id': ids, 'value': torch.ones(B, F, device=dev())}).reshape(-1), target)
        loss.backward()
        fa.step()
        losses.append(loss.item())
    assert losses[-1] < 0.7 * losses[0]


@pytest.mark.parametrize('B,C,L', [(4096, 512, 16), (257, 40, 10), (64, 3, 1), (2, 7, 5)])
def test_train_mode_arm_bn_matches_torch(B, C, L):
    """csrc/bn.cu against nn.BatchNorm1d in train mode: output, running statistics, dx / dweight / dbias."""
    from armnet_b200 import ops
    torch.manual_seed(B + C)
    # like the interaction output at initialisation: exp(small) ~ 1 with a tiny spread (cancellation-prone)
    z = torch.exp(torch.randn(B, C, L, device=dev()) * 1e-3 + torch.randn(1, C, 1, device=dev()) * 0.05)
    ref = nn.BatchNorm1d(C).to(dev()).train()
    ours = nn.BatchNorm1d(C).to(dev()).train()
    with torch.no_grad():
        ref.weight.uniform_(0.5, 1.5)
        ref.bias.normal_(0, 0.3)
        ours.load_state_dict(ref.state_dict())
    z1 = z.clone().requires_grad_(True)
    z2 = z.clone().requires_grad_(True)
    y1 = ref(z1.double()).float() if False else ref(z1)
    y2 = ops.batch_norm_train(z2, ours)
    # fp64 evaluation is the judge: our error must not exceed cuDNN's by much
    z64 = z.double()
    m64 = z64.mean(dim=(0, 2), keepdim=True)
    v64 = z64.var(dim=(0, 2), unbiased=False, keepdim=True)
    y64 = (z64 - m64) / torch.sqrt(v64 + ref.eps) * ref.weight.double().view(1, C, 1) + ref.bias.double().view(1, C, 1)
    e_ref = (y1.double() - y64).abs().max().item()
    e_ours = (y2.double() - y64).abs().max().item()
    print(f'B={B} C={C} L={L}: max abs err vs fp64: ours {e_ours:.2e}, torch/cuDNN {e_ref:.2e}')
    assert e_ours <= max(2 * e_ref, 1e-4)
    assert torch.allclose(ours.running_mean, ref.running_mean, rtol=1e-5, atol=1e-6)
    assert torch.allclose(ours.running_var, ref.running_var, rtol=1e-4, atol=1e-7)
    assert int(ours.num_batches_tracked) == 1
    gy = torch.randn(B, C, L, device=dev())
    y1.backward(gy)
    y2.backward(gy)
    scale = z1.grad.abs().max().item()
    assert (z2.grad - z1.grad).abs().max().item() <= 2e-3 * scale
    assert torch.allclose(ours.weight.grad, ref.weight.grad, rtol=2e-3, atol=2e-3 * ref.weight.grad.abs().max().item())
    assert torch.allclose(ours.bias.grad, ref.bias.grad, rtol=1e-4, atol=1e-4 * ref.bias.grad.abs().max().item())


def test_train_mode_bn_single_value_raises():
    from armnet_b200 import ops
    bn = nn.BatchNorm1d(4).to(dev()).train()
    with pytest.raises(ValueError):
        ops.batch_norm_train(torch.ones(1, 4, 1, device=dev()), bn)


@pytest.mark.parametrize('case', ['c1_1h', 'c4_mh'])
@pytest.mark.parametrize('optimizer', ['flat_adam', 'torch_adam'])
def test_training_trajectory_matches_reference(case, optimizer):
    """Replays tests/golden/traj_<case>.npz (the unmodified reference's train.py steps on CPU: BCE, Adam, clamp hooks):
    same initial state, same batches -> same losses step by step and the same parameters at the end, through the fused
    forward / backward kernels, the train-mode arm_bn kernels and (flat_adam) the fused clamp+Adam kernel."""
    import os
    import numpy as np
    import armnet_b200 as ab
    from armnet_b200.parallel import FlatAdam, GradAllReducer
    z = np.load(os.path.join(os.path.dirname(__file__), 'golden', f'traj_{case}.npz'))
    c = {k[4:]: z[k].item() for k in z.files if k.startswith('cfg/')}
    if c['model'] == 'armnet':
        model = ab.ARMNetModel(c['nfield'], c['nfeat'], c['nemb'], c['nhead'], c['alpha'], c['nhid'], c['mlp_nlayer'],
                               c['mlp_nhid'], 0.0, False, 2, 16)
    else:
        model = ab.ARMNet1H(c['nfield'], c['nfeat'], c['nemb'], c['alpha'], c['nhid'], c['d_k'], c['mlp_nlayer'],
                            c['mlp_nhid'], 0.0, False, 2, 16)
    model.load_state_dict({k[7:]: torch.from_numpy(z[k]) for k in z.files if k.startswith('state0/')})
    model = model.to(dev()).train()
    ids, vals, target = (torch.from_numpy(z[k]).to(dev()) for k in ('ids', 'values', 'target'))
    crit = nn.BCEWithLogitsLoss(reduction='mean')
    if optimizer == 'flat_adam':
        fa = FlatAdam(model.parameters(), lr=c['lr'], clamp=1.0)
    else:
        opt = torch.optim.Adam(model.parameters(), lr=c['lr'])
        red = GradAllReducer(model.parameters(), clamp=1.0)
    losses = []
    for t in range(c['steps']):
        y = model({'id': ids[t], 'value': vals[t].clone()})
        loss = crit(y.reshape(-1), target[t])
        if optimizer == 'flat_adam':
            fa.zero_grad()
            loss.backward()
            fa.step()
        else:
            opt.zero_grad(set_to_none=True)
            loss.backward()
            red.step()
            opt.step()
        losses.append(loss.item())
    ref = z['losses']
    print(case, optimizer, 'loss - ref per step:', np.array2string(np.array(losses) - ref, precision=1))
    # One-head / 10 fields: the whole trajectory agrees to fp32 rounding (1e-7 per step). 39 fields x 64 neurons: a few of
    # the 320k gate entries per batch sit within rounding of the entmax support boundary, where the Jacobian factor
    # p^(2-alpha) (entmax.py:73-74) is not Lipschitz (p = 0 vs 1e-10 -> 0 vs 1e-3); Adam's g / sqrt(v) then amplifies those
    # differences step by step (the reference in fp64 vs fp32 drifts the same way). The first steps must agree tightly,
    # the rest to 2e-3 of a loss of ~0.7.
    assert np.allclose(losses[:2], ref[:2], rtol=0, atol=2e-6)
    if case == 'c1_1h':
        assert np.allclose(losses, ref, rtol=0, atol=2e-6)
    else:
        assert np.allclose(losses, ref, rtol=0, atol=2e-3)
        return
    final = model.state_dict()
    # A constant shift in front of a train-mode BatchNorm is removed by its mean subtraction: arm_bn.bias and the bias of
    # every Linear that feeds a BatchNorm have an analytically ZERO gradient. What reaches Adam is rounding noise
    # (|g| ~ 1e-10), which m / (sqrt(v) + eps) turns into a random walk -- in the reference as much as here.
    n_lin = c['mlp_nlayer']
    # The running_mean of those BatchNorms tracks the walked bias, so it is excluded with them.
    noise_only = {'arm_bn.bias'} | {f'mlp.mlp.{4 * i}.bias' for i in range(n_lin)} | \
                 {f'mlp.mlp.{4 * i + 1}.running_mean' for i in range(n_lin)}
    errs = {}
    for k in z.files:
        if not k.startswith('final/') or 'num_batches' in k or k[6:] in noise_only:
            continue
        a, b = final[k[6:]].detach().cpu().double(), torch.from_numpy(z[k]).double()
        errs[k[6:]] = (a - b).abs().max().item() / max(b.abs().max().item(), 1e-6)
    print({k: float(f'{v:.1e}') for k, v in sorted(errs.items(), key=lambda kv: -kv[1])[:6]})
    assert max(errs.values()) <= 2e-3, errs    # Adam divides by sqrt(v): tiny gradient differences are amplified


@pytest.mark.parametrize('B,K,N', [(4096, 8192, 256), (512, 1024, 64), (1028, 1500, 36)])
def test_linear_tf32x3_forward_backward_match_fp64(B, K, N):
    """The training-mode first Linear on tcgen05 (ops.linear_tf32x3: y, dx, dW, db) against an fp64 evaluation; the stock
    fp32 cuBLAS path is measured next to it."""
    from armnet_b200 import ops
    torch.manual_seed(B + N)
    x = torch.randn(B, K, device=dev())
    W = (torch.randn(N, K, device=dev()) * K ** -0.5)
    b = torch.randn(N, device=dev()) * 0.1
    gy = torch.randn(B, N, device=dev())

    def run(fn, dtype):
        xx, ww, bb = (t.detach().to(dtype).requires_grad_(True) for t in (x, W, b))
        y = fn(xx, ww, bb)
        y.backward(gy.to(dtype))
        return y.detach(), xx.grad, ww.grad, bb.grad

    old = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    ref = run(torch.nn.functional.linear, torch.float64)
    stock = run(torch.nn.functional.linear, torch.float32)
    torch.backends.cuda.matmul.allow_tf32 = old
    ours = run(ops.linear_tf32x3, torch.float32)
    for name, o, s_, r in zip(('y', 'dx', 'dW', 'db'), ours, stock, ref):
        eo = ((o.double() - r).abs().max() / r.abs().max()).item()
        es = ((s_.double() - r).abs().max() / r.abs().max()).item()
        print(f'B={B} K={K} N={N} {name}: tcgen05 3xTF32 err {eo:.2e}, cuBLAS fp32 err {es:.2e}')
        assert eo <= 1e-5, name


@pytest.mark.parametrize('rows,cols', [(4096, 8192), (257, 130), (64, 64), (1, 7), (1000, 36)])
def test_transpose2d(rows, cols):
    from armnet_b200 import ops
    x = torch.randn(rows, cols, device=dev())
    assert torch.equal(ops.transpose2d(x), x.t().contiguous())


# ------------------------------------------------------------------ input pipeline + metrics on the device
def test_device_split_is_resident_covers_every_row_and_shards():
    """DeviceSplit (data_loader.py:12-73 replacement): tensors live on the GPU, one epoch touches every row exactly once,
    ranks take disjoint contiguous shares of every global batch, batches come out on the device with int ids."""
    from armnet_b200.data import DeviceSplit
    n, F = 1000, 7
    ids = torch.arange(n * F, dtype=torch.int32).reshape(n, F)
    vals = torch.rand(n, F)
    y = torch.arange(n, dtype=torch.float32)
    sp = DeviceSplit(ids.numpy(), vals.numpy(), y.numpy(), device=dev())
    assert sp.ids.is_cuda and sp.values.is_cuda and sp.y.is_cuda and len(sp) == n
    assert sp.nbytes() == n * F * 8 + n * 4
    seen = []
    for b in sp.batches(128, shuffle=True, generator=torch.Generator().manual_seed(3)):
        assert b['id'].is_cuda and b['id'].shape[1] == F and b['value'].dtype == torch.float32
        assert torch.equal(b['id'][:, 0].long(), (b['y'] * F).long())          # rows stay intact
        seen.append(b['y'])
    assert torch.equal(torch.cat(seen).sort().values.cpu(), y)
    # two ranks: same global order, disjoint shares whose union is the global batch
    for (b0, b1, bg) in zip(sp.batches(128, True, torch.Generator().manual_seed(3), 0, 2),
                            sp.batches(128, True, torch.Generator().manual_seed(3), 1, 2),
                            sp.batches(128, True, torch.Generator().manual_seed(3))):
        assert torch.equal(torch.cat([b0['y'], b1['y']]), bg['y'])
        assert abs(b0['y'].numel() - b1['y'].numel()) <= 1 and b0['global_rows'] == bg['y'].numel()


def test_auc_on_device_cuda_matches_sklearn():
    import importlib.util
    import os
    from sklearn.metrics import roc_auc_score
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location('armnet_train_cli', os.path.join(root, 'train.py'))
    tr = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(tr)
    g = torch.Generator().manual_seed(0)
    logits = ((torch.randn(4096, generator=g) * 4).round() / 4).to(dev())      # ties
    target = (torch.rand(4096, generator=g) < 0.3).float().to(dev())
    got = tr.auc_on_device(logits, target)
    assert got.is_cuda and got.dim() == 0
    assert abs(float(got) - roc_auc_score(target.cpu().numpy(), logits.cpu().numpy())) < 1e-9
    assert float(tr.auc_on_device(logits, torch.ones_like(target))) == 0.0


def test_eval_caches_follow_flat_adam_and_bn_updates():
    """FlatAdam and csrc/bn.cu write parameters / running statistics through raw pointers; every eval-time cache
    (padded table, attention workspace, folded arm_bn, split MLP weights) is keyed on tensor versions and must be
    rebuilt: train -> eval -> train -> eval, second eval compared with a fresh module holding the same state."""
    import armnet_b200 as ab
    from armnet_b200.parallel import FlatAdam
    torch.manual_seed(0)
    F, V, B = 10, 300, 256
    args = (F, V, 10, 2, 1.7, 8, 2, 16, 0.0, False, 1, 8)
    model = ab.ARMNetModel(*args).to(dev())
    fa = FlatAdam(model.parameters(), lr=5e-2, clamp=1.0)
    ids = torch.randint(0, V, (B, F), device=dev())
    target = (torch.rand(B, device=dev()) < 0.3).float()
    crit = nn.BCEWithLogitsLoss()

    def train_steps(k):
        model.train()
        for _ in range(k):
            fa.zero_grad()
            crit(model({'id': ids, 'value': torch.ones(B, F, device=dev())}).reshape(-1), target).backward()
            fa.step()

    def eval_once(m):
        m.eval()
        with torch.no_grad():
            return m({'id': ids, 'value': torch.ones(B, F, device=dev())}).clone()

    train_steps(3)
    y1 = eval_once(model)                  # builds every cache
    train_steps(3)
    y2 = eval_once(model)                  # must see the new parameters and running statistics
    fresh = ab.ARMNetModel(*args).to(dev())
    fresh.load_state_dict({k: v.clone() for k, v in model.state_dict().items()})
    y_ref = eval_once(fresh)
    assert (y1 - y2).abs().max().item() > 1e-4, 'training did not move the output: the test is vacuous'
    assert torch.allclose(y2, y_ref, rtol=1e-6, atol=1e-6), (y2 - y_ref).abs().max().item()


def _sharded_adam_worker(rank, world, port, out):
    import os
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    d = torch.device('cuda', rank)
    dist.init_process_group('nccl', device_id=d)
    from armnet_b200.parallel import FlatAdam
    res = {}
    for shard in (False, True):
        torch.manual_seed(5)
        net = torch.nn.Sequential(torch.nn.Linear(37, 19), torch.nn.ReLU(), torch.nn.Linear(19, 3)).to(d)   # odd sizes: padding
        opt = FlatAdam(net.parameters(), lr=1e-2, clamp=0.05, shard_state='force' if shard else False)
        assert opt.sharded == shard
        g = torch.Generator().manual_seed(100 + rank)
        for _ in range(4):
            x = torch.randn(16, 37, generator=g).to(d)
            opt.zero_grad()
            net(x).square().mean().backward()
            opt.step()
        res[shard] = torch.cat([p.detach().reshape(-1) for p in net.parameters()]).cpu()
        if shard:
            assert opt.exp_avg.numel() * world == opt.flat_p.numel()      # the state really is 1 / world per rank
    if rank == 0:
        torch.save(res, out)
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs two GPUs (NCCL reduce-scatter / all-gather)')
def test_flat_adam_sharded_state_matches_all_reduce(tmp_path):
    """FlatAdam(shard_state=...): reduce-scatter + Adam on 1/world of the bucket + all-gather gives the parameters of the
    all-reduce + full Adam path (two ranks, different data per rank, gradient clamp active, bucket size not a multiple of
    the world size)."""
    import torch.multiprocessing as mp
    out = str(tmp_path / 'res.pt')
    mp.spawn(_sharded_adam_worker, args=(2, 29577, out), nprocs=2, join=True)
    res = torch.load(out)
    assert torch.allclose(res[False], res[True], rtol=0, atol=1e-7)
    assert not torch.equal(res[False], torch.zeros_like(res[False]))
