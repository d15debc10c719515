import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.nn as nn
import armnet_b200 as ab
dev = torch.device('cuda:0')
z = np.load('tests/golden/traj_c4_mh.npz')
c = {k[4:]: z[k].item() for k in z.files if k.startswith('cfg/')}
ids, vals, target = (torch.from_numpy(z[k]).to(dev) for k in ('ids', 'values', 'target'))
crit = nn.BCEWithLogitsLoss()
G = {}
for name, kw in [('fused', {}), ('unfused', {'fused_backward': False})]:
    model = ab.ARMNetModel(c['nfield'], c['nfeat'], c['nemb'], c['nhead'], c['alpha'], c['nhid'], c['mlp_nlayer'], c['mlp_nhid'], 0.0, False, 2, 16)
    model.load_state_dict({k[7:]: torch.from_numpy(z[k]) for k in z.files if k.startswith('state0/')})
    model = model.to(dev).train().double() if name == 'f64' else model.to(dev).train()
    for k, v in kw.items():
        setattr(model, k, v)
    y = model({'id': ids[0], 'value': vals[0].clone()})
    loss = crit(y.reshape(-1), target[0])
    loss.backward()
    G[name] = {n: p.grad.detach().double().cpu() for n, p in model.named_parameters()}
# fp64 reference of the same graph with plain torch ops (unfused formulas), on CPU
def f64_grads():
    st = {k[7:]: torch.from_numpy(z[k]).double() for k in z.files if k.startswith('state0/')}
    P = {k: v.clone().requires_grad_(True) for k, v in st.items() if v.dtype.is_floating_point and 'running' not in k}
    i0, v0, t0 = ids[0].cpu(), vals[0].cpu().double().clamp(0.001, 1.0), target[0].cpu().double()
    e = P['embedding.embedding.weight'][i0] * v0[..., None]
    g = torch.einsum('bfx,kxy,koy->bkof', e, P['attn_layer.bilinear_w'], P['attn_layer.query']) * c['d_k'] ** -0.5
    # entmax by bisection in fp64
    al = c['alpha']; X = g * (al - 1); mx = X.max(-1, keepdim=True).values
    lo = mx - 1; hi = mx - (1.0 / c['nfield']) ** (al - 1)
    with torch.no_grad():
        for _ in range(100):
            mid = (lo + hi) / 2
            f = (torch.clamp(X - mid, min=0) ** (1 / (al - 1))).sum(-1, keepdim=True) - 1
            lo = torch.where(f >= 0, mid, lo); hi = torch.where(f >= 0, hi, mid)
    tau = lo
    # differentiable gates through the implicit function: use the closed-form Jacobian by writing p as a function with custom grad
    class Ent(torch.autograd.Function):
        @staticmethod
        def forward(ctx, X):
            p = torch.clamp(X - tau, min=0) ** (1 / (al - 1)); p = p / p.sum(-1, keepdim=True)
            ctx.save_for_backward(p); return p
        @staticmethod
        def backward(ctx, dY):
            p, = ctx.saved_tensors
            gp = torch.where(p > 0, p ** (2 - al), torch.zeros_like(p))
            dX = dY * gp; q = dX.sum(-1, keepdim=True) / gp.sum(-1, keepdim=True)
            return (dX - q * gp) / (al - 1)      # dp/dg = p^(2-alpha): the (alpha-1) of X = (alpha-1) g cancels (entmax.py:71-80)
    p = Ent.apply(X)
    w = p * P['attn_layer.values']
    zz = torch.exp(torch.einsum('bfe,bkof->bkoe', e, w)).reshape(e.shape[0], -1, c['nemb'])
    bn = lambda x, wt, b, dims: (x - x.mean(dims, keepdim=True)) / torch.sqrt(x.var(dims, unbiased=False, keepdim=True) + 1e-5) * wt + b
    h = bn(zz, P['arm_bn.weight'].view(1, -1, 1), P['arm_bn.bias'].view(1, -1, 1), (0, 2)).reshape(e.shape[0], -1)
    for i in range(c['mlp_nlayer']):
        h = h @ P[f'mlp.mlp.{4*i}.weight'].t() + P[f'mlp.mlp.{4*i}.bias']
        h = torch.relu(bn(h, P[f'mlp.mlp.{4*i+1}.weight'], P[f'mlp.mlp.{4*i+1}.bias'], (0,)))
    y = (h @ P[f"mlp.mlp.{4*c['mlp_nlayer']}.weight"].t() + P[f"mlp.mlp.{4*c['mlp_nlayer']}.bias"]).reshape(-1)
    nn.BCEWithLogitsLoss()(y, t0).backward()
    return {k: v.grad for k, v in P.items()}
R = f64_grads()
for n in G['fused']:
    r = R[n]; s = r.abs().max().item()
    ef = (G['fused'][n] - r).abs().max().item() / s; eu = (G['unfused'][n] - r).abs().max().item() / s
    print(f'{n:34s} |g|max {s:.2e}  err/max: fused {ef:.1e}  unfused {eu:.1e}')
    if n == 'embedding.embedding.weight':
        d = (G['fused'][n] - r).abs(); i = d.argmax().item(); row = i // r.shape[1]
        print('   worst element row', row, 'ref', r.flatten()[i].item(), 'fused', G['fused'][n].flatten()[i].item(), 'unfused', G['unfused'][n].flatten()[i].item())
        small = (r.abs() < 1e-6) & (r.abs() > 0)
        print('   elements with 0<|g|<1e-6:', int(small.sum()), 'sign flips fused', int(((G['fused'][n] * r) < 0).sum()), 'unfused', int(((G['unfused'][n] * r) < 0).sum()))
