import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.nn as nn
import armnet_b200 as ab
from armnet_b200 import ops
dev = torch.device('cuda:0')
z = np.load('tests/golden/traj_c4_mh.npz')
c = {k[4:]: z[k].item() for k in z.files if k.startswith('cfg/')}
ids, vals, target = (torch.from_numpy(z[k]).to(dev) for k in ('ids', 'values', 'target'))
crit = nn.BCEWithLogitsLoss()
for name, kw in [('all', {}), ('stock_bn', {'cuda_bn': False}), ('unfused_bwd', {'fused_backward': False}),
                 ('unfused_bwd+stock_bn', {'fused_backward': False, 'cuda_bn': False}), ('bisect', {'solver': ops.SOLVER_BISECT})]:
    model = ab.ARMNetModel(c['nfield'], c['nfeat'], c['nemb'], c['nhead'], c['alpha'], c['nhid'], c['mlp_nlayer'], c['mlp_nhid'], 0.0, False, 2, 16)
    model.load_state_dict({k[7:]: torch.from_numpy(z[k]) for k in z.files if k.startswith('state0/')})
    model = model.to(dev).train()
    for k, v in kw.items():
        setattr(model, k, v)
    opt = torch.optim.Adam(model.parameters(), lr=c['lr'])
    for p in model.parameters():
        p.register_hook(lambda g: g.clamp(-1.0, 1.0))
    losses = []
    for t in range(c['steps']):
        y = model({'id': ids[t], 'value': vals[t].clone()})
        loss = crit(y.reshape(-1), target[t])
        opt.zero_grad()
        loss.backward()
        opt.step()
        losses.append(loss.item())
    print(name, np.array2string(np.array(losses) - z['losses'], precision=1, max_line_width=200))
