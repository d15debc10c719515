"""A few launches of the fused hot path for ncu: python tools/prof_hot.py --workload c2a --regime init|trained [--tmem 0|1]
(ncu --set full -k regex:armnet_fwd --launch-skip 2 -c 1 python tools/prof_hot.py ...)."""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from armnet_b200 import ops

ap = argparse.ArgumentParser()
ap.add_argument('--workload', default='c2a')
ap.add_argument('--regime', default='init')
ap.add_argument('--tmem', type=int, default=-1)
ap.add_argument('--mma', type=int, default=-1)
ap.add_argument('--iters', type=int, default=4)
a = ap.parse_args()
ops.set_tuning('tmem', a.tmem)
ops.set_tuning('mma', a.mma)
w = bench.WORKLOADS[a.workload]
dev = torch.device('cuda:0')
model = bench.build_module(w).to(dev).eval()
if a.regime == 'trained':
    bench.trained_like_(model)
ids, vals = bench.make_batches(w, 1, seed=1000)[0]
ids, vals = ids.to(dev), vals.to(dev)
table = model.embedding.embedding.weight
tab, ld = model._shadow.get(table)
W, Q, Vv = (t.detach() for t in model._attn_weights())
ws = ops.fused_prepare(W, Q, Vv, w['alpha'], w['nfield'], one_head=model.one_head)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for _ in range(a.iters):
    flush.zero_()
    z, _ = ops.fused_forward(ids, vals, tab, W, Q, Vv, w['alpha'], one_head=model.one_head, ld=ld, nemb=table.shape[1],
                             prepared=ws)
torch.cuda.synchronize()
print('kind', ops.fused_fwd_kernel_kind(w['nfield'], w['nemb'], w['nhead'], w['nhid'], w['alpha']), 'z mean', float(z.mean()))
