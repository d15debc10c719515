"""Event-timed launches of the fused hot path (L2 flushed between launches), both regimes; for A/B runs of tuning builds:
ARMNET_B200_LIB=armnet_b200/tuning/libX.so python tools/time_hot.py [--workload c2a] [--iters 40]"""
import argparse, os, statistics, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from armnet_b200 import ops

ap = argparse.ArgumentParser()
ap.add_argument('--workload', default='c2a')
ap.add_argument('--iters', type=int, default=40)
ap.add_argument('--tag', default=os.path.basename(os.environ.get('ARMNET_B200_LIB', 'default')))
a = ap.parse_args()
w = bench.WORKLOADS[a.workload]
dev = torch.device('cuda:0')
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
out = []
for regime in ('init', 'trained'):
    model = bench.build_module(w).to(dev).eval()
    if regime == 'trained':
        bench.trained_like_(model)
    ids, vals = bench.make_batches(w, 1, seed=1000)[0]
    ids, vals = ids.to(dev), vals.to(dev)
    table = model.embedding.embedding.weight
    tab, ld = model._shadow.get(table)
    W, Q, Vv = (t.detach() for t in model._attn_weights())
    ws = ops.fused_prepare(W, Q, Vv, w['alpha'], w['nfield'], one_head=model.one_head)
    ts = []
    for i in range(a.iters + 5):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        z, _ = ops.fused_forward(ids, vals, tab, W, Q, Vv, w['alpha'], one_head=model.one_head, ld=ld, nemb=table.shape[1],
                                 prepared=ws)
        e1.record()
        torch.cuda.synchronize()
        if i >= 5:
            ts.append(1000 * e0.elapsed_time(e1))
    out.append('%s: med %.1f us min %.1f (%.2f M/s)' % (regime, statistics.median(ts), min(ts), w['bsz'] / statistics.median(ts)))
print('%-16s' % a.tag, ' | '.join(out), ' zsum %.6f' % float(z.double().sum()))
