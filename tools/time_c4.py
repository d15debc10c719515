"""Event-timed launches of the fused hot path at the config-4 shape (nemb 16) through each kernel kind (A/B of the wide
tensor-memory instance against armnet_fwd_mma_kernel): python tools/time_c4.py"""
import os, statistics, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from armnet_b200 import ops
w = bench.WORKLOADS['c4']
dev = torch.device('cuda:0')
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for regime in ('init', 'trained'):
    model = bench.build_module(w).to(dev).eval()
    if regime == 'trained':
        bench.trained_like_(model)
    ids, vals = bench.make_batches(w, 1, seed=1000)[0]
    ids, vals = ids.to(dev), vals.to(dev)
    table = model.embedding.embedding.weight
    tab, ld = model._shadow.get(table)
    W, Q, Vv = (t.detach() for t in model._attn_weights())
    res = {}
    for name, tun in (('mma', dict(tmem=0, mma=1)), ('tmem16', dict(tmem=1)), ('fp32', dict(tmem=0, mma=0))):
        with ops.tuning(**tun):
            kind = ops.fused_fwd_kernel_kind(w['nfield'], w['nemb'], w['nhead'], w['nhid'], w['alpha'])
            ws = ops.fused_prepare(W, Q, Vv, w['alpha'], w['nfield'], one_head=model.one_head)
            ts = []
            for i in range(25):
                flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                z, _ = ops.fused_forward(ids, vals, tab, W, Q, Vv, w['alpha'], one_head=model.one_head, ld=ld,
                                         nemb=table.shape[1], prepared=ws)
                e1.record()
                torch.cuda.synchronize()
                if i >= 5:
                    ts.append(1000 * e0.elapsed_time(e1))
            res[name] = (kind, statistics.median(ts), z.double().sum().item(), z)
    ref = res['fp32'][3]
    print(regime, ' | '.join('%s(kind %d): %.1f us (%.2f M/s) maxrel %.1e' % (k, v[0], v[1], w['bsz'] / v[1],
          ((v[3] - ref).abs().max() / ref.abs().max()).item()) for k, v in res.items()))
