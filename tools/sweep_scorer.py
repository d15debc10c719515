"""BatchScorer pipeline shapes (slots x compute streams) at C2a: samples/s from pinned host batches (wall clock, 300 batches)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from armnet_b200 import BatchScorer
w = bench.WORKLOADS['c2a']
dev = torch.device('cuda:0')
model = bench.build_module(w).to(dev).eval()
host = [(i.pin_memory(), v.pin_memory()) for i, v in bench.make_batches(w, 8, seed=3)]
import sys as _s
model.mlp.first_splits = int(_s.argv[1]) if len(_s.argv) > 1 else None   # optional: split-K slices of the first Linear
for depth, streams in ((6, 3), (4, 2), (8, 4), (8, 2), (9, 3), (12, 4), (2, 1)):
    scorer = BatchScorer(model, w['bsz'], w['nfield'], depth=depth, compute_streams=streams)
    def run(steps):
        pending, acc = [], 0.0
        for i in range(steps):
            if len(pending) == scorer.depth:
                acc += float(scorer.result(pending.pop(0))[0])
            pending.append(scorer.submit(*host[i % len(host)]))
        for t in pending:
            acc += float(scorer.result(t)[0])
        return acc
    run(20)
    torch.cuda.synchronize()
    best = 0
    for _ in range(3):
        t0 = time.perf_counter()
        run(300)
        torch.cuda.synchronize()
        best = max(best, 300 * w['bsz'] / (time.perf_counter() - t0))
    print(f'depth {depth:2d} streams {streams}: {best / 1e6:.2f} M samples/s ({w["bsz"] / best * 1e6:.1f} us / batch)')
    del scorer
