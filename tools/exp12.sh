#!/bin/bash
timeout 300 python - <<'PY'
import sys, time, torch
sys.path.insert(0, '.')
import bench, armnet_b200 as ab
w = bench.WORKLOADS['c2a']; dev = torch.device('cuda:0')
model = bench.build_module(w).to(dev).eval()
host = [(i.pin_memory(), v.pin_memory()) for i, v in bench.make_batches(w, 4, seed=1000)]
for depth, ncs in ((4, 2), (4, 4), (6, 2), (6, 3), (8, 2), (8, 4), (4, 1)):
    sc = ab.BatchScorer(model, w['bsz'], w['nfield'], depth=depth, compute_streams=ncs)
    def run(steps):
        pend = []
        for i in range(steps):
            if len(pend) == depth: sc.result(pend.pop(0))
            pend.append(sc.submit(*host[i % 4]))
        for t in pend: sc.result(t)
    run(10); torch.cuda.synchronize(); t0 = time.perf_counter(); run(100); torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 100
    print('depth %d compute_streams %d: %.1f us/step  %.2f M samples/s' % (depth, ncs, dt * 1e6, w['bsz'] / dt / 1e6))
    del sc
PY
