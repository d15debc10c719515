"""Timeline of armnet_fwd_tmem_kernel from a -DARMNET_TMEM_TRACE build (tools/build_variant.sh trace "-DARMNET_TMEM_TRACE"):
ARMNET_B200_LIB=armnet_b200/tuning/libtrace.so python tools/trace_tmem.py [--regime init|trained] [--no-flush]
Prints, per trace slot, min / median / max over the CTAs of (clock64 at the event - clock64 at kernel entry) in us."""
import argparse, ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import bench
from armnet_b200 import ops, _capi

ap = argparse.ArgumentParser()
ap.add_argument('--workload', default='c2a')
ap.add_argument('--regime', default='init')
ap.add_argument('--no-flush', action='store_true')
ap.add_argument('--mhz', type=float, default=1965.0)
a = ap.parse_args()
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
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for _ in range(5):
    if not a.no_flush:
        flush.zero_()
    e0.record()
    z, _ = ops.fused_forward(ids, vals, tab, W, Q, Vv, w['alpha'], one_head=model.one_head, ld=ld, nemb=table.shape[1],
                             prepared=ws)
    e1.record()
torch.cuda.synchronize()
print('event time of the last launch: %.1f us' % (1000 * e0.elapsed_time(e1)))
lib = _capi.lib if hasattr(_capi, 'lib') else ctypes.CDLL(_capi.LIB_PATH)
buf = (ctypes.c_longlong * (148 * 64))()
rc = lib.armnet_debug_tmem_trace(buf)
assert rc == 0, rc
t = np.frombuffer(buf, dtype=np.int64).reshape(148, 64).astype(np.float64)
rel = (t - t[:, :1]) / a.mhz
names = {0: 'entry', 1: 'after first __syncthreads', 2: 'A operand in TMEM (warp 0)', 3: 'gather: tile 0 issued',
         4: 'convert: tile 0 rows landed', 5: 'convert: tile 0 written', 6: 'mma: a_ready seen', 7: 'mma: item 0 committed',
         8: 'consumer 0: V table landed', 9: 'consumer 0: first logits ready', 10: 'consumer 0: first unit solved',
         53: 'gather: branch entered', 54: 'gather: prefetch(0) issued', 55: 'gather: t=0 ring wait passed', 56: 'gather: t=0 values in smem (LDG landed)', 57: 'gather: t=0 expect_tx armed',
         12: 'gather warp done', 13: 'convert warp done', 14: 'mma warp done', 15: 'after last __syncthreads'}
for s in sorted(names):
    v = rel[:, s]
    print('%-34s min %7.2f  med %7.2f  max %7.2f us' % (names[s], v.min(), np.median(v), v.max()))
cons = rel[:, 16:32]
print('consumer warps, all units done    min %7.2f  med %7.2f  max %7.2f us  (per CTA: last - first consumer: med %.2f, max %.2f)' %
      (cons.min(), np.median(cons), cons.max(), np.median(cons.max(1) - cons.min(1)), (cons.max(1) - cons.min(1)).max()))
st = rel[:, 36:52]
print('consumer warps, first logits      min %7.2f  med %7.2f  max %7.2f us' % (st.min(), np.median(st), st.max()))
end = rel[:, 15]
print('CTA lifetime: min %.2f med %.2f max %.2f us; CTAs with 14 tiles: %d' % (end.min(), np.median(end), end.max(),
      int(((w['bsz'] + 1) // 2) - 148 * (((w['bsz'] + 1) // 2) // 148))))
