"""BASELINE config 5: model-zoo sweep (afm / dcn+ / xdfm / afn) sharing the fused embedding-gather kernel, synthetic Criteo
shape (39 fields, 1M vocabulary, nemb 10), bsz 4096, eval forward, inputs resident on one B200.
Per model: samples/s with the CUDA lookups of this repo vs the same model body with the stock torch lookups
(table[ids] * value[..., None], layers.py:20-21), plus the gather kernels alone against the HBM roofline."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from armnet_b200 import ops, zoo

dev = torch.device('cuda:0')
F, V, E, B = 39, 1000000, 10, 4096
g = torch.Generator().manual_seed(0)
batches = [(torch.randint(0, V, (B, F), generator=g).to(dev), torch.rand(B, F, generator=g).to(dev)) for _ in range(4)]
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timeit(fn, n=20, warm=3):
    for i in range(warm):
        fn(i)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
    torch.cuda.synchronize()
    for i in range(n):
        flush.zero_()
        ev[i][0].record()
        fn(i)
        ev[i][1].record()
    torch.cuda.synchronize()
    return sum(a.elapsed_time(b) for a, b in ev) / n


def stock_embed(ids, values, table, clamp=None, **kw):
    if clamp is not None:
        values.clamp_(*clamp)
    return table[ids] * values.unsqueeze(2)


def stock_linear(ids, values, weight, bias=None, err_flag=None):
    y = (weight.reshape(-1)[ids] * values).sum(1)
    return y + bias if bias is not None else y


models = {
    'afm': lambda: zoo.AFMModel(V, E, 32, 0.0),
    'dcn+': lambda: zoo.DCNModel(F, V, E, 3, 2, 256, 0.0),
    'xdfm': lambda: zoo.xDeepFMModel(F, V, E, 2, 32, 2, 256, 0.0),
    'afn': lambda: zoo.AFNModel(F, V, E, 128, 2, 256, 0.0, False, 2, 256),
}
res = {}
table = torch.randn(V, E, device=dev)
w1 = torch.randn(V, 1, device=dev)
t_g = timeit(lambda i: ops.embed_gather(batches[i % 4][0], batches[i % 4][1], table))
t_l = timeit(lambda i: ops.linear_gather(batches[i % 4][0], batches[i % 4][1], w1))
t_gs = timeit(lambda i: stock_embed(batches[i % 4][0], batches[i % 4][1], table))
gb = B * F * (12 + 8 * E) / 1e9
res['embed_gather_kernel'] = {'us': t_g * 1e3, 'algorithmic_GBps': gb / (t_g * 1e-3), 'stock_torch_us': t_gs * 1e3}
res['linear_gather_kernel'] = {'us': t_l * 1e3}
real = (ops.embed_gather, ops.linear_gather)
for name, mk in models.items():
    torch.manual_seed(1)
    m = mk().to(dev).eval()
    with torch.no_grad():
        ops.embed_gather, ops.linear_gather = real
        t_ours = timeit(lambda i: m({'id': batches[i % 4][0], 'value': batches[i % 4][1]}))
        ops.embed_gather, ops.linear_gather = stock_embed, stock_linear
        t_stock = timeit(lambda i: m({'id': batches[i % 4][0], 'value': batches[i % 4][1]}))
        ops.embed_gather, ops.linear_gather = real
    res[name] = {'samples_per_s': B / (t_ours * 1e-3), 'ms': t_ours, 'stock_lookups_samples_per_s': B / (t_stock * 1e-3),
                 'stock_lookups_ms': t_stock}
    del m
print(json.dumps(res))
