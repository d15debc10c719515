"""Times the tcgen05 3xTF32 first-Linear GEMM and the tail kernel against cuBLAS (fp32 SIMT and TF32) on one B200."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from armnet_b200 import ops
from armnet_b200.layers import MLP

dev = torch.device('cuda:0')
B, K, N = 4096, 5120, 256
x = torch.randn(B, K, device=dev)
w = torch.randn(N, K, device=dev) * K ** -0.5
hi, lo = ops.mlp_split_weight(w)
ref = x.double() @ w.double().t()


def timeit(fn, n=50, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e3


for s in (1, 2, 4, 8, 9, 16):
    t = timeit(lambda: ops.mlp_first_linear(x, hi, lo, splits=s))
    y = ops.partials_to_dense(ops.mlp_first_linear(x, hi, lo, splits=s), B)
    err = ((y.double() - ref).abs().max() / ref.abs().max()).item()
    print(f'3xTF32 tcgen05 splits={s:2d}: {t:7.1f} us  {2*B*K*N/t/1e6:7.1f} TFLOP/s (algorithmic)  err {err:.2e}')
torch.backends.cuda.matmul.allow_tf32 = False
t = timeit(lambda: torch.nn.functional.linear(x, w))
err = (((x @ w.t()).double() - ref).abs().max() / ref.abs().max()).item()
print(f'cuBLAS fp32: {t:7.1f} us  err {err:.2e}')
torch.backends.cuda.matmul.allow_tf32 = True
t = timeit(lambda: torch.nn.functional.linear(x, w))
err = (((x @ w.t()).double() - ref).abs().max() / ref.abs().max()).item()
print(f'cuBLAS tf32: {t:7.1f} us  err {err:.2e}')
torch.backends.cuda.matmul.allow_tf32 = False
m = MLP(K, 2, N, 0.0).to(dev).eval()
with torch.no_grad():
    t1 = timeit(lambda: m(x))
    m.tensor_core = False
    t2 = timeit(lambda: m(x))
    m.tensor_core = True
    w_hi, w_lo, packed, (ac, splits2, last) = m._prepared()
    part = ops.mlp_first_linear(x, w_hi, w_lo)
    t3 = timeit(lambda: ops.mlp_tail(part, packed, 1, 1, B))
    print(f'MLP eval forward: fast path {t1:.1f} us (CUDA-core tail kernel alone {t3:.1f} us), stock torch modules {t2:.1f} us')
    th = timeit(lambda: ops.mlp_hidden_tc(part, B, ac, splits2, last))
    print(f'  hidden layer 2 + output on tcgen05 (armnet_mlp_hidden_tc_f32): {th:.1f} us; whole MLP {timeit(lambda: m(x)):.1f} us')
    m.hidden_tensor_core = False
    print(f'  whole MLP with the CUDA-core tail: {timeit(lambda: m(x)):.1f} us')
