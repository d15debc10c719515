"""A few launches of armnet_mlp_hidden_tc_f32 at the C2a shape for ncu (-k regex:mlp_hidden)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from armnet_b200 import ops
from armnet_b200.layers import MLP
dev = torch.device('cuda:0')
B, K, N = 4096, 5120, 256
x = torch.randn(B, K, device=dev)
m = MLP(K, 2, N, 0.0).to(dev).eval()
with torch.no_grad():
    w_hi, w_lo, packed, (ac, splits2, last) = m._prepared()
    part = ops.mlp_first_linear(x, w_hi, w_lo)
    for _ in range(4):
        y = ops.mlp_hidden_tc(part, B, ac, splits2, last)
torch.cuda.synchronize()
print(float(y.sum()))
