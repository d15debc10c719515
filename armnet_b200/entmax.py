"""Drop-in for utils/entmax.py's public surface on this path: entmax_bisect() and the EntmaxBisect module
(entmax.py:134-175, :238-275), computed by the CUDA row solver (csrc/entmax.cu)."""
import torch.nn as nn

from . import ops

entmax_bisect = ops.entmax


class EntmaxBisect(nn.Module):
    def __init__(self, alpha=1.5, dim=-1, n_iter=50):
        super().__init__()
        self.dim = dim
        self.n_iter = n_iter
        self.alpha = alpha

    def forward(self, X):
        return ops.entmax(X, alpha=self.alpha, dim=self.dim, n_iter=self.n_iter)
