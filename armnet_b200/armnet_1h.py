"""Drop-in for models/armnet_1h.py (ARM-Net with one shared key projection): same constructor signature,
parameter names/shapes/init order and forward contract (armnet_1h.py:8-98). The hot path is the K == 1 case of the
fused kernel with the nn.Linear weight layout [d_k, nemb]."""
import torch
import torch.nn as nn

from . import ops
from .armnet import ARMNetModel as _MultiHead
from .armnet import _PaddedTable
from .layers import MLP, Embedding


class SparseAttention(nn.Module):
    """armnet_1h.py:8-34. forward() is the unfused composition used when autograd is recording."""

    def __init__(self, nfield, d_k, nhid, nemb, alpha=1.5):
        super().__init__()
        self.alpha = alpha
        self.scale = d_k ** -0.5
        self.bilinear_w = nn.Linear(nemb, d_k, bias=False)
        self.query = nn.Parameter(torch.zeros(nhid, d_k))
        self.values = nn.Parameter(torch.zeros(nhid, nfield))
        self.reset_parameters()

    def reset_parameters(self):
        nn.init.xavier_uniform_(self.query, gain=1.414)
        nn.init.xavier_uniform_(self.values, gain=1.414)

    def forward(self, x):
        """x [B,F,E] -> gates*values [B,O,F]"""
        keys = self.bilinear_w(x)
        g = torch.einsum('bfe,oe->bof', keys, self.query) * self.scale
        p = ops.entmax(g, alpha=self.alpha, dim=-1)
        return torch.einsum('bof,of->bof', p, self.values)


class ARMNetModel(_MultiHead):
    """Adaptive Relation Modeling Network, one head (armnet_1h.py:37-98)."""

    one_head = True

    def __init__(self, nfield, nfeat, nemb, alpha, nhid, d_k, mlp_nlayer, mlp_nhid, dropout, ensemble,
                 deep_nlayer, deep_nhid, noutput=1):
        nn.Module.__init__(self)
        self.alpha = alpha
        self.embedding = Embedding(nfeat, nemb)
        self.attn_layer = SparseAttention(nfield, d_k, nhid, nemb, alpha)
        self.arm_bn = nn.BatchNorm1d(nhid)
        self.mlp = MLP(nhid * nemb, mlp_nlayer, mlp_nhid, dropout, noutput=noutput)
        if ensemble:
            self.deep_embedding = Embedding(nfeat, nemb)
            self.deep_mlp = MLP(nfield * nemb, deep_nlayer, deep_nhid, dropout, noutput=noutput)
            self.ensemble_layer = nn.Linear(2 * noutput, noutput)
            nn.init.constant_(self.ensemble_layer.weight, 0.5)
            nn.init.constant_(self.ensemble_layer.bias, 0.)
        self.padded_table = True
        self.validate_ids = 'lazy'
        self.solver = ops.SOLVER_AUTO
        self.fuse_bn = True
        self.fused_backward = True
        self.cuda_bn = True
        self.cache_attention = True
        self._attn_key = None
        self._attn_ws = None
        self._shadow = _PaddedTable()
        self._err_flag = None
        self._bn_key = None

    def _attn_weights(self):
        a = self.attn_layer
        return a.bilinear_w.weight, a.query, a.values

    def _interaction_unfused(self, x):
        x['value'].clamp_(1e-3, 1.)
        e = self.embedding(x)
        w = self.attn_layer(e)
        return torch.exp(torch.einsum('bfe,bof->boe', e, w))
