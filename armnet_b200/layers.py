"""Host-side mirrors of the reference's shared layers on the hot path (models/layers.py), same names and
parameter layout, backed by the CUDA gather kernel."""
import torch
import torch.nn as nn

from . import ops


class _GatherFn(torch.autograd.Function):
    """e = table[ids] * values with the dense [V,E] table gradient the reference's nn.Embedding(sparse=False)
    produces (layers.py:12,20-21). No gradient flows to ids / values (leaf inputs in train.py:104-109)."""

    @staticmethod
    def forward(ctx, table, ids, values):
        ctx.save_for_backward(ids, values)
        ctx.table_shape = table.shape
        return ops.embed_gather(ids, values, table)

    @staticmethod
    def backward(ctx, de):
        ids, values = ctx.saved_tensors
        V, E = ctx.table_shape
        g = torch.zeros(V, E, dtype=de.dtype, device=de.device)
        g.index_add_(0, ids.reshape(-1).long(), (de * values.unsqueeze(2)).reshape(-1, E))
        return g, None, None


class Embedding(nn.Module):
    """layers.Embedding (layers.py:8-21): nn.Embedding(nfeat, nemb) + Xavier-uniform, forward = rows * value."""

    def __init__(self, nfeat, nemb):
        super().__init__()
        self.embedding = nn.Embedding(nfeat, nemb)
        nn.init.xavier_uniform_(self.embedding.weight)

    def forward(self, x):
        """x = {'id': Long[B,F], 'value': Float[B,F]} -> [B,F,E]"""
        w = self.embedding.weight
        if torch.is_grad_enabled() and w.requires_grad:
            return _GatherFn.apply(w, x['id'], x['value'])
        return ops.embed_gather(x['id'], x['value'], w)


class MLP(nn.Module):
    """layers.MLP (layers.py:68-88): [Linear -> BatchNorm1d -> ReLU -> Dropout] x nlayers -> Linear(., noutput).
    Dense GEMMs: stock torch (cuBLAS) modules, identical state_dict keys ('mlp.<i>.weight', ...)."""

    def __init__(self, ninput, nlayers, nhid, dropout, noutput=1):
        super().__init__()
        mods = []
        for _ in range(nlayers):
            mods += [nn.Linear(ninput, nhid), nn.BatchNorm1d(nhid), nn.ReLU(), nn.Dropout(p=dropout)]
            ninput = nhid
        if nlayers == 0:
            nhid = ninput
        mods.append(nn.Linear(nhid, noutput))
        self.mlp = nn.Sequential(*mods)

    def forward(self, x):
        return self.mlp(x)
