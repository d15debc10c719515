"""The four baseline models of BASELINE.json config 5 (afm / dcn / xdfm / afn) on top of the SAME CUDA gather kernels the
ARM-Net path uses: every one of them starts with layers.Embedding (reference afm.py:38, dcn.py:34,56, xdfm.py:45,66,
afn.py:20,37), AFM / CIN / xDeepFM add the 1-wide layers.Linear term (layers.py:24-37).  Only those two lookups are
kernels of this repo (armnet_embed_gather_f32, armnet_linear_gather_f32); the model bodies are a handful of dense
torch ops each, written here as einsum formulations of the papers' equations.  Constructor signatures, parameter names
and shapes follow the reference, so `state_dict`s are interchangeable and `create_model` can route to them
(model_utils.py:35-46,74-79).
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops
from .layers import MLP, Embedding

__all__ = ['Linear', 'AFMModel', 'CrossNetModel', 'DCNModel', 'CINModel', 'xDeepFMModel', 'AFNModel']


class _LinearGatherFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, weight, bias, ids, values):
        ctx.save_for_backward(ids, values)
        ctx.V = weight.shape[0]
        return ops.linear_gather(ids, values, weight, bias)

    @staticmethod
    def backward(ctx, dy):
        ids, values = ctx.saved_tensors
        gw = torch.zeros(ctx.V, dtype=dy.dtype, device=dy.device)
        gw.index_add_(0, ids.reshape(-1).long(), (dy.unsqueeze(1) * values).reshape(-1))
        return gw.unsqueeze(1), dy.sum().reshape(1), None, None


class Linear(nn.Module):
    """layers.Linear (layers.py:24-37): y[b] = sum_f weight[id[b,f]] * value[b,f] + bias, one gather-reduce kernel."""

    def __init__(self, nfeat):
        super().__init__()
        self.weight = nn.Embedding(nfeat, 1)
        self.bias = nn.Parameter(torch.zeros(1))

    def forward(self, x):
        w = self.weight.weight
        if torch.is_grad_enabled() and w.requires_grad:
            return _LinearGatherFn.apply(w, self.bias, x['id'], x['value'])
        return ops.linear_gather(x['id'], x['value'], w, self.bias.detach())


# --------------------------------------------------------------------------------------------------------- AFM
class AttentionalFactorizationMachine(nn.Module):
    """Attention-weighted sum of the pairwise element-wise products e_i * e_j, i < j (afm.py:5-28; Xiao et al. 2017)."""

    def __init__(self, nemb, nattn, dropout):
        super().__init__()
        self.attn_w = nn.Linear(nemb, nattn)
        self.attn_h = nn.Linear(nattn, 1)
        self.attn_p = nn.Linear(nemb, 1)
        self.dropout = nn.Dropout(p=dropout)

    def forward(self, e):
        nf = e.shape[1]
        i, j = torch.triu_indices(nf, nf, offset=1, device=e.device)      # row-major upper triangle, as np.triu_indices
        pair = e[:, i] * e[:, j]                                          # [B, P, E]
        score = self.attn_h(F.relu(self.attn_w(pair)))                    # [B, P, 1]
        a = self.dropout(torch.softmax(score, dim=1))
        pooled = self.dropout((a * pair).sum(dim=1))                      # [B, E]
        return self.attn_p(pooled).squeeze(1)


class AFMModel(nn.Module):
    def __init__(self, nfeat, nemb, nattn, dropout):
        super().__init__()
        self.embedding = Embedding(nfeat, nemb)
        self.linear = Linear(nfeat)
        self.afm = AttentionalFactorizationMachine(nemb, nattn, dropout)

    def forward(self, x):
        return self.linear(x) + self.afm(self.embedding(x))


# --------------------------------------------------------------------------------------------------------- DCN
class CrossNetwork(nn.Module):
    """x_{l+1} = x_0 (w_l . x_l) + b_l + x_l (dcn.py:4-26; Wang et al. 2017)."""

    def __init__(self, ninput, nlayers):
        super().__init__()
        self.nlayers = nlayers
        self.w = nn.ModuleList([nn.Linear(ninput, 1, bias=False) for _ in range(nlayers)])
        self.b = nn.ParameterList([nn.Parameter(torch.zeros(ninput)) for _ in range(nlayers)])

    def forward(self, x0):
        x = x0
        for w, b in zip(self.w, self.b):
            x = x0 * w(x) + b + x
        return x


class CrossNetModel(nn.Module):
    """--model dcn: cross network only (dcn.py:28-49)."""

    def __init__(self, nfield, nfeat, nemb, cn_layers):
        super().__init__()
        self.embedding = Embedding(nfeat, nemb)
        self.ninput = nfield * nemb
        self.cross_net = CrossNetwork(self.ninput, cn_layers)
        self.w = nn.Linear(self.ninput, 1, bias=False)

    def forward(self, x):
        flat = self.embedding(x).reshape(-1, self.ninput)
        return self.w(self.cross_net(flat)).squeeze(1)


class DCNModel(nn.Module):
    """--model dcn+: cross network next to an MLP (dcn.py:51-70)."""

    def __init__(self, nfield, nfeat, nemb, cn_layers, mlp_layers, mlp_hid, dropout):
        super().__init__()
        self.embedding = Embedding(nfeat, nemb)
        self.ninput = nfield * nemb
        self.cross_net = CrossNetwork(self.ninput, cn_layers)
        self.mlp = MLP(self.ninput, mlp_layers, mlp_hid, dropout, noutput=mlp_hid)
        self.w = nn.Linear(mlp_hid + self.ninput, 1, bias=False)

    def forward(self, x):
        flat = self.embedding(x).reshape(-1, self.ninput)
        return self.w(torch.cat([self.cross_net(flat), self.mlp(flat)], dim=1)).squeeze(1)


# --------------------------------------------------------------------------------------------------------- CIN / xDeepFM
class CompressedInteraction(nn.Module):
    """x^k[h] = relu(sum_{i,j} W^k[h,i,j] x^0[i] * x^{k-1}[j]) per embedding lane, sum-pooled over lanes, then one
    affine layer over all layers' feature maps (xdfm.py:5-34; Lian et al. 2018). The 1x1 Conv1d weights keep the
    reference's [nfilter, nfield * n_prev, 1] shape."""

    def __init__(self, nfield, nlayers, nfilter):
        super().__init__()
        self.nlayers = nlayers
        self.filters = nn.ModuleList()
        n_prev, n_fc = nfield, 0
        for _ in range(nlayers):
            self.filters.append(nn.Conv1d(nfield * n_prev, nfilter, kernel_size=1, bias=False))
            n_prev = nfilter
            n_fc += nfilter
        self.affine = nn.Linear(n_fc, 1, bias=False)

    def forward(self, e):
        B, nf, _ = e.shape
        xk, pooled = e, []
        for conv in self.filters:
            w = conv.weight.view(conv.out_channels, nf, xk.shape[1])              # [H, F, n_prev]
            xk = F.relu(torch.einsum('bie,bje,hij->bhe', e, xk, w))                # [B, H, E]
            pooled.append(xk.sum(dim=-1))
        return self.affine(torch.cat(pooled, dim=1)).squeeze(1)


class CINModel(nn.Module):
    """--model cin (xdfm.py:37-57)."""

    def __init__(self, nfield, nfeat, nemb, cin_layers, nfilter):
        super().__init__()
        self.embedding = Embedding(nfeat, nemb)
        self.linear = Linear(nfeat)
        self.cin = CompressedInteraction(nfield, cin_layers, nfilter)

    def forward(self, x):
        return self.linear(x) + self.cin(self.embedding(x))


class xDeepFMModel(nn.Module):
    """--model xdfm (xdfm.py:59-79)."""

    def __init__(self, nfield, nfeat, nemb, cin_layers, nfilter, mlp_layers, mlp_hid, dropout):
        super().__init__()
        self.embedding = Embedding(nfeat, nemb)
        self.linear = Linear(nfeat)
        self.cin = CompressedInteraction(nfield, cin_layers, nfilter)
        self.ninput = nfield * nemb
        self.mlp = MLP(self.ninput, mlp_layers, mlp_hid, dropout)

    def forward(self, x):
        e = self.embedding(x)
        return self.linear(x) + self.cin(e) + self.mlp(e.reshape(-1, self.ninput)).squeeze(1)


# --------------------------------------------------------------------------------------------------------- AFN
class AFNModel(nn.Module):
    """Logarithmic neurons: exp(W . log e) learns adaptive-order products of the (positive) embeddings
    (afn.py:5-77; Cheng et al. 2020). Like the reference, forward() clamps x['value'] in place and REWRITES the embedding
    table in place (abs, then >= 1e-4) before the lookup (afn.py:52-54,74-77)."""

    def __init__(self, nfield, nfeat, nemb, afn_hid, mlp_layers, mlp_hid, dropout, ensemble, deep_layers, deep_hid):
        super().__init__()
        self.nfield, self.nfeat, self.nemb, self.afn_hid, self.ensemble = nfield, nfeat, nemb, afn_hid, ensemble
        self.dropout = nn.Dropout(p=dropout)
        self.embedding = Embedding(nfeat, nemb)
        self.emb_bn = nn.BatchNorm1d(nfield)
        self.afn = nn.Linear(nfield, afn_hid)
        self.afn_bn = nn.BatchNorm1d(afn_hid)
        nn.init.normal_(self.afn.weight, std=0.1)
        nn.init.constant_(self.afn.bias, 0.)
        self.mlp = MLP(afn_hid * nemb, mlp_layers, mlp_hid, dropout)
        if ensemble:
            self.deep_embedding = Embedding(nfeat, nemb)
            self.deep_mlp = MLP(nfield * nemb, deep_layers, deep_hid, dropout)
            self.ensemble_layer = nn.Linear(2, 1)
            nn.init.constant_(self.ensemble_layer.weight, 0.5)
            nn.init.constant_(self.ensemble_layer.bias, 0.)

    def embedding_clip(self):
        with torch.no_grad():
            self.embedding.embedding.weight.abs_().clamp_(min=1e-4)

    def forward(self, x):
        x['value'].clamp_(0.001, 1.)
        self.embedding_clip()
        e = self.embedding(x)                                               # [B, F, E], strictly positive
        logs = self.emb_bn(torch.log(e)).transpose(1, 2)                    # [B, E, F]
        neurons = torch.exp(self.afn(logs)).transpose(1, 2)                 # [B, O, E]
        h = self.dropout(self.afn_bn(neurons).reshape(-1, self.afn_hid * self.nemb))
        y = self.mlp(h)
        if self.ensemble:
            deep = self.deep_mlp(self.deep_embedding(x).reshape(-1, self.nfield * self.nemb))
            y = self.ensemble_layer(torch.cat([y, deep], dim=1))
        return y.squeeze(1)
