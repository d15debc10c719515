"""CPU oracle for the ARM-Net forward hot path  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this module; armnet_b200/ never does (the product path is CUDA-only and fails
loudly without its extension).

What it is: a functional, stage-by-stage restatement (torch CPU fp32 ops, same op order)
of the reference's algorithm for
    embedding lookup  ->  attention logits  ->  alpha-entmax gates  ->  gates*values
    ->  exponential-neuron interaction  ->  BatchNorm1d  ->  MLP  [-> DNN ensemble]
Every function cites the reference file:line (nusdbsystem/ARM-Net @ 7aeb3a4) it follows.
All arithmetic on that path is executed by PyTorch ATen (reference pins torch==1.6.0,
requirements.txt:1; this image runs 2.11.0), so the oracle uses the same ATen calls.

Parity PINNED: the reference ships no tests / golden vectors, so the
oracle is pinned against outputs of the reference itself, generated in the dev container
by tests/golden/make_golden.py and committed under tests/golden/*.npz;
tests/test_oracle_golden.py asserts the oracle reproduces every stage of every fixture
bit-for-bit (same ATen kernels, same op order).
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import torch
import torch.nn.functional as F_

Tensor = torch.Tensor


# --------------------------------------------------------------------------- stages

def clamp_values_(values: Tensor) -> Tensor:
    """armnet.py:82 / armnet_1h.py:81 -- in-place clip of the caller's tensor to [1e-3, 1]."""
    return values.clamp_(0.001, 1.)


def embed(ids: Tensor, values: Tensor, table: Tensor) -> Tensor:
    """layers.py:15-21 -- e[b,f,:] = table[ids[b,f],:] * values[b,f].  [B,F] -> [B,F,E]."""
    rows = F_.embedding(ids, table)
    return rows * values.unsqueeze(2)


def attn_logits_mh(e: Tensor, bilinear_w: Tensor, query: Tensor) -> Tensor:
    """armnet.py:15,31-34 -- g[b,k,o,f] = (sum_xy e[b,f,x] W[k,x,y] Q[k,o,y]) * d_k**-0.5."""
    d_k = bilinear_w.shape[2]
    return torch.einsum('bfx,kxy,koy->bkof', e, bilinear_w, query) * (d_k ** -0.5)


def attn_logits_1h(e: Tensor, lin_w: Tensor, query: Tensor) -> Tensor:
    """armnet_1h.py:14,30-32 -- keys = e @ lin_w.T (nn.Linear, no bias); g[b,o,f] = keys.Q * d_k**-0.5."""
    d_k = lin_w.shape[0]
    keys = F_.linear(e, lin_w)
    return torch.einsum('bfe,oe->bof', keys, query) * (d_k ** -0.5)


def entmax_bisect(g: Tensor, alpha: float, n_iter: int = 50) -> Tensor:
    """entmax.py:29-68 -- alpha-entmax along the last axis by n_iter bisection steps on tau.

    Kept quirks: alpha is materialised as a tensor of g's dtype, so alpha-1
    and 1/(alpha-1) are fp32 roundings; f_lo is never refreshed; the returned p is the one
    evaluated at the LAST midpoint tau_m (not at tau_lo) and is renormalised by its sum.
    """
    a = torch.tensor(alpha, dtype=g.dtype, device=g.device).expand(*g.shape[:-1], 1)   # entmax.py:31-36
    am1 = a - 1
    inv = 1 / am1
    d = g.shape[-1]

    def p_of(t):                                   # entmax.py:24-26
        return torch.clamp(t, min=0) ** inv

    x = g * am1                                    # entmax.py:42
    mx, _ = x.max(dim=-1, keepdim=True)            # entmax.py:44
    tau_lo = mx - 1 ** am1                         # entmax.py:46   (_gp(1, alpha) == 1)
    tau_hi = mx - (1 / d) ** am1                   # entmax.py:47
    f_lo = p_of(x - tau_lo).sum(-1) - 1            # entmax.py:49
    dm = tau_hi - tau_lo                           # entmax.py:51
    p_m = None
    for _ in range(n_iter):                        # entmax.py:53-61
        dm = dm / 2
        tau_m = tau_lo + dm
        p_m = p_of(x - tau_m)
        f_m = p_m.sum(-1) - 1
        keep = (f_m * f_lo >= 0).unsqueeze(-1)
        tau_lo = torch.where(keep, tau_m, tau_lo)
    return p_m / p_m.sum(dim=-1).unsqueeze(-1)     # entmax.py:63-64


def gates(g: Tensor, alpha: float, n_iter: int = 50) -> Tensor:
    """armnet.py:12-13,35 -- softmax when alpha == 1., otherwise entmax bisection."""
    if alpha == 1.:
        return torch.softmax(g, dim=-1)
    return entmax_bisect(g, alpha, n_iter)


def entmax_backward(p: Tensor, dp: Tensor, alpha: float) -> Tensor:
    """entmax.py:71-80 -- dX = dY*gppr - (sum dY*gppr / sum gppr)*gppr, gppr = p^(2-alpha)[p>0]."""
    a = torch.tensor(alpha, dtype=p.dtype)
    gppr = torch.where(p > 0, p ** (2 - a), p.new_zeros(1))
    dx = dp * gppr
    q = (dx.sum(-1) / gppr.sum(-1)).unsqueeze(-1)
    return dx - q * gppr


def interaction_mh(e: Tensor, w: Tensor) -> Tensor:
    """armnet.py:87 -- s[b,k,o,:] = sum_f w[b,k,o,f] e[b,f,:]  (the log-space product)."""
    return torch.einsum('bfe,bkof->bkoe', e, w)


def interaction_1h(e: Tensor, w: Tensor) -> Tensor:
    """armnet_1h.py:86."""
    return torch.einsum('bfe,bof->boe', e, w)


def batchnorm1d(x: Tensor, st: Dict[str, Tensor], prefix: str, training: bool) -> Tensor:
    """nn.BatchNorm1d defaults (eps 1e-5, momentum 0.1) as used at armnet.py:67,89 and layers.py:75.
    Functional: running statistics in `st` are NOT updated (eval-mode parity and train-mode forward only)."""
    return F_.batch_norm(x, st[prefix + 'running_mean'].clone(), st[prefix + 'running_var'].clone(),
                         st[prefix + 'weight'], st[prefix + 'bias'], training, 0.1, 1e-5)


def mlp(x: Tensor, st: Dict[str, Tensor], prefix: str, training: bool) -> Tensor:
    """layers.py:68-88 -- [Linear -> BatchNorm1d -> ReLU -> Dropout(p=0)] x n, then Linear."""
    idx = sorted({int(k[len(prefix):].split('.')[0]) for k in st if k.startswith(prefix)})
    last = idx[-1]
    i = 0
    while i < last:
        x = F_.linear(x, st[f'{prefix}{i}.weight'], st[f'{prefix}{i}.bias'])
        x = batchnorm1d(x, st, f'{prefix}{i + 1}.', training)
        x = torch.relu(x)
        i += 4
    return F_.linear(x, st[f'{prefix}{last}.weight'], st[f'{prefix}{last}.bias'])


# --------------------------------------------------------------------------- whole path

def hot_path(st: Dict[str, Tensor], alpha: float, ids: Tensor, values: Tensor,
             n_iter: int = 50) -> Dict[str, Tensor]:
    """The path BASELINE.json names: armnet.py:82-87 / armnet_1h.py:81-86 up to z = exp(s).
    `values` is clamped IN PLACE like the reference does. Returns every stage."""
    one_head = 'attn_layer.bilinear_w.weight' in st
    clamp_values_(values)
    e = embed(ids, values, st['embedding.embedding.weight'])
    if one_head:
        g = attn_logits_1h(e, st['attn_layer.bilinear_w.weight'], st['attn_layer.query'])
    else:
        g = attn_logits_mh(e, st['attn_layer.bilinear_w'], st['attn_layer.query'])
    p = gates(g, alpha, n_iter)
    if one_head:
        w = torch.einsum('bof,of->bof', p, st['attn_layer.values'])        # armnet_1h.py:34
        s = interaction_1h(e, w)
    else:
        w = torch.einsum('bkof,kof->bkof', p, st['attn_layer.values'])     # armnet.py:36
        s = interaction_mh(e, w)
    z = torch.exp(s)                                                       # armnet.py:86
    B = ids.shape[0]
    z = z.reshape(B, -1, z.shape[-1])                                      # armnet.py:88 'b k o e -> b (k o) e'
    return {'e': e, 'g': g, 'p': p, 'w': w, 's': s, 'z': z}


def forward(st: Dict[str, Tensor], alpha: float, ids: Tensor, values: Tensor,
            training: bool = False, n_iter: int = 50) -> Dict[str, Tensor]:
    """Full ARMNetModel.forward (armnet.py:77-101 / armnet_1h.py:76-98): hot path + arm_bn + MLP
    [+ ensemble DNN]. Returns the stage dict with 'y' added."""
    out = hot_path(st, alpha, ids, values, n_iter)
    B = ids.shape[0]
    x = batchnorm1d(out['z'], st, 'arm_bn.', training)                     # armnet.py:89
    x = x.reshape(B, -1)                                                   # armnet.py:90
    y = mlp(x, st, 'mlp.mlp.', training)                                   # armnet.py:92
    if 'ensemble_layer.weight' in st:                                      # armnet.py:93-99
        xd = embed(ids, values, st['deep_embedding.embedding.weight']).reshape(B, -1)
        yd = mlp(xd, st, 'deep_mlp.mlp.', training)
        y = F_.linear(torch.cat([y, yd], dim=1), st['ensemble_layer.weight'], st['ensemble_layer.bias'])
    out['y'] = y.squeeze()                                                 # armnet.py:101
    return out


# --------------------------------------------------------------------------- helpers for tests / bench

def reference_init_state(model: str, nfield: int, nfeat: int, nemb: int, nhead: int, nhid: int,
                         d_k: Optional[int] = None, mlp_nlayer: int = 2, mlp_nhid: int = 256,
                         seed: int = 2025) -> Dict[str, Tensor]:
    """Parameters drawn with the reference's init distributions (layers.py:12-13 Xavier-uniform table;
    armnet.py:21-24 Xavier-uniform gain 1.414 attention params; nn.Linear / BatchNorm1d defaults).
    Used for synthetic benchmark workloads where the reference itself is not importable (GPU box);
    it is NOT bit-identical to constructing the reference module under the same seed."""
    gen = torch.Generator().manual_seed(seed)
    d_k = nemb if d_k is None else d_k

    def xavier(*shape, gain=1.0):
        t = torch.empty(*shape)
        if t.dim() == 2:
            fan_in, fan_out = shape[1], shape[0]
        else:
            rf = math.prod(shape[2:])
            fan_in, fan_out = shape[1] * rf, shape[0] * rf
        a = gain * math.sqrt(6.0 / (fan_in + fan_out))
        return t.uniform_(-a, a, generator=gen)

    def linear(out_f, in_f, prefix, st):
        bound = 1.0 / math.sqrt(in_f)
        st[prefix + 'weight'] = torch.empty(out_f, in_f).uniform_(-bound, bound, generator=gen)
        st[prefix + 'bias'] = torch.empty(out_f).uniform_(-bound, bound, generator=gen)

    def bn(n, prefix, st):
        st[prefix + 'weight'] = torch.ones(n)
        st[prefix + 'bias'] = torch.zeros(n)
        st[prefix + 'running_mean'] = torch.zeros(n)
        st[prefix + 'running_var'] = torch.ones(n)

    st: Dict[str, Tensor] = {'embedding.embedding.weight': xavier(nfeat, nemb)}
    if model == 'armnet_1h':
        bound = 1.0 / math.sqrt(nemb)
        st['attn_layer.bilinear_w.weight'] = torch.empty(d_k, nemb).uniform_(-bound, bound, generator=gen)
        st['attn_layer.query'] = xavier(nhid, d_k, gain=1.414)
        st['attn_layer.values'] = xavier(nhid, nfield, gain=1.414)
        rows = nhid
    else:
        st['attn_layer.bilinear_w'] = xavier(nhead, nemb, d_k, gain=1.414)
        st['attn_layer.query'] = xavier(nhead, nhid, d_k, gain=1.414)
        st['attn_layer.values'] = xavier(nhead, nhid, nfield, gain=1.414)
        rows = nhead * nhid
    bn(rows, 'arm_bn.', st)
    nin = rows * nemb
    for i in range(mlp_nlayer):
        linear(mlp_nhid, nin, f'mlp.mlp.{4 * i}.', st)
        bn(mlp_nhid, f'mlp.mlp.{4 * i + 1}.', st)
        nin = mlp_nhid
    linear(1, nin, f'mlp.mlp.{4 * mlp_nlayer}.', st)
    return st
