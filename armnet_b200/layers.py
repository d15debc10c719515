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
    """layers.MLP (layers.py:68-88): [Linear -> BatchNorm1d -> ReLU -> Dropout] x nlayers -> Linear(., noutput),
    identical state_dict keys ('mlp.<i>.weight', ...).

    Eval mode without autograd on CUDA: the first Linear -- the one true GEMM of the path -- runs on the tcgen05
    tensor cores (3xTF32 split, fp32 parity; armnet_mlp_linear_tf32x3); the layers after it run on tcgen05 as well
    (armnet_mlp_hidden_tc_f32: one launch per hidden layer, 128 samples per CTA, the last one fused with the output
    Linear) when nhid <= 256, noutput <= 4 and the batch has >= 512 samples, else in one CUDA-core kernel
    (armnet_mlp_tail_f32). Training, or shapes TMA cannot address (ninput % 4 != 0), use the stock torch modules
    (cuBLAS fp32)."""

    def __init__(self, ninput, nlayers, nhid, dropout, noutput=1):
        super().__init__()
        self.ninput, self.nlayers, self.noutput = ninput, nlayers, noutput
        mods = []
        for _ in range(nlayers):
            mods += [nn.Linear(ninput, nhid), nn.BatchNorm1d(nhid), nn.ReLU(), nn.Dropout(p=dropout)]
            ninput = nhid
        if nlayers == 0:
            nhid = ninput
        mods.append(nn.Linear(nhid, noutput))
        self.mlp = nn.Sequential(*mods)
        self.nhid = nhid
        self.tensor_core = True     # host-side knobs (not in state_dict)
        self.train_tensor_core = True   # training: first Linear's three GEMMs on tcgen05 (large batches only)
        self.hidden_tensor_core = True  # eval: hidden layers 2..n + output Linear on tcgen05 (armnet_mlp_hidden_tc_f32)
        self.hidden_tensor_core_min_batch = 512    # below it the CUDA-core tail kernel (one CTA per 32 samples) wins
        self.first_splits = None        # split-K slices of the first Linear's GEMM (None: armnet_mlp_linear_splits)
        self._cache_key = None
        self._cache = None

    def _fast_ok(self, x):
        return (self.tensor_core and not self.training and not torch.is_grad_enabled() and x.is_cuda
                and x.dtype == torch.float32 and x.dim() == 2 and self.nlayers >= 1 and self.ninput % 4 == 0
                and self.nhid % 4 == 0 and self.nhid <= 512
                # limits of armnet_mlp_tail_f32 (csrc/mlp.cu): at most 64 outputs, a non-empty batch
                and self.noutput <= 64 and x.shape[0] > 0)

    def _prepared(self):
        """(w_hi, w_lo, packed tail parameters), rebuilt when any parameter / buffer changed."""
        tensors = list(self.mlp.parameters()) + [b for b in self.mlp.buffers() if b.dtype.is_floating_point]
        key = tuple((t.data_ptr(), t._version) for t in tensors)
        if key != self._cache_key:
            linears = [m for m in self.mlp if isinstance(m, nn.Linear)]
            bns = [m for m in self.mlp if isinstance(m, nn.BatchNorm1d)]
            w_hi, w_lo = ops.mlp_split_weight(linears[0].weight)
            self._cache = (w_hi, w_lo, ops.mlp_pack_tail(linears, bns), ops.mlp_pack_layers(linears, bns))
            self._cache_key = key
        return self._cache

    def forward(self, x):
        if self._fast_ok(x):
            w_hi, w_lo, packed, (ac, splits, last) = self._prepared()
            B = x.shape[0]
            partials = ops.mlp_first_linear(x, w_hi, w_lo, splits=self.first_splits)
            if (self.hidden_tensor_core and self.nlayers >= 2 and B >= self.hidden_tensor_core_min_batch
                    and self.nhid <= 256 and self.noutput <= 4 and partials.shape[0] <= 4):
                # every further hidden Linear on tcgen05 too, fused with the activations around it and the output Linear
                return ops.mlp_hidden_tc(partials, B, ac, splits, last)
            return ops.mlp_tail(partials, packed, self.nlayers - 1, self.noutput, B)
        if (self.tensor_core and self.train_tensor_core and torch.is_grad_enabled() and x.is_cuda and x.dim() == 2
                and x.dtype == torch.float32 and self.nlayers >= 1 and self.ninput % 4 == 0 and self.nhid % 4 == 0
                and x.shape[0] % 4 == 0 and x.shape[0] >= 512 and self.ninput >= 1024):
            # training at scale: the wide first Linear (forward, dX, dW) on the tensor cores, the rest stock torch
            h = ops.linear_tf32x3(x, self.mlp[0].weight, self.mlp[0].bias)
            for mod in list(self.mlp)[1:]:
                h = mod(h)
            return h
        return self.mlp(x)
