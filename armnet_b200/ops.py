"""torch-facing wrappers of the C ABI (include/armnet_b200.h): device pointers in, torch tensors out.

PyTorch is plumbing here (device memory, streams, autograd bookkeeping); the arithmetic of the hot path runs in
libarmnet_b200.so.  Every op raises on CPU tensors -- there is no fallback path.
"""
from typing import Optional, Tuple

import torch

from . import _capi
from ._capi import SOLVER_AUTO, SOLVER_BISECT, check, lib

__all__ = ['fused_prepare', 'transpose2d', 'linear_tf32x3', 'linear_dense', 'linear_gather', 'batch_norm_train', 'clamp_adam', 'partials_to_dense', 'mlp_split_weight', 'mlp_first_linear', 'mlp_tail', 'mlp_pack_tail', 'embed_gather', 'entmax', 'fused_forward', 'fused_backward', 'fused_backward_finish', 'fused_interaction', 'fused_bwd_supported', 'new_error_flag', 'raise_if_bad_ids',
           'SOLVER_AUTO', 'SOLVER_BISECT', 'last_launch_count']


def _need_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError('armnet_b200 ops run on CUDA tensors only (no CPU fallback); got a '
                               f'{t.device} tensor')


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _ids_arg(ids):
    if ids.dtype == torch.int64:
        return 0
    if ids.dtype == torch.int32:
        return 1
    raise TypeError(f'ids must be int64 or int32, got {ids.dtype}')


def _f32c(t, name):
    if t.dtype != torch.float32:
        raise TypeError(f'{name} must be float32, got {t.dtype}')
    return t.contiguous()


def new_error_flag(device):
    """Device word the kernels OR bit 0 into when they meet an id outside [0, nfeat)."""
    return torch.zeros(1, dtype=torch.int32, device=device)


def raise_if_bad_ids(flag):
    """Synchronising check of an error flag; mirrors nn.Embedding's IndexError (layers.py:20)."""
    if int(flag.item()) & 1:
        flag.zero_()
        raise IndexError('index out of range in self (armnet_b200: id outside [0, nfeat))')


def last_launch_count():
    return lib.armnet_last_launch_count()


def _values_inplace(values, clamp=None, clamp_inplace=True):
    """The reference clamps the caller's tensor in place (armnet.py:82). For a contiguous fp32 tensor the kernel
    writes into it directly; any other layout / float dtype the reference accepts (a strided view, float64) is
    clamped in place with torch first and the kernel gets a contiguous fp32 copy."""
    if values.dtype == torch.float32 and values.is_contiguous():
        return values, True
    if not values.dtype.is_floating_point:
        raise TypeError('values must be a floating-point tensor')
    if clamp is not None and clamp_inplace:
        values.clamp_(clamp[0], clamp[1])
    return values.to(torch.float32).contiguous(), False


def embed_gather(ids, values, table, clamp: Optional[Tuple[float, float]] = None, clamp_inplace=True,
                 ld: Optional[int] = None, nemb: Optional[int] = None, err_flag=None):
    """layers.Embedding.forward (layers.py:15-21) [+ the value clamp of armnet.py:82 when clamp=(lo, hi)].
    ids [B,F] int64/int32, values [B,F] f32, table [V, ld] f32 -> [B,F,E] f32 (bit-exact vs the reference)."""
    _need_cuda(ids, values, table)
    ids_c = ids.contiguous()
    values, _ = _values_inplace(values, clamp, clamp_inplace)
    table = _f32c(table, 'table')
    V = table.shape[0]
    ld = table.shape[1] if ld is None else ld
    E = table.shape[1] if nemb is None else nemb
    B, F = ids_c.shape
    out = torch.empty(B, F, E, dtype=torch.float32, device=table.device)
    lo, hi = clamp if clamp is not None else (0.0, 0.0)
    rc = lib.armnet_embed_gather_f32(ids_c.data_ptr(), _ids_arg(ids_c), values.data_ptr(), table.data_ptr(), V, ld,
                                     B, F, E, out.data_ptr(), int(clamp is not None), lo, hi, int(clamp_inplace),
                                     err_flag.data_ptr() if err_flag is not None else None, _stream())
    check(rc, 'armnet_embed_gather_f32')
    return out


def entmax_forward(x, alpha, n_iter=50, solver=SOLVER_AUTO):
    """Row-wise alpha-entmax over the last axis of a contiguous fp32 CUDA tensor (entmax.py:29-68)."""
    _need_cuda(x)
    x = _f32c(x, 'x')
    F = x.shape[-1]
    rows = x.numel() // F if F else 0
    p = torch.empty_like(x)
    check(lib.armnet_entmax_f32(x.data_ptr(), rows, F, float(alpha), solver, n_iter, p.data_ptr(), _stream()),
          'armnet_entmax_f32')
    return p


def entmax_backward(p, dp, alpha):
    """EntmaxBisectFunction.backward (entmax.py:71-80)."""
    _need_cuda(p, dp)
    p = _f32c(p, 'p')
    dp = _f32c(dp, 'dp')
    F = p.shape[-1]
    rows = p.numel() // F if F else 0
    dx = torch.empty_like(p)
    check(lib.armnet_entmax_bwd_f32(p.data_ptr(), dp.data_ptr(), rows, F, float(alpha), dx.data_ptr(), _stream()),
          'armnet_entmax_bwd_f32')
    return dx


class _EntmaxFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, alpha, n_iter, solver):
        p = entmax_forward(x, alpha, n_iter, solver)
        ctx.alpha = alpha
        ctx.save_for_backward(p)
        return p

    @staticmethod
    def backward(ctx, dp):
        p, = ctx.saved_tensors
        return entmax_backward(p, dp.contiguous(), ctx.alpha), None, None, None


def entmax(x, alpha=1.5, dim=-1, n_iter=50, solver=SOLVER_AUTO):
    """entmax_bisect(X, alpha, dim, n_iter) (entmax.py:134-175); alpha == 1 gives softmax."""
    if dim not in (-1, x.dim() - 1):
        return entmax(x.transpose(dim, -1), alpha, -1, n_iter, solver).transpose(dim, -1)
    return _EntmaxFn.apply(x, float(alpha), int(n_iter), int(solver))


def fused_forward(ids, values, table, bilinear_w, query, att_values, alpha, one_head=False, n_iter=50,
                  solver=SOLVER_AUTO, clamp: Optional[Tuple[float, float]] = (0.001, 1.0), clamp_inplace=True,
                  ld: Optional[int] = None, nemb: Optional[int] = None, want_tau=False, want_p=False,
                  want_g=False, want_s=False, err_flag=None, post=None, prepared=None):
    """The fused hot path (armnet.py:82-87 / armnet_1h.py:81-86): returns z [B, K*O, E] and a dict of the optional
    outputs ('tau' [B,K*O,2], 'p' [B,K*O,F], 'g' [B,K*O,F], 's' [B,K*O,E]).
    post=(mean, scale, shift), each [K*O]: eval-mode arm_bn epilogue z <- (z - mean) * scale + shift (armnet.py:89)."""
    _need_cuda(ids, values, table, bilinear_w, query, att_values)
    ids_c = ids.contiguous()
    values, _ = _values_inplace(values, clamp, clamp_inplace)
    table = _f32c(table, 'table')
    W = _f32c(bilinear_w, 'bilinear_w')
    Q = _f32c(query, 'query')
    Vv = _f32c(att_values, 'values')
    B, F = ids_c.shape
    V = table.shape[0]
    ld = table.shape[1] if ld is None else ld
    E = table.shape[1] if nemb is None else nemb
    if one_head:
        D, K, O = W.shape[0], 1, Q.shape[0]
        assert W.shape == (D, E) and Q.shape == (O, D) and Vv.shape == (O, F)
    else:
        K, O, D = Q.shape
        assert W.shape == (K, E, D) and Vv.shape == (K, O, F)
    R = K * O
    dev = table.device
    ws_bytes = lib.armnet_fused_workspace_bytes(F, E, K, O)
    if ws_bytes == 0:
        raise _capi.ArmnetError(f'armnet_fused_fwd_f32: unsupported shape F={F} E={E} (F <= 64, E <= 128)')
    if prepared is not None:
        assert prepared.numel() * 4 == ws_bytes and prepared.is_cuda
    ws = prepared if prepared is not None else torch.empty(ws_bytes // 4, dtype=torch.float32, device=dev)
    z = torch.empty(B, R, E, dtype=torch.float32, device=dev)
    extra = {}
    if want_tau:
        extra['tau'] = torch.empty(B, R, 2, dtype=torch.float32, device=dev)
    if want_p:
        extra['p'] = torch.empty(B, R, F, dtype=torch.float32, device=dev)
    if want_g:
        extra['g'] = torch.empty(B, R, F, dtype=torch.float32, device=dev)
    if want_s:
        extra['s'] = torch.empty(B, R, E, dtype=torch.float32, device=dev)
    lo, hi = clamp if clamp is not None else (0.0, 0.0)
    pm = ps = pb = None
    if post is not None:
        pm, ps, pb = (_f32c(t, 'post') for t in post)
        _need_cuda(pm, ps, pb)
        assert pm.numel() == R and ps.numel() == R and pb.numel() == R

    def ptr(k):
        return extra[k].data_ptr() if k in extra else None

    if prepared is not None:     # serving: the pre-contracted attention parameters are reused (read-only) across batches
        rc = lib.armnet_fused_fwd_prepared_f32(
            ids_c.data_ptr(), _ids_arg(ids_c), values.data_ptr(), table.data_ptr(), V, ld, float(alpha), int(solver),
            int(n_iter), B, F, E, D, K, O, int(clamp is not None), lo, hi, int(clamp_inplace),
            pm.data_ptr() if post is not None else None, ps.data_ptr() if post is not None else None,
            pb.data_ptr() if post is not None else None, z.data_ptr(), ptr('tau'), ptr('p'), ptr('g'), ptr('s'),
            ws.data_ptr(), err_flag.data_ptr() if err_flag is not None else None, _stream())
        check(rc, 'armnet_fused_fwd_prepared_f32')
        return z, extra
    rc = lib.armnet_fused_fwd_f32(
        ids_c.data_ptr(), _ids_arg(ids_c), values.data_ptr(), table.data_ptr(), V, ld, W.data_ptr(), Q.data_ptr(),
        Vv.data_ptr(), int(one_head), float(alpha), int(solver), int(n_iter), B, F, E, D, K, O,
        int(clamp is not None), lo, hi, int(clamp_inplace),
        pm.data_ptr() if post is not None else None, ps.data_ptr() if post is not None else None,
        pb.data_ptr() if post is not None else None, z.data_ptr(), ptr('tau'), ptr('p'), ptr('g'),
        ptr('s'), ws.data_ptr(), err_flag.data_ptr() if err_flag is not None else None, _stream())
    check(rc, 'armnet_fused_fwd_f32')
    return z, extra


def fused_prepare(bilinear_w, query, att_values, alpha, F, one_head=False):
    """armnet_fused_prepare_f32: the pre-contracted attention parameters for fused_forward(..., prepared=ws)."""
    _need_cuda(bilinear_w, query, att_values)
    W, Q, Vv = _f32c(bilinear_w, 'bilinear_w'), _f32c(query, 'query'), _f32c(att_values, 'values')
    if one_head:
        D, E = W.shape
        K, O = 1, Q.shape[0]
    else:
        K, E, D = W.shape
        O = Q.shape[1]
    ws_bytes = lib.armnet_fused_workspace_bytes(F, E, K, O)
    if ws_bytes == 0:
        raise _capi.ArmnetError(f'armnet_fused_prepare_f32: unsupported shape F={F} E={E} (F <= 64, E <= 128)')
    ws = torch.empty(ws_bytes // 4, dtype=torch.float32, device=W.device)
    check(lib.armnet_fused_prepare_f32(W.data_ptr(), Q.data_ptr(), Vv.data_ptr(), int(one_head), float(alpha), F, E, D, K, O,
                                       ws.data_ptr(), _stream()), 'armnet_fused_prepare_f32')
    return ws


def fused_bwd_supported(F, E):
    return bool(lib.armnet_fused_bwd_supported(int(F), int(E)))


def set_tuning(key, value):
    """armnet_set_tuning: process-global experiment switch ('tmem', 'mma', ...; -1 = library default)."""
    check(lib.armnet_set_tuning(key.encode(), int(value)), 'armnet_set_tuning')


class tuning:
    """with ops.tuning(tmem=0, mma=1): ...  -- switches restored to the library defaults (-1 / 0) on exit."""
    _DEFAULTS = {'tmem': -1, 'tmem_rows': -1, 'mma': -1, 'mma_warps': -1, 'force_nw': -1, 'force_look': -1}

    def __init__(self, **kw):
        self.kw = kw

    def __enter__(self):
        for k, v in self.kw.items():
            set_tuning(k, v)
        return self

    def __exit__(self, *exc):
        for k in self.kw:
            set_tuning(k, self._DEFAULTS.get(k, 0))
        return False


def fused_fwd_kernel_kind(F, E, K, O, alpha, solver=0):
    """0: no instance, 1: armnet_fwd_kernel (FP32 pipe), 2: armnet_fwd_mma_kernel (TF32 warp-MMA products),
    3: armnet_fwd_tmem_kernel (logits on tcgen05 / tensor memory)."""
    return int(lib.armnet_fused_fwd_kernel_kind(int(F), int(E), int(K), int(O), float(alpha), int(solver)))


def fused_backward(ids, values, table, bilinear_w, query, att_values, alpha, z, dz, tau, one_head=False,
                   ld: Optional[int] = None, nemb: Optional[int] = None):
    """armnet_fused_bwd_f32: returns (w [B,F,R], dg [B,F,R], dvalues [R,F], dm [R,E]); see include/armnet_b200.h."""
    _need_cuda(ids, values, table, bilinear_w, query, att_values, z, dz, tau)
    ids_c = ids.contiguous()
    values, _ = _values_inplace(values)
    table = _f32c(table, 'table')
    W, Q, Vv = _f32c(bilinear_w, 'bilinear_w'), _f32c(query, 'query'), _f32c(att_values, 'values')
    z, dz, tau = _f32c(z, 'z'), _f32c(dz, 'dz'), _f32c(tau, 'tau')
    B, F = ids_c.shape
    V = table.shape[0]
    ld = table.shape[1] if ld is None else ld
    E = table.shape[1] if nemb is None else nemb
    if one_head:
        D, K, O = W.shape[0], 1, Q.shape[0]
    else:
        K, O, D = Q.shape
    R = K * O
    dev = table.device
    ws = torch.empty(max(lib.armnet_fused_workspace_bytes(F, E, K, O) // 4, 4), dtype=torch.float32, device=dev)
    w = torch.empty(B, F, R, dtype=torch.float32, device=dev)
    dg = torch.empty(B, F, R, dtype=torch.float32, device=dev)
    dvals = torch.zeros(R, F, dtype=torch.float32, device=dev)
    dm = torch.zeros(R, E, dtype=torch.float32, device=dev)
    rc = lib.armnet_fused_bwd_f32(ids_c.data_ptr(), _ids_arg(ids_c), values.data_ptr(), table.data_ptr(), V, ld,
                                  W.data_ptr(), Q.data_ptr(), Vv.data_ptr(), int(one_head), float(alpha), B, F, E, D,
                                  K, O, z.data_ptr(), dz.data_ptr(), tau.data_ptr(), w.data_ptr(), dg.data_ptr(),
                                  dvals.data_ptr(), dm.data_ptr(), ws.data_ptr(), None, _stream())
    check(rc, 'armnet_fused_bwd_f32')
    return w, dg, dvals, dm


def fused_backward_finish(ids, values, V, bilinear_w, query, w, dg, z, dz, dm, one_head=False, dT=None):
    """armnet_fused_bwd_finish_f32: (dT [V,E], dW, dQ) from the partials of fused_backward; see include/armnet_b200.h.
    dT: optional [V,E] tensor to accumulate into (zeroed here when not given)."""
    _need_cuda(ids, values, bilinear_w, query, w, dg, z, dz, dm)
    ids_c = ids.contiguous()
    values = _f32c(values, 'values')
    W, Q = _f32c(bilinear_w.detach(), 'bilinear_w'), _f32c(query.detach(), 'query')
    w, dg, z, dz, dm = (_f32c(t, n) for t, n in ((w, 'w'), (dg, 'dg'), (z, 'z'), (dz, 'dz'), (dm, 'dm')))
    B, F = ids_c.shape
    if one_head:
        D, K, O = W.shape[0], 1, Q.shape[0]
        E = W.shape[1]
    else:
        K, O, D = Q.shape
        E = W.shape[1]
    dev = w.device
    if dT is None:
        dT = torch.zeros(V, E, dtype=torch.float32, device=dev)
    dW, dQ = torch.empty_like(W), torch.empty_like(Q)
    ws = torch.empty(max(lib.armnet_fused_bwd_finish_workspace_bytes(E, K, O) // 4, 4), dtype=torch.float32, device=dev)
    check(lib.armnet_fused_bwd_finish_f32(ids_c.data_ptr(), _ids_arg(ids_c), values.data_ptr(), V, B, F, E, D, K, O,
                                          W.data_ptr(), Q.data_ptr(), int(one_head), w.data_ptr(), dg.data_ptr(),
                                          z.data_ptr(), dz.data_ptr(), dm.data_ptr(), dT.data_ptr(), dW.data_ptr(),
                                          dQ.data_ptr(), ws.data_ptr(), _stream()), 'armnet_fused_bwd_finish_f32')
    return dT, dW, dQ


class _FusedInteractionFn(torch.autograd.Function):
    """z = exp(sum_f entmax(g)_f V_f e_f) with the fused forward and backward kernels; gradients for the embedding
    table (dense, like nn.Embedding(sparse=False)), bilinear_w, query and values. ids / values get none
    (leaf inputs in train.py:104-109)."""

    @staticmethod
    def forward(ctx, table, bilinear_w, query, att_values, ids, values, alpha, one_head, solver):
        z, extra = fused_forward(ids, values, table, bilinear_w, query, att_values, alpha, one_head=one_head,
                                 solver=solver, want_tau=True)
        ctx.save_for_backward(table, bilinear_w, query, att_values, ids, values, z, extra['tau'])
        ctx.alpha, ctx.one_head = alpha, one_head
        return z

    @staticmethod
    def backward(ctx, dz):
        table, W, Q, Vv, ids, values, z, tau = ctx.saved_tensors
        one_head = ctx.one_head
        w, dg, dvals, dm = fused_backward(ids, values, table, W, Q, Vv, ctx.alpha, z, dz.contiguous(), tau,
                                          one_head=one_head)
        # FlatAdam points table.grad at its slice of the flat gradient bucket (zeroed every step): scatter straight into
        # it instead of materialising a second [V,E] tensor that autograd would add to it (64 MB at config 4)
        g = table.grad if getattr(table, '_armnet_flat_grad', False) else None
        inplace = g is not None and g.is_contiguous() and g.dtype == torch.float32 and g.shape == table.shape
        dT, dW, dQ = fused_backward_finish(ids, values, table.shape[0], W, Q, w, dg, z, dz.contiguous(), dm,
                                           one_head=one_head, dT=g if inplace else None)
        return (None if inplace else dT), dW, dQ, dvals.reshape(Vv.shape), None, None, None, None, None


def fused_interaction(table, bilinear_w, query, att_values, ids, values, alpha, one_head=False, solver=SOLVER_AUTO):
    """Differentiable fused hot path (armnet.py:82-87): clamps `values` in place, returns z [B, K*O, E]."""
    return _FusedInteractionFn.apply(table, bilinear_w, query, att_values, ids, values, float(alpha), bool(one_head),
                                     int(solver))


# ---------------------------------------------------------------------------------------------- trailing MLP (eval)
def mlp_split_weight(w):
    """nn.Linear weight [N,K] -> (w_hi, w_lo): TF32-rounded part and exact remainder (armnet_mlp_split_weight_f32)."""
    _need_cuda(w)
    w = _f32c(w.detach(), 'weight')
    hi, lo = torch.empty_like(w), torch.empty_like(w)
    check(lib.armnet_mlp_split_weight_f32(w.data_ptr(), w.numel(), hi.data_ptr(), lo.data_ptr(), _stream()),
          'armnet_mlp_split_weight_f32')
    return hi, lo


def mlp_first_linear(x, w_hi, w_lo, splits=None):
    """x [B,K] . w^T on the tcgen05 tensor cores (3xTF32, fp32 accumulate) -> split-K partial sums in blocks of 32
    samples, [splits, ceil(B/32), N, 32] (layers.py:73 without bias; partials_to_dense() gives F.linear(x, w))."""
    _need_cuda(x, w_hi, w_lo)
    x = _f32c(x, 'x')
    B, K = x.shape
    N = w_hi.shape[0]
    assert w_hi.shape == (N, K) and w_lo.shape == (N, K) and w_hi.is_contiguous() and w_lo.is_contiguous()
    if splits is None:
        splits = lib.armnet_mlp_linear_splits(B, K, N)
    out = torch.empty(splits, (B + 31) // 32, N, 32, dtype=torch.float32, device=x.device)
    check(lib.armnet_mlp_linear_tf32x3(x.data_ptr(), B, K, w_hi.data_ptr(), w_lo.data_ptr(), N, splits,
                                       out.data_ptr(), _stream()), 'armnet_mlp_linear_tf32x3')
    return out


def mlp_pack_tail(linears, bns):
    """Pack the eval-mode parameters every layer after the first GEMM needs (include/armnet_b200.h,
    armnet_mlp_tail_f32): linears = [L1, ..., Ln, Lout], bns = [BN1, ..., BNn]."""
    with torch.no_grad():
        parts = []
        for i, (lin, bn) in enumerate(zip(linears[:-1], bns)):
            a = bn.weight * torch.rsqrt(bn.running_var + bn.eps)
            c = (lin.bias - bn.running_mean) * a + bn.bias
            if i > 0:
                parts.append(lin.weight.t().contiguous().reshape(-1))
            parts += [a, c]
        parts += [linears[-1].weight.reshape(-1), linears[-1].bias]
        return torch.cat([p.reshape(-1).float() for p in parts]).contiguous()


def mlp_pack_layers(linears, bns):
    """Per-layer eval-mode parameters of the tensor-core hidden-layer path (armnet_mlp_hidden_tc_f32):
    [(a_i, c_i)] for every hidden layer (a = bn.weight / sqrt(running_var + eps), c = (bias - running_mean) a + bn.bias),
    [(w_hi, w_lo)] for hidden layers 2..n, and (wf, bf) of the output Linear."""
    with torch.no_grad():
        ac = []
        for lin, bn in zip(linears[:-1], bns):
            a = (bn.weight * torch.rsqrt(bn.running_var + bn.eps)).float().contiguous()
            c = ((lin.bias - bn.running_mean) * a + bn.bias).float().contiguous()
            ac.append((a, c))
        splits = [mlp_split_weight(lin.weight) for lin in linears[1:-1]]
        return ac, splits, (linears[-1].weight.detach().float().contiguous(), linears[-1].bias.detach().float().contiguous())


def mlp_hidden_tc(partials, B, ac, splits, out):
    """Hidden layers 2..n and the output Linear on tcgen05: one armnet_mlp_hidden_tc_f32 launch per hidden layer
    (partials = split-K sums of layer 1, [S, ceil(B/32), H, 32]) -> y [B, noutput]."""
    _need_cuda(partials)
    wf, bf = out
    NO = wf.shape[0]
    MB = (B + 31) // 32
    for l, (hi, lo) in enumerate(splits):
        S, mb, H_in, _ = partials.shape
        assert mb == MB and hi.shape[1] == H_in
        H_out = hi.shape[0]
        a_in, c_in = ac[l]
        last = l == len(splits) - 1
        if last:
            y = torch.empty(B, NO, dtype=torch.float32, device=partials.device)
            a_out, c_out = ac[l + 1]
            check(lib.armnet_mlp_hidden_tc_f32(partials.data_ptr(), S, B, H_in, a_in.data_ptr(), c_in.data_ptr(),
                                               hi.data_ptr(), lo.data_ptr(), H_out, a_out.data_ptr(), c_out.data_ptr(),
                                               wf.data_ptr(), bf.data_ptr(), NO, y.data_ptr(), None, _stream()),
                  'armnet_mlp_hidden_tc_f32')
            return y
        nxt = torch.empty(1, MB, H_out, 32, dtype=torch.float32, device=partials.device)
        check(lib.armnet_mlp_hidden_tc_f32(partials.data_ptr(), S, B, H_in, a_in.data_ptr(), c_in.data_ptr(), hi.data_ptr(),
                                           lo.data_ptr(), H_out, None, None, None, None, 0, None, nxt.data_ptr(),
                                           _stream()), 'armnet_mlp_hidden_tc_f32')
        partials = nxt
    raise AssertionError('mlp_hidden_tc needs at least one hidden layer after the first')


def partials_to_dense(partials, B):
    """[splits, MB, N, 32] partial sums -> the dense product [B, N] (test / validation helper)."""
    S, MB, N, _ = partials.shape
    return partials.sum(0).permute(0, 2, 1).reshape(MB * 32, N)[:B]


def mlp_tail(partials, packed, n_rest, noutput, B):
    """Everything after the first Linear's GEMM in one launch (armnet_mlp_tail_f32) -> y [B, noutput]."""
    _need_cuda(partials, packed)
    S, MB, H, _ = partials.shape
    assert MB == (B + 31) // 32
    assert packed.numel() == lib.armnet_mlp_tail_packed_floats(H, n_rest, noutput)
    y = torch.empty(B, noutput, dtype=torch.float32, device=partials.device)
    check(lib.armnet_mlp_tail_f32(partials.data_ptr(), S, B, H, n_rest, noutput, packed.data_ptr(), y.data_ptr(),
                                  _stream()), 'armnet_mlp_tail_f32')
    return y


def linear_dense(x, w_hi, w_lo, bias=None):
    """armnet_linear_tf32x3_dense: y [B,N] = x [B,K] . w^T (+ bias) on tcgen05 (3xTF32), row-major output."""
    _need_cuda(x, w_hi, w_lo)
    x = _f32c(x, 'x')
    B, K = x.shape
    N = w_hi.shape[0]
    assert w_hi.shape == (N, K) and w_lo.shape == (N, K) and w_hi.is_contiguous() and w_lo.is_contiguous()
    y = torch.empty(B, N, dtype=torch.float32, device=x.device)
    check(lib.armnet_linear_tf32x3_dense(x.data_ptr(), B, K, w_hi.data_ptr(), w_lo.data_ptr(), N,
                                         bias.data_ptr() if bias is not None else None, y.data_ptr(), _stream()),
          'armnet_linear_tf32x3_dense')
    return y


def transpose2d(x):
    """armnet_transpose_f32: contiguous [cols, rows] copy of a contiguous fp32 [rows, cols] CUDA tensor."""
    _need_cuda(x)
    x = _f32c(x, 'x')
    rows, cols = x.shape
    out = torch.empty(cols, rows, dtype=torch.float32, device=x.device)
    check(lib.armnet_transpose_f32(x.data_ptr(), rows, cols, out.data_ptr(), _stream()), 'armnet_transpose_f32')
    return out


class _LinearTF32x3Fn(torch.autograd.Function):
    """F.linear(x, W, b) for training with all three GEMMs on the tcgen05 3xTF32 kernel (fp32 parity):
         y  [B,N] = x . W^T + b        split-K partial sums (K = ninput is long), summed here
         dx [B,K] = dy . W             dense output (reduction over N = nhid is short)
         dW [N,K] = dy^T . x           computed as (x^T . dy)^T: the big operand is the one converted in-kernel
    Needs K % 4 == 0, N % 4 == 0, B % 4 == 0 (TMA row pitches)."""

    @staticmethod
    def forward(ctx, x, W, b):
        x = _f32c(x, 'x')
        hi, lo = mlp_split_weight(W)
        y = partials_to_dense(mlp_first_linear(x, hi, lo), x.shape[0])
        if b is not None:
            y = y + b
        ctx.save_for_backward(x, W)
        ctx.has_bias = b is not None
        return y

    @staticmethod
    def backward(ctx, dy):
        x, W = ctx.saved_tensors
        dy = _f32c(dy, 'dy')
        B = x.shape[0]
        dx = dW = db = None
        if ctx.needs_input_grad[0]:
            wt_hi, wt_lo = mlp_split_weight(transpose2d(W.detach()))              # [K, N]
            dx = linear_dense(dy, wt_hi, wt_lo)
        if ctx.needs_input_grad[1]:
            dyt_hi, dyt_lo = mlp_split_weight(transpose2d(dy))                    # [N, B]
            xt = transpose2d(x)                                                   # [K, B]
            dWt = partials_to_dense(mlp_first_linear(xt, dyt_hi, dyt_lo), xt.shape[0])   # [K, N]
            dW = dWt.t()
        if ctx.has_bias and ctx.needs_input_grad[2]:
            db = dy.sum(0)
        return dx, dW, db


def linear_tf32x3(x, weight, bias=None):
    """Differentiable F.linear on the tensor cores with fp32 parity (see _LinearTF32x3Fn)."""
    return _LinearTF32x3Fn.apply(x, weight, bias)


def clamp_adam(param, grad, exp_avg, exp_avg_sq, grad_scale, clamp, lr, beta1, beta2, eps, step):
    """armnet_clamp_adam_f32: in-place dense Adam over flat fp32 buffers, fused with gradient scaling and clamping."""
    _need_cuda(param, grad, exp_avg, exp_avg_sq)
    n = param.numel()
    assert grad.numel() == n and exp_avg.numel() == n and exp_avg_sq.numel() == n
    for t in (param, grad, exp_avg, exp_avg_sq):
        assert t.dtype == torch.float32 and t.is_contiguous()
    check(lib.armnet_clamp_adam_f32(param.data_ptr(), grad.data_ptr(), exp_avg.data_ptr(), exp_avg_sq.data_ptr(), n,
                                    float(grad_scale), float(clamp if clamp is not None else 0.0), float(lr), float(beta1),
                                    float(beta2), float(eps), int(step), _stream()), 'armnet_clamp_adam_f32')


# ---------------------------------------------------------------------------------------------- train-mode arm_bn
class _BatchNormTrainFn(torch.autograd.Function):
    """F.batch_norm(x [B,C,L], training=True) with the CUDA kernels of csrc/bn.cu (armnet.py:89 in train mode)."""

    @staticmethod
    def forward(ctx, x, weight, bias, running_mean, running_var, momentum, eps):
        _need_cuda(x)
        x = _f32c(x, 'x')
        B, C, L = x.shape
        ws = torch.empty(lib.armnet_bn_workspace_floats(B, C, L), dtype=torch.float32, device=x.device)
        out = torch.empty_like(x)
        mean = torch.empty(C, dtype=torch.float32, device=x.device)
        invstd = torch.empty(C, dtype=torch.float32, device=x.device)
        rc = lib.armnet_bn_train_fwd_f32(
            x.data_ptr(), B, C, L, weight.data_ptr() if weight is not None else None,
            bias.data_ptr() if bias is not None else None,
            running_mean.data_ptr() if running_mean is not None else None,
            running_var.data_ptr() if running_var is not None else None, float(momentum), float(eps), out.data_ptr(),
            mean.data_ptr(), invstd.data_ptr(), ws.data_ptr(), _stream())
        if rc == -2:
            raise ValueError(f'Expected more than 1 value per channel when training, got input size {tuple(x.shape)}')
        check(rc, 'armnet_bn_train_fwd_f32')
        ctx.save_for_backward(x, weight, mean, invstd)
        ctx.has_bias = bias is not None
        # running statistics were written through raw pointers: bump their versions so caches keyed on them rebuild
        written = [t for t in (running_mean, running_var) if t is not None]
        if written:
            torch.autograd.graph.increment_version(written)
        return out

    @staticmethod
    def backward(ctx, dy):
        x, weight, mean, invstd = ctx.saved_tensors
        dy = _f32c(dy, 'dy')
        B, C, L = x.shape
        ws = torch.empty(lib.armnet_bn_workspace_floats(B, C, L), dtype=torch.float32, device=x.device)
        dx = torch.empty_like(x)
        dw = torch.empty(C, dtype=torch.float32, device=x.device)
        db = torch.empty(C, dtype=torch.float32, device=x.device)
        check(lib.armnet_bn_train_bwd_f32(x.data_ptr(), dy.data_ptr(), B, C, L,
                                          weight.data_ptr() if weight is not None else None, mean.data_ptr(),
                                          invstd.data_ptr(), dx.data_ptr(), dw.data_ptr(), db.data_ptr(), ws.data_ptr(),
                                          _stream()), 'armnet_bn_train_bwd_f32')
        return dx, (dw if weight is not None else None), (db if ctx.has_bias else None), None, None, None, None


def batch_norm_train(x, bn):
    """Train-mode forward of an nn.BatchNorm1d module `bn` on x [B,C,L] (statistics over B and L), including the
    running-statistics update and num_batches_tracked, through armnet_bn_train_fwd_f32 / _bwd_f32."""
    momentum = bn.momentum
    if bn.track_running_stats and bn.num_batches_tracked is not None:
        bn.num_batches_tracked.add_(1)
        if momentum is None:
            momentum = 1.0 / float(bn.num_batches_tracked)
    rm, rv = (bn.running_mean, bn.running_var) if bn.track_running_stats else (None, None)
    return _BatchNormTrainFn.apply(x, bn.weight, bn.bias, rm, rv, 0.0 if momentum is None else momentum, bn.eps)


def linear_gather(ids, values, weight, bias=None, err_flag=None):
    """layers.Linear.forward (layers.py:31-37): y[b] = sum_f weight[ids[b,f]] * values[b,f] + bias. weight [V,1] or [V]."""
    _need_cuda(ids, values, weight)
    ids_c = ids.contiguous()
    values = _f32c(values, 'values')
    w = _f32c(weight.detach(), 'weight').reshape(-1)
    B, F = ids_c.shape
    y = torch.empty(B, dtype=torch.float32, device=w.device)
    check(lib.armnet_linear_gather_f32(ids_c.data_ptr(), _ids_arg(ids_c), values.data_ptr(), w.data_ptr(), w.numel(), B, F,
                                       bias.data_ptr() if bias is not None else None, y.data_ptr(),
                                       err_flag.data_ptr() if err_flag is not None else None, _stream()),
          'armnet_linear_gather_f32')
    return y
