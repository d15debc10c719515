"""Design experiment (not shipped, not imported by the product): numpy fp32 emulation of the threshold solver of
csrc/entmax_rows.cuh (Hoelder-bound pre-solve without MUFU + Newton on the q-norm + fused last sweep) against the
reference's 50-step bisection (utils/entmax.py:29-68, through the oracle). Prints max |p - p_ref| and the number of
2-MUFU-per-element evaluations per row / per 64-row warp. Run here: python tools/entmax_holder_proto.py"""
import sys
import numpy as np
import torch
sys.path.insert(0, '/root/repo')
from oracle import armnet_oracle as oracle

f32 = np.float32
LN2 = f32(0.6931471805599453)


def lg2(x):
    with np.errstate(divide='ignore'):
        return np.log2(x, dtype=f32)


def ex2(x):
    return np.exp2(x, dtype=f32)


def solve(g, alpha, holder_it=3, fuse_thr=3e-3, tol_fused=5e-7, tol_plain=2e-5, warp=64):
    """Returns p [N,F], evals per row (warp-uniform control flow emulated over groups of `warp` rows)."""
    N, F = g.shape
    a = f32(alpha)
    am1 = f32(a - f32(1))
    q = f32(f32(1) / am1)
    qm1 = f32(q - f32(1))
    th = f32(f32(2) - q)
    cF = f32(float(F) ** (-(float(am1))))
    X = (g * am1).astype(f32)
    P = np.zeros_like(X)
    evals = np.zeros(N, int)
    uni_k = f32(qm1 / (f32(2) * cF))
    uni_var = f32((0.2 * cF) ** 2 / F)
    for w0 in range(0, N, warp):
        x = X[w0:w0 + warp]
        mx = x.max(1)
        mean = (x.sum(1, dtype=f32) / f32(F)).astype(f32)
        n_ev = 0
        uni = bool(np.all(mx - mean <= f32(0.2) * cF))
        if uni:
            d = x - mx[:, None]
            sd2 = (d * d).sum(1, dtype=f32)
            md = mean - mx
            var = np.maximum(sd2 / f32(F) - md * md, 0).astype(f32)
            tau = (mean - cF + uni_k * var).astype(f32)
            uni = bool(np.all(var <= uni_var))
        if uni:
            prev_small = True
        else:
            tau = np.maximum(mx - f32(1), mean - cF).astype(f32)
            hi = (mx - cF).astype(f32)
            for it in range(holder_it if (q > 1 and q < 2) else 0):
                u = np.maximum(x - tau[:, None], 0).astype(f32)
                A = np.maximum(u.sum(1, dtype=f32), f32(1e-30))
                Bq = np.maximum((u * u).sum(1, dtype=f32), f32(1e-30))
                c = (u > 0).sum(1).astype(f32)
                lphi = LN2 * (th * lg2(A) + (f32(1) - th) * lg2(Bq))
                dl = -(th * c / A + (f32(1) - th) * f32(2) * A / Bq)
                tau = np.minimum(tau - lphi / dl, hi).astype(f32)
            prev_small = False
        done = False
        for it in range(12):
            fused = prev_small
            u = np.maximum(x - tau[:, None], 0).astype(f32)
            gq = np.where(u > 0, ex2(qm1 * lg2(u)), 0).astype(f32)
            S1 = gq.sum(1, dtype=f32)
            S = (gq * u).sum(1, dtype=f32)
            n_ev += 1
            Nn = ex2(am1 * lg2(S))
            with np.errstate(invalid='ignore', divide='ignore'):
                d = ((Nn - f32(1)) * S / (Nn * S1)).astype(f32)
            d = np.where((S1 > 0) & (np.abs(d) < 1e30), d, 0).astype(f32)
            rel = f32(2.4e-7) * np.abs(tau)
            if fused and np.all(np.abs(d) <= np.maximum(tol_fused, rel)):
                p = gq * u
                done = True
                break
            tau = (tau + d).astype(f32)
            if not fused and np.all(np.abs(d) <= np.maximum(tol_plain, rel)):
                break
            prev_small = bool(np.all(np.abs(d) <= fuse_thr))
        if not done:
            u = np.maximum(x - tau[:, None], 0).astype(f32)
            p = np.where(u > 0, ex2(q * lg2(u)), 0).astype(f32)
            n_ev += 1
        P[w0:w0 + warp] = p / p.sum(1, dtype=f32)[:, None]
        evals[w0:w0 + warp] = n_ev
    return P, evals


def trial(name, g, alpha):
    ref = oracle.entmax_bisect(torch.from_numpy(g), alpha).numpy()
    ref64 = oracle.entmax_bisect(torch.from_numpy(g).double(), alpha).numpy()
    p, ev = solve(g, alpha)
    supp = (ref > 0).sum(1).mean()
    print(f'{name:34s} alpha {alpha:<4} F {g.shape[1]:3d} support {supp:5.1f}  |p-ref| {np.abs(p - ref).max():.2e}  '
          f'|p-ref64| {np.abs(p - ref64).max():.2e}  |ref-ref64| {np.abs(ref - ref64).max():.2e}  evals/warp {ev.mean():.2f} '
          f'max {ev.max()}  nan {int(np.isnan(p).sum())}')


if __name__ == '__main__':
    rng = np.random.default_rng(0)
    for alpha in (1.7, 1.2, 1.9, 1.05, 1.99):
        for F in (39, 40, 10):
            for scale in (1e-3, 0.1, 0.5, 1.0, 3.0, 10.0, 30.0):
                g = (rng.standard_normal((64 * 64, F)) * scale).astype(f32)
                trial(f'normal x{scale}', g, alpha)
        g = (rng.standard_t(2, (4096, 39)) * 2).astype(f32)
        trial('student-t x2', g, alpha)
        g = np.zeros((4096, 39), f32); g[:, 3] = 5.0
        trial('one-hot', g, alpha)
        g = np.zeros((4096, 39), f32); g[:, :2] = 1.0
        trial('two tied maxima', g, alpha)
        g = np.tile(np.linspace(-3, 3, 39, dtype=f32), (4096, 1))
        trial('linear ramp', g, alpha)
        g = (rng.standard_normal((4096, 39)) * np.repeat(rng.choice([1e-3, 1.0, 10.0], 4096), 1)[:, None]).astype(f32)
        trial('mixed scales inside a warp', g, alpha)
