"""Design experiment (not shipped, not imported by the product): safeguarded Newton for
alpha-entmax in fp32 vs the reference's 50-step bisection (utils/entmax.py:29-68).
Run here with /root/reference importable; prints max |p - p_ref| and pass counts."""
import sys
import numpy as np
import torch
sys.path.insert(0, '/root/reference')
from utils.entmax import entmax_bisect

f32 = np.float32


def lg2(x):
    with np.errstate(divide='ignore'):
        return np.log2(x, dtype=f32)


def ex2(x):
    return np.exp2(x, dtype=f32)


def newton_entmax(g, alpha, eps_eval=2e-7, eps_skip=2e-5, maxit=16):
    """g: [N,F] fp32 logits. Returns p [N,F], passes [N] (evaluation passes incl. final)."""
    N, F = g.shape
    a = f32(alpha)
    am1 = f32(a - f32(1))
    q = f32(f32(1) / am1)
    qm1 = f32(q - f32(1))
    cF = f32(float(F) ** (-(float(am1))))
    X = (g * am1).astype(f32)
    mx = X.max(1)
    mean = (X.sum(1, dtype=f32) / f32(F)).astype(f32)
    lo = np.maximum(mx - f32(1), mean - cF).astype(f32)
    hi = (mx - cF).astype(f32)
    tau = lo.copy()
    done = np.zeros(N, bool)
    passes = np.zeros(N, int)
    for it in range(maxit):
        u = np.maximum(X - tau[:, None], f32(0)).astype(f32)
        w = ex2(qm1 * lg2(u)) if qm1 != 0 else (u > 0).astype(f32)
        w = np.where(u > 0, w, f32(0)).astype(f32)
        S1 = w.sum(1, dtype=f32)
        S = (w * u).sum(1, dtype=f32)
        passes += ~done
        fv = S - f32(1)
        d = (fv / (q * S1)).astype(f32)
        lo = np.where(fv >= 0, tau, lo)
        hi = np.where(fv < 0, tau, hi)
        tn = (tau + d).astype(f32)
        bad = ~((tn >= lo) & (tn <= hi))
        tn = np.where(bad, f32(0.5) * (lo + hi), tn).astype(f32)
        conv = np.abs(d) <= eps_skip          # apply the step, skip re-evaluation
        upd = ~done
        tau = np.where(upd, tn, tau)
        done |= conv
        if done.all():
            break
    u = np.maximum(X - tau[:, None], f32(0)).astype(f32)
    p = np.where(u > 0, ex2(q * lg2(u)), f32(0)).astype(f32)
    p = p / p.sum(1, dtype=f32)[:, None]
    return p.astype(f32), passes + 1


def trial(F, alpha, scale, N=20000, seed=0, dist='normal'):
    rng = np.random.default_rng(seed)
    if dist == 'normal':
        g = (rng.standard_normal((N, F)) * scale).astype(f32)
    else:
        g = (rng.standard_t(2, (N, F)) * scale).astype(f32)
    ref = entmax_bisect(torch.from_numpy(g), alpha=alpha, dim=-1).numpy()
    ref64 = entmax_bisect(torch.from_numpy(g).double(), alpha=alpha, dim=-1).numpy()
    p, passes = newton_entmax(g, alpha)
    return (np.abs(p - ref).max(), np.abs(p - ref64).max(), np.abs(ref - ref64).max(),
            passes.mean(), passes.max())


if __name__ == '__main__':
    for alpha in (1.7, 1.5, 2.0, 1.2, 2.5):
        for F in (10, 39):
            for scale in (1e-3, 0.1, 1.0, 5.0, 30.0):
                r = trial(F, alpha, scale)
                print(f'alpha={alpha} F={F} scale={scale:g}: |p-ref32|={r[0]:.2e} |p-ref64|={r[1]:.2e} '
                      f'|ref32-ref64|={r[2]:.2e} passes mean={r[3]:.2f} max={r[4]}')
