// alpha-entmax threshold solvers for a thread that owns NR rows (1 or 2) whose F logits are packed along the FIELD axis:
//   X[n][j] = (x[n][2j], x[n][2j+1]),  n < NR,  j < NP = ceil(F/2);  a padded last element (odd F) holds -inf.
// This is the register layout tcgen05.ld hands out (adjacent TMEM columns -> adjacent registers), so the packed f32x2
// instructions work on (f, f+1) pairs of ONE row (entmax_pair.cuh packs the same f of TWO rows instead).
//
// Reference: utils/entmax.py:29-68 (50 bisection halvings; p evaluated at the last midpoint and renormalised).  As in
// entmax.cuh the kernels find the same root tau* of f(tau) = sum_f [X_f - tau]_+^q - 1, q = 1/(alpha-1), with a faster
// iteration and end with the reference's renormalisation.  New in this header (1 < q < 2, i.e. 1.5 < alpha < 2, the
// range of every ARM-Net configuration with a general alpha, default 1.7):
//   * MUFU-free pre-solve.  For u in [0,1] and q = theta*1 + (1-theta)*2 Hoelder's inequality gives
//       sum u^q <= (sum u)^theta (sum u^2)^(1-theta)      (theta = 2 - q),
//     and the right-hand side needs no transcendental per element.  Three Newton steps on log(rhs) = 0 (per ROW: two lg2
//     and two divisions) land within ~1e-2 of tau*, on its right; tau is clamped to the reference's own bracket
//     [max-1, max-F^-(alpha-1)] (entmax.py:46-47).
//   * Newton on the q-norm N(tau) = (sum u^q)^(1/q) - 1 instead of on f: N is convex, decreasing and exactly linear when
//     the support is a set of equal logits, so far from the root a step is much longer than Newton's on f, from either
//     side the next iterate lies left of the root, and from there the iteration is monotone.
//   * The last sweep is the cross pass.  Once a step is small enough that the next one is predicted to be below the parity
//     budget, the sweep also accumulates gates*values*e (the caller's functor); if its own step then is <= 5e-7 (or 2 ulp
//     of tau) the row is finished without a separate evaluation.
// tools/entmax_holder_proto.py emulates this in fp32 numpy against the oracle: 3-4 sweeps of 2 MUFU per element per
// 64-row warp where the bound-started Newton of entmax.cuh needs 6-7, same gate error (at the level of the reference's
// own fp32-vs-fp64 difference).
#pragma once

#include "entmax_pair.cuh"

namespace armnet {

// Row maxima and means.  ODD: element (NP-1).y is padding (-inf) and excluded from the mean.
template <int NR, int NP, bool ODD>
__device__ __forceinline__ void rows_max_mean(const float2 (&X)[NR][NP], const EntmaxParams &ep, float (&mx)[NR],
                                              float (&mean)[NR]) {
    // two interleaved chains per reduction (even / odd pairs): the chains are latency-bound, not issue-bound
#pragma unroll
    for (int n = 0; n < NR; ++n) {
        float2 ma = X[n][0], mb = make_float2(neg_inf(), neg_inf());
        float2 sa = X[n][0], sb = make_float2(0.f, 0.f);
#pragma unroll
        for (int j = 1; j < NP - 1; ++j) {
            if (j & 1) {
                mb = make_float2(fmaxf(mb.x, X[n][j].x), fmaxf(mb.y, X[n][j].y));
                sb = fadd2(sb, X[n][j]);
            } else {
                ma = make_float2(fmaxf(ma.x, X[n][j].x), fmaxf(ma.y, X[n][j].y));
                sa = fadd2(sa, X[n][j]);
            }
        }
        if (NP > 1) {
            mb.x = fmaxf(mb.x, X[n][NP - 1].x);
            sb.x += X[n][NP - 1].x;
            if (!ODD) {
                mb.y = fmaxf(mb.y, X[n][NP - 1].y);
                sb.y += X[n][NP - 1].y;
            }
        }
        const float2 s = fadd2(sa, sb);
        mx[n] = fmaxf(fmaxf(ma.x, ma.y), fmaxf(mb.x, mb.y));
        mean[n] = (s.x + s.y) * ep.inv_F;
    }
}

// Closed-form start for near-uniform rows (entmax.cuh: entmax_uniform_start). Returns whether both rows qualify.
template <int NR, int NP, bool ODD>
__device__ __forceinline__ bool rows_uniform_start(const float2 (&X)[NR][NP], const EntmaxParams &ep, const float (&mx)[NR],
                                                   const float (&mean)[NR], float (&tau0)[NR]) {
    bool ok = true;
#pragma unroll
    for (int n = 0; n < NR; ++n) {
        const float2 nm = splat2(-mx[n]);
        float2 sd2 = make_float2(0.f, 0.f), sd2b = make_float2(0.f, 0.f);
#pragma unroll
        for (int j = 0; j < NP - 1; ++j) {
            const float2 d = fadd2(X[n][j], nm);
            if (j & 1)
                sd2b = ffma2(d, d, sd2b);
            else
                sd2 = ffma2(d, d, sd2);
        }
        sd2 = fadd2(sd2, sd2b);
        {
            const float d0 = X[n][NP - 1].x - mx[n];
            sd2.x = fmaf(d0, d0, sd2.x);
            if (!ODD) {
                const float d1 = X[n][NP - 1].y - mx[n];
                sd2.y = fmaf(d1, d1, sd2.y);
            }
        }
        const float md = mean[n] - mx[n];
        const float var = fmaxf(fmaf(-md, md, (sd2.x + sd2.y) * ep.inv_F), 0.f);
        tau0[n] = mean[n] - ep.cF + ep.uni_k * var;
        ok = ok && (var <= ep.uni_var);
    }
    return ok;
}

// Mean and variance of each row about a pivot (its first logit): no max pass, no cancellation for near-uniform rows.
// Returns whether every row passes the near-uniform test  F var <= (0.2 cF)^2  (which bounds every |X_f - mean| by
// 0.2 cF) and, for those, the closed-form start  tau0 = mean - cF + (q-1) var / (2 cF)  (entmax.cuh).
template <int NR, int NP, bool ODD>
__device__ __forceinline__ bool rows_moments_uniform(const float2 (&X)[NR][NP], const EntmaxParams &ep, float (&mean)[NR],
                                                     float (&tau0)[NR]) {
    bool ok = true;
#pragma unroll
    for (int n = 0; n < NR; ++n) {
        const float c = X[n][0].x;
        const float2 nc = splat2(-c);
        float2 sa = make_float2(0.f, 0.f), sb = sa, qa = sa, qb = sa;   // two interleaved chains per sum
#pragma unroll
        for (int j = 0; j < NP - 1; ++j) {
            const float2 d = fadd2(X[n][j], nc);
            if (j & 1) {
                sb = fadd2(sb, d);
                qb = ffma2(d, d, qb);
            } else {
                sa = fadd2(sa, d);
                qa = ffma2(d, d, qa);
            }
        }
        {
            const float d0 = X[n][NP - 1].x - c;
            sb.x += d0;
            qb.x = fmaf(d0, d0, qb.x);
            if (!ODD) {
                const float d1 = X[n][NP - 1].y - c;
                sb.y += d1;
                qb.y = fmaf(d1, d1, qb.y);
            }
        }
        const float2 sd = fadd2(sa, sb), sq = fadd2(qa, qb);
        const float md = (sd.x + sd.y) * ep.inv_F;
        const float var = fmaxf(fmaf(-md, md, (sq.x + sq.y) * ep.inv_F), 0.f);
        mean[n] = c + md;
        tau0[n] = mean[n] - ep.cF + ep.uni_k * var;
        ok = ok && (var <= ep.uni_var);
    }
    return ok;
}

// Row maxima only (the mean is known).
template <int NR, int NP, bool ODD>
__device__ __forceinline__ void rows_max(const float2 (&X)[NR][NP], float (&mx)[NR]) {
#pragma unroll
    for (int n = 0; n < NR; ++n) {
        float2 ma = X[n][0], mb = make_float2(neg_inf(), neg_inf());
#pragma unroll
        for (int j = 1; j < NP - 1; ++j) {
            if (j & 1)
                mb = make_float2(fmaxf(mb.x, X[n][j].x), fmaxf(mb.y, X[n][j].y));
            else
                ma = make_float2(fmaxf(ma.x, X[n][j].x), fmaxf(ma.y, X[n][j].y));
        }
        if (NP > 1) {
            mb.x = fmaxf(mb.x, X[n][NP - 1].x);
            if (!ODD) mb.y = fmaxf(mb.y, X[n][NP - 1].y);
        }
        mx[n] = fmaxf(fmaxf(ma.x, ma.y), fmaxf(mb.x, mb.y));
    }
}

// Hoelder-bound pre-solve (see the header): tau of both rows, MUFU only per row.  Valid for 1 < q < 2.
template <int NR, int NP, int ITERS>
__device__ __forceinline__ void rows_holder_presolve(const float2 (&X)[NR][NP], const EntmaxParams &ep,
                                                     const float (&mx)[NR], const float (&mean)[NR], float (&tau)[NR]) {
    const float th = 2.f - ep.q, omt = ep.q - 1.f;
#pragma unroll
    for (int n = 0; n < NR; ++n) tau[n] = fmaxf(mx[n] - 1.f, mean[n] - ep.cF);
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
        float2 A[NR], Bq[NR], C[NR], nt[NR];
#pragma unroll
        for (int n = 0; n < NR; ++n) nt[n] = splat2(-tau[n]);
#pragma unroll
        for (int n = 0; n < NR; ++n) A[n] = Bq[n] = C[n] = make_float2(0.f, 0.f);
#pragma unroll
        for (int j = 0; j < NP; ++j) {
#pragma unroll
            for (int n = 0; n < NR; ++n) {
                const float2 d = fadd2(X[n][j], nt[n]);
                const float2 u = relu2(d);
                A[n] = fadd2(A[n], u);
                Bq[n] = ffma2(u, u, Bq[n]);
                C[n] = fadd2(C[n], make_float2(d.x > 0.f ? 1.f : 0.f, d.y > 0.f ? 1.f : 0.f));
            }
        }
#pragma unroll
        for (int n = 0; n < NR; ++n) {
            const float a = fmaxf(A[n].x + A[n].y, 1e-30f), b = fmaxf(Bq[n].x + Bq[n].y, 1e-30f);
            const float c = C[n].x + C[n].y;
            const float lphi = 0.6931471805599453f * (th * fast_lg2(a) + omt * fast_lg2(b));
            const float dl = th * __fdividef(c, a) + 2.f * omt * __fdividef(a, b);  // -d(log rhs)/d tau > 0
            float d = __fdividef(lphi, dl);
            if (!(fabsf(d) < 1e30f)) d = 0.f;
            tau[n] = fminf(tau[n] + d, mx[n] - ep.cF);
        }
    }
}

// Step of Newton's iteration on the q-norm from the sums S = sum u^q, S1 = sum u^(q-1) at tau:
//   N = S^(1/q), N' = -(N/S) S1  ->  d = (1 - 1/N) S / S1,   1/N = ex2(-(alpha-1) lg2 S).
__device__ __forceinline__ float qnorm_newton_step(float S, float S1, const EntmaxParams &ep) {
    const float rN = fast_ex2(-ep.am1 * fast_lg2(fmaxf(S, 1e-30f)));
    float d = __fdividef((1.f - rN) * S, S1);
    if (!(S1 > 0.f) || !(fabsf(d) < 1e30f)) d = 0.f;
    return d;
}

// One evaluation sweep of both rows at tau for POW_GENERAL: S[n] = sum u^q, S1[n] = sum u^(q-1).
// FUSED: `cross(j, w0, w1)` is called for every field pair j with w_n = (p_n[2j] V_n[2j], p_n[2j+1] V_n[2j+1]) -- the
// caller accumulates the log-space product with them; V comes from `vrow(n, j)`.
template <int NR, int NP, bool FUSED, class VRow, class Cross>
__device__ __forceinline__ void rows_general_sweep(const float2 (&X)[NR][NP], const float (&tau)[NR],
                                                   const EntmaxParams &ep, float (&S)[NR], float (&S1)[NR], VRow vrow,
                                                   Cross cross) {
    float2 nt[NR];
#pragma unroll
    for (int n = 0; n < NR; ++n) nt[n] = splat2(-tau[n]);
    const float2 qm1 = splat2(ep.qm1);
    float2 s[NR], s1[NR];
#pragma unroll
    for (int n = 0; n < NR; ++n) s[n] = s1[n] = make_float2(0.f, 0.f);
#pragma unroll
    for (int j = 0; j < NP; ++j) {
        float2 w[NR];
#pragma unroll
        for (int n = 0; n < NR; ++n) {
            const float2 u = relu2(fadd2(X[n][j], nt[n]));
            const float2 t = fmul2(make_float2(fast_lg2(u.x), fast_lg2(u.y)), qm1);
            const float2 g = make_float2(fast_ex2(t.x), fast_ex2(t.y));  // u^(q-1); u = 0 -> 0 because q - 1 > 0
            s1[n] = fadd2(s1[n], g);
            if (FUSED) {
                const float2 p = fmul2(g, u);  // u^q
                s[n] = fadd2(s[n], p);
                w[n] = fmul2(p, vrow(n, j));   // armnet.py:36; the normalisation by S is applied once, at the end
            } else {
                s[n] = ffma2(g, u, s[n]);
            }
        }
        if (FUSED) cross(j, w);
    }
#pragma unroll
    for (int n = 0; n < NR; ++n) {
        S[n] = s[n].x + s[n].y;
        S1[n] = s1[n].x + s1[n].y;
    }
}

// Final pass at a solved tau for any mode: unnormalised gates, their sums, and the cross callback.
template <int MODE, int NR, int NP, class VRow, class Cross>
__device__ __forceinline__ void rows_cross_pass(const float2 (&X)[NR][NP], const float (&tau)[NR], const EntmaxParams &ep,
                                                float (&S)[NR], VRow vrow, Cross cross) {
    float2 nt[NR], s[NR];
#pragma unroll
    for (int n = 0; n < NR; ++n) {
        nt[n] = splat2(-tau[n]);
        s[n] = make_float2(0.f, 0.f);
    }
#pragma unroll
    for (int j = 0; j < NP; ++j) {
        float2 w[NR];
#pragma unroll
        for (int n = 0; n < NR; ++n) {
            const float2 p = gate_unnorm2<MODE>(X[n][j], nt[n], ep);
            s[n] = fadd2(s[n], p);
            w[n] = fmul2(p, vrow(n, j));
        }
        cross(j, w);
    }
#pragma unroll
    for (int n = 0; n < NR; ++n) S[n] = s[n].x + s[n].y;
}

// alpha = 1.5 (u^2, u) and alpha = 2 (Michelot): the iterations of entmax.cuh on the field-packed layout.
template <int NR, int NP>
__device__ __forceinline__ void rows_solve_simple(const float2 (&X)[NR][NP], const EntmaxParams &ep, const float (&mx)[NR],
                                                  const float (&mean)[NR], float (&tau)[NR]) {
#pragma unroll
    for (int n = 0; n < NR; ++n) tau[n] = fmaxf(mx[n] - 1.f, mean[n] - ep.cF);
    constexpr int kMaxIt = 12;
#pragma unroll 1
    for (int it = 0; it < kMaxIt; ++it) {
        bool done = true;
#pragma unroll
        for (int n = 0; n < NR; ++n) {
            const float2 nt = splat2(-tau[n]);
            float2 s = make_float2(0.f, 0.f), s1 = make_float2(0.f, 0.f);
            if (ep.mode == POW_SQUARE) {
#pragma unroll
                for (int j = 0; j < NP; ++j) {
                    const float2 u = relu2(fadd2(X[n][j], nt));
                    s1 = fadd2(s1, u);
                    s = ffma2(u, u, s);
                }
                const float den = 2.f * (s1.x + s1.y);
                float d = __fdividef(s.x + s.y - 1.f, den);
                if (!(den > 0.f)) d = 0.f;
                tau[n] += d;
                done = done && fabsf(d) <= fmaxf(2e-5f, 2.4e-7f * fabsf(tau[n]));
            } else {  // POW_LINEAR
#pragma unroll
                for (int j = 0; j < NP; ++j) {
                    const float2 u = fadd2(X[n][j], nt);
                    s = fadd2(s, relu2(u));
                    s1 = fadd2(s1, make_float2(u.x > 0.f ? 1.f : 0.f, u.y > 0.f ? 1.f : 0.f));
                }
                const float den = s1.x + s1.y;
                float d = __fdividef(s.x + s.y - 1.f, den);
                if (!(den > 0.f)) d = 0.f;
                tau[n] += d;
                done = done && fabsf(d) <= 2.4e-7f * fmaxf(1.f, fabsf(tau[n]));
            }
        }
        if (__all_sync(0xffffffffu, done)) break;
    }
}

// Complete gates + cross product of two rows: solves tau, calls `reset()` before every sweep that feeds `cross`, and
// returns S[n] (the normaliser of the gates the LAST cross sweep used).  All 32 lanes must call this together.
template <int NR, int NP, bool ODD, class VRow, class Cross, class Reset>
__device__ __forceinline__ void rows_entmax_cross(const float2 (&X)[NR][NP], const EntmaxParams &ep, float (&tau)[NR],
                                                  float (&S)[NR], VRow vrow, Cross cross, Reset reset) {
    constexpr unsigned kFull = 0xffffffffu;
    float mx[NR], mean[NR];
    if (ep.mode != POW_GENERAL) {
        rows_max_mean<NR, NP, ODD>(X, ep, mx, mean);
        if (ep.mode == POW_SOFTMAX) {
#pragma unroll
            for (int n = 0; n < NR; ++n) tau[n] = mx[n];
        } else {
            rows_solve_simple<NR, NP>(X, ep, mx, mean, tau);
        }
        reset();
        switch (ep.mode) {
            case POW_SOFTMAX: rows_cross_pass<POW_SOFTMAX, NR, NP>(X, tau, ep, S, vrow, cross); break;
            case POW_LINEAR: rows_cross_pass<POW_LINEAR, NR, NP>(X, tau, ep, S, vrow, cross); break;
            default: rows_cross_pass<POW_SQUARE, NR, NP>(X, tau, ep, S, vrow, cross); break;
        }
        return;
    }
    // ---- POW_GENERAL
    bool fuse = false;  // warp-uniform: the next sweep also accumulates the cross product
    if (__all_sync(kFull, rows_moments_uniform<NR, NP, ODD>(X, ep, mean, tau))) {
        fuse = true;  // random-init weights / weakly attending neurons: the closed form is usually exact to 1e-7
    } else {
        rows_max<NR, NP, ODD>(X, mx);
        if (ep.q < 2.f) {
            rows_holder_presolve<NR, NP, 3>(X, ep, mx, mean, tau);
        } else {
#pragma unroll
            for (int n = 0; n < NR; ++n) tau[n] = fmaxf(mx[n] - 1.f, mean[n] - ep.cF);
        }
    }
    constexpr int kMaxIt = 12;
    float S1[NR];
#pragma unroll 1
    for (int it = 0; it < kMaxIt; ++it) {
        if (fuse) {
            reset();
            rows_general_sweep<NR, NP, true>(X, tau, ep, S, S1, vrow, cross);
        } else {
            rows_general_sweep<NR, NP, false>(X, tau, ep, S, S1, vrow, cross);
        }
        float d[NR], amax = 0.f;
        bool tight = true, loose = true;
#pragma unroll
        for (int n = 0; n < NR; ++n) {
            d[n] = qnorm_newton_step(S[n], S1[n], ep);
            const float a = fabsf(d[n]), rel = 2.4e-7f * fabsf(tau[n]);
            tight = tight && a <= fmaxf(5e-7f, rel);
            loose = loose && a <= fmaxf(2e-5f, rel);
            amax = fmaxf(amax, a);
        }
        // |dp| <= q u^(q-1) |d| before the renormalisation: inside the parity budget (gates 2e-6 abs)
        if (fuse && __all_sync(kFull, tight)) return;
#pragma unroll
        for (int n = 0; n < NR; ++n) tau[n] += d[n];
        // the step is applied even when it is the last: |f(tau + d)| = O(d^2), and the caller renormalises
        if (!fuse && __all_sync(kFull, loose)) break;
        fuse = __all_sync(kFull, amax <= 3e-3f);
    }
    reset();
    rows_cross_pass<POW_GENERAL, NR, NP>(X, tau, ep, S, vrow, cross);
}

}  // namespace armnet
