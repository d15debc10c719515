// alpha-entmax threshold solvers for a thread that owns TWO rows whose F logits are packed along the FIELD axis:
//   X[n][j] = (x[n][2j], x[n][2j+1]),  n in {0,1},  j < NP = ceil(F/2);  a padded last element (odd F) holds -inf.
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
template <int NP, bool ODD>
__device__ __forceinline__ void rows_max_mean(const float2 (&X)[2][NP], const EntmaxParams &ep, float (&mx)[2],
                                              float (&mean)[2]) {
#pragma unroll
    for (int n = 0; n < 2; ++n) {
        float m0 = X[n][0].x, m1 = X[n][0].y;
        float2 s = X[n][0];
#pragma unroll
        for (int j = 1; j < NP - 1; ++j) {
            m0 = fmaxf(m0, X[n][j].x);
            m1 = fmaxf(m1, X[n][j].y);
            s = fadd2(s, X[n][j]);
        }
        if (NP > 1) {
            m0 = fmaxf(m0, X[n][NP - 1].x);
            s.x += X[n][NP - 1].x;
            if (!ODD) {
                m1 = fmaxf(m1, X[n][NP - 1].y);
                s.y += X[n][NP - 1].y;
            }
        }
        mx[n] = fmaxf(m0, m1);
        mean[n] = (s.x + s.y) * ep.inv_F;
    }
}

// Closed-form start for near-uniform rows (entmax.cuh: entmax_uniform_start). Returns whether both rows qualify.
template <int NP, bool ODD>
__device__ __forceinline__ bool rows_uniform_start(const float2 (&X)[2][NP], const EntmaxParams &ep, const float (&mx)[2],
                                                   const float (&mean)[2], float (&tau0)[2]) {
    bool ok = true;
#pragma unroll
    for (int n = 0; n < 2; ++n) {
        const float2 nm = splat2(-mx[n]);
        float2 sd2 = make_float2(0.f, 0.f);
#pragma unroll
        for (int j = 0; j < NP - 1; ++j) {
            const float2 d = fadd2(X[n][j], nm);
            sd2 = ffma2(d, d, sd2);
        }
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

// Hoelder-bound pre-solve (see the header): tau of both rows, MUFU only per row.  Valid for 1 < q < 2.
template <int NP, int ITERS>
__device__ __forceinline__ void rows_holder_presolve(const float2 (&X)[2][NP], const EntmaxParams &ep,
                                                     const float (&mx)[2], const float (&mean)[2], float (&tau)[2]) {
    const float th = 2.f - ep.q, omt = ep.q - 1.f;
#pragma unroll
    for (int n = 0; n < 2; ++n) tau[n] = fmaxf(mx[n] - 1.f, mean[n] - ep.cF);
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
        float2 A[2], Bq[2], C[2];
        const float2 nt[2] = {splat2(-tau[0]), splat2(-tau[1])};
#pragma unroll
        for (int n = 0; n < 2; ++n) A[n] = Bq[n] = C[n] = make_float2(0.f, 0.f);
#pragma unroll
        for (int j = 0; j < NP; ++j) {
#pragma unroll
            for (int n = 0; n < 2; ++n) {
                const float2 d = fadd2(X[n][j], nt[n]);
                const float2 u = relu2(d);
                A[n] = fadd2(A[n], u);
                Bq[n] = ffma2(u, u, Bq[n]);
                C[n] = fadd2(C[n], make_float2(d.x > 0.f ? 1.f : 0.f, d.y > 0.f ? 1.f : 0.f));
            }
        }
#pragma unroll
        for (int n = 0; n < 2; ++n) {
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
template <int NP, bool FUSED, class VRow, class Cross>
__device__ __forceinline__ void rows_general_sweep(const float2 (&X)[2][NP], const float (&tau)[2],
                                                   const EntmaxParams &ep, float (&S)[2], float (&S1)[2], VRow vrow,
                                                   Cross cross) {
    const float2 nt[2] = {splat2(-tau[0]), splat2(-tau[1])};
    const float2 qm1 = splat2(ep.qm1);
    float2 s[2], s1[2];
#pragma unroll
    for (int n = 0; n < 2; ++n) s[n] = s1[n] = make_float2(0.f, 0.f);
#pragma unroll
    for (int j = 0; j < NP; ++j) {
        float2 w[2];
#pragma unroll
        for (int n = 0; n < 2; ++n) {
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
        if (FUSED) cross(j, w[0], w[1]);
    }
#pragma unroll
    for (int n = 0; n < 2; ++n) {
        S[n] = s[n].x + s[n].y;
        S1[n] = s1[n].x + s1[n].y;
    }
}

// Final pass at a solved tau for any mode: unnormalised gates, their sums, and the cross callback.
template <int MODE, int NP, class VRow, class Cross>
__device__ __forceinline__ void rows_cross_pass(const float2 (&X)[2][NP], const float (&tau)[2], const EntmaxParams &ep,
                                                float (&S)[2], VRow vrow, Cross cross) {
    const float2 nt[2] = {splat2(-tau[0]), splat2(-tau[1])};
    float2 s[2] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
#pragma unroll
    for (int j = 0; j < NP; ++j) {
        float2 w[2];
#pragma unroll
        for (int n = 0; n < 2; ++n) {
            const float2 p = gate_unnorm2<MODE>(X[n][j], nt[n], ep);
            s[n] = fadd2(s[n], p);
            w[n] = fmul2(p, vrow(n, j));
        }
        cross(j, w[0], w[1]);
    }
    S[0] = s[0].x + s[0].y;
    S[1] = s[1].x + s[1].y;
}

// alpha = 1.5 (u^2, u) and alpha = 2 (Michelot): the iterations of entmax.cuh on the field-packed layout.
template <int NP>
__device__ __forceinline__ void rows_solve_simple(const float2 (&X)[2][NP], const EntmaxParams &ep, const float (&mx)[2],
                                                  const float (&mean)[2], float (&tau)[2]) {
#pragma unroll
    for (int n = 0; n < 2; ++n) tau[n] = fmaxf(mx[n] - 1.f, mean[n] - ep.cF);
    constexpr int kMaxIt = 12;
#pragma unroll 1
    for (int it = 0; it < kMaxIt; ++it) {
        bool done = true;
#pragma unroll
        for (int n = 0; n < 2; ++n) {
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
template <int NP, bool ODD, class VRow, class Cross, class Reset>
__device__ __forceinline__ void rows_entmax_cross(const float2 (&X)[2][NP], const EntmaxParams &ep, float (&tau)[2],
                                                  float (&S)[2], VRow vrow, Cross cross, Reset reset) {
    constexpr unsigned kFull = 0xffffffffu;
    float mx[2], mean[2];
    rows_max_mean<NP, ODD>(X, ep, mx, mean);
    if (ep.mode != POW_GENERAL) {
        if (ep.mode == POW_SOFTMAX) {
            tau[0] = mx[0];
            tau[1] = mx[1];
        } else {
            rows_solve_simple<NP>(X, ep, mx, mean, tau);
        }
        reset();
        switch (ep.mode) {
            case POW_SOFTMAX: rows_cross_pass<POW_SOFTMAX, NP>(X, tau, ep, S, vrow, cross); break;
            case POW_LINEAR: rows_cross_pass<POW_LINEAR, NP>(X, tau, ep, S, vrow, cross); break;
            default: rows_cross_pass<POW_SQUARE, NP>(X, tau, ep, S, vrow, cross); break;
        }
        return;
    }
    // ---- POW_GENERAL
    bool fuse = false;  // warp-uniform: the next sweep also accumulates the cross product
    if (__all_sync(kFull, fmaxf(mx[0] - mean[0], mx[1] - mean[1]) <= 0.2f * ep.cF) &&
        __all_sync(kFull, rows_uniform_start<NP, ODD>(X, ep, mx, mean, tau))) {
        fuse = true;  // random-init weights / weakly attending neurons: the closed form is usually exact to 1e-7
    } else if (ep.q < 2.f) {
        rows_holder_presolve<NP, 3>(X, ep, mx, mean, tau);
    } else {
#pragma unroll
        for (int n = 0; n < 2; ++n) tau[n] = fmaxf(mx[n] - 1.f, mean[n] - ep.cF);
    }
    constexpr int kMaxIt = 12;
    float S1[2];
#pragma unroll 1
    for (int it = 0; it < kMaxIt; ++it) {
        if (fuse) {
            reset();
            rows_general_sweep<NP, true>(X, tau, ep, S, S1, vrow, cross);
        } else {
            rows_general_sweep<NP, false>(X, tau, ep, S, S1, vrow, cross);
        }
        const float d0 = qnorm_newton_step(S[0], S1[0], ep), d1 = qnorm_newton_step(S[1], S1[1], ep);
        const float a0 = fabsf(d0), a1 = fabsf(d1);
        const float rel0 = 2.4e-7f * fabsf(tau[0]), rel1 = 2.4e-7f * fabsf(tau[1]);
        // |dp| <= q u^(q-1) |d| before the renormalisation: inside the parity budget (gates 2e-6 abs)
        if (fuse && __all_sync(kFull, a0 <= fmaxf(5e-7f, rel0) && a1 <= fmaxf(5e-7f, rel1))) return;
        tau[0] += d0;
        tau[1] += d1;
        // the step is applied even when it is the last: |f(tau + d)| = O(d^2), and the caller renormalises
        if (!fuse && __all_sync(kFull, a0 <= fmaxf(2e-5f, rel0) && a1 <= fmaxf(2e-5f, rel1))) break;
        fuse = __all_sync(kFull, fmaxf(a0, a1) <= 3e-3f);
    }
    reset();
    rows_cross_pass<POW_GENERAL, NP>(X, tau, ep, S, vrow, cross);
}

}  // namespace armnet
