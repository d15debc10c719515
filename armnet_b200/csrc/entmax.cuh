// Per-row alpha-entmax threshold solvers, one row of FP (padded) logits held in one thread's registers.
//
// Reference: utils/entmax.py:29-68 finds tau with 50 bisection halvings of [max-1, max-(1/d)^(alpha-1)]
// and returns p = [X - tau]_+^(1/(alpha-1)) / sum.  In fp32 that bisection reaches a bitwise fixed point after
// 20-40 halvings; any converged root-finder lands inside the reference's own fp32-vs-fp64 noise
// (tools/entmax_newton_proto.py, DESIGN.md "entmax").  The solvers here:
//   POW_GENERAL (1<alpha<2): f(tau) = sum u^q - 1 is convex and decreasing, so Newton from a lower bound
//       converges monotonically. Start at max(max-1, mean - F^-(alpha-1)) (max-element bound and Jensen bound);
//       u^(q-1) = ex2((q-1) lg2 u) gives f' and, times u, f: two MUFU per element per pass.
//   POW_SQUARE  (alpha=1.5): same Newton with u^2 / u, no MUFU.
//   POW_LINEAR  (alpha=2)  : Michelot's iteration tau += (sum u - 1)/|support| (Newton on a piecewise-linear f).
//   POW_BISECT  : the reference's loop, literally, with exit at the fixed point tau_lo + dm == tau_lo.
//   POW_SOFTMAX (alpha=1)  : tau := max, p = exp(X - max).
// Every solver returns tau; the caller evaluates p_f = pow_mode(X_f, tau) and divides by their sum, which is the
// reference's final renormalisation (entmax.py:63-64).
#pragma once

#include "common.cuh"

namespace armnet {

__device__ __forceinline__ float neg_inf() { return __int_as_float(0xff800000); }

// Unnormalised gate for one element, given the solved tau (mode is warp-uniform).
template <int MODE>
__device__ __forceinline__ float gate_unnorm(float x, float tau, const EntmaxParams &ep) {
    if (MODE == POW_SOFTMAX) {
        return fast_ex2((x - tau) * 1.4426950408889634f);
    } else if (MODE == POW_LINEAR) {
        return fmaxf(x - tau, 0.f);
    } else if (MODE == POW_SQUARE) {
        const float u = fmaxf(x - tau, 0.f);
        return u * u;
    } else {  // POW_GENERAL, POW_BISECT: u^q, u = 0 -> lg2 = -inf -> ex2(-inf) = 0
        const float u = fmaxf(x - tau, 0.f);
        return fast_ex2(ep.q * fast_lg2(u));
    }
}

// Runtime-mode variant for cold paths (debug outputs, standalone kernels' epilogues).
__device__ __forceinline__ float gate_unnorm_rt(float x, float tau, const EntmaxParams &ep) {
    switch (ep.mode) {
        case POW_SOFTMAX: return gate_unnorm<POW_SOFTMAX>(x, tau, ep);
        case POW_LINEAR: return gate_unnorm<POW_LINEAR>(x, tau, ep);
        case POW_SQUARE: return gate_unnorm<POW_SQUARE>(x, tau, ep);
        default: return gate_unnorm<POW_GENERAL>(x, tau, ep);
    }
}

// Row maximum and mean over the real fields (padded entries are -inf and excluded from the mean).
template <int FP, bool EXACT>
__device__ __forceinline__ void row_max_mean(const float (&X)[FP], int F, const EntmaxParams &ep, float &mx,
                                             float &mean) {
    mx = X[0];
#pragma unroll
    for (int f = 1; f < FP; ++f) mx = fmaxf(mx, X[f]);
    float sum = 0.f;
#pragma unroll
    for (int f = 0; f < FP; ++f)
        if (EXACT || f < F) sum += X[f];
    mean = sum * ep.inv_F;
}

// X[n][f] = (alpha-1) * g[f] of row n for f < F (g itself for softmax), -inf for padded f >= F.
// NR rows are solved together in one loop (their element-wise work interleaves: twice the ILP per thread); a row that
// has converged keeps taking Newton steps until every row of every lane has, which only tightens it.
// All 32 lanes of the warp must call this together (warp-uniform exit votes).
// mx / mean: from row_max_mean.  `warm` (POW_GENERAL only): tau[] already holds a Newton iterate (from the fused
// first pass of the caller); the bound-based start is skipped.
template <int NR, int FP, bool EXACT>
__device__ __forceinline__ void entmax_solve_tau(const float (&X)[NR][FP], int F, const EntmaxParams &ep,
                                                 const float (&mx)[NR], const float (&mean)[NR], float (&tau)[NR],
                                                 bool warm = false) {
    if (!warm) {
#pragma unroll
        for (int n = 0; n < NR; ++n) tau[n] = mx[n];
    }
    if (ep.mode == POW_SOFTMAX) return;

    if (ep.mode == POW_BISECT) {
        // entmax.py:46-61, step by step.
        float tau_lo[NR], f_lo[NR], dm[NR];
#pragma unroll
        for (int n = 0; n < NR; ++n) {
            tau_lo[n] = mx[n] - 1.f;
            const float tau_hi = mx[n] - ep.cF;
            float s = 0.f;
#pragma unroll
            for (int f = 0; f < FP; ++f) s += gate_unnorm<POW_BISECT>(X[n][f], tau_lo[n], ep);
            f_lo[n] = s - 1.f;
            dm[n] = tau_hi - tau_lo[n];
        }
        for (int it = 0; it < ep.n_iter; ++it) {
            bool fixed = true;
            float s[NR];
#pragma unroll
            for (int n = 0; n < NR; ++n) {
                dm[n] *= 0.5f;
                tau[n] = tau_lo[n] + dm[n];
                // fixed point: every later midpoint equals tau_lo, so every later p_m equals this one
                fixed = fixed && (tau[n] == tau_lo[n]);
                s[n] = 0.f;
            }
#pragma unroll
            for (int f = 0; f < FP; ++f)
#pragma unroll
                for (int n = 0; n < NR; ++n) s[n] += gate_unnorm<POW_BISECT>(X[n][f], tau[n], ep);
#pragma unroll
            for (int n = 0; n < NR; ++n)
                if ((s[n] - 1.f) * f_lo[n] >= 0.f) tau_lo[n] = tau[n];
            if (__all_sync(0xffffffffu, fixed)) break;
        }
        return;
    }

    // Lower bounds on the root: the max element alone gives tau >= max - 1; Jensen on the convex u^q
    // (q >= 1) gives tau >= mean - F^-(1/q) = mean - F^-(alpha-1).
    if (!warm) {
#pragma unroll
        for (int n = 0; n < NR; ++n) tau[n] = fmaxf(mx[n] - 1.f, mean[n] - ep.cF);
    }

    constexpr int kMaxIt = 12;
    if (ep.mode == POW_GENERAL) {
        const float qm1 = ep.qm1;
        for (int it = 0; it < kMaxIt; ++it) {
            float s[NR], s1[NR];
#pragma unroll
            for (int n = 0; n < NR; ++n) s[n] = s1[n] = 0.f;
#pragma unroll
            for (int f = 0; f < FP; ++f) {
#pragma unroll
                for (int n = 0; n < NR; ++n) {
                    const float u = fmaxf(X[n][f] - tau[n], 0.f);
                    const float w = fast_ex2(qm1 * fast_lg2(u));  // u^(q-1); u = 0 -> 0 because q-1 > 0
                    s1[n] += w;
                    s[n] = fmaf(w, u, s[n]);
                }
            }
            bool done = true;
#pragma unroll
            for (int n = 0; n < NR; ++n) {
                float d = __fdividef(s[n] - 1.f, ep.q * s1[n]);
                if (!(s1[n] > 0.f)) d = 0.f;
                // the step is applied even when it is the last: |f(tau+d)| = O(d^2), and the caller renormalises
                tau[n] += d;
                done = done && (fabsf(d) <= 2e-5f);
            }
            if (__all_sync(0xffffffffu, done)) break;
        }
    } else if (ep.mode == POW_SQUARE) {
        for (int it = 0; it < kMaxIt; ++it) {
            float s[NR], s1[NR];
#pragma unroll
            for (int n = 0; n < NR; ++n) s[n] = s1[n] = 0.f;
#pragma unroll
            for (int f = 0; f < FP; ++f) {
#pragma unroll
                for (int n = 0; n < NR; ++n) {
                    const float u = fmaxf(X[n][f] - tau[n], 0.f);
                    s1[n] += u;
                    s[n] = fmaf(u, u, s[n]);
                }
            }
            bool done = true;
#pragma unroll
            for (int n = 0; n < NR; ++n) {
                float d = __fdividef(s[n] - 1.f, 2.f * s1[n]);
                if (!(s1[n] > 0.f)) d = 0.f;
                tau[n] += d;
                done = done && (fabsf(d) <= 2e-5f);
            }
            if (__all_sync(0xffffffffu, done)) break;
        }
    } else {  // POW_LINEAR
        for (int it = 0; it < kMaxIt; ++it) {
            float s[NR], cnt[NR];
#pragma unroll
            for (int n = 0; n < NR; ++n) s[n] = cnt[n] = 0.f;
#pragma unroll
            for (int f = 0; f < FP; ++f) {
#pragma unroll
                for (int n = 0; n < NR; ++n) {
                    const float u = X[n][f] - tau[n];
                    s[n] += fmaxf(u, 0.f);
                    cnt[n] += (u > 0.f) ? 1.f : 0.f;
                }
            }
            bool done = true;
#pragma unroll
            for (int n = 0; n < NR; ++n) {
                float d = __fdividef(s[n] - 1.f, cnt[n]);
                if (!(cnt[n] > 0.f)) d = 0.f;
                tau[n] += d;
                done = done && (fabsf(d) <= 2.4e-7f * fmaxf(1.f, fabsf(tau[n])));
            }
            if (__all_sync(0xffffffffu, done)) break;
        }
    }
}


// Near-uniform rows (X_f = m + eps_f with |eps| << c = F^-(alpha-1)): expanding sum (c' + eps_f)^q = 1 to second order
// gives tau = m - c + (q-1) var(eps) / (2c) + O(eps^3 / c^2).  Returns that start and whether every |eps_f| <= 0.2 c
// (checked through F*var >= max eps^2).  The caller verifies the start with a Newton residual before trusting it.
template <int FP, bool EXACT>
__device__ __forceinline__ bool entmax_uniform_start(const float (&X)[FP], int F, const EntmaxParams &ep, float mx,
                                                     float mean, float &tau0) {
    float sd2 = 0.f;  // second moment of d_f = X_f - max (small numbers: no cancellation at large |X|)
#pragma unroll
    for (int f = 0; f < FP; ++f) {
        if (EXACT || f < F) {
            const float d = X[f] - mx;
            sd2 = fmaf(d, d, sd2);
        }
    }
    const float md = mean - mx;
    const float var = fmaxf(fmaf(-md, md, sd2 * ep.inv_F), 0.f);
    tau0 = mean - ep.cF + ep.uni_k * var;
    return var <= ep.uni_var;
}

}  // namespace armnet
