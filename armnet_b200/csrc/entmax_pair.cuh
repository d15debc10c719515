// Packed (f32x2) variant of the entmax solvers in entmax.cuh for a thread that owns TWO rows: element f of both rows
// sits in one 64-bit register pair X[f] = (row0, row1), so subtractions, scalings and the running sums are one
// FADD2 / FMUL2 / FFMA2 (Blackwell packed-fp32 instructions) for both rows; FMNMX and MUFU stay scalar.
// Same algorithms, same constants, same stopping rules as entmax.cuh (see there for the maths and the reference
// lines, utils/entmax.py:29-68).
#pragma once

#include "entmax.cuh"

namespace armnet {

__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
    float2 d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;"
        : "=l"(reinterpret_cast<uint64_t &>(d))
        : "l"(reinterpret_cast<uint64_t &>(a)), "l"(reinterpret_cast<uint64_t &>(b)),
          "l"(reinterpret_cast<uint64_t &>(c)));
    return d;
}
__device__ __forceinline__ float2 fadd2(float2 a, float2 b) {
    float2 d;
    asm("add.rn.f32x2 %0, %1, %2;"
        : "=l"(reinterpret_cast<uint64_t &>(d))
        : "l"(reinterpret_cast<uint64_t &>(a)), "l"(reinterpret_cast<uint64_t &>(b)));
    return d;
}
__device__ __forceinline__ float2 fmul2(float2 a, float2 b) {
    float2 d;
    asm("mul.rn.f32x2 %0, %1, %2;"
        : "=l"(reinterpret_cast<uint64_t &>(d))
        : "l"(reinterpret_cast<uint64_t &>(a)), "l"(reinterpret_cast<uint64_t &>(b)));
    return d;
}
__device__ __forceinline__ float2 splat2(float a) { return make_float2(a, a); }
__device__ __forceinline__ float2 relu2(float2 a) { return make_float2(fmaxf(a.x, 0.f), fmaxf(a.y, 0.f)); }

// Unnormalised gates of element x of both rows; ntau = -tau.
template <int MODE>
__device__ __forceinline__ float2 gate_unnorm2(float2 x, float2 ntau, const EntmaxParams &ep) {
    if (MODE == POW_SOFTMAX) {
        const float2 t = fmul2(fadd2(x, ntau), splat2(1.4426950408889634f));
        return make_float2(fast_ex2(t.x), fast_ex2(t.y));
    } else if (MODE == POW_LINEAR) {
        return relu2(fadd2(x, ntau));
    } else if (MODE == POW_SQUARE) {
        const float2 u = relu2(fadd2(x, ntau));
        return fmul2(u, u);
    } else {  // POW_GENERAL, POW_BISECT: u^q; u = 0 -> lg2 = -inf -> ex2(-inf) = 0
        const float2 u = relu2(fadd2(x, ntau));
        const float2 t = fmul2(make_float2(fast_lg2(u.x), fast_lg2(u.y)), splat2(ep.q));
        return make_float2(fast_ex2(t.x), fast_ex2(t.y));
    }
}

__device__ __forceinline__ float2 gate_unnorm2_rt(float2 x, float2 ntau, const EntmaxParams &ep) {
    switch (ep.mode) {
        case POW_SOFTMAX: return gate_unnorm2<POW_SOFTMAX>(x, ntau, ep);
        case POW_LINEAR: return gate_unnorm2<POW_LINEAR>(x, ntau, ep);
        case POW_SQUARE: return gate_unnorm2<POW_SQUARE>(x, ntau, ep);
        default: return gate_unnorm2<POW_GENERAL>(x, ntau, ep);
    }
}

// Row maxima and means over the real fields (padded entries are -inf and excluded from the mean).
template <int FP, bool EXACT>
__device__ __forceinline__ void row_max_mean2(const float2 (&X)[FP], int F, const EntmaxParams &ep, float2 &mx,
                                              float2 &mean) {
    mx = X[0];
#pragma unroll
    for (int f = 1; f < FP; ++f) {
        mx.x = fmaxf(mx.x, X[f].x);
        mx.y = fmaxf(mx.y, X[f].y);
    }
    float2 sum = make_float2(0.f, 0.f);
#pragma unroll
    for (int f = 0; f < FP; ++f)
        if (EXACT || f < F) sum = fadd2(sum, X[f]);
    mean = fmul2(sum, splat2(ep.inv_F));
}

// Closed-form start for near-uniform rows (entmax.cuh: entmax_uniform_start), both rows at once.
template <int FP, bool EXACT>
__device__ __forceinline__ bool entmax_uniform_start2(const float2 (&X)[FP], int F, const EntmaxParams &ep, float2 mx,
                                                      float2 mean, float2 &tau0) {
    const float2 nmx = make_float2(-mx.x, -mx.y);
    float2 sd2 = make_float2(0.f, 0.f);
#pragma unroll
    for (int f = 0; f < FP; ++f) {
        if (EXACT || f < F) {
            const float2 d = fadd2(X[f], nmx);
            sd2 = ffma2(d, d, sd2);
        }
    }
    const float2 md = fadd2(mean, nmx);
    const float v0 = fmaxf(fmaf(-md.x, md.x, sd2.x * ep.inv_F), 0.f);
    const float v1 = fmaxf(fmaf(-md.y, md.y, sd2.y * ep.inv_F), 0.f);
    tau0 = make_float2(mean.x - ep.cF + ep.uni_k * v0, mean.y - ep.cF + ep.uni_k * v1);
    return fmaxf(v0, v1) <= ep.uni_var;
}

// tau of both rows. `warm`: tau already holds a Newton iterate (POW_GENERAL), skip the bound-based start.
// All 32 lanes of the warp must call this together (warp-uniform exit votes).
template <int FP, bool EXACT>
__device__ __forceinline__ void entmax_solve_tau2(const float2 (&X)[FP], int F, const EntmaxParams &ep, float2 mx,
                                                  float2 mean, float2 &tau, bool warm) {
    if (!warm) tau = mx;
    if (ep.mode == POW_SOFTMAX) return;

    if (ep.mode == POW_BISECT) {
        // entmax.py:46-61, step by step.
        float2 tau_lo = make_float2(mx.x - 1.f, mx.y - 1.f);
        float2 s = make_float2(0.f, 0.f);
        {
            const float2 nt = make_float2(-tau_lo.x, -tau_lo.y);
#pragma unroll
            for (int f = 0; f < FP; ++f) s = fadd2(s, gate_unnorm2<POW_BISECT>(X[f], nt, ep));
        }
        const float2 f_lo = make_float2(s.x - 1.f, s.y - 1.f);
        float2 dm = make_float2((mx.x - ep.cF) - tau_lo.x, (mx.y - ep.cF) - tau_lo.y);
        for (int it = 0; it < ep.n_iter; ++it) {
            dm = fmul2(dm, splat2(0.5f));
            tau = fadd2(tau_lo, dm);
            // fixed point: every later midpoint equals tau_lo, so every later p_m equals this one
            const bool fixed = (tau.x == tau_lo.x) && (tau.y == tau_lo.y);
            const float2 nt = make_float2(-tau.x, -tau.y);
            s = make_float2(0.f, 0.f);
#pragma unroll
            for (int f = 0; f < FP; ++f) s = fadd2(s, gate_unnorm2<POW_BISECT>(X[f], nt, ep));
            if ((s.x - 1.f) * f_lo.x >= 0.f) tau_lo.x = tau.x;
            if ((s.y - 1.f) * f_lo.y >= 0.f) tau_lo.y = tau.y;
            if (__all_sync(0xffffffffu, fixed)) break;
        }
        return;
    }

    // Lower bounds on the root: max element alone -> tau >= max - 1; Jensen on the convex u^q -> tau >= mean - cF.
    if (!warm) tau = make_float2(fmaxf(mx.x - 1.f, mean.x - ep.cF), fmaxf(mx.y - 1.f, mean.y - ep.cF));

    constexpr int kMaxIt = 12;
    if (ep.mode == POW_GENERAL) {
        const float2 qm1 = splat2(ep.qm1);
        for (int it = 0; it < kMaxIt; ++it) {
            const float2 nt = make_float2(-tau.x, -tau.y);
            float2 s = make_float2(0.f, 0.f), s1 = make_float2(0.f, 0.f);
#pragma unroll
            for (int f = 0; f < FP; ++f) {
                const float2 u = relu2(fadd2(X[f], nt));
                const float2 t = fmul2(make_float2(fast_lg2(u.x), fast_lg2(u.y)), qm1);
                const float2 w = make_float2(fast_ex2(t.x), fast_ex2(t.y));  // u^(q-1); u = 0 -> 0 because q-1 > 0
                s1 = fadd2(s1, w);
                s = ffma2(w, u, s);
            }
            float d0 = __fdividef(s.x - 1.f, ep.q * s1.x);
            float d1 = __fdividef(s.y - 1.f, ep.q * s1.y);
            if (!(s1.x > 0.f)) d0 = 0.f;
            if (!(s1.y > 0.f)) d1 = 0.f;
            // the step is applied even when it is the last: |f(tau+d)| = O(d^2), and the caller renormalises
            tau.x += d0;
            tau.y += d1;
            if (__all_sync(0xffffffffu, fmaxf(fabsf(d0), fabsf(d1)) <= 2e-5f)) break;
        }
    } else if (ep.mode == POW_SQUARE) {
        for (int it = 0; it < kMaxIt; ++it) {
            const float2 nt = make_float2(-tau.x, -tau.y);
            float2 s = make_float2(0.f, 0.f), s1 = make_float2(0.f, 0.f);
#pragma unroll
            for (int f = 0; f < FP; ++f) {
                const float2 u = relu2(fadd2(X[f], nt));
                s1 = fadd2(s1, u);
                s = ffma2(u, u, s);
            }
            float d0 = __fdividef(s.x - 1.f, 2.f * s1.x);
            float d1 = __fdividef(s.y - 1.f, 2.f * s1.y);
            if (!(s1.x > 0.f)) d0 = 0.f;
            if (!(s1.y > 0.f)) d1 = 0.f;
            tau.x += d0;
            tau.y += d1;
            if (__all_sync(0xffffffffu, fmaxf(fabsf(d0), fabsf(d1)) <= 2e-5f)) break;
        }
    } else {  // POW_LINEAR
        for (int it = 0; it < kMaxIt; ++it) {
            const float2 nt = make_float2(-tau.x, -tau.y);
            float2 s = make_float2(0.f, 0.f), cnt = make_float2(0.f, 0.f);
#pragma unroll
            for (int f = 0; f < FP; ++f) {
                const float2 u = fadd2(X[f], nt);
                s = fadd2(s, relu2(u));
                cnt.x += (u.x > 0.f) ? 1.f : 0.f;
                cnt.y += (u.y > 0.f) ? 1.f : 0.f;
            }
            float d0 = __fdividef(s.x - 1.f, cnt.x);
            float d1 = __fdividef(s.y - 1.f, cnt.y);
            if (!(cnt.x > 0.f)) d0 = 0.f;
            if (!(cnt.y > 0.f)) d1 = 0.f;
            tau.x += d0;
            tau.y += d1;
            const bool done = fabsf(d0) <= 2.4e-7f * fmaxf(1.f, fabsf(tau.x)) &&
                              fabsf(d1) <= 2.4e-7f * fmaxf(1.f, fabsf(tau.y));
            if (__all_sync(0xffffffffu, done)) break;
        }
    }
}

}  // namespace armnet
