// alpha-entmax + cross product for a thread that owns TWO rows whose logits stay in TENSOR MEMORY: every pass over a
// row's F logits streams them from TMEM in chunks of 8 columns (tcgen05.ld.32x32b.x8, the next chunk in flight while
// the current one is processed) instead of holding 2 x F registers.  Same algorithms, constants and stopping rules as
// entmax_rows.cuh (reference: utils/entmax.py:29-68); see there for the maths.  What it buys: two rows per thread share
// every shared-memory read of an embedding row (the kernel is bound by the shared-memory -> register return path when a
// thread owns one row), at ~100 registers per thread.
#pragma once

#include <type_traits>

#include "entmax_rows.cuh"

namespace armnet {

__device__ __forceinline__ void tms_ld8(uint32_t taddr, float2 (&X)[4]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=f"(X[0].x), "=f"(X[0].y), "=f"(X[1].x), "=f"(X[1].y), "=f"(X[2].x), "=f"(X[2].y), "=f"(X[3].x),
                   "=f"(X[3].y)
                 : "r"(taddr)
                 : "memory");
}
__device__ __forceinline__ void tms_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// Calls f(j, x0, x1) for j = 0 .. NP-1 with x_n = (X_n[2j], X_n[2j+1]) of the rows at TMEM addresses t0 / t1.
// ODD: element (NP-1).y is padding and arrives as -inf.
template <int NP, bool ODD, class F>
__device__ __forceinline__ void tm_stream_rows2(uint32_t t0, uint32_t t1, F f) {
    static_assert(NP % 4 == 0, "chunks of 4 field pairs");
    constexpr int NC = NP / 4;
    float2 a[2][4], b[2][4];
    tms_ld8(t0, a[0]);
    tms_ld8(t1, a[1]);
    tms_ld_wait();
#pragma unroll
    for (int c = 0; c < NC; ++c) {
        if (c + 1 < NC) {
            tms_ld8(t0 + 8 * (c + 1), b[0]);
            tms_ld8(t1 + 8 * (c + 1), b[1]);
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            float2 x0 = a[0][k], x1 = a[1][k];
            if (ODD && 4 * c + k == NP - 1) x0.y = x1.y = neg_inf();
            f(4 * c + k, x0, x1);
        }
        if (c + 1 < NC) {
            tms_ld_wait();
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                a[0][k] = b[0][k];
                a[1][k] = b[1][k];
            }
        }
    }
}

// Complete gates + cross product of two rows.  X holds their logits in registers for the passes that need no
// accumulators (moments, maxima, pre-solve, plain Newton sweeps: entmax_rows.cuh at register speed); the passes that
// accumulate the cross product re-stream the logits from tensor memory (t0 / t1) instead, so that the 2 x F logit
// registers and the 2 x E accumulators are never live together.  Calls reset() before every pass that feeds
// cross(j, w0, w1) (w_n = gates * values of row n for the field pair j); returns S[n], the normaliser of the gates the
// LAST cross pass used.  vrow(n, j) = (V_n[2j], V_n[2j+1]).  All 32 lanes must call this together.
template <int NP, bool ODD, class VRow, class Cross, class Reset>
__device__ __forceinline__ void stream_entmax_cross(const float2 (&X)[2][NP], uint32_t t0, uint32_t t1,
                                                    const EntmaxParams &ep, float (&tau)[2], float (&S)[2], VRow vrow,
                                                    Cross cross, Reset reset) {
    constexpr unsigned kFull = 0xffffffffu;
    float mx[2], mean[2];
    auto pass_cross = [&](auto mode_tag) {
        constexpr int MODE = decltype(mode_tag)::value;
        const float2 nt0 = splat2(-tau[0]), nt1 = splat2(-tau[1]);
        float2 s0 = make_float2(0.f, 0.f), s1 = s0;
        reset();
        tm_stream_rows2<NP, ODD>(t0, t1, [&](int j, float2 x0, float2 x1) {
            const float2 p0 = gate_unnorm2<MODE>(x0, nt0, ep), p1 = gate_unnorm2<MODE>(x1, nt1, ep);
            s0 = fadd2(s0, p0);
            s1 = fadd2(s1, p1);
            cross(j, fmul2(p0, vrow(0, j)), fmul2(p1, vrow(1, j)));
        });
        S[0] = s0.x + s0.y;
        S[1] = s1.x + s1.y;
    };
    if (ep.mode != POW_GENERAL) {
        rows_max_mean<2, NP, ODD>(X, ep, mx, mean);
        if (ep.mode == POW_SOFTMAX) {
            tau[0] = mx[0];
            tau[1] = mx[1];
            pass_cross(std::integral_constant<int, POW_SOFTMAX>{});
        } else {
            rows_solve_simple<2, NP>(X, ep, mx, mean, tau);
            if (ep.mode == POW_SQUARE)
                pass_cross(std::integral_constant<int, POW_SQUARE>{});
            else
                pass_cross(std::integral_constant<int, POW_LINEAR>{});
        }
        return;
    }
    // ---- POW_GENERAL, phase 1 (logits in registers)
    // near-uniform rows (random-init weights / weakly attending neurons): closed-form start, every gate is positive
    const bool dense = __all_sync(kFull, rows_moments_uniform<2, NP, ODD>(X, ep, mean, tau));
    bool fuse = dense;
    int it = 0;
    if (!dense) {
        rows_max<2, NP, ODD>(X, mx);
        if (ep.q < 2.f) {
            rows_holder_presolve<2, NP, 3>(X, ep, mx, mean, tau);
        } else {
            tau[0] = fmaxf(mx[0] - 1.f, mean[0] - ep.cF);
            tau[1] = fmaxf(mx[1] - 1.f, mean[1] - ep.cF);
        }
        float S1[2];
#pragma unroll 1
        for (; it < 12; ++it) {
            rows_general_sweep<2, NP, false>(X, tau, ep, S, S1, vrow, [](int, const float2 (&)[2]) {});
            const float d0 = qnorm_newton_step(S[0], S1[0], ep), d1 = qnorm_newton_step(S[1], S1[1], ep);
            const float a0 = fabsf(d0), a1 = fabsf(d1);
            const float rel0 = 2.4e-7f * fabsf(tau[0]), rel1 = 2.4e-7f * fabsf(tau[1]);
            tau[0] += d0;
            tau[1] += d1;
            // the step is applied even when it is the last: |f(tau + d)| = O(d^2), and the caller renormalises
            if (__all_sync(kFull, a0 <= fmaxf(2e-5f, rel0) && a1 <= fmaxf(2e-5f, rel1))) break;
            if (__all_sync(kFull, fmaxf(a0, a1) <= 3e-3f)) {   // the next sweep is predicted to be the last: fuse it
                fuse = true;
                ++it;
                break;
            }
        }
    }
    // ---- phase 2 (logits streamed from tensor memory): sweeps that also accumulate the cross product
    if (fuse) {
        const float2 qm1 = splat2(ep.qm1);
#pragma unroll 1
        for (; it < 12; ++it) {
            const float2 nt0 = splat2(-tau[0]), nt1 = splat2(-tau[1]);
            float2 s0 = make_float2(0.f, 0.f), s1 = s0, g0 = s0, g1 = s0;   // sums of u^q and u^(q-1)
            reset();
            tm_stream_rows2<NP, ODD>(t0, t1, [&](int j, float2 x0, float2 x1) {
                float2 u0 = fadd2(x0, nt0), u1 = fadd2(x1, nt1);
                if (!dense || (ODD && j == NP - 1)) {   // dense rows: every X_f - tau >= 0.8 cF > 0, no clamp needed
                    u0 = relu2(u0);
                    u1 = relu2(u1);
                }
                const float2 l0 = fmul2(make_float2(fast_lg2(u0.x), fast_lg2(u0.y)), qm1);
                const float2 l1 = fmul2(make_float2(fast_lg2(u1.x), fast_lg2(u1.y)), qm1);
                const float2 e0 = make_float2(fast_ex2(l0.x), fast_ex2(l0.y));   // u^(q-1); u = 0 -> 0 because q - 1 > 0
                const float2 e1 = make_float2(fast_ex2(l1.x), fast_ex2(l1.y));
                g0 = fadd2(g0, e0);
                g1 = fadd2(g1, e1);
                const float2 p0 = fmul2(e0, u0), p1 = fmul2(e1, u1);          // u^q
                s0 = fadd2(s0, p0);
                s1 = fadd2(s1, p1);
                cross(j, fmul2(p0, vrow(0, j)), fmul2(p1, vrow(1, j)));        // armnet.py:36; normalised once, at the end
            });
            S[0] = s0.x + s0.y;
            S[1] = s1.x + s1.y;
            const float d0 = qnorm_newton_step(S[0], g0.x + g0.y, ep), d1 = qnorm_newton_step(S[1], g1.x + g1.y, ep);
            const float a0 = fabsf(d0), a1 = fabsf(d1);
            const float rel0 = 2.4e-7f * fabsf(tau[0]), rel1 = 2.4e-7f * fabsf(tau[1]);
            // |dp| <= q u^(q-1) |d| before the renormalisation: inside the parity budget (gates 2e-6 abs)
            if (__all_sync(kFull, a0 <= fmaxf(5e-7f, rel0) && a1 <= fmaxf(5e-7f, rel1))) return;
            tau[0] += d0;
            tau[1] += d1;
        }
    }
    pass_cross(std::integral_constant<int, POW_GENERAL>{});
}

}  // namespace armnet
