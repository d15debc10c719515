// Dense-row fast path for a thread that owns TWO rows whose logits stay in TENSOR MEMORY: every pass over a row's F
// logits streams them from TMEM in chunks of 8 columns (tcgen05.ld.32x32b.x8) instead of holding 2 x F registers, so two
// rows per thread share every shared-memory read of an embedding row at ~100 registers per thread (with one row per
// thread the init-weight regime is bound by the shared-memory -> register return path).  Rows that are not near-uniform
// take the register path of entmax_rows.cuh, one row at a time (measured: the streamed plain / pre-solve passes lose to
// the register-resident ones, 13.5 M vs 15.5 M samples/s at the trained-like C2a regime).  Same algorithms, constants and
// stopping rules as entmax_rows.cuh (reference: utils/entmax.py:29-68).
#pragma once

#include <type_traits>

#include "entmax_rows.cuh"

namespace armnet {

#ifndef ARMNET_STREAM_UNROLL
#define ARMNET_STREAM_UNROLL 5
#endif

__device__ __forceinline__ void tms_ld8(uint32_t taddr, float2 (&X)[4]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=f"(X[0].x), "=f"(X[0].y), "=f"(X[1].x), "=f"(X[1].y), "=f"(X[2].x), "=f"(X[2].y), "=f"(X[3].x),
                   "=f"(X[3].y)
                 : "r"(taddr)
                 : "memory");
}
__device__ __forceinline__ void tms_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// Calls f(c, k, x0, x1) for chunk c = 0 .. NP/4-1 (a RUNTIME loop: the code of a pass is one chunk long, which keeps the
// kernel inside the instruction cache) and pair k = 0 .. 3 (unrolled), field pair j = 4c + k, with
// x_n = (X_n[2j], X_n[2j+1]) of the rows at TMEM addresses t0 / t1.  ODD: element (NP-1).y is padding and arrives as -inf.
template <int NP, bool ODD, class F>
__device__ __forceinline__ void tm_stream_rows2(uint32_t t0, uint32_t t1, F f) {
    static_assert(NP % 4 == 0, "chunks of 4 field pairs");
    constexpr int NC = NP / 4;
    constexpr int kUnroll = ARMNET_STREAM_UNROLL;
#pragma unroll kUnroll
    for (int c = 0; c < NC; ++c) {
        // load, wait, use: the registers of an in-flight tcgen05.ld are never live across other work (the compiler does
        // not know the load is asynchronous); the latency of a chunk is hidden by the CTA's other warps
        float2 a[2][4];
        tms_ld8(t0 + 8 * c, a[0]);
        tms_ld8(t1 + 8 * c, a[1]);
        tms_ld_wait();
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            float2 x0 = a[0][k], x1 = a[1][k];
            if (ODD && k == 3 && c == NC - 1) x0.y = x1.y = neg_inf();
            f(c, k, x0, x1);
        }
    }
}

// DENSE rows only (near-uniform logits: random-init weights, weakly attending neurons): gates + cross product of the two
// rows at t0 / t1 with the logits streamed from tensor memory.  Returns false -- nothing accumulated -- when the rows are
// not near-uniform or the mode is not POW_GENERAL; the caller then runs the register path (entmax_rows.cuh) row by row.
// Otherwise: closed-form start tau0 = mean - cF + (q-1) var / (2 cF) (entmax.cuh), sweeps that accumulate
// cross(c, k, w0, w1) (w_n = gates * values of row n for the field pair 4c + k; reset() is called before each) and verify
// themselves with the q-norm Newton step, S[n] = the normaliser of the gates of the last sweep.
// vrow(n, c, k) = (V_n[2j], V_n[2j+1]), j = 4c + k.  All 32 lanes call this together.
template <int NP, bool ODD, class VRow, class Cross, class Reset>
__device__ __forceinline__ bool stream_dense_entmax_cross(uint32_t t0, uint32_t t1, const EntmaxParams &ep,
                                                          float (&tau)[2], float (&S)[2], VRow vrow, Cross cross,
                                                          Reset reset) {
    constexpr unsigned kFull = 0xffffffffu;
    constexpr int NC = NP / 4;
    if (ep.mode != POW_GENERAL) return false;
    // mean and variance about a pivot (each row's first logit): no cancellation for near-uniform rows
    {
        float c0 = 0.f, c1 = 0.f;
        float2 nc0 = make_float2(0.f, 0.f), nc1 = nc0, sa0 = nc0, sa1 = nc0, q0 = nc0, q1 = nc0;
        tm_stream_rows2<NP, ODD>(t0, t1, [&](int c, int k, float2 x0, float2 x1) {
            if (k == 0 && c == 0) {
                c0 = x0.x;
                c1 = x1.x;
                nc0 = splat2(-c0);
                nc1 = splat2(-c1);
            }
            float2 d0 = fadd2(x0, nc0), d1 = fadd2(x1, nc1);
            if (ODD && k == 3 && c == NC - 1) d0.y = d1.y = 0.f;
            sa0 = fadd2(sa0, d0);
            sa1 = fadd2(sa1, d1);
            q0 = ffma2(d0, d0, q0);
            q1 = ffma2(d1, d1, q1);
        });
        const float md0 = (sa0.x + sa0.y) * ep.inv_F, md1 = (sa1.x + sa1.y) * ep.inv_F;
        const float v0 = fmaxf(fmaf(-md0, md0, (q0.x + q0.y) * ep.inv_F), 0.f);
        const float v1 = fmaxf(fmaf(-md1, md1, (q1.x + q1.y) * ep.inv_F), 0.f);
        // F var <= (0.2 cF)^2 bounds every |X_f - mean| by 0.2 cF: every X_f - tau >= 0.8 cF > 0, no clamp in the sweep
        if (!__all_sync(kFull, fmaxf(v0, v1) <= ep.uni_var)) return false;
        tau[0] = c0 + md0 - ep.cF + ep.uni_k * v0;
        tau[1] = c1 + md1 - ep.cF + ep.uni_k * v1;
    }
    const float2 qm1 = splat2(ep.qm1);
#pragma unroll 1
    for (int it = 0; it < 6; ++it) {
        const float2 nt0 = splat2(-tau[0]), nt1 = splat2(-tau[1]);
        float2 s0 = make_float2(0.f, 0.f), s1 = s0, g0 = s0, g1 = s0;   // sums of u^q and u^(q-1)
        reset();
        tm_stream_rows2<NP, ODD>(t0, t1, [&](int c, int k, float2 x0, float2 x1) {
            float2 u0 = fadd2(x0, nt0), u1 = fadd2(x1, nt1);
            if (ODD && k == 3) {   // only the padded element (-inf) needs the clamp
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
            cross(c, k, fmul2(p0, vrow(0, c, k)), fmul2(p1, vrow(1, c, k)));   // armnet.py:36; normalised at the end
        });
        S[0] = s0.x + s0.y;
        S[1] = s1.x + s1.y;
        const float d0 = qnorm_newton_step(S[0], g0.x + g0.y, ep), d1 = qnorm_newton_step(S[1], g1.x + g1.y, ep);
        const float a0 = fabsf(d0), a1 = fabsf(d1);
        const float rel0 = 2.4e-7f * fabsf(tau[0]), rel1 = 2.4e-7f * fabsf(tau[1]);
        // |dp| <= q u^(q-1) |d| before the renormalisation: inside the parity budget (gates 2e-6 abs)
        if (__all_sync(kFull, a0 <= fmaxf(5e-7f, rel0) && a1 <= fmaxf(5e-7f, rel1))) return true;
        // a Newton step moves tau by a tiny fraction of cF here, the rows stay dense
        tau[0] += d0;
        tau[1] += d1;
    }
    return true;   // six self-correcting sweeps: the last accumulators are within the parity budget
}

}  // namespace armnet
