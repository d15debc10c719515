// Fused ARM-Net forward, tensor-core formulation of the two per-sample E x F products (sm_100a).
//
// Same function, same ring / gather / store pipeline and the same FwdParams as armnet_fwd_kernel (fused_fwd.cuh); what
// changes is who multiplies.  armnet_fwd_kernel spends ~72 % of its FP32-pipe cycles in the FFMA2s of
//   X[r,f] = sum_x M'[x,r] e[f,x]      (attention logits, armnet.py:33-34)            and
//   s[r,x] = sum_f w[r,f] e[f,x]       (log-space cross product, armnet.py:86-87),
// both tiny per-sample GEMMs (16 rows x 40 fields x 10 lanes per warp step).  Here they run as warp-level
// mma.sync.m16n8k8 TF32 instructions with the 3xTF32 split (a_lo*b_hi + a_hi*b_lo + a_hi*b_hi: fp32-level accuracy, the
// parity budget on g / s / z is 1e-5 norm-relative), chained through registers like an attention kernel's P*V:
//   * logits: A = M'^T (rows = 16 exponential neurons), B = e^T of the sample, D[r, f].  A thread ends up with TWO rows
//     x 2 adjacent fields per 8-field block = 10 of a row's 40 logits; the entmax reductions over f are a thread-private
//     sum plus two quad shuffles, the element-wise work is packed f32x2 over the (f, f+1) pairs.
//   * cross: the D fragment of the logits (now gates*values) IS the A fragment of the second MMA when k-slot t maps to
//     field 8j+2t and k-slot t+4 to field 8j+2t+1; B = e with the same row permutation.  No shuffles, no shared memory.
//   * Warp-level mma.sync rather than tcgen05 for this step: the products are block-diagonal per sample (N = 40 columns
//     of one sample per 16 rows) and the gates between them need the logits in registers, so the whole chain stays in
//     one warp's registers with no TMEM round trip.  Measured issue rate on B200: 8.8 cycles per m16n8k8 per SM
//     sub-partition (tools/ubench/mma_rate.cu), overlapping with the FP32 pipe; the tensor pipe is ~45 % busy at nemb 16.
//     What this formulation cannot remove is the instruction stream around the MMAs (per-thread operand splits, quad
//     shuffles); the TMEM formulation with one row per thread is the next step and reuses this numerical core.
//   * nemb = 8*EK + ER: EK MMA steps cover 8 embedding lanes each, the ER (0 or 2) leftover lanes are done with FP32
//     FMAs (a whole MMA step for 2 of 8 lanes would double the tensor work at nemb = 10).
// The MMA row i of a 16-row step is neuron 2i (i < 8) / 2(i-8)+1, so a thread's two rows are an adjacent pair: the
// pre-contracted tables keep the row-pair layout of armnet_fwd_kernel (same workspace, same attn_prepare_kernel).
// Requirements (checked by the host): K*O % 64 == 0 (one sample per tile, whole units), F in (8(NT-1), 8NT],
// solver != literal bisection.  Everything else goes to armnet_fwd_kernel.
// Measured (profiles/r1_v7_summary.md): nemb 16, 39 fields: 172 us per 4096 x 512 rows vs 346 us for armnet_fwd_kernel
// (default there); nemb 10: 152 us vs 146 us (opt-in with ARMNET_MMA=1).
#pragma once

#include "fused_fwd.cuh"

namespace armnet {

#ifndef ARMNET_MMA_WARPS
#define ARMNET_MMA_WARPS 16  // 16 warps -> 128 registers per thread
#endif
constexpr int kMmaWarps = ARMNET_MMA_WARPS;

// x = hi + lo for the 3xTF32 products.  RAW: the tensor core reads only the top 19 bits of an operand register, so x
// itself serves as hi = trunc_tf32(x) and lo = x - trunc_tf32(x) costs one LOP3 + one FADD (|lo| < 2^-10 |x|, dropped
// lo*lo and the truncation of lo are ~2^-20 relative).  !RAW: hi = cvt.rna.tf32 (three integer/FP ops, |lo| <= 2^-11 |x|).
template <bool RAW>
__device__ __forceinline__ void split_tf32(float x, uint32_t &hi, uint32_t &lo) {
    if (RAW) {
        hi = __float_as_uint(x);
        lo = __float_as_uint(x - __uint_as_float(hi & 0xffffe000u));
    } else {
        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hi) : "f"(x));
        lo = __float_as_uint(x - __uint_as_float(hi));  // exact; the tensor core reads its top 19 bits
    }
}
// D(16x8) += A(16x8, row) * B(8x8, col).  lane = 4g + t:
//   a0 (row g, k t)  a1 (row g+8, k t)  a2 (row g, k t+4)  a3 (row g+8, k t+4);  b0 (k t, n g)  b1 (k t+4, n g);
//   d0 (row g, n 2t)  d1 (row g, n 2t+1)  d2 (row g+8, n 2t)  d3 (row g+8, n 2t+1).
__device__ __forceinline__ void mma_tf32(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                         uint32_t b1) {
    asm("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void mma_3xtf32(float (&d)[4], const uint32_t (&ah)[4], const uint32_t (&al)[4],
                                           const uint32_t (&bh)[2], const uint32_t (&bl)[2]) {
    mma_tf32(d, al[0], al[1], al[2], al[3], bh[0], bh[1]);
    mma_tf32(d, ah[0], ah[1], ah[2], ah[3], bl[0], bl[1]);
    mma_tf32(d, ah[0], ah[1], ah[2], ah[3], bh[0], bh[1]);
}
// Sum / max over the 4 lanes of a quad (they hold the 4 field-slices of one row pair); every lane gets the result, and
// the same bits (the butterfly adds the same two partial sums in either order).
__device__ __forceinline__ float quad_sum(float v) {
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    v += __shfl_xor_sync(0xffffffffu, v, 2);
    return v;
}
__device__ __forceinline__ float quad_max(float v) {
    v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 1));
    v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 2));
    return v;
}

// Unnormalised gates of this thread's slice of both rows at tau, and their row sums (entmax.py:61-64 numerators).
template <int MODE, int NT>
__device__ __forceinline__ void gates_at_tau(const float2 (&X)[2][NT], const float (&tau)[2], const EntmaxParams &ep,
                                             float2 (&G)[2][NT], float (&S)[2]) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const float2 nt = splat2(-tau[h]);
        float2 s = make_float2(0.f, 0.f);
#pragma unroll
        for (int j = 0; j < NT; ++j) {
            G[h][j] = gate_unnorm2<MODE>(X[h][j], nt, ep);
            s = fadd2(s, G[h][j]);
        }
        S[h] = quad_sum(s.x + s.y);
    }
}

// tau of both rows (same algorithms, constants and stopping rules as entmax_solve_tau2; utils/entmax.py:44-61).
// `warm`: tau holds a Newton iterate already.  All 32 lanes call this together.
template <int NT>
__device__ __forceinline__ void solve_tau_quad(const float2 (&X)[2][NT], const EntmaxParams &ep, const float (&mx)[2],
                                               const float (&mean)[2], float (&tau)[2], bool warm) {
    if (ep.mode == POW_SOFTMAX) {
        tau[0] = mx[0];
        tau[1] = mx[1];
        return;
    }
    if (!warm) {
        tau[0] = fmaxf(mx[0] - 1.f, mean[0] - ep.cF);
        tau[1] = fmaxf(mx[1] - 1.f, mean[1] - ep.cF);
    }
    constexpr int kMaxIt = 12;
    for (int it = 0; it < kMaxIt; ++it) {
        float num[2], den[2];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const float2 nt = splat2(-tau[h]);
            float2 s = make_float2(0.f, 0.f), s1 = make_float2(0.f, 0.f);
            if (ep.mode == POW_GENERAL) {
                const float2 qm1 = splat2(ep.qm1);
#pragma unroll
                for (int j = 0; j < NT; ++j) {
                    const float2 u = relu2(fadd2(X[h][j], nt));
                    const float2 l = fmul2(make_float2(fast_lg2(u.x), fast_lg2(u.y)), qm1);
                    const float2 w = make_float2(fast_ex2(l.x), fast_ex2(l.y));  // u^(q-1); u = 0 -> 0
                    s1 = fadd2(s1, w);
                    s = ffma2(w, u, s);
                }
                num[h] = quad_sum(s.x + s.y) - 1.f;
                den[h] = ep.q * quad_sum(s1.x + s1.y);
            } else if (ep.mode == POW_SQUARE) {
#pragma unroll
                for (int j = 0; j < NT; ++j) {
                    const float2 u = relu2(fadd2(X[h][j], nt));
                    s1 = fadd2(s1, u);
                    s = ffma2(u, u, s);
                }
                num[h] = quad_sum(s.x + s.y) - 1.f;
                den[h] = 2.f * quad_sum(s1.x + s1.y);
            } else {  // POW_LINEAR: Michelot
#pragma unroll
                for (int j = 0; j < NT; ++j) {
                    const float2 u = fadd2(X[h][j], nt);
                    s = fadd2(s, relu2(u));
                    s1.x += (u.x > 0.f) ? 1.f : 0.f;
                    s1.y += (u.y > 0.f) ? 1.f : 0.f;
                }
                num[h] = quad_sum(s.x + s.y) - 1.f;
                den[h] = quad_sum(s1.x + s1.y);
            }
        }
        float d0 = __fdividef(num[0], den[0]);
        float d1 = __fdividef(num[1], den[1]);
        if (!(den[0] > 0.f)) d0 = 0.f;
        if (!(den[1] > 0.f)) d1 = 0.f;
        // the step is applied even when it is the last: |f(tau+d)| = O(d^2), and the caller renormalises
        tau[0] += d0;
        tau[1] += d1;
        bool done;
        if (ep.mode == POW_LINEAR)
            done = fabsf(d0) <= 2.4e-7f * fmaxf(1.f, fabsf(tau[0])) && fabsf(d1) <= 2.4e-7f * fmaxf(1.f, fabsf(tau[1]));
        else
            done = fmaxf(fabsf(d0), fabsf(d1)) <= 2e-5f;
        if (__all_sync(0xffffffffu, done)) break;
    }
}

// DBG: the optional outputs (out_g / out_p / out_s / out_tau) are compiled in; the plain instance has no trace of them in
// its hot loop.  WARPS: CTA size the instance is compiled for (16 -> 128 registers per thread, 12 -> 168).
template <int NT, int EK, int ER, int E_STRIDE, bool RAW, bool DBG, int WARPS>
__global__ void __launch_bounds__(WARPS * 32, 1) armnet_fwd_mma_kernel(const __grid_constant__ FwdParams P) {
    static_assert(ER == 0 || ER == 2, "leftover embedding lanes: 0 or 2");
    static_assert(8 * EK + ER <= E_STRIDE, "rows are E_STRIDE floats apart");
    constexpr int MT = 4;  // 16-row MMA steps per warp-unit (a unit = 32 row pairs, as in armnet_fwd_kernel)
    extern __shared__ __align__(128) unsigned char smem[];
    const SmemLayout L(P.tabFP, P.tabEL, E_STRIDE, 1, P, false);
    uint64_t *bar_full = reinterpret_cast<uint64_t *>(smem + L.off_bar);
    uint64_t *bar_empty = bar_full + kMaxSlots;
    uint64_t *bar_raw = bar_empty + kMaxSlots;
    uint64_t *bar_par = bar_raw + kMaxSlots;
    int *next_unit = reinterpret_cast<int *>(bar_par + 1);
    const float *Mf = reinterpret_cast<const float *>(smem + L.off_M);
    const float2 *Vs2 = reinterpret_cast<const float2 *>(smem + L.off_V);
    float *es = reinterpret_cast<float *>(smem + L.off_e);
    float *vals = reinterpret_cast<float *>(smem + L.off_vals);
    int *idsm = reinterpret_cast<int *>(smem + L.off_ids);
    float *outs = reinterpret_cast<float *>(smem + L.off_out);

    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int NW = P.NW, NS = P.n_slots, UPG = P.UPG, LOOK = P.look;
    const int F = P.F, E = P.E, R = P.R;
    const int EL = P.tabEL;
    const EntmaxParams ep = P.ep;

    if (tid == 0) {
        for (int s = 0; s < NS; ++s) {
            mbar_init(&bar_full[s], 1);
            mbar_init(&bar_raw[s], 1);
            mbar_init(&bar_empty[s], UPG);
        }
        mbar_init(bar_par, 1);
        mbar_fence_init();
        mbar_arrive_expect_tx(bar_par, (uint32_t)(L.m_bytes + L.v_bytes));
        tma_load_bulk(smem + L.off_M, P.Mg2, (uint32_t)L.m_bytes, bar_par);
        tma_load_bulk(smem + L.off_V, P.Vg2, (uint32_t)L.v_bytes, bar_par);
    }
    for (int i = tid; i < NS * L.slot_floats; i += blockDim.x) es[i] = 0.f;  // pad lanes stay zero for good
    fence_proxy_async_smem();

    const int n_local = (P.n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    const int rows_per_tile = F;  // one sample per tile
    float *ost_base = outs + warp * L.out_floats;
    // fields of this thread's slice that are padding (only the last 8-field block can hold any)
    const bool vx = 8 * (NT - 1) + 2 * t < F;
    const bool vy = 8 * (NT - 1) + 2 * t + 1 < F;

    for (int ep0 = 0; ep0 < n_local; ep0 += P.TPE) {
    const int ep_tiles = min(P.TPE, n_local - ep0);
    // ---- epoch start: preload ids / clamped values (armnet.py:82 in place; layers.py:20 range check)
    __syncthreads();
    for (int q = tid; q < ep_tiles * rows_per_tile; q += blockDim.x) {
        const int tl = q / rows_per_tile, idx = q - tl * rows_per_tile;
        const long long row = ((long long)blockIdx.x + (long long)(ep0 + tl) * gridDim.x) * rows_per_tile + idx;
        long long id = 0;
        float v = 0.f;
        if (row < P.B * F) {
            id = P.ids_i32 ? (long long)reinterpret_cast<const int *>(P.ids)[row]
                           : reinterpret_cast<const long long *>(P.ids)[row];
            v = P.values[row];
            if (P.clamp) {
                const float vc = fminf(fmaxf(v, P.clamp_lo), P.clamp_hi);
                if (P.clamp_inplace && vc != v) P.values[row] = vc;
                v = vc;
            }
            if ((unsigned long long)id >= (unsigned long long)P.V) {
                if (P.err_flag) atomicOr(P.err_flag, 1);
                id = 0;
                v = 0.f;
            }
        }
        idsm[tl * L.rows_pad + idx] = (int)id;
        vals[tl * L.rows_pad + idx] = v;
    }
    if (tid == 0) *next_unit = ep0 * UPG;
    __syncthreads();
    if (ep0 == 0) mbar_wait(bar_par, 0);
    for (int tt = warp; tt < LOOK && tt < ep_tiles; tt += NW) {
        const int j = ep0 + tt;
        const int sj = j % NS, use = j / NS;
        if (use > 0) mbar_wait(&bar_empty[sj], (uint32_t)(use - 1) & 1u);
        issue_tile_gather<E_STRIDE>(P, L, lane, (long long)blockIdx.x + (long long)j * gridDim.x, sj,
                                    idsm + tt * L.rows_pad, vals + tt * L.rows_pad, es, bar_raw, bar_full);
    }
    const int u_end = (ep0 + ep_tiles) * UPG;

    for (;;) {
        int u = 0;
        if (lane == 0) u = atomicAdd(next_unit, 1);
        u = __shfl_sync(0xffffffffu, u, 0);
        if (u >= u_end) break;
        const int i = u / UPG;
        const int k = u - i * UPG;
        const int slot = i % NS;
        const uint32_t ph = (uint32_t)(i / NS) & 1u;
        const long long b = (long long)blockIdx.x + (long long)i * gridDim.x;  // the tile's sample

        if (k == 0) {
            if (P.tma_gather) {  // rows landed -> e = row * v (layers.py:21) in place, then release the tile
                mbar_wait(&bar_raw[slot], ph);
                constexpr int C4 = E_STRIDE / 4;
                float4 *e4 = reinterpret_cast<float4 *>(es + slot * L.slot_floats);
                const float *v_st = vals + (i - ep0) * L.rows_pad;
                for (int q = lane; q < F * C4; q += 32) {
                    const float v = v_st[q / C4];
                    float4 r4 = e4[q];
                    r4.x = __fmul_rn(r4.x, v);
                    r4.y = __fmul_rn(r4.y, v);
                    r4.z = __fmul_rn(r4.z, v);
                    r4.w = __fmul_rn(r4.w, v);
                    e4[q] = r4;
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&bar_full[slot]);
            }
            const int j = i + LOOK;
            if (j < ep0 + ep_tiles) {
                const int sj = j % NS;
                const int use = j / NS;
                if (use > 0) mbar_wait(&bar_empty[sj], (uint32_t)(use - 1) & 1u);
                issue_tile_gather<E_STRIDE>(P, L, lane, (long long)blockIdx.x + (long long)j * gridDim.x, sj,
                                            idsm + (j - ep0) * L.rows_pad, vals + (j - ep0) * L.rows_pad, es, bar_raw,
                                            bar_full);
            }
        }
        mbar_wait(&bar_full[slot], ph);

        // ---- the sample's e tile as MMA B fragments, split into TF32 hi / lo once per unit (4 MMA row steps use them)
        const float *eb = es + slot * L.slot_floats;
        uint32_t bh[NT][EK][2], bl[NT][EK][2];  // logits:  B[k = x][n = f]   rows f = 8j + g
        uint32_t ch[NT][EK][2], cl[NT][EK][2];  // cross:   B[k = f][n = x]   rows f = 8j + 2t, 8j + 2t + 1
        float2 ex8[NT], ex9[NT];                // leftover lanes x = 8 EK / 8 EK + 1 of the field pair (8j + 2t, 8j + 2t + 1)
#pragma unroll
        for (int j = 0; j < NT; ++j) {
            const float *rowB = eb + min(8 * j + g, F - 1) * E_STRIDE;  // padded fields re-read a real row; masked below
            const float *row0 = eb + min(8 * j + 2 * t, F - 1) * E_STRIDE;
            const float *row1 = eb + min(8 * j + 2 * t + 1, F - 1) * E_STRIDE;
#pragma unroll
            for (int ks = 0; ks < EK; ++ks) {
                split_tf32<RAW>(rowB[8 * ks + t], bh[j][ks][0], bl[j][ks][0]);
                split_tf32<RAW>(rowB[8 * ks + t + 4], bh[j][ks][1], bl[j][ks][1]);
                split_tf32<RAW>(row0[8 * ks + g], ch[j][ks][0], cl[j][ks][0]);
                split_tf32<RAW>(row1[8 * ks + g], ch[j][ks][1], cl[j][ks][1]);
            }
            if (ER == 2) {  // scalar loads straight into the (f0, f1) pairs the packed ops read: no re-pairing moves
                ex8[j] = make_float2(row0[8 * EK], row1[8 * EK]);
                ex9[j] = make_float2(row0[8 * EK + 1], row1[8 * EK + 1]);
            } else {
                ex8[j] = ex9[j] = make_float2(0.f, 0.f);
            }
        }
        // every lane holds what it needs of the tile in registers: hand the slot back (one arrival per unit)
        __syncwarp();
        if (lane == 0) mbar_arrive(&bar_empty[slot]);

        const bool stage = P.tma_store != 0;
        if (stage) {
            if (lane == 0) tma_store_wait_read<0>();  // this warp's previous bulk store has drained the buffer
            __syncwarp();
        }

#pragma unroll 1
        for (int m = 0; m < MT; ++m) {
            const int j2 = k * 32 + m * 8 + g;  // row pair of this thread inside the sample: rows r0, r0 + 1
            const int r0 = 2 * j2;
            const long long grow = b * R + r0;
            const float *mrow = Mf + j2 * (2 * L.mstr);  // [n][x]: M'[x][r0 + n]

            // ---- X = (alpha-1) g = M'^T e^T (armnet.py:33-34, entmax.py:42): c[j] = {X[r0][f0], X[r0][f1], X[r1][f0], X[r1][f1]}
            float c[NT][4];
            if (ER == 2) {
                const float2 m0 = *reinterpret_cast<const float2 *>(mrow + 8 * EK);
                const float2 m1 = *reinterpret_cast<const float2 *>(mrow + EL + 8 * EK);
                const float2 m08 = splat2(m0.x), m09 = splat2(m0.y), m18 = splat2(m1.x), m19 = splat2(m1.y);
#pragma unroll
                for (int j = 0; j < NT; ++j) {
                    const float2 x0 = ffma2(ex9[j], m09, fmul2(ex8[j], m08));  // row r0, fields f0 / f1
                    const float2 x1 = ffma2(ex9[j], m19, fmul2(ex8[j], m18));  // row r1
                    c[j][0] = x0.x;
                    c[j][1] = x0.y;
                    c[j][2] = x1.x;
                    c[j][3] = x1.y;
                }
            } else {
#pragma unroll
                for (int j = 0; j < NT; ++j) c[j][0] = c[j][1] = c[j][2] = c[j][3] = 0.f;
            }
#pragma unroll
            for (int ks = 0; ks < EK; ++ks) {
                uint32_t ah[4], al[4];
                split_tf32<RAW>(mrow[8 * ks + t], ah[0], al[0]);
                split_tf32<RAW>(mrow[EL + 8 * ks + t], ah[1], al[1]);
                split_tf32<RAW>(mrow[8 * ks + t + 4], ah[2], al[2]);
                split_tf32<RAW>(mrow[EL + 8 * ks + t + 4], ah[3], al[3]);
#pragma unroll
                for (int j = 0; j < NT; ++j) mma_3xtf32(c[j], ah, al, bh[j][ks], bl[j][ks]);
            }
            if (!vx) c[NT - 1][0] = c[NT - 1][2] = neg_inf();
            if (!vy) c[NT - 1][1] = c[NT - 1][3] = neg_inf();
            if (DBG && P.out_g != nullptr) {  // validation output: g = X / (alpha-1)
#pragma unroll
                for (int j = 0; j < NT; ++j) {
#pragma unroll
                    for (int cc = 0; cc < 2; ++cc) {
                        const int f = 8 * j + 2 * t + cc;
                        if (f < F) {
                            P.out_g[grow * F + f] = c[j][cc] * P.g_unscale;
                            P.out_g[(grow + 1) * F + f] = c[j][2 + cc] * P.g_unscale;
                        }
                    }
                }
            }
            float2 X[2][NT];
#pragma unroll
            for (int j = 0; j < NT; ++j) {
                X[0][j] = make_float2(c[j][0], c[j][1]);
                X[1][j] = make_float2(c[j][2], c[j][3]);
            }

            // ---- thresholds (entmax.py:44-61): row maxima / means over the real fields, then the solvers
            float mx[2], mean[2];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                float2 m2 = X[h][0];
                float2 s2 = make_float2(0.f, 0.f);
#pragma unroll
                for (int j = 1; j < NT; ++j) {
                    m2.x = fmaxf(m2.x, X[h][j].x);
                    m2.y = fmaxf(m2.y, X[h][j].y);
                }
#pragma unroll
                for (int j = 0; j < NT - 1; ++j) s2 = fadd2(s2, X[h][j]);
                s2 = fadd2(s2, make_float2(vx ? X[h][NT - 1].x : 0.f, vy ? X[h][NT - 1].y : 0.f));
                mx[h] = quad_max(fmaxf(m2.x, m2.y));
                mean[h] = quad_sum(s2.x + s2.y) * ep.inv_F;
            }
            float tau[2], S[2];
            float2 G[2][NT];  // unnormalised gates
            bool finished = false, warm = false;
            // near-uniform rows: closed-form start verified by a Newton residual (entmax.cuh: entmax_uniform_start)
            if (ep.mode == POW_GENERAL &&
                __all_sync(0xffffffffu, fmaxf(mx[0] - mean[0], mx[1] - mean[1]) <= 0.2f * ep.cF)) {
                float var[2];
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const float2 nmx = splat2(-mx[h]);
                    float2 sd2 = make_float2(0.f, 0.f);
#pragma unroll
                    for (int j = 0; j < NT - 1; ++j) {
                        const float2 d = fadd2(X[h][j], nmx);
                        sd2 = ffma2(d, d, sd2);
                    }
                    {
                        const float2 d = fadd2(X[h][NT - 1], nmx);
                        const float2 dm = make_float2(vx ? d.x : 0.f, vy ? d.y : 0.f);
                        sd2 = ffma2(dm, dm, sd2);
                    }
                    const float md = mean[h] - mx[h];
                    var[h] = fmaxf(fmaf(-md, md, quad_sum(sd2.x + sd2.y) * ep.inv_F), 0.f);
                    tau[h] = mean[h] - ep.cF + ep.uni_k * var[h];
                }
                if (__all_sync(0xffffffffu, fmaxf(var[0], var[1]) <= ep.uni_var)) {
                    float d[2];
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const float2 nt = splat2(-tau[h]);
                        const float2 qm1 = splat2(ep.qm1);
                        float2 s = make_float2(0.f, 0.f), s1 = make_float2(0.f, 0.f);
#pragma unroll
                        for (int j = 0; j < NT; ++j) {
                            const float2 uu = relu2(fadd2(X[h][j], nt));
                            const float2 l = fmul2(make_float2(fast_lg2(uu.x), fast_lg2(uu.y)), qm1);
                            const float2 w = make_float2(fast_ex2(l.x), fast_ex2(l.y));  // u^(q-1)
                            s1 = fadd2(s1, w);
                            G[h][j] = fmul2(w, uu);  // u^q
                            s = fadd2(s, G[h][j]);
                        }
                        S[h] = quad_sum(s.x + s.y);
                        const float S1 = quad_sum(s1.x + s1.y);
                        d[h] = __fdividef(S[h] - 1.f, ep.q * S1);
                    }
                    if (__all_sync(0xffffffffu, fmaxf(fabsf(d[0]), fabsf(d[1])) <= 1e-6f)) {
                        finished = true;  // |dp| <= q * 1e-6 before renormalisation: inside the parity budget
                    } else {
                        tau[0] += d[0];
                        tau[1] += d[1];
                        warm = true;
                    }
                }
            }
            if (!finished) {
                solve_tau_quad<NT>(X, ep, mx, mean, tau, warm);
                switch (ep.mode) {
                    case POW_SOFTMAX: gates_at_tau<POW_SOFTMAX, NT>(X, tau, ep, G, S); break;
                    case POW_LINEAR: gates_at_tau<POW_LINEAR, NT>(X, tau, ep, G, S); break;
                    case POW_SQUARE: gates_at_tau<POW_SQUARE, NT>(X, tau, ep, G, S); break;
                    default: gates_at_tau<POW_GENERAL, NT>(X, tau, ep, G, S); break;
                }
            }
            if (DBG && P.out_tau != nullptr && t == 0) {
                P.out_tau[2 * grow + 0] = tau[0];
                P.out_tau[2 * grow + 1] = S[0];
                P.out_tau[2 * grow + 2] = tau[1];
                P.out_tau[2 * grow + 3] = S[1];
            }
            if (DBG && P.out_p != nullptr) {
#pragma unroll
                for (int j = 0; j < NT; ++j) {
                    const int f = 8 * j + 2 * t;
                    if (f < F) {
                        P.out_p[grow * F + f] = __fdiv_rn(G[0][j].x, S[0]);
                        P.out_p[(grow + 1) * F + f] = __fdiv_rn(G[1][j].x, S[1]);
                    }
                    if (f + 1 < F) {
                        P.out_p[grow * F + f + 1] = __fdiv_rn(G[0][j].y, S[0]);
                        P.out_p[(grow + 1) * F + f + 1] = __fdiv_rn(G[1][j].y, S[1]);
                    }
                }
            }

            // ---- w = gates * values (armnet.py:36) as the A fragment; s = w e (armnet.py:87); normalised by S once
            const float2 *vrow = Vs2 + j2 * L.vstr;  // [f] -> (V[r0][f], V[r1][f])
            float acc[EK][4];
            float2 accr[2];     // (s[x = 8 EK], s[x = 8 EK + 1]) of rows r0 / r1
            float2 a8[2], a9[2];  // their partial sums over this thread's (f0, f1) pairs
#pragma unroll
            for (int n = 0; n < EK; ++n) acc[n][0] = acc[n][1] = acc[n][2] = acc[n][3] = 0.f;
            a8[0] = a8[1] = a9[0] = a9[1] = make_float2(0.f, 0.f);
#pragma unroll
            for (int j = 0; j < NT; ++j) {
                const float2 v0 = vrow[min(8 * j + 2 * t, F - 1)];
                const float2 v1 = vrow[min(8 * j + 2 * t + 1, F - 1)];
                const float w00 = G[0][j].x * v0.x, w01 = G[0][j].y * v1.x;  // row r0: fields f0, f1
                const float w10 = G[1][j].x * v0.y, w11 = G[1][j].y * v1.y;  // row r1
                uint32_t ah[4], al[4];
                split_tf32<RAW>(w00, ah[0], al[0]);
                split_tf32<RAW>(w10, ah[1], al[1]);
                split_tf32<RAW>(w01, ah[2], al[2]);
                split_tf32<RAW>(w11, ah[3], al[3]);
#pragma unroll
                for (int n = 0; n < EK; ++n) mma_3xtf32(acc[n], ah, al, ch[j][n], cl[j][n]);
                if (ER == 2) {
                    const float2 wr0 = make_float2(w00, w01), wr1 = make_float2(w10, w11);
                    a8[0] = ffma2(wr0, ex8[j], a8[0]);
                    a9[0] = ffma2(wr0, ex9[j], a9[0]);
                    a8[1] = ffma2(wr1, ex8[j], a8[1]);
                    a9[1] = ffma2(wr1, ex9[j], a9[1]);
                }
            }
            if (ER == 2) {
                accr[0] = make_float2(quad_sum(a8[0].x + a8[0].y), quad_sum(a9[0].x + a9[0].y));
                accr[1] = make_float2(quad_sum(a8[1].x + a8[1].y), quad_sum(a9[1].x + a9[1].y));
            } else {
                accr[0] = accr[1] = make_float2(0.f, 0.f);
            }

            // ---- s = acc / S, z = exp(s) (armnet.py:86) [then eval-mode arm_bn, armnet.py:89]
            const float inv0 = __frcp_rn(S[0]), inv1 = __frcp_rn(S[1]);  // entmax.py:63-64 renormalisation
            if (DBG && P.out_s != nullptr) {
#pragma unroll
                for (int n = 0; n < EK; ++n) {
                    float *d0 = P.out_s + grow * E + 8 * n + 2 * t;
                    d0[0] = acc[n][0] * inv0;
                    d0[1] = acc[n][1] * inv0;
                    d0[E] = acc[n][2] * inv1;
                    d0[E + 1] = acc[n][3] * inv1;
                }
                if (ER == 2 && t == 0) {
                    float *d0 = P.out_s + grow * E + 8 * EK;
                    d0[0] = accr[0].x * inv0;
                    d0[1] = accr[0].y * inv0;
                    d0[E] = accr[1].x * inv1;
                    d0[E + 1] = accr[1].y * inv1;
                }
            }
            const float k0 = inv0 * 1.4426950408889634f, k1 = inv1 * 1.4426950408889634f;
            float z[EK][4];
#pragma unroll
            for (int n = 0; n < EK; ++n) {
                z[n][0] = fast_ex2(acc[n][0] * k0);
                z[n][1] = fast_ex2(acc[n][1] * k0);
                z[n][2] = fast_ex2(acc[n][2] * k1);
                z[n][3] = fast_ex2(acc[n][3] * k1);
            }
            float2 zr[2];
            zr[0] = make_float2(fast_ex2(accr[0].x * k0), fast_ex2(accr[0].y * k0));
            zr[1] = make_float2(fast_ex2(accr[1].x * k1), fast_ex2(accr[1].y * k1));
            if (P.post_scale != nullptr) {
                const float pm0 = __ldg(P.post_mean + r0), pa0 = __ldg(P.post_scale + r0), ph0 = __ldg(P.post_shift + r0);
                const float pm1 = __ldg(P.post_mean + r0 + 1), pa1 = __ldg(P.post_scale + r0 + 1),
                            ph1 = __ldg(P.post_shift + r0 + 1);
#pragma unroll
                for (int n = 0; n < EK; ++n) {
                    z[n][0] = fmaf(z[n][0] - pm0, pa0, ph0);
                    z[n][1] = fmaf(z[n][1] - pm0, pa0, ph0);
                    z[n][2] = fmaf(z[n][2] - pm1, pa1, ph1);
                    z[n][3] = fmaf(z[n][3] - pm1, pa1, ph1);
                }
                zr[0].x = fmaf(zr[0].x - pm0, pa0, ph0);
                zr[0].y = fmaf(zr[0].y - pm0, pa0, ph0);
                zr[1].x = fmaf(zr[1].x - pm1, pa1, ph1);
                zr[1].y = fmaf(zr[1].y - pm1, pa1, ph1);
            }
            // rows r0, r0+1 of the unit's 64 contiguous output rows ([B, R, E] layout)
            float *dst = stage ? ost_base + (m * 16 + 2 * g) * E : P.out_z + grow * E;
            if (stage) {
#pragma unroll
                for (int n = 0; n < EK; ++n) {
                    *reinterpret_cast<float2 *>(dst + 8 * n + 2 * t) = make_float2(z[n][0], z[n][1]);
                    *reinterpret_cast<float2 *>(dst + E + 8 * n + 2 * t) = make_float2(z[n][2], z[n][3]);
                }
                if (ER == 2) {
                    if (t == 0) *reinterpret_cast<float2 *>(dst + 8 * EK) = zr[0];
                    if (t == 1) *reinterpret_cast<float2 *>(dst + E + 8 * EK) = zr[1];
                }
            } else {
#pragma unroll
                for (int n = 0; n < EK; ++n) {
                    dst[8 * n + 2 * t] = z[n][0];
                    dst[8 * n + 2 * t + 1] = z[n][1];
                    dst[E + 8 * n + 2 * t] = z[n][2];
                    dst[E + 8 * n + 2 * t + 1] = z[n][3];
                }
                if (ER == 2) {
                    if (t == 0) {
                        dst[8 * EK] = zr[0].x;
                        dst[8 * EK + 1] = zr[0].y;
                    }
                    if (t == 1) {
                        dst[E + 8 * EK] = zr[1].x;
                        dst[E + 8 * EK + 1] = zr[1].y;
                    }
                }
            }
        }  // MMA row steps of the unit
        if (stage) {
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
                tma_store_bulk(P.out_z + (b * R + 64LL * k) * (long long)E, ost_base, (uint32_t)(64 * E * 4));
                tma_store_commit();
            }
        }
    }  // units of the epoch
    }  // epochs
    if (P.tma_store && lane == 0) tma_store_wait_all<0>();
}

// One compiled shape of the tensor-core kernel.
struct MmaInstance {
    int NT, EK, ER, E_STRIDE;
    int default_on;            // 1: used without ARMNET_MMA=1 (set where it measured faster than armnet_fwd_kernel)
    const void *kernel16;      // 16 warps x 128 registers
    const void *kernel12;      // 12 warps x 168 registers (ARMNET_MMA_WARPS=12)
    const void *kernel_dbg;    // with the optional outputs (16 warps)
    const void *kernel_rna;    // cvt.rna.tf32 operand split (ARMNET_MMA_SPLIT=rna), with the optional outputs
};
#define ARMNET_MMA_INSTANCE(NT, EK, ER, ESTR, ON)                                                      \
    { NT, EK, ER, ESTR, ON, (const void *)&armnet_fwd_mma_kernel<NT, EK, ER, ESTR, true, false, 16>, \
      (const void *)&armnet_fwd_mma_kernel<NT, EK, ER, ESTR, true, false, 12>,                   \
      (const void *)&armnet_fwd_mma_kernel<NT, EK, ER, ESTR, true, true, 16>,                    \
      (const void *)&armnet_fwd_mma_kernel<NT, EK, ER, ESTR, false, true, 16> }

}  // namespace armnet
