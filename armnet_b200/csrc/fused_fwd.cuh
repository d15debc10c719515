// Fused ARM-Net forward hot path for sm_100a:
//   value clamp -> embedding row gather -> attention logits -> alpha-entmax gates -> gates*values
//   -> log-space cross-feature product -> exp [-> eval-mode arm_bn]   (models/armnet.py:82-89, armnet_1h.py:81-85)
//
// Shape of the work.  A "row" is one (sample b, exponential neuron r = k*O + o).  A row needs the sample's F gathered
// embedding vectors e[f, 0:E] (shared by all R rows of the sample), its own column of the pre-contracted attention
// matrix M'[x, r] = (alpha-1) d_k^-0.5 sum_y W[k,x,y] Q[k,o,y], and its own row of the value matrix V[r, f];
// it produces E outputs.  The kernel is latency-/issue-bound (entmax), not HBM-bound: measured throughput scales
// linearly with resident warps, so the design maximises warps at a register budget that holds 2 x F logits.
//
// Mapping.  Persistent CTAs (one per SM) of NW symmetric warps; no CTA-wide barrier after start-up.
//   * tile  = SPG consecutive samples whose F embedding rows sit in one shared-memory slot of a ring (mbarriers
//     raw/full/empty per slot).  unit = the PPW = 32/ES row pairs one warp processes at a time; a tile has UPG units.
//     Warps take units in order from a shared counter and walk them independently of each other.
//   * gather: the warp that takes a tile's first unit (a) scales the rows of that tile, which landed long ago, by
//     their values so every consumer sees e = T[id]*v exactly as layers.py:21 computes it, and releases the tile on
//     the full mbarrier; then (b) issues the gather of the tile two ahead -- ids/values are read coalesced, values
//     clamped (and written back in place like the reference), rows fetched with TMA bulk copies (cp.async.bulk ->
//     UBLKCP, byte-counted on the slot's raw mbarrier) when they are 16-byte aligned, else with 8-/4-byte loads.
//   * compute: ES adjacent lanes own one PAIR of adjacent rows (ES = 1 when E <= 16).  Every e value read from shared
//     memory (warp-broadcast LDS.128) feeds both rows; M' and V come as float2 (one value per row of the pair) from
//     conflict-free odd-stride tables.  The 2 x F logits live in registers (thread-private: the entmax reductions
//     need no shuffles), tau comes from the solvers in entmax.cuh, the same registers feed s[x] = sum_f p_f V_f e[f,x].
//   * output: a unit's z rows are contiguous in global memory ([B, R, E] layout): they are staged in the warp's own
//     shared-memory buffer and written with one TMA bulk store per unit, or with direct stores when the 16-byte
//     granularity of bulk copies does not fit the shape.
#pragma once

#include "entmax_pair.cuh"

namespace armnet {

#ifndef ARMNET_MAX_WARPS
#define ARMNET_MAX_WARPS 12  // 12 warps -> 3 per SM sub-partition -> 168 registers per thread
#endif
constexpr int kMaxWarps = ARMNET_MAX_WARPS;
constexpr int kMaxThreads = kMaxWarps * 32;
constexpr int kMaxSlots = 24;
constexpr int kNR = 2;    // rows per thread
constexpr int kBwdWarps = 8;  // backward keeps two F-long register arrays: 8 warps -> 255 registers per thread

struct FwdParams {
    const void *ids;
    float *values;
    const float *table;
    const float2 *Mg2;        // [R2][MSTR]  pre-contracted attention matrix, row pairs, scale*(alpha-1) folded in
    const float2 *Vg2;        // [R2][VSTR]  att_values, row pairs
    const float *post_mean;   // [R] or null: eval-mode arm_bn, out = (z - mean) * scale + shift
    const float *post_scale;  // [R]
    const float *post_shift;  // [R]
    float *out_z;             // [B][R][E]
    float *out_tau;           // [B][R][2] or null
    float *out_p;             // [B][R][F] or null
    float *out_g;             // [B][R][F] or null
    float *out_s;             // [B][R][E] or null
    // backward (BWD kernel): inputs saved by the forward, per-row outputs, batch-accumulated parameter gradients
    const float *in_z;        // [B][R][E]  forward output (without the arm_bn epilogue)
    const float *in_dz;       // [B][R][E]  gradient w.r.t. z
    const float *in_tau;      // [B][R][2]  (tau, sum of unnormalised gates) saved by the forward
    float *out_wA;            // [B][F][R]  w = p * values                       (r fastest: coalesced, bmm-ready)
    float *out_gA;            // [B][F][R]  dg = gradient w.r.t. the logits g
    float *acc_dV;            // [R][F]     += p * dw  over the batch (caller zeroes)
    float *acc_dM;            // [R][E]     += sum_f dg_f e_f over the batch (caller zeroes)
    int *err_flag;
    long long V, ld, B;
    int F, E, R, R2;
    int R_out, r_off;  // out_z holds R_out rows per sample; this launch computes rows [r_off, r_off + R) (K*O tiling)
    int ids_i32;
    int clamp, clamp_inplace;
    float clamp_lo, clamp_hi;
    float g_unscale;  // 1 / (alpha-1): logits g = X * g_unscale for the validation output
    EntmaxParams ep;
    int SPG;         // samples per tile
    int UPG;         // warp-units per tile
    int NW;          // warps per CTA
    int n_slots;     // ring size
    int look;        // tiles of gather look-ahead
    int TPE;         // tiles per epoch (ids/values of one epoch are preloaded into shared memory)
    int n_tiles;
    int tma_gather;  // 1: rows fetched with cp.async.bulk of row_bytes each
    int row_bytes;
    int tma_store;   // 1: z staged in shared memory and written with cp.async.bulk
    int lockstep;    // 1: static unit assignment with a CTA barrier per round; 0: dynamic unit counter
    int tabFP, tabEL;  // armnet_fwd_mma_kernel: field / embedding-lane counts the pair tables in `workspace` were laid out for
};

__host__ __device__ constexpr int round_up_c(int x, int a) { return (x + a - 1) / a * a; }

// Shared-memory carve-up, identical on host (sizing) and device (pointers). Offsets in bytes.
struct SmemLayout {
    int off_bar, off_M, off_V, off_e, off_vals, off_ids, off_out, total;
    int m_bytes, v_bytes;  // bytes of the M / V pair tables (padded to 16 for the bulk copy)
    int mstr, vstr;    // float2 strides of the M / V pair tables (odd: conflict-free LDS.64 across a half-warp)
    int slot_floats;   // floats per e slot
    int rows_pad;      // ids / values entries per tile
    int out_floats;    // floats per warp staging buffer
    __host__ __device__ static int up(int x, int a) { return (x + a - 1) / a * a; }
    // bwd: the V region holds the dV accumulation table (V itself is read through L1), a dM table follows it, and
    // there is no output staging.
    int off_dM;
    __host__ __device__ SmemLayout(int FP, int E_lanes, int E_stride, int ES, const FwdParams &P, bool bwd = false) {
        mstr = E_lanes | 1;
        vstr = FP | 1;
        off_bar = 0;                                   // 3 * kMaxSlots + 1 mbarriers, then the unit counter
        off_M = up((3 * kMaxSlots + 1) * 8 + 16, 128);
        m_bytes = up(P.R2 * mstr * 8, 16);
        v_bytes = up(P.R2 * vstr * 8, 16);
        off_V = off_M + m_bytes;
        off_dM = off_V + v_bytes;
        off_e = up(off_dM + (bwd ? m_bytes : 0), 128);
        slot_floats = P.SPG * P.F * E_stride;
        rows_pad = up(P.SPG * P.F, 4);
        off_vals = up(off_e + P.n_slots * slot_floats * 4, 16);
        off_ids = up(off_vals + P.TPE * rows_pad * 4, 16);
        off_out = up(off_ids + P.TPE * rows_pad * 4, 128);
        out_floats = up((32 / ES) * kNR * P.E, 4);
        total = off_out + ((P.tma_store && !bwd) ? P.NW * out_floats * 4 : 0);
    }
};

// Phase boundary (logits | fused first pass | cross pass).  Each phase re-reads the e tile from shared memory.  Without
// a real fence between them nvcc/ptxas value-number the LDS across phases, keep F*E values alive through the solver
// and spill them to local memory (kilobytes of stack); a CTA-scope fence costs a few dozen cycles per unit.
__device__ __forceinline__ void phase_fence() { __threadfence_block(); }

// EC consecutive floats of one e row as EC/2 register pairs (x, x+1): float4 loads plus a float2 tail (EC even).
template <int EC>
__device__ __forceinline__ void load_e_chunk(const float *src, float2 (&e)[EC / 2]) {
#pragma unroll
    for (int j = 0; j < EC / 4; ++j) {
        const float4 t = reinterpret_cast<const float4 *>(src)[j];
        e[2 * j + 0] = make_float2(t.x, t.y);
        e[2 * j + 1] = make_float2(t.z, t.w);
    }
    if (EC % 4 == 2) e[EC / 2 - 1] = *reinterpret_cast<const float2 *>(src + (EC / 4) * 4);
}

// Final pass: gates of both rows at the solved tau, gates*values (armnet.py:36) and the log-space product
// s[x] += w_f e[f,x] (armnet.py:87).  acc[n][x/2] holds (s[x], s[x+1]) of row n: one FFMA2 per two embedding lanes.
template <int MODE, int FP, bool EXACT, int EC, int E_STRIDE>
__device__ __forceinline__ void cross_pass(const float2 (&X)[FP], float2 tau, const EntmaxParams &ep, const float *eb,
                                           const float2 *vrow, int F, float2 (&acc)[kNR][EC / 2], float2 &S) {
    const float2 nt = make_float2(-tau.x, -tau.y);
#pragma unroll
    for (int f = 0; f < FP; ++f) {
        if (EXACT || f < F) {
            float2 e[EC / 2];
            load_e_chunk<EC>(eb + f * E_STRIDE, e);
            const float2 p = gate_unnorm2<MODE>(X[f], nt, ep);
            S = fadd2(S, p);
            const float2 w = fmul2(p, vrow[f]);  // normalisation by S is applied once, after the sum
            const float2 w0 = splat2(w.x), w1 = splat2(w.y);
#pragma unroll
            for (int x = 0; x < EC / 2; ++x) {
                acc[0][x] = ffma2(w0, e[x], acc[0][x]);
                acc[1][x] = ffma2(w1, e[x], acc[1][x]);
            }
        }
    }
}

// Fused first pass for POW_GENERAL: gates, cross product AND the Newton residual sums at the same tau, so that a
// start which is already converged costs one sweep (2 MUFU per element) instead of a solver sweep plus a cross sweep.
template <int FP, bool EXACT, int EC, int E_STRIDE>
__device__ __forceinline__ void fused_first_pass(const float2 (&X)[FP], float2 tau, const EntmaxParams &ep,
                                                 const float *eb, const float2 *vrow, int F,
                                                 float2 (&acc)[kNR][EC / 2], float2 &S, float2 &S1) {
    const float2 nt = make_float2(-tau.x, -tau.y);
    const float2 qm1 = splat2(ep.qm1);
#pragma unroll
    for (int f = 0; f < FP; ++f) {
        if (EXACT || f < F) {
            float2 e[EC / 2];
            load_e_chunk<EC>(eb + f * E_STRIDE, e);
            const float2 u = relu2(fadd2(X[f], nt));
            const float2 t = fmul2(make_float2(fast_lg2(u.x), fast_lg2(u.y)), qm1);
            const float2 g = make_float2(fast_ex2(t.x), fast_ex2(t.y));  // u^(q-1)
            S1 = fadd2(S1, g);
            const float2 p = fmul2(g, u);  // u^q
            S = fadd2(S, p);
            const float2 w = fmul2(p, vrow[f]);
            const float2 w0 = splat2(w.x), w1 = splat2(w.y);
#pragma unroll
            for (int x = 0; x < EC / 2; ++x) {
                acc[0][x] = ffma2(w0, e[x], acc[0][x]);
                acc[1][x] = ffma2(w1, e[x], acc[1][x]);
            }
        }
    }
}

// Backward of one row pair (SURVEY.md 8a formulas; entmax.py:71-80 for the gate Jacobian).  Recomputes the gates
// from the saved (tau, S), then
//   ds = dz * z ;  dw_f = ds . e_f ;  w_f = p_f V_f -> out_wA ;  dV += p_f dw_f ;  dp_f = dw_f V_f
//   gppr_f = p_f^(2-alpha) [p_f > 0] ;  q = sum dp gppr / sum gppr ;  dg_f = (dp_f - q) gppr_f -> out_gA
//   dM'[x] += dg_f e_f[x]
// X comes in as (alpha-1) g of both rows and is overwritten.
template <int FP, bool EXACT, int EC, int ES, int E_STRIDE>
__device__ __forceinline__ void backward_pair(const FwdParams &P, float2 (&X)[FP], const float *eb, const float2 *vrow,
                                              float2 *dV_row, float *dM_row0, float *dM_row1, long long grow,
                                              long long b, int r0, bool has1, bool valid, int c) {
    const EntmaxParams &ep = P.ep;
    const int F = P.F, E = P.E, R = P.R;
    // saved threshold and normaliser
    float2 tau, S;
    tau.x = __ldg(P.in_tau + 2 * grow);
    S.x = __ldg(P.in_tau + 2 * grow + 1);
    tau.y = has1 ? __ldg(P.in_tau + 2 * grow + 2) : tau.x;
    S.y = has1 ? __ldg(P.in_tau + 2 * grow + 3) : S.x;
    const float2 invS = make_float2(__frcp_rn(S.x), __frcp_rn(S.y));
    const float2 nt = make_float2(-tau.x, -tau.y);
    // ds = dz * z on this lane's chunk of the embedding axis, as (x, x+1) pairs.  16-byte loads when the rows allow it
    // (nemb % 4 == 0): the kernel runs 8 warps per SM at 255 registers, so bytes in flight per load instruction matter.
    float2 ds[kNR][EC / 2];
    if (EC % 4 == 0 && E % 4 == 0) {
        const float4 *dz0 = reinterpret_cast<const float4 *>(P.in_dz + grow * E + c * EC);
        const float4 *z0 = reinterpret_cast<const float4 *>(P.in_z + grow * E + c * EC);
        const float4 *dz1 = reinterpret_cast<const float4 *>(P.in_dz + (grow + 1) * E + c * EC);
        const float4 *z1 = reinterpret_cast<const float4 *>(P.in_z + (grow + 1) * E + c * EC);
#pragma unroll
        for (int x4 = 0; x4 < EC / 4; ++x4) {
            float4 a = make_float4(0.f, 0.f, 0.f, 0.f), bz = a, a1 = a, bz1 = a;
            if (c * EC + 4 * x4 < E) {
                a = __ldg(dz0 + x4);
                bz = __ldg(z0 + x4);
                if (has1) {
                    a1 = __ldg(dz1 + x4);
                    bz1 = __ldg(z1 + x4);
                }
            }
            ds[0][2 * x4] = make_float2(a.x * bz.x, a.y * bz.y);
            ds[0][2 * x4 + 1] = make_float2(a.z * bz.z, a.w * bz.w);
            ds[1][2 * x4] = make_float2(a1.x * bz1.x, a1.y * bz1.y);
            ds[1][2 * x4 + 1] = make_float2(a1.z * bz1.z, a1.w * bz1.w);
        }
    } else {
#pragma unroll
        for (int x = 0; x < EC; ++x) {
            const int xx = c * EC + x;
            float d0 = 0.f, d1 = 0.f;
            if (xx < E) {
                d0 = __ldg(P.in_dz + grow * E + xx) * __ldg(P.in_z + grow * E + xx);
                if (has1) d1 = __ldg(P.in_dz + (grow + 1) * E + xx) * __ldg(P.in_z + (grow + 1) * E + xx);
            }
            if (x & 1) {
                ds[0][x / 2].y = d0;
                ds[1][x / 2].y = d1;
            } else {
                ds[0][x / 2].x = d0;
                ds[1][x / 2].x = d1;
            }
        }
    }
    // mode constants for gppr = p^(2-alpha)
    const float tma = 1.f - ep.am1;                     // 2 - alpha
    const float c2 = tma * ep.q;                        // p^(2-alpha) = ex2(c2 lg2(u) - (2-alpha) lg2(S))
    const float2 c3 = make_float2(-tma * fast_lg2(S.x), -tma * fast_lg2(S.y));
    const float2 rsq = make_float2(rsqrtf(S.x), rsqrtf(S.y));

    float2 G[FP];  // gppr; X is reused for t = dp * gppr
    float2 A = make_float2(0.f, 0.f), Bq = make_float2(0.f, 0.f);
#pragma unroll
    for (int f = 0; f < FP; ++f) {
        if (EXACT || f < F) {
            float2 e[EC / 2];
            load_e_chunk<EC>(eb + f * E_STRIDE, e);
            float2 p, g;  // normalised gate and its Jacobian factor
            if (ep.mode == POW_SOFTMAX) {
                p = fmul2(gate_unnorm2<POW_SOFTMAX>(X[f], nt, ep), invS);
                g = p;
            } else if (ep.mode == POW_LINEAR) {
                const float2 u = relu2(fadd2(X[f], nt));
                p = fmul2(u, invS);
                g = make_float2(u.x > 0.f ? 1.f : 0.f, u.y > 0.f ? 1.f : 0.f);
            } else if (ep.mode == POW_SQUARE) {
                const float2 u = relu2(fadd2(X[f], nt));
                p = fmul2(fmul2(u, u), invS);
                g = fmul2(u, rsq);
            } else {
                const float2 u = relu2(fadd2(X[f], nt));
                const float2 l = make_float2(fast_lg2(u.x), fast_lg2(u.y));
                const float2 t1 = fmul2(l, splat2(ep.q));
                p = fmul2(make_float2(fast_ex2(t1.x), fast_ex2(t1.y)), invS);
                const float2 t2 = ffma2(l, splat2(c2), c3);
                g = make_float2(u.x > 0.f ? fast_ex2(t2.x) : 0.f, u.y > 0.f ? fast_ex2(t2.y) : 0.f);
            }
            // dw_f = ds . e_f for both rows
            float2 a0 = make_float2(0.f, 0.f), a1 = make_float2(0.f, 0.f);
#pragma unroll
            for (int x = 0; x < EC / 2; ++x) {
                a0 = ffma2(e[x], ds[0][x], a0);
                a1 = ffma2(e[x], ds[1][x], a1);
            }
            float dw0 = a0.x + a0.y, dw1 = a1.x + a1.y;
#pragma unroll
            for (int m = 1; m < ES; m <<= 1) {
                dw0 += __shfl_xor_sync(0xffffffffu, dw0, m);
                dw1 += __shfl_xor_sync(0xffffffffu, dw1, m);
            }
            const float2 dw = make_float2(dw0, dw1);
            const float2 v = __ldg(vrow + f);
            if (valid && c == 0) {
                const float2 w = fmul2(p, v);  // armnet.py:36
                float *wdst = P.out_wA + (b * F + f) * (long long)R + r0;
                if (has1 && (R & 1) == 0) {
                    *reinterpret_cast<float2 *>(wdst) = w;      // r0 even, R even: 8-byte aligned
                } else {
                    wdst[0] = w.x;
                    if (has1) wdst[1] = w.y;
                }
                const float2 pv = fmul2(p, dw);  // d values (summed over the batch)
                atomicAdd(&dV_row[f].x, pv.x);
                if (has1) atomicAdd(&dV_row[f].y, pv.y);
            }
            const float2 t = fmul2(fmul2(dw, v), g);  // dp * gppr
            A = fadd2(A, t);
            Bq = fadd2(Bq, g);
            X[f] = t;
            G[f] = g;
        } else {
            G[f] = make_float2(0.f, 0.f);
        }
    }
    const float2 nq = make_float2(-__fdividef(A.x, Bq.x), -__fdividef(A.y, Bq.y));  // entmax.py:77
    phase_fence();
    float2 dM[kNR][EC / 2];
#pragma unroll
    for (int x = 0; x < EC / 2; ++x) dM[0][x] = dM[1][x] = make_float2(0.f, 0.f);
#pragma unroll
    for (int f = 0; f < FP; ++f) {
        if (EXACT || f < F) {
            const float2 dg = ffma2(nq, G[f], X[f]);  // dY*gppr - q*gppr (entmax.py:76,79)
            if (valid && c == 0) {
                float *gdst = P.out_gA + (b * F + f) * (long long)R + r0;
                if (has1 && (R & 1) == 0) {
                    *reinterpret_cast<float2 *>(gdst) = dg;
                } else {
                    gdst[0] = dg.x;
                    if (has1) gdst[1] = dg.y;
                }
            }
            float2 e[EC / 2];
            load_e_chunk<EC>(eb + f * E_STRIDE, e);
            const float2 g0 = splat2(dg.x), g1 = splat2(dg.y);
#pragma unroll
            for (int x = 0; x < EC / 2; ++x) {
                dM[0][x] = ffma2(g0, e[x], dM[0][x]);
                dM[1][x] = ffma2(g1, e[x], dM[1][x]);
            }
        }
    }
    if (valid) {
#pragma unroll
        for (int x = 0; x < EC / 2; ++x) {
            atomicAdd(dM_row0 + 2 * x, dM[0][x].x);
            atomicAdd(dM_row0 + 2 * x + 1, dM[0][x].y);
            if (has1) {
                atomicAdd(dM_row1 + 2 * x, dM[1][x].x);
                atomicAdd(dM_row1 + 2 * x + 1, dM[1][x].y);
            }
        }
    }
}

// Gather of one tile into its slot, by one whole warp; ids / clamped values come from the epoch's preloaded arrays.
template <int E_STRIDE>
__device__ __forceinline__ void issue_tile_gather(const FwdParams &P, const SmemLayout &L, int lane, long long tile,
                                                  int slot, const int *i_ep, const float *v_ep, float *es,
                                                  uint64_t *bar_raw, uint64_t *bar_full) {
    const int F = P.F, E = P.E;
    const long long b0 = tile * P.SPG;
    const int ts = (int)min((long long)P.SPG, P.B - b0);
    const int n = ts * F;
    float *e_st = es + slot * L.slot_floats;
    if (P.tma_gather) {
        if (lane == 0) mbar_arrive_expect_tx(&bar_raw[slot], (uint32_t)(n * P.row_bytes));
        __syncwarp();
        for (int idx = lane; idx < n; idx += 32)
            tma_load_bulk(e_st + idx * E_STRIDE, P.table + (long long)i_ep[idx] * P.ld, (uint32_t)P.row_bytes,
                          &bar_raw[slot]);
    } else {
        for (int idx = lane; idx < n; idx += 32) {
            const float v = v_ep[idx];
            const float *src = P.table + (long long)i_ep[idx] * P.ld;
            float *dst = e_st + idx * E_STRIDE;
            if (((E | (int)P.ld) & 1) == 0 && (reinterpret_cast<uintptr_t>(P.table) & 7) == 0) {
                for (int x = 0; x < E; x += 2) {
                    const float2 t = __ldg(reinterpret_cast<const float2 *>(src + x));
                    *reinterpret_cast<float2 *>(dst + x) = make_float2(__fmul_rn(t.x, v), __fmul_rn(t.y, v));
                }
            } else {
                for (int x = 0; x < E; ++x) dst[x] = __fmul_rn(__ldg(src + x), v);
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&bar_full[slot]);  // rows are already scaled: the tile is ready
    }
}

template <int FP, bool EXACT, int EC, int ES, bool BWD>
__global__ void __launch_bounds__(BWD ? kBwdWarps * 32 : kMaxThreads, 1)
    armnet_fwd_kernel(const __grid_constant__ FwdParams P) {
    static_assert(EC % 2 == 0, "EC must be even (float2/float4 shared-memory reads)");
    static_assert(ES == 1 || EC % 4 == 0, "split rows need 16-byte aligned chunks");
    static_assert(ES == 1 || ES == 2 || ES == 4 || ES == 8, "ES must be a power of two <= 8");
    constexpr int E_LANES = EC * ES;
    constexpr int E_STRIDE = round_up_c(E_LANES, 4);
    constexpr int PPW = 32 / ES;  // row pairs per warp-unit
    extern __shared__ __align__(128) unsigned char smem[];
    const SmemLayout L(FP, E_LANES, E_STRIDE, ES, P, BWD);
    uint64_t *bar_full = reinterpret_cast<uint64_t *>(smem + L.off_bar);
    uint64_t *bar_empty = bar_full + kMaxSlots;
    uint64_t *bar_raw = bar_empty + kMaxSlots;
    uint64_t *bar_par = bar_raw + kMaxSlots;
    int *next_unit = reinterpret_cast<int *>(bar_par + 1);
    float2 *Ms2 = reinterpret_cast<float2 *>(smem + L.off_M);
    float2 *Vs2 = reinterpret_cast<float2 *>(smem + L.off_V);
    float *es = reinterpret_cast<float *>(smem + L.off_e);
    float *vals = reinterpret_cast<float *>(smem + L.off_vals);
    int *idsm = reinterpret_cast<int *>(smem + L.off_ids);
    float *outs = reinterpret_cast<float *>(smem + L.off_out);

    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;
    const int NW = P.NW, NS = P.n_slots, UPG = P.UPG, LOOK = P.look;
    const int F = P.F, E = P.E, R = P.R, R2 = P.R2;
    const EntmaxParams ep = P.ep;

    if (tid == 0) {
        for (int s = 0; s < NS; ++s) {
            mbar_init(&bar_full[s], 1);
            mbar_init(&bar_raw[s], 1);
            mbar_init(&bar_empty[s], UPG);
        }
        mbar_init(bar_par, 1);
        mbar_fence_init();
        // the pre-contracted parameter tables arrive by TMA while the ids are being preloaded
        mbar_arrive_expect_tx(bar_par, (uint32_t)(L.m_bytes + (BWD ? 0 : L.v_bytes)));
        tma_load_bulk(Ms2, P.Mg2, (uint32_t)L.m_bytes, bar_par);
        if (!BWD) tma_load_bulk(Vs2, P.Vg2, (uint32_t)L.v_bytes, bar_par);
    }
    float *dMs = reinterpret_cast<float *>(smem + L.off_dM);  // BWD: [R2][2][E_LANES] (+pad), same layout as Ms2
    if (BWD) {  // Vs2 is the dV accumulation table in the backward kernel
        for (int i = tid; i < (L.v_bytes + L.m_bytes) / 4; i += blockDim.x) reinterpret_cast<float *>(Vs2)[i] = 0.f;
    }
    for (int i = tid; i < NS * L.slot_floats; i += blockDim.x) es[i] = 0.f;  // pad lanes stay zero for good
    fence_proxy_async_smem();

    // tiles of this CTA: blockIdx.x + i * gridDim.x, i in [0, n_local)
    const int n_local = (P.n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    const int rows_per_tile = P.SPG * F;

    const int c = lane % ES;   // which E-chunk of the pair this lane owns
    const int pl = lane / ES;  // pair slot inside the unit
    float *ost_base = outs + warp * L.out_floats;

    for (int ep0 = 0; ep0 < n_local; ep0 += P.TPE) {
    const int ep_tiles = min(P.TPE, n_local - ep0);
    // ---- epoch start: every warp is done with the previous epoch; preload ids / clamped values of this one
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
            if (P.clamp) {  // armnet.py:82 -- in place on the caller's tensor
                const float vc = fminf(fmaxf(v, P.clamp_lo), P.clamp_hi);
                if (P.clamp_inplace && vc != v) P.values[row] = vc;
                v = vc;
            }
            if ((unsigned long long)id >= (unsigned long long)P.V) {  // reference: IndexError (layers.py:20)
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
    if (ep0 == 0) mbar_wait(bar_par, 0);  // parameter tables have landed
    // prologue of the epoch: its first LOOK tiles are fetched right away, one warp per tile
    for (int t = warp; t < LOOK && t < ep_tiles; t += NW) {
        const int j = ep0 + t;
        const int sj = j % NS, use = j / NS;
        if (use > 0) mbar_wait(&bar_empty[sj], (uint32_t)(use - 1) & 1u);
        issue_tile_gather<E_STRIDE>(P, L, lane, (long long)blockIdx.x + (long long)j * gridDim.x, sj,
                                    idsm + t * L.rows_pad, vals + t * L.rows_pad, es, bar_raw, bar_full);
    }
    const int u_end = (ep0 + ep_tiles) * UPG;

    // Units are handed out in increasing order, either from a shared counter (a warp delayed by bookkeeping simply
    // takes fewer units) or round-robin with a CTA barrier per round (lockstep).
    const bool lockstep = P.lockstep != 0;
    for (int round = 0;; ++round) {
        int u;
        if (lockstep) {
            if (ep0 * UPG + round * NW >= u_end) break;
            named_bar_sync(1, NW * 32);
            u = ep0 * UPG + round * NW + warp;
            if (u >= u_end) continue;
        } else {
            u = 0;
            if (lane == 0) u = atomicAdd(next_unit, 1);
            u = __shfl_sync(0xffffffffu, u, 0);
            if (u >= u_end) break;
        }
        const int i = u / UPG;      // local tile index
        const int k = u - i * UPG;  // unit inside the tile
        const int slot = i % NS;
        const uint32_t ph = (uint32_t)(i / NS) & 1u;
        const long long tile = (long long)blockIdx.x + (long long)i * gridDim.x;
        const long long b0 = tile * P.SPG;
        const int ts = (int)min((long long)P.SPG, P.B - b0);

        if (k == 0) {
            // (a) own tile: rows landed -> e = row * v (layers.py:21) in place, then release it to every consumer
            if (P.tma_gather) {
                mbar_wait(&bar_raw[slot], ph);
                constexpr int C4 = E_STRIDE / 4;
                float4 *e4 = reinterpret_cast<float4 *>(es + slot * L.slot_floats);
                const float *v_st = vals + (i - ep0) * L.rows_pad;
                const int n = ts * F;
                for (int q = lane; q < n * C4; q += 32) {
                    const float v = v_st[q / C4];
                    float4 t = e4[q];
                    t.x = __fmul_rn(t.x, v);
                    t.y = __fmul_rn(t.y, v);
                    t.z = __fmul_rn(t.z, v);
                    t.w = __fmul_rn(t.w, v);
                    e4[q] = t;
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&bar_full[slot]);
            }
            // (b) fetch the tile LOOK ahead into its slot, once that slot's previous tile is fully consumed
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

        const int n_pairs = ts * R2;    // row pairs in this tile
        const int p_first = k * PPW;    // first pair of this unit
        const int pi = p_first + pl;
        const bool valid = pi < n_pairs;
        const int pic = valid ? pi : n_pairs - 1;  // idle lanes redo the last pair (votes/shuffles stay full-warp)
        const int bl = pic / R2;
        const int j2 = pic - bl * R2;
        const int r0 = 2 * j2;
        const bool has1 = (r0 + 1 < R);             // odd R: the last pair's second row is a dummy
        const long long grow = (b0 + bl) * R + r0;  // global index of the pair's first row
        const int e_off = slot * L.slot_floats + bl * F * E_STRIDE + c * EC;
        const float *eb = es + e_off;

        // ---- X = (alpha-1) * g = e . M' (armnet.py:33-34 and entmax.py:42; both scalings are folded into M').
        // X[f] = (row0, row1); the dot product over x runs as EC/2 FFMA2 on (even, odd) partial sums.
        float2 X[FP];
        {
            float2 Mr[kNR][EC / 2];
            const float2 *mrow = Ms2 + j2 * L.mstr + (c * EC) / 2;
#pragma unroll
            for (int x = 0; x < EC / 2; ++x) {
                Mr[0][x] = mrow[x];
                Mr[1][x] = mrow[E_LANES / 2 + x];
            }
#pragma unroll
            for (int f = 0; f < FP; ++f) {
                if (EXACT || f < F) {
                    float2 e[EC / 2];
                    load_e_chunk<EC>(eb + f * E_STRIDE, e);
                    float2 a0 = make_float2(0.f, 0.f), a1 = make_float2(0.f, 0.f);
#pragma unroll
                    for (int x = 0; x < EC / 2; ++x) {
                        a0 = ffma2(e[x], Mr[0][x], a0);
                        a1 = ffma2(e[x], Mr[1][x], a1);
                    }
                    float s0 = a0.x + a0.y, s1 = a1.x + a1.y;
#pragma unroll
                    for (int m = 1; m < ES; m <<= 1) {
                        s0 += __shfl_xor_sync(0xffffffffu, s0, m);
                        s1 += __shfl_xor_sync(0xffffffffu, s1, m);
                    }
                    X[f] = make_float2(s0, s1);
                } else {
                    X[f] = make_float2(neg_inf(), neg_inf());
                }
            }
        }
        if (P.out_g != nullptr && valid && c == 0) {  // validation output: g = X / (alpha-1)
#pragma unroll
            for (int f = 0; f < FP; ++f) {
                if (EXACT || f < F) {
                    P.out_g[grow * F + f] = X[f].x * P.g_unscale;
                    if (has1) P.out_g[(grow + 1) * F + f] = X[f].y * P.g_unscale;
                }
            }
        }
        phase_fence();

        if constexpr (BWD) {
            backward_pair<FP, EXACT, EC, ES, E_STRIDE>(P, X, eb, P.Vg2 + j2 * L.vstr, Vs2 + j2 * L.vstr,
                                                       dMs + j2 * (L.mstr * 2) + c * EC,
                                                       dMs + j2 * (L.mstr * 2) + E_LANES + c * EC, grow, b0 + bl, r0,
                                                       has1, valid, c);
            __syncwarp();
            if (lane == 0) mbar_arrive(&bar_empty[slot]);
            continue;
        }

        // ---- thresholds (entmax.py:44-61) and gates*values / log-space product (armnet.py:36,87), both rows together
        float2 tau;
        float2 acc[kNR][EC / 2];
#pragma unroll
        for (int x = 0; x < EC / 2; ++x) acc[0][x] = acc[1][x] = make_float2(0.f, 0.f);
        float2 S = make_float2(0.f, 0.f);
        const float2 *vrow = Vs2 + j2 * L.vstr;
        bool finished = false, warm = false;
        float2 mx, mean;
        row_max_mean2<FP, EXACT>(X, F, ep, mx, mean);
        // near-uniform rows (random-init weights, weakly attending neurons): closed-form start, verified by the Newton
        // residual of a fused first pass. The quick test costs nothing; the variance is only computed when it passes.
        if (ep.mode == POW_GENERAL &&
            __all_sync(0xffffffffu, fmaxf(mx.x - mean.x, mx.y - mean.y) <= 0.2f * ep.cF)) {
            if (__all_sync(0xffffffffu, entmax_uniform_start2<FP, EXACT>(X, F, ep, mx, mean, tau))) {
                float2 S1 = make_float2(0.f, 0.f);
                fused_first_pass<FP, EXACT, EC, E_STRIDE>(X, tau, ep, eb, vrow, F, acc, S, S1);
                const float d0 = __fdividef(S.x - 1.f, ep.q * S1.x);
                const float d1 = __fdividef(S.y - 1.f, ep.q * S1.y);
                if (__all_sync(0xffffffffu, fmaxf(fabsf(d0), fabsf(d1)) <= 1e-6f)) {
                    finished = true;  // |dp| <= q * 1e-6 before renormalisation: inside the parity budget
                } else {            // keep the Newton step, continue with the regular solver
                    tau.x += d0;
                    tau.y += d1;
                    warm = true;
#pragma unroll
                    for (int x = 0; x < EC / 2; ++x) acc[0][x] = acc[1][x] = make_float2(0.f, 0.f);
                    S = make_float2(0.f, 0.f);
                }
            }
        }
        if (!finished) {
            entmax_solve_tau2<FP, EXACT>(X, F, ep, mx, mean, tau, warm);
            phase_fence();
            const float *ebc = eb;
            switch (ep.mode) {
                case POW_SOFTMAX: cross_pass<POW_SOFTMAX, FP, EXACT, EC, E_STRIDE>(X, tau, ep, ebc, vrow, F, acc, S); break;
                case POW_LINEAR: cross_pass<POW_LINEAR, FP, EXACT, EC, E_STRIDE>(X, tau, ep, ebc, vrow, F, acc, S); break;
                case POW_SQUARE: cross_pass<POW_SQUARE, FP, EXACT, EC, E_STRIDE>(X, tau, ep, ebc, vrow, F, acc, S); break;
                default: cross_pass<POW_GENERAL, FP, EXACT, EC, E_STRIDE>(X, tau, ep, ebc, vrow, F, acc, S); break;
            }
        }
        // every lane is done reading the tile: hand the slot back (one arrival per unit)
        __syncwarp();
        if (lane == 0) mbar_arrive(&bar_empty[slot]);

        const float inv0 = __frcp_rn(S.x);  // entmax.py:63-64 renormalisation
        const float inv1 = __frcp_rn(S.y);

        if (valid && c == 0 && (P.out_tau != nullptr || P.out_p != nullptr)) {
            if (P.out_tau != nullptr) {
                P.out_tau[2 * grow + 0] = tau.x;
                P.out_tau[2 * grow + 1] = S.x;
                if (has1) {
                    P.out_tau[2 * grow + 2] = tau.y;
                    P.out_tau[2 * grow + 3] = S.y;
                }
            }
            if (P.out_p != nullptr) {
                const float2 nt = make_float2(-tau.x, -tau.y);
#pragma unroll
                for (int f = 0; f < FP; ++f) {  // static indices only: X must stay in registers
                    if (EXACT || f < F) {
                        const float2 pu = gate_unnorm2_rt(X[f], nt, ep);
                        P.out_p[grow * F + f] = __fdiv_rn(pu.x, S.x);
                        if (has1) P.out_p[(grow + 1) * F + f] = __fdiv_rn(pu.y, S.y);
                    }
                }
            }
        }
        // s = acc / S, z = exp(s) (armnet.py:86) [then eval-mode arm_bn, armnet.py:89].  exp(s) = ex2(acc * (log2(e) / S)):
        // one FMUL2 + two MUFU.EX2 per element pair (ex2.approx: relative error <= 2^-22, against a 1e-5 budget on z).
        float z[kNR][EC];
        if (P.out_s != nullptr && valid) {
#pragma unroll
            for (int x = 0; x < EC; ++x) {
                if (c * EC + x < E) {
                    const float a0 = (x & 1) ? acc[0][x / 2].y : acc[0][x / 2].x;
                    const float a1 = (x & 1) ? acc[1][x / 2].y : acc[1][x / 2].x;
                    P.out_s[grow * E + c * EC + x] = a0 * inv0;
                    if (has1) P.out_s[(grow + 1) * E + c * EC + x] = a1 * inv1;
                }
            }
        }
        {
            const float2 k0 = splat2(inv0 * 1.4426950408889634f), k1 = splat2(inv1 * 1.4426950408889634f);
#pragma unroll
            for (int x = 0; x < EC / 2; ++x) {
                const float2 t0 = fmul2(acc[0][x], k0);
                const float2 t1 = fmul2(acc[1][x], k1);
                z[0][2 * x] = fast_ex2(t0.x);
                z[0][2 * x + 1] = fast_ex2(t0.y);
                z[1][2 * x] = fast_ex2(t1.x);
                z[1][2 * x + 1] = fast_ex2(t1.y);
            }
        }
        if (P.post_scale != nullptr) {
            const int r1 = has1 ? r0 + 1 : r0;
            const float m0 = __ldg(P.post_mean + r0), a0 = __ldg(P.post_scale + r0), h0 = __ldg(P.post_shift + r0);
            const float m1 = __ldg(P.post_mean + r1), a1 = __ldg(P.post_scale + r1), h1 = __ldg(P.post_shift + r1);
#pragma unroll
            for (int x = 0; x < EC; ++x) {
                z[0][x] = fmaf(z[0][x] - m0, a0, h0);
                z[1][x] = fmaf(z[1][x] - m1, a1, h1);
            }
        }
        // the unit's rows [2*p_first, 2*min(p_first+PPW, n_pairs)) of the tile are contiguous in out_z when R is even
        const int pairs_here = min(PPW, n_pairs - p_first);
        float *gdst = P.out_z + (b0 * P.R_out + P.r_off + 2LL * p_first) * (long long)E;   // SPG == 1 whenever R < R_out
        const uint32_t bytes = (uint32_t)(pairs_here * kNR * E * 4);
        if (P.tma_store && ((bytes | (uint32_t)reinterpret_cast<uintptr_t>(gdst)) & 15u) == 0) {
            if (lane == 0) tma_store_wait_read<0>();  // this warp's previous bulk store has drained the buffer
            __syncwarp();
            float *ost = ost_base + pl * (kNR * E) + c * EC;
            if (valid) {
#pragma unroll
                for (int x = 0; x < EC; ++x) {
                    if (c * EC + x < E) {
                        ost[x] = z[0][x];
                        ost[E + x] = z[1][x];
                    }
                }
            }
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
                tma_store_bulk(gdst, ost_base, bytes);
                tma_store_commit();
            }
        } else if (valid) {
            float *dst = P.out_z + ((b0 + bl) * (long long)P.R_out + P.r_off + r0) * E + c * EC;
#pragma unroll
            for (int x = 0; x < EC; ++x) {
                if (c * EC + x < E) {
                    dst[x] = z[0][x];
                    if (has1) dst[E + x] = z[1][x];
                }
            }
        }
    }  // units of the epoch
    }  // epochs
    if constexpr (BWD) {
        // flush this CTA's dV / dM' tables into the batch accumulators (one atomic per entry per CTA)
        __syncthreads();
        for (int i = tid; i < R2 * L.vstr; i += blockDim.x) {
            const int j = i / L.vstr, f = i - j * L.vstr;
            if (f < F) {
                const float2 t = Vs2[i];
                atomicAdd(P.acc_dV + (long long)(2 * j) * F + f, t.x);
                if (2 * j + 1 < R) atomicAdd(P.acc_dV + (long long)(2 * j + 1) * F + f, t.y);
            }
        }
        for (int i = tid; i < R2 * 2 * E_LANES; i += blockDim.x) {
            const int j = i / (2 * E_LANES), rem = i - j * (2 * E_LANES);
            const int n = rem / E_LANES, x = rem - n * E_LANES;
            const int r = 2 * j + n;
            if (x < E && r < R) atomicAdd(P.acc_dM + (long long)r * E + x, dMs[j * (L.mstr * 2) + n * E_LANES + x]);
        }
    } else {
        if (P.tma_store && lane == 0) tma_store_wait_all<0>();
    }
}

// One compiled shape of the kernel (forward, and optionally backward).
struct FwdInstance {
    int FP;
    int exact;  // FP == F required
    int EC;
    int ES;
    const void *kernel;
    const void *kernel_bwd;  // null: no fused backward for this shape
};

#define ARMNET_FWD_INSTANCE(FP, EXACT, EC, ES) \
    { FP, EXACT, EC, ES, (const void *)&armnet_fwd_kernel<FP, (EXACT) != 0, EC, ES, false>, nullptr }
#define ARMNET_FWD_BWD_INSTANCE(FP, EXACT, EC, ES)                                            \
    { FP, EXACT, EC, ES, (const void *)&armnet_fwd_kernel<FP, (EXACT) != 0, EC, ES, false>, \
      (const void *)&armnet_fwd_kernel<FP, (EXACT) != 0, EC, ES, true> }

}  // namespace armnet
