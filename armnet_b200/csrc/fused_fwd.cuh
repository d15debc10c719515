// Fused ARM-Net forward hot path for sm_100a:
//   value clamp -> embedding row gather -> attention logits -> alpha-entmax gates -> gates*values
//   -> log-space cross-feature product -> exp [-> eval-mode arm_bn]   (models/armnet.py:82-89, armnet_1h.py:81-85)
//
// Shape of the work.  A "row" is one (sample b, exponential neuron r = k*O + o).  A row needs the sample's F gathered
// embedding vectors e[f, 0:E] (shared by all R rows of the sample), its own column of the pre-contracted attention
// matrix M[x, r] = (alpha-1) d_k^-0.5 sum_y W[k,x,y] Q[k,o,y], and its own column of the value matrix V[r, f];
// it produces E outputs.
//
// Mapping.  Persistent CTAs (one per SM), warp-specialised:
//   * 1 producer warp: per tile of TS samples it reads ids/values (coalesced), clamps the values (writing them back
//     in place like the reference), and gathers the TS*F embedding rows into a shared-memory stage -- with TMA bulk
//     copies (cp.async.bulk -> UBLKCP, completion on an mbarrier) when rows are 16-byte aligned, otherwise with
//     8-/4-byte loads -- then scales each row by its value so consumers see e = T[id]*v exactly as layers.py:21
//     computes it.  A 2-3 stage full/empty mbarrier ring decouples it from the consumers.
//   * NT <= 256 consumer threads.  ES adjacent lanes own one PAIR of adjacent rows (ES = 1 when E <= 16): every e value
//     read from shared memory (warp-broadcast LDS.128) feeds both rows, M and V come as float2 (one value per row of
//     the pair) from conflict-free odd-stride tables.  The 2 x F logits live in registers (thread-private, so the
//     entmax reductions need no shuffles), tau is found by the solvers in entmax.cuh, and the same registers feed the
//     cross product s[x] = sum_f p_f V_f e[f,x].  9 warps per SM -> 168 registers per thread, no spills.
//   * output: z rows of one pass are contiguous in global memory ([B, R, E] layout), so they are staged in shared
//     memory and written with one TMA bulk store per pass (double-buffered), or with direct stores when the
//     16-byte granularity of bulk copies does not fit the shape.
#pragma once

#include "entmax.cuh"

namespace armnet {

constexpr int kMaxConsumerThreads = 256;
constexpr int kProducerThreads = 32;
constexpr int kMaxThreads = kMaxConsumerThreads + kProducerThreads;
constexpr int kMaxStages = 3;
constexpr int kNR = 2;  // rows per thread

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
    int *err_flag;
    long long V, ld, B;
    int F, E, R, R2;
    int ids_i32;
    int clamp, clamp_inplace;
    float clamp_lo, clamp_hi;
    float g_unscale;  // 1 / (alpha-1): logits g = X * g_unscale for the validation output
    EntmaxParams ep;
    int TS;          // samples per tile
    int NT;          // consumer threads (multiple of 32)
    int n_tiles;
    int n_stages;    // 2 or 3
    int tma_gather;  // 1: rows fetched with cp.async.bulk of row_bytes each
    int row_bytes;
    int tma_store;   // 1: z written with cp.async.bulk from the staging buffers
};

__host__ __device__ constexpr int round_up_c(int x, int a) { return (x + a - 1) / a * a; }

// Shared-memory carve-up, identical on host (sizing) and device (pointers). Offsets in bytes.
struct SmemLayout {
    int off_bar, off_M, off_V, off_e, off_vals, off_out, total;
    int mstr, vstr;    // float2 strides of the M / V pair tables (odd: conflict-free LDS.64 across a half-warp)
    int stage_floats;  // floats per e stage
    int vals_floats;   // floats per value stage
    int out_floats;    // floats per output staging buffer
    __host__ __device__ static int up(int x, int a) { return (x + a - 1) / a * a; }
    __host__ __device__ SmemLayout(int FP, int E_lanes, int E_stride, int ES, const FwdParams &P) {
        mstr = E_lanes | 1;
        vstr = FP | 1;
        off_bar = 0;
        off_M = 128;
        off_V = up(off_M + P.R2 * mstr * 8, 16);
        off_e = up(off_V + P.R2 * vstr * 8, 128);
        stage_floats = P.TS * P.F * E_stride;
        off_vals = up(off_e + P.n_stages * stage_floats * 4, 16);
        vals_floats = up(P.TS * P.F, 4);
        off_out = up(off_vals + P.n_stages * vals_floats * 4, 128);
        out_floats = up((P.NT / ES) * kNR * P.E, 4);
        total = off_out + (P.tma_store ? 2 * out_floats * 4 : 0);
    }
};

// EC consecutive floats of one e row, as float4s plus a float2 tail (EC even).
template <int EC>
__device__ __forceinline__ void load_e_chunk(const float *src, float (&e)[EC]) {
#pragma unroll
    for (int j = 0; j < EC / 4; ++j) {
        const float4 t = reinterpret_cast<const float4 *>(src)[j];
        e[4 * j + 0] = t.x;
        e[4 * j + 1] = t.y;
        e[4 * j + 2] = t.z;
        e[4 * j + 3] = t.w;
    }
    if (EC % 4 == 2) {
        const float2 t = *reinterpret_cast<const float2 *>(src + (EC / 4) * 4);
        e[EC - 2] = t.x;
        e[EC - 1] = t.y;
    }
}

template <int MODE, int FP, bool EXACT, int EC, int E_STRIDE>
__device__ __forceinline__ void cross_pass(const float (&X)[kNR][FP], const float (&tau)[kNR], const EntmaxParams &ep,
                                           const float *eb, const float2 *vrow, int F, float (&acc)[kNR][EC],
                                           float (&S)[kNR]) {
#pragma unroll
    for (int f = 0; f < FP; ++f) {
        if (EXACT || f < F) {
            float e[EC];
            load_e_chunk<EC>(eb + f * E_STRIDE, e);
            const float2 v = vrow[f];
            const float p0 = gate_unnorm<MODE>(X[0][f], tau[0], ep);
            const float p1 = gate_unnorm<MODE>(X[1][f], tau[1], ep);
            S[0] += p0;
            S[1] += p1;
            const float w0 = p0 * v.x;  // armnet.py:36 (normalisation by S is applied once, after the sum)
            const float w1 = p1 * v.y;
#pragma unroll
            for (int x = 0; x < EC; ++x) {
                acc[0][x] = fmaf(w0, e[x], acc[0][x]);
                acc[1][x] = fmaf(w1, e[x], acc[1][x]);
            }
        }
    }
}

template <int FP, bool EXACT, int EC, int ES>
__global__ void __launch_bounds__(kMaxThreads, 1) armnet_fwd_kernel(const __grid_constant__ FwdParams P) {
    static_assert(EC % 2 == 0, "EC must be even (float2/float4 shared-memory reads)");
    static_assert(ES == 1 || EC % 4 == 0, "split rows need 16-byte aligned chunks");
    static_assert(ES == 1 || ES == 2 || ES == 4 || ES == 8, "ES must be a power of two <= 8");
    constexpr int E_LANES = EC * ES;
    constexpr int E_STRIDE = round_up_c(E_LANES, 4);
    extern __shared__ __align__(128) unsigned char smem[];
    const SmemLayout L(FP, E_LANES, E_STRIDE, ES, P);
    uint64_t *bar_full = reinterpret_cast<uint64_t *>(smem + L.off_bar);
    uint64_t *bar_empty = bar_full + kMaxStages;
    uint64_t *bar_raw = bar_empty + kMaxStages;
    float2 *Ms2 = reinterpret_cast<float2 *>(smem + L.off_M);
    float2 *Vs2 = reinterpret_cast<float2 *>(smem + L.off_V);
    float *es = reinterpret_cast<float *>(smem + L.off_e);
    float *vals = reinterpret_cast<float *>(smem + L.off_vals);
    float *outs = reinterpret_cast<float *>(smem + L.off_out);

    const int tid = threadIdx.x;
    const int NT = P.NT;
    const int F = P.F, E = P.E, R = P.R, R2 = P.R2;
    const EntmaxParams ep = P.ep;

    if (tid == 0) {
        for (int s = 0; s < kMaxStages; ++s) {
            mbar_init(&bar_full[s], 1);
            mbar_init(&bar_raw[s], 1);
            mbar_init(&bar_empty[s], NT / 32);
        }
        mbar_fence_init();
    }
    for (int i = tid; i < R2 * L.mstr; i += blockDim.x) Ms2[i] = P.Mg2[i];
    for (int i = tid; i < R2 * L.vstr; i += blockDim.x) Vs2[i] = P.Vg2[i];
    for (int i = tid; i < P.n_stages * L.stage_floats; i += blockDim.x) es[i] = 0.f;  // pad lanes stay zero for good
    fence_proxy_async_smem();
    __syncthreads();

    if (tid >= NT) {
        // ===================================================== producer warp
        const int lane = tid - NT;
        int it = 0;
        for (int tile = blockIdx.x; tile < P.n_tiles; tile += gridDim.x, ++it) {
            const int s = it % P.n_stages;
            const uint32_t ph = (uint32_t)(it / P.n_stages) & 1u;
            mbar_wait(&bar_empty[s], ph ^ 1u);
            const long long b0 = (long long)tile * P.TS;
            const int ts = (int)min((long long)P.TS, P.B - b0);
            const int n = ts * F;
            const long long base = b0 * F;
            float *e_st = es + s * L.stage_floats;
            float *v_st = vals + s * L.vals_floats;
            if (P.tma_gather && lane == 0) mbar_arrive_expect_tx(&bar_raw[s], (uint32_t)(n * P.row_bytes));
            __syncwarp();
            for (int idx = lane; idx < n; idx += 32) {
                long long id = P.ids_i32 ? (long long)reinterpret_cast<const int *>(P.ids)[base + idx]
                                         : reinterpret_cast<const long long *>(P.ids)[base + idx];
                float v = P.values[base + idx];
                if (P.clamp) {  // armnet.py:82 -- in place on the caller's tensor
                    const float vc = fminf(fmaxf(v, P.clamp_lo), P.clamp_hi);
                    if (P.clamp_inplace && vc != v) P.values[base + idx] = vc;
                    v = vc;
                }
                if ((unsigned long long)id >= (unsigned long long)P.V) {  // reference: IndexError (layers.py:20)
                    if (P.err_flag) atomicOr(P.err_flag, 1);
                    id = 0;
                    v = 0.f;
                }
                const float *src = P.table + id * P.ld;
                float *dst = e_st + idx * E_STRIDE;
                if (P.tma_gather) {
                    v_st[idx] = v;
                    tma_load_bulk(dst, src, (uint32_t)P.row_bytes, &bar_raw[s]);
                } else if (((E | (int)P.ld) & 1) == 0 && (reinterpret_cast<uintptr_t>(P.table) & 7) == 0) {
                    for (int x = 0; x < E; x += 2) {
                        const float2 t = __ldg(reinterpret_cast<const float2 *>(src + x));
                        *reinterpret_cast<float2 *>(dst + x) = make_float2(__fmul_rn(t.x, v), __fmul_rn(t.y, v));
                    }
                } else {
                    for (int x = 0; x < E; ++x) dst[x] = __fmul_rn(__ldg(src + x), v);
                }
            }
            if (P.tma_gather) {
                mbar_wait(&bar_raw[s], ph);
                __syncwarp();
                // e = row * v (layers.py:21), in place, float4 chunks of the rows that just landed
                constexpr int C4 = E_STRIDE / 4;
                float4 *e4 = reinterpret_cast<float4 *>(e_st);
                for (int j = lane; j < n * C4; j += 32) {
                    const float v = v_st[j / C4];
                    float4 t = e4[j];
                    t.x = __fmul_rn(t.x, v);
                    t.y = __fmul_rn(t.y, v);
                    t.z = __fmul_rn(t.z, v);
                    t.w = __fmul_rn(t.w, v);
                    e4[j] = t;
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&bar_full[s]);
        }
        return;
    }

    // ========================================================= consumer threads
    const int ppp = NT / ES;  // row pairs per pass
    const int c = tid % ES;   // which E-chunk of the pair this lane owns
    const int rl = tid / ES;  // pair slot inside the pass
    int it = 0;
    int ob = 0;
    for (int tile = blockIdx.x; tile < P.n_tiles; tile += gridDim.x, ++it) {
        const int s = it % P.n_stages;
        const uint32_t ph = (uint32_t)(it / P.n_stages) & 1u;
        const long long b0 = (long long)tile * P.TS;
        const int ts = (int)min((long long)P.TS, P.B - b0);
        const int n_pairs = ts * R2;
        const float *e_st = es + s * L.stage_floats;
        mbar_wait(&bar_full[s], ph);

        for (int pb = 0; pb < n_pairs; pb += ppp) {
            const int pi = pb + rl;
            const bool valid = pi < n_pairs;
            const int pic = valid ? pi : n_pairs - 1;  // idle lanes redo the last pair (votes/shuffles stay full-warp)
            const int bl = pic / R2;
            const int j = pic - bl * R2;
            const int r0 = 2 * j;
            const bool has1 = (r0 + 1 < R);             // odd R: the last pair's second row is a dummy
            const long long grow = (b0 + bl) * R + r0;  // global index of the pair's first row
            const float *eb = e_st + bl * F * E_STRIDE + c * EC;

            // ---- X = (alpha-1) * g = e . M' (armnet.py:33-34 and entmax.py:42; both scalings are folded into M')
            float X[kNR][FP];
            {
                float2 Mr[EC];
                const float2 *mrow = Ms2 + j * L.mstr + c * EC;
#pragma unroll
                for (int x = 0; x < EC; ++x) Mr[x] = mrow[x];
#pragma unroll
                for (int f = 0; f < FP; ++f) {
                    if (EXACT || f < F) {
                        float e[EC];
                        load_e_chunk<EC>(eb + f * E_STRIDE, e);
                        float a0 = 0.f, a1 = 0.f;
#pragma unroll
                        for (int x = 0; x < EC; ++x) {
                            a0 = fmaf(e[x], Mr[x].x, a0);
                            a1 = fmaf(e[x], Mr[x].y, a1);
                        }
#pragma unroll
                        for (int m = 1; m < ES; m <<= 1) {
                            a0 += __shfl_xor_sync(0xffffffffu, a0, m);
                            a1 += __shfl_xor_sync(0xffffffffu, a1, m);
                        }
                        X[0][f] = a0;
                        X[1][f] = a1;
                    } else {
                        X[0][f] = neg_inf();
                        X[1][f] = neg_inf();
                    }
                }
            }
            if (P.out_g != nullptr && valid && c == 0) {  // validation output: g = X / (alpha-1)
#pragma unroll
                for (int f = 0; f < FP; ++f) {
                    if (EXACT || f < F) {
                        P.out_g[grow * F + f] = X[0][f] * P.g_unscale;
                        if (has1) P.out_g[(grow + 1) * F + f] = X[1][f] * P.g_unscale;
                    }
                }
            }
            // The cross pass re-reads e from shared memory; without this compiler barrier nvcc keeps all loaded e
            // values alive across the solver (and spills them) instead of re-issuing 29-cycle LDS.
            asm volatile("" ::: "memory");

            // ---- thresholds (entmax.py:44-61), both rows of the pair together
            float tau[kNR];
            entmax_solve_tau<kNR, FP, EXACT>(X, F, ep, tau);
            asm volatile("" ::: "memory");

            // ---- gates, gates*values and the log-space product s = sum_f w_f e_f (armnet.py:36,87)
            float acc[kNR][EC];
#pragma unroll
            for (int x = 0; x < EC; ++x) acc[0][x] = acc[1][x] = 0.f;
            float S[kNR] = {0.f, 0.f};
            const float2 *vrow = Vs2 + j * L.vstr;
            switch (ep.mode) {
                case POW_SOFTMAX: cross_pass<POW_SOFTMAX, FP, EXACT, EC, E_STRIDE>(X, tau, ep, eb, vrow, F, acc, S); break;
                case POW_LINEAR: cross_pass<POW_LINEAR, FP, EXACT, EC, E_STRIDE>(X, tau, ep, eb, vrow, F, acc, S); break;
                case POW_SQUARE: cross_pass<POW_SQUARE, FP, EXACT, EC, E_STRIDE>(X, tau, ep, eb, vrow, F, acc, S); break;
                default: cross_pass<POW_GENERAL, FP, EXACT, EC, E_STRIDE>(X, tau, ep, eb, vrow, F, acc, S); break;
            }
            const float inv0 = __frcp_rn(S[0]);  // entmax.py:63-64 renormalisation
            const float inv1 = __frcp_rn(S[1]);

            if (valid && c == 0 && (P.out_tau != nullptr || P.out_p != nullptr)) {
                if (P.out_tau != nullptr) {
                    P.out_tau[2 * grow + 0] = tau[0];
                    P.out_tau[2 * grow + 1] = S[0];
                    if (has1) {
                        P.out_tau[2 * grow + 2] = tau[1];
                        P.out_tau[2 * grow + 3] = S[1];
                    }
                }
                if (P.out_p != nullptr) {
#pragma unroll
                    for (int f = 0; f < FP; ++f) {  // static indices only: X must stay in registers
                        if (EXACT || f < F) {
                            P.out_p[grow * F + f] = __fdiv_rn(gate_unnorm_rt(X[0][f], tau[0], ep), S[0]);
                            if (has1) P.out_p[(grow + 1) * F + f] = __fdiv_rn(gate_unnorm_rt(X[1][f], tau[1], ep), S[1]);
                        }
                    }
                }
            }
#pragma unroll
            for (int x = 0; x < EC; ++x) {
                acc[0][x] *= inv0;
                acc[1][x] *= inv1;
            }
            if (P.out_s != nullptr && valid) {
#pragma unroll
                for (int x = 0; x < EC; ++x) {
                    if (c * EC + x < E) {
                        P.out_s[grow * E + c * EC + x] = acc[0][x];
                        if (has1) P.out_s[(grow + 1) * E + c * EC + x] = acc[1][x];
                    }
                }
            }

            // ---- z = exp(s) (armnet.py:86) [then eval-mode arm_bn, armnet.py:89], written as [b][r][0:E]
#pragma unroll
            for (int x = 0; x < EC; ++x) {
                acc[0][x] = expf(acc[0][x]);
                acc[1][x] = expf(acc[1][x]);
            }
            if (P.post_scale != nullptr) {
                const int r1 = has1 ? r0 + 1 : r0;
                const float m0 = __ldg(P.post_mean + r0), a0 = __ldg(P.post_scale + r0), h0 = __ldg(P.post_shift + r0);
                const float m1 = __ldg(P.post_mean + r1), a1 = __ldg(P.post_scale + r1), h1 = __ldg(P.post_shift + r1);
#pragma unroll
                for (int x = 0; x < EC; ++x) {
                    acc[0][x] = fmaf(acc[0][x] - m0, a0, h0);
                    acc[1][x] = fmaf(acc[1][x] - m1, a1, h1);
                }
            }
            if (P.tma_store) {  // R is even here: the pair's 2E outputs are contiguous
                float *ost = outs + ob * L.out_floats + rl * (kNR * E) + c * EC;
                if (valid) {
#pragma unroll
                    for (int x = 0; x < EC; ++x) {
                        if (c * EC + x < E) {
                            ost[x] = acc[0][x];
                            ost[E + x] = acc[1][x];
                        }
                    }
                }
                if (tid == 0) tma_store_wait_read<0>();  // the buffer the NEXT pass writes is free again
                fence_proxy_async_smem();
                named_bar_sync(1, NT);
                if (tid == 0) {
                    const int pairs_here = min(ppp, n_pairs - pb);
                    tma_store_bulk(P.out_z + ((b0 * R + 2LL * pb) * (long long)E), outs + ob * L.out_floats,
                                   (uint32_t)(pairs_here * kNR * E * 4));
                    tma_store_commit();
                }
                ob ^= 1;
            } else if (valid) {
                float *dst = P.out_z + grow * E + c * EC;
#pragma unroll
                for (int x = 0; x < EC; ++x) {
                    if (c * EC + x < E) {
                        dst[x] = acc[0][x];
                        if (has1) dst[E + x] = acc[1][x];
                    }
                }
            }
        }
        __syncwarp();
        if ((tid & 31) == 0) mbar_arrive(&bar_empty[s]);
    }
    if (P.tma_store && tid == 0) tma_store_wait_all<0>();
}

// One compiled shape of the kernel.
struct FwdInstance {
    int FP;
    int exact;  // FP == F required
    int EC;
    int ES;
    const void *kernel;
};

#define ARMNET_FWD_INSTANCE(FP, EXACT, EC, ES) \
    { FP, EXACT, EC, ES, (const void *)&armnet_fwd_kernel<FP, (EXACT) != 0, EC, ES> }

}  // namespace armnet
