// Fused ARM-Net forward hot path for sm_100a:
//   value clamp -> embedding row gather -> attention logits -> alpha-entmax gates -> gates*values
//   -> log-space cross-feature product -> exp                      (models/armnet.py:82-87, armnet_1h.py:81-86)
//
// Shape of the work.  A "row" is one (sample b, exponential neuron r = k*O + o).  A row needs the sample's F gathered
// embedding vectors e[f, 0:E] (shared by all R rows of the sample), its own column of the pre-contracted attention
// matrix M[x, r] = sum_y W[k,x,y] Q[k,o,y], and its own column of the value matrix Vt[f, r]; it produces E outputs.
//
// Mapping.  Persistent CTAs (one per SM), warp-specialised:
//   * 1 producer warp: per tile of TS samples it reads ids/values (coalesced), clamps the values (writing them back
//     in place like the reference), and gathers the TS*F embedding rows into a shared-memory stage -- with TMA bulk
//     copies (cp.async.bulk -> UBLKCP, completion on an mbarrier) when rows are 16-byte aligned, otherwise with
//     8-/4-byte loads -- then scales each row by its value so consumers see e = T[id]*v exactly as layers.py:21
//     computes it.  A 3-stage full/empty mbarrier ring decouples it from the consumers.
//   * NT consumer threads: ES adjacent lanes own one row (ES = 1 when E <= 16).  The row's F logits live in
//     registers (thread-private, so the entmax reductions need no shuffles), tau is found by the solvers in
//     entmax.cuh, and the same registers feed the cross product s[x] = sum_f p_f V_f e[f,x]; e is read from shared
//     memory as warp-broadcast float4, M and Vt from conflict-free shared-memory columns.
//   * output: z rows of one pass are contiguous in global memory ([B, R, E] layout), so they are staged in shared
//     memory and written with one TMA bulk store per pass (double-buffered), or with direct stores when the
//     16-byte granularity of bulk copies does not fit the shape.
#pragma once

#include "entmax.cuh"

namespace armnet {

constexpr int kMaxConsumerThreads = 512;
constexpr int kProducerThreads = 32;
constexpr int kMaxThreads = kMaxConsumerThreads + kProducerThreads;
constexpr int kMaxStages = 3;

struct FwdParams {
    const void *ids;
    float *values;
    const float *table;
    const float *Mg;   // [E_pad][R]  pre-contracted attention matrix (rows >= E are zero)
    const float *Vtg;  // [F][R]      att_values transposed
    float *out_z;      // [B][R][E]
    float *out_tau;    // [B][R][2] or null
    float *out_p;      // [B][R][F] or null
    float *out_g;      // [B][R][F] or null
    float *out_s;      // [B][R][E] or null
    int *err_flag;
    long long V, ld, B;
    int F, E, R;
    int ids_i32;
    int clamp, clamp_inplace;
    float clamp_lo, clamp_hi;
    float scale;  // d_k^-0.5 as fp32 (armnet.py:15,34)
    EntmaxParams ep;
    int TS;          // samples per tile
    int NT;          // consumer threads (multiple of 32)
    int n_tiles;
    int n_stages;    // 2 or 3
    int tma_gather;  // 1: rows fetched with cp.async.bulk of row_bytes each
    int row_bytes;
    int tma_store;   // 1: z written with cp.async.bulk from the staging buffers
};

// Shared-memory carve-up, identical on host (sizing) and device (pointers). Offsets in bytes.
struct SmemLayout {
    int off_bar, off_M, off_V, off_e, off_vals, off_out, total;
    int stage_floats;  // floats per e stage
    int vals_floats;   // floats per value stage
    int out_floats;    // floats per output staging buffer
    __host__ __device__ static int up(int x, int a) { return (x + a - 1) / a * a; }
    __host__ __device__ SmemLayout(int FP, int E_pad, int ES, const FwdParams &P) {
        off_bar = 0;
        off_M = 128;
        off_V = up(off_M + E_pad * P.R * 4, 16);
        off_e = up(off_V + FP * P.R * 4, 128);
        stage_floats = P.TS * P.F * E_pad;
        off_vals = up(off_e + P.n_stages * stage_floats * 4, 16);
        vals_floats = up(P.TS * P.F, 4);
        off_out = up(off_vals + P.n_stages * vals_floats * 4, 128);
        out_floats = up((P.NT / ES) * P.E, 4);
        total = off_out + (P.tma_store ? 2 * out_floats * 4 : 0);
    }
};

template <int MODE, int FP, bool EXACT, int EC, int E_PAD>
__device__ __forceinline__ void cross_pass(const float (&X)[FP], float tau, const EntmaxParams &ep, const float *eb,
                                           const float *vcol, int R, int F, float (&acc)[EC], float &S) {
#pragma unroll
    for (int f = 0; f < FP; ++f) {
        if (EXACT || f < F) {
            const float pf = gate_unnorm<MODE>(X[f], tau, ep);
            S += pf;
            const float wv = pf * vcol[f * R];  // armnet.py:36 (normalisation by S is applied once, after the sum)
            const float4 *e4 = reinterpret_cast<const float4 *>(eb + f * E_PAD);
#pragma unroll
            for (int j = 0; j < EC / 4; ++j) {
                const float4 t = e4[j];
                acc[4 * j + 0] = fmaf(wv, t.x, acc[4 * j + 0]);
                acc[4 * j + 1] = fmaf(wv, t.y, acc[4 * j + 1]);
                acc[4 * j + 2] = fmaf(wv, t.z, acc[4 * j + 2]);
                acc[4 * j + 3] = fmaf(wv, t.w, acc[4 * j + 3]);
            }
        }
    }
}

template <int FP, bool EXACT, int EC, int ES>
__global__ void __launch_bounds__(kMaxThreads, 1) armnet_fwd_kernel(const __grid_constant__ FwdParams P) {
    static_assert(EC % 4 == 0, "EC must be a multiple of 4 (float4 shared-memory reads)");
    static_assert(ES == 1 || ES == 2 || ES == 4 || ES == 8, "ES must be a power of two <= 8");
    constexpr int E_PAD = EC * ES;
    extern __shared__ __align__(128) unsigned char smem[];
    const SmemLayout L(FP, E_PAD, ES, P);
    uint64_t *bar_full = reinterpret_cast<uint64_t *>(smem + L.off_bar);
    uint64_t *bar_empty = bar_full + kMaxStages;
    uint64_t *bar_raw = bar_empty + kMaxStages;
    float *Ms = reinterpret_cast<float *>(smem + L.off_M);
    float *Vs = reinterpret_cast<float *>(smem + L.off_V);
    float *es = reinterpret_cast<float *>(smem + L.off_e);
    float *vals = reinterpret_cast<float *>(smem + L.off_vals);
    float *outs = reinterpret_cast<float *>(smem + L.off_out);

    const int tid = threadIdx.x;
    const int NT = P.NT;
    const int F = P.F, E = P.E, R = P.R;
    const EntmaxParams ep = P.ep;

    if (tid == 0) {
        for (int s = 0; s < kMaxStages; ++s) {
            mbar_init(&bar_full[s], 1);
            mbar_init(&bar_raw[s], 1);
            mbar_init(&bar_empty[s], NT / 32);
        }
        mbar_fence_init();
    }
    for (int i = tid; i < E_PAD * R; i += blockDim.x) Ms[i] = P.Mg[i];
    for (int i = tid; i < FP * R; i += blockDim.x) Vs[i] = (i < F * R) ? P.Vtg[i] : 0.f;
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
                float *dst = e_st + idx * E_PAD;
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
                constexpr int C4 = E_PAD / 4;
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
    const int rpp = NT / ES;  // rows per pass
    const int c = tid % ES;   // which E-chunk of the row this lane owns
    const int rl = tid / ES;  // row slot inside the pass
    int it = 0;
    int ob = 0;
    for (int tile = blockIdx.x; tile < P.n_tiles; tile += gridDim.x, ++it) {
        const int s = it % P.n_stages;
        const uint32_t ph = (uint32_t)(it / P.n_stages) & 1u;
        const long long b0 = (long long)tile * P.TS;
        const int ts = (int)min((long long)P.TS, P.B - b0);
        const int n_rows = ts * R;
        const float *e_st = es + s * L.stage_floats;
        mbar_wait(&bar_full[s], ph);

        for (int pb = 0; pb < n_rows; pb += rpp) {
            const int ri = pb + rl;
            const bool valid = ri < n_rows;
            const int ric = valid ? ri : n_rows - 1;  // idle lanes redo the last row (votes/shuffles stay full-warp)
            const int bl = ric / R;
            const int r = ric - bl * R;
            const long long grow = (b0 + bl) * R + r;  // global row index
            const float *eb = e_st + bl * F * E_PAD + c * EC;

            float Mr[EC];
#pragma unroll
            for (int x = 0; x < EC; ++x) Mr[x] = Ms[(c * EC + x) * R + r];

            // ---- attention logits g = scale * e.M (armnet.py:33-34), X = (alpha-1) g (entmax.py:42)
            float X[FP];
#pragma unroll
            for (int f = 0; f < FP; ++f) {
                if (EXACT || f < F) {
                    const float4 *e4 = reinterpret_cast<const float4 *>(eb + f * E_PAD);
                    float a = 0.f;
#pragma unroll
                    for (int j = 0; j < EC / 4; ++j) {
                        const float4 t = e4[j];
                        a = fmaf(t.x, Mr[4 * j + 0], a);
                        a = fmaf(t.y, Mr[4 * j + 1], a);
                        a = fmaf(t.z, Mr[4 * j + 2], a);
                        a = fmaf(t.w, Mr[4 * j + 3], a);
                    }
#pragma unroll
                    for (int m = 1; m < ES; m <<= 1) a += __shfl_xor_sync(0xffffffffu, a, m);
                    const float g = a * P.scale;
                    if (P.out_g != nullptr && valid && c == 0) P.out_g[grow * F + f] = g;
                    X[f] = g * ep.am1;
                } else {
                    X[f] = neg_inf();
                }
            }

            // The cross pass re-reads e from shared memory; without this compiler barrier nvcc keeps all F*E_PAD
            // loaded values alive across the solver (and spills them) instead of re-issuing 29-cycle LDS.
            asm volatile("" ::: "memory");

            // ---- threshold (entmax.py:44-61)
            const float tau = entmax_solve_tau<FP, EXACT>(X, F, ep);
            asm volatile("" ::: "memory");

            // ---- gates, gates*values and the log-space product s = sum_f w_f e_f (armnet.py:36,87)
            float acc[EC];
#pragma unroll
            for (int x = 0; x < EC; ++x) acc[x] = 0.f;
            float S = 0.f;
            const float *vcol = Vs + r;
            switch (ep.mode) {
                case POW_SOFTMAX: cross_pass<POW_SOFTMAX, FP, EXACT, EC, E_PAD>(X, tau, ep, eb, vcol, R, F, acc, S); break;
                case POW_LINEAR: cross_pass<POW_LINEAR, FP, EXACT, EC, E_PAD>(X, tau, ep, eb, vcol, R, F, acc, S); break;
                case POW_SQUARE: cross_pass<POW_SQUARE, FP, EXACT, EC, E_PAD>(X, tau, ep, eb, vcol, R, F, acc, S); break;
                default: cross_pass<POW_GENERAL, FP, EXACT, EC, E_PAD>(X, tau, ep, eb, vcol, R, F, acc, S); break;
            }
            const float inv = __frcp_rn(S);  // entmax.py:63-64 renormalisation

            if (valid && c == 0) {
                if (P.out_tau != nullptr) {
                    P.out_tau[2 * grow + 0] = tau;
                    P.out_tau[2 * grow + 1] = S;
                }
                if (P.out_p != nullptr) {
#pragma unroll
                    for (int f = 0; f < FP; ++f)  // static indices only: X must stay in registers
                        if (EXACT || f < F) P.out_p[grow * F + f] = __fdiv_rn(gate_unnorm_rt(X[f], tau, ep), S);
                }
            }

            if (P.out_s != nullptr && valid) {
#pragma unroll
                for (int x = 0; x < EC; ++x)
                    if (c * EC + x < E) P.out_s[grow * E + c * EC + x] = acc[x] * inv;
            }

            // ---- z = exp(s) (armnet.py:86), written as [b][r][0:E]
            if (P.tma_store) {
                float *ost = outs + ob * L.out_floats + rl * E + c * EC;
                if (valid) {
#pragma unroll
                    for (int x = 0; x < EC; ++x)
                        if (c * EC + x < E) ost[x] = expf(acc[x] * inv);
                }
                if (tid == 0) tma_store_wait_read<0>();  // the buffer the NEXT pass writes is free again
                fence_proxy_async_smem();
                named_bar_sync(1, NT);
                if (tid == 0) {
                    const int rows_here = min(rpp, n_rows - pb);
                    tma_store_bulk(P.out_z + ((b0 * R + pb) * (long long)E), outs + ob * L.out_floats,
                                   (uint32_t)(rows_here * E * 4));
                    tma_store_commit();
                }
                ob ^= 1;
            } else if (valid) {
                float *dst = P.out_z + grow * E + c * EC;
#pragma unroll
                for (int x = 0; x < EC; ++x)
                    if (c * EC + x < E) dst[x] = expf(acc[x] * inv);
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
