// armnet_fused_fwd_f32: host side of the fused forward (instance selection, launch geometry, parameter
// pre-contraction).  Kernel: fused_fwd.cuh.  Instances: fused_fwd_inst_*.cu.
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "fused_fwd_mma.cuh"

namespace armnet {

// Tensor-core instances (fused_fwd_mma.cu).
extern const MmaInstance kMmaInstances[];
extern const int kNumMmaInstances;

// Tensor-memory kernel (fused_fwd_tmem.cu): logits on tcgen05, the default where the shape allows it.
bool tmem_shape_supported(int F, int E, int R);
size_t tmem_workspace_bytes(int F, int E, int R);
int tmem_prepare(const float *bilinear_w, const float *query, const float *att_values, int w_is_linear_layout, float am1,
                 int F, int E, int D, int K, int O, void *workspace, cudaStream_t st);
int tmem_launch(const char *who, const void *ids, int ids_i32, float *values, const float *table, int64_t V, int64_t ld,
                const EntmaxParams &ep, int64_t B, int F, int E, int K, int O, int clamp, float clamp_lo, float clamp_hi,
                int clamp_inplace, const float *post_mean, const float *post_scale, const float *post_shift, float *out_z,
                const void *workspace, int *err_flag, cudaStream_t st);

// Tables of compiled shapes, one per translation unit so they build in parallel.
extern const FwdInstance kFwdInstancesA[];
extern const int kNumFwdInstancesA;
extern const FwdInstance kFwdInstancesB[];
extern const int kNumFwdInstancesB;
extern const FwdInstance kFwdInstancesC[];
extern const int kNumFwdInstancesC;
extern const FwdInstance kFwdInstancesD[];
extern const int kNumFwdInstancesD;
extern const FwdInstance kFwdInstancesE[];
extern const int kNumFwdInstancesE;
extern const FwdInstance kFwdInstancesF[];
extern const int kNumFwdInstancesF;

static const FwdInstance *select_instance(int F, int E) {
    const FwdInstance *tabs[] = {kFwdInstancesA, kFwdInstancesB, kFwdInstancesC,
                                 kFwdInstancesD, kFwdInstancesE, kFwdInstancesF};
    const int ns[] = {kNumFwdInstancesA, kNumFwdInstancesB, kNumFwdInstancesC,
                      kNumFwdInstancesD, kNumFwdInstancesE, kNumFwdInstancesF};
    const FwdInstance *best = nullptr;
    long best_cost = 0;
    for (int t = 0; t < 6; ++t) {
        for (int i = 0; i < ns[t]; ++i) {
            const FwdInstance &I = tabs[t][i];
            if (I.exact ? (I.FP != F) : (I.FP < F)) continue;
            if (I.EC * I.ES < E) continue;
            // padded fields cost entmax passes, padded embedding lanes cost FMAs, splitting a row costs shuffles
            const long cost = (long)I.FP * 10000 + (long)I.EC * I.ES * 10 + I.ES - (I.exact ? 1 : 0);
            if (!best || cost < best_cost) {
                best = &I;
                best_cost = cost;
            }
        }
    }
    return best;
}

// Pre-contracted attention parameters in the row-pair layout the kernel reads (pair j = rows 2j, 2j+1):
//   Mg2 (floats) [j][n][x], pair stride 2*mstr floats: M'[x][2j+n] for x < E_lanes, n in {0,1}
//       M'[x][r] = (alpha-1) * d_k^-0.5 * sum_y W[k,x,y] Q[k,o,y]   (r = k*O + o)
//       armnet.py:33-34 'bfx,kxy,koy->bkof' * scale, contracted over y first; entmax.py:42 X = g*(alpha-1).
//       one-head: keys = e W_lin^T then keys.Q (armnet_1h.py:30-32), i.e. W[x,y] = W_lin[y,x].
//   Vg2 (float2) [j][f] = att_values[2j..2j+1][f]                                            (armnet.py:36)
// Entries with x >= E, f >= F or row >= R (and the stride padding) are zero.
__global__ void attn_prepare_kernel(const float *__restrict__ W, const float *__restrict__ Q,
                                    const float *__restrict__ Vals, int lin_layout, int F, int E, int D, int O, int R,
                                    int R2, int E_lanes, int mstr, int vstr, float scale, float am1,
                                    float *__restrict__ Mg2, float *__restrict__ Vg2) {
    const int nM = R2 * mstr * 2;
    const int total = nM + R2 * vstr * 2;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        if (i < nM) {
            const int j = i / (mstr * 2), rem = i - j * (mstr * 2);
            const int n = rem / E_lanes, x = rem - n * E_lanes;
            const int r = 2 * j + n;
            float a = 0.f;
            if (n < 2 && x < E && r < R) {
                const int k = r / O;
                const float *q = Q + (long long)r * D;
                if (lin_layout) {
                    for (int y = 0; y < D; ++y) a = fmaf(W[y * E + x], q[y], a);
                } else {
                    const float *w = W + ((long long)k * E + x) * D;
                    for (int y = 0; y < D; ++y) a = fmaf(w[y], q[y], a);
                }
                a = (a * scale) * am1;
            }
            Mg2[i] = a;
        } else {
            const int i2 = i - nM;
            const int half = i2 & 1, jf = i2 >> 1;
            const int j = jf / vstr, f = jf - j * vstr;
            const int r = 2 * j + half;
            Vg2[i2] = (f < F && r < R) ? Vals[(long long)r * F + f] : 0.f;
        }
    }
}

static inline int round_up(int x, int a) { return (x + a - 1) / a * a; }

// The tensor-core kernel for this shape, or null (see the requirements in fused_fwd_mma.cuh).  `I` is the
// armnet_fwd_kernel instance of the shape: the pair tables in the workspace are laid out for it.
static const MmaInstance *select_mma_instance(const FwdInstance *I, int F, int E, int R, int mode) {
    // ARMNET_MMA=1 / 0 forces it on / off; unset: the instance's own default (on where it measured faster than
    // armnet_fwd_kernel on B200; at C2a it is 4 % slower, 27.0 M vs 27.9 M samples/s: profiles/r1_v7_mma_experiment.md)
    const int force = tuning().mma;
    if (force == 0 || I->ES != 1 || R % 64 != 0 || mode == POW_BISECT) return nullptr;
    for (int i = 0; i < kNumMmaInstances; ++i) {
        const MmaInstance &M = kMmaInstances[i];
        if (F > 8 * (M.NT - 1) && F <= 8 * M.NT && E == 8 * M.EK + M.ER && round_up(I->EC, 4) == M.E_STRIDE)
            return (force == 1 || M.default_on) ? &M : nullptr;
    }
    return nullptr;
}

}  // namespace armnet

namespace armnet {
// Bytes of the pair tables armnet_fwd_kernel / armnet_fwd_mma_kernel read; the operands of armnet_fwd_tmem_kernel
// follow them in the same workspace when the shape has such an instance.
static size_t pair_tables_bytes(const FwdInstance *I, int K, int O) {
    const size_t R2 = ((size_t)K * O + 1) / 2;
    const size_t mstr = (size_t)(I->EC * I->ES) | 1, vstr = (size_t)I->FP | 1;
    return (R2 * mstr * 8 + 15) / 16 * 16 + (R2 * vstr * 8 + 15) / 16 * 16;  // bulk-copy granularity
}
// Does this call run on armnet_fwd_tmem_kernel?  (shape, solver, 16-byte table rows, no validation outputs)
static bool use_tmem_kernel(int F, int E, int R, int mode) {
    // default: only for a general alpha (MUFU-bound entmax), where it measured faster than armnet_fwd_kernel on B200
    // (C2a: 33.6 M vs 28.1 M samples/s); at alpha = 2 / 1.5 (no MUFU in the solver) armnet_fwd_kernel wins
    // (C2b: 49.5 M vs 41.5 M).  tuning "tmem" = 1 forces it wherever the shape allows, 0 switches it off.
    if (tuning().tmem == 0 || mode == POW_BISECT || !tmem_shape_supported(F, E, R)) return false;
    // nemb 11..16 (wide packed layout): opt-in until it measures faster than armnet_fwd_mma_kernel
    return tuning().tmem == 1 || (mode == POW_GENERAL && E <= 10);
}
}  // namespace armnet

extern "C" size_t armnet_fused_workspace_bytes(int F, int E, int K, int O) {
    using namespace armnet;
    if (F <= 0 || E <= 0 || K <= 0 || O <= 0) return 0;
    const FwdInstance *I = select_instance(F, E);
    if (!I) return 0;
    return pair_tables_bytes(I, K, O) + tmem_workspace_bytes(F, E, K * O);
}

namespace armnet {

struct BwdArgs {
    const float *z, *dz, *tau;
    float *wA, *gA, *dV, *dM;
};

// Shared by the forward and backward entries: validation, instance selection, geometry, shared-memory budget,
// parameter pre-contraction and the launch.
static int launch_fused(const char *who, const BwdArgs *bwd, const void *ids, int ids_i32, float *values,
                        const float *table, int64_t V, int64_t ld, const float *bilinear_w, const float *query,
                        const float *att_values, int w_is_linear_layout, float alpha, int solver, int n_iter, int64_t B,
                        int F, int E, int D, int K, int O, int clamp, float clamp_lo, float clamp_hi, int clamp_inplace,
                        const float *post_mean, const float *post_scale, const float *post_shift, float *out_z,
                        float *out_tau, float *out_p, float *out_g, float *out_s, void *workspace, int *err_flag,
                        void *stream, int mode = 0) {
    // mode 0: pre-contract the attention parameters into `workspace`, then run; 1: pre-contract only; 2: run only
    // (`workspace` was filled by an earlier mode-1 call with the same parameters, alpha and shape)
    note_launches(0);
    if (B == 0) return ARMNET_OK;  // empty batch: nothing to read or write (tensors may have null data pointers)
    if (!ids || !values || !table || (mode != 2 && (!bilinear_w || !query || !att_values)) || !workspace ||
        (!bwd && !out_z) ||
        (bwd && (!bwd->z || !bwd->dz || !bwd->tau || !bwd->wA || !bwd->gA || !bwd->dV || !bwd->dM))) {
        set_error("%s: null pointer", who);
        return ARMNET_ERR_NULL;
    }
    if (V <= 0 || B < 0 || F <= 0 || E <= 0 || D <= 0 || K <= 0 || O <= 0 || ld < E ||
        (w_is_linear_layout && K != 1) || B * (int64_t)K * O > (int64_t)1 << 40) {
        set_error("%s: bad shape V=%lld ld=%lld B=%lld F=%d E=%d D=%d K=%d O=%d", who, (long long)V, (long long)ld,
                  (long long)B, F, E, D, K, O);
        return ARMNET_ERR_SHAPE;
    }
    if ((uintptr_t)workspace % 16 || (uintptr_t)table % 4 || (uintptr_t)out_z % 4 || (uintptr_t)values % 4 ||
        (uintptr_t)ids % (ids_i32 ? 4 : 8)) {
        set_error("%s: misaligned pointer", who);
        return ARMNET_ERR_ALIGN;
    }
    if (V > 0x7fffffffLL) {
        set_error("%s: nfeat=%lld exceeds the 2^31-1 rows the kernel indexes", who, (long long)V);
        return ARMNET_ERR_UNSUPPORTED;
    }
    const FwdInstance *I = select_instance(F, E);
    if (!I || (bwd && !I->kernel_bwd)) {
        set_error("%s: no compiled kernel instance for F=%d E=%d (forward: F <= 64, E <= 128; backward: E <= 16 and "
                  "the BASELINE shapes in this build)", who, F, E);
        return ARMNET_ERR_UNSUPPORTED;
    }
    const void *kernel = bwd ? I->kernel_bwd : I->kernel;
    int max_warps = bwd ? kBwdWarps : kMaxWarps;
    FwdParams P;
    memset(&P, 0, sizeof(P));
    int rc = make_entmax_params(alpha, F, solver, n_iter, &P.ep);
    if (rc != ARMNET_OK) return rc;
    if (!bwd) {
        if (const MmaInstance *M = select_mma_instance(I, F, E, K * O, P.ep.mode)) {
            const bool dbg = out_tau || out_p || out_g || out_s;
            max_warps = kMmaWarps;
            if (tuning().mma_split_rna) {
                kernel = M->kernel_rna;
            } else if (dbg) {
                kernel = M->kernel_dbg;
            } else if (tuning().mma_warps == 12) {
                kernel = M->kernel12;
                max_warps = 12;
            } else {
                kernel = M->kernel16;
            }
        }
    }
    P.tabFP = I->FP;
    P.tabEL = I->EC * I->ES;
    DeviceInfo di;
    rc = get_device_info(&di);
    if (rc != ARMNET_OK) return rc;

    const int R = K * O;
    const int ES = I->ES;
    const int E_lanes = I->EC * ES;
    const int E_stride = round_up(E_lanes, 4);
    const int mstr = E_lanes | 1, vstr = I->FP | 1;
    const int R2_all = (R + 1) / 2;
    float *Mg2 = (float *)workspace;
    float *Vg2 = Mg2 + ((size_t)R2_all * mstr * 8 + 15) / 16 * 4;  // 16-byte padded M table, then V
    if ((post_scale != nullptr) != (post_mean != nullptr) || (post_scale != nullptr) != (post_shift != nullptr)) {
        set_error("%s: post_mean/post_scale/post_shift must be given together", who);
        return ARMNET_ERR_NULL;
    }

    P.ids = ids;
    P.values = values;
    P.table = table;
    P.out_z = out_z;
    P.out_tau = out_tau;
    P.out_p = out_p;
    P.out_g = out_g;
    P.out_s = out_s;
    if (bwd) {
        P.in_z = bwd->z;
        P.in_dz = bwd->dz;
        P.in_tau = bwd->tau;
        P.out_wA = bwd->wA;
        P.out_gA = bwd->gA;
        P.acc_dV = bwd->dV;
        P.acc_dM = bwd->dM;
    }
    P.err_flag = err_flag;
    P.V = V;
    P.ld = ld;
    P.B = B;
    P.F = F;
    P.E = E;
    P.ids_i32 = ids_i32;
    P.clamp = clamp;
    P.clamp_inplace = clamp_inplace;
    P.clamp_lo = clamp_lo;
    P.clamp_hi = clamp_hi;
    const float scale = (float)pow((double)D, -0.5);  // armnet.py:15 `d_k ** -0.5`, rounded to fp32 at use (:34)
    P.g_unscale = 1.f / P.ep.am1;
    P.R_out = R;

    cudaStream_t st = (cudaStream_t)stream;
    // armnet_fwd_tmem_kernel: logits on tcgen05 (fused_fwd_tmem.cuh).  Its operands live after the pair tables.
    void *tmem_ws = (char *)workspace + pair_tables_bytes(I, K, O);
    const bool tmem_ok = tmem_shape_supported(F, E, R) && P.ep.mode != POW_BISECT;
    const bool tmem_run = !bwd && use_tmem_kernel(F, E, R, P.ep.mode) && !out_tau && !out_p && !out_g && !out_s &&
                          (ld * 4) % 16 == 0 && (uintptr_t)table % 16 == 0 && ((E * 4 + 15) / 16 * 16) <= ld * 4;

    // ---- launch geometry of armnet_fwd_kernel / armnet_fwd_mma_kernel for the rows [r_off, r_off + Rc) of every sample.
    // Returns ARMNET_ERR_UNSUPPORTED (without an error text) when the tables of Rc rows do not fit shared memory.
    const int PPW = 32 / ES;  // row pairs per warp-unit
    unsigned grid = 0;
    int smem_total = 0;
    auto plan = [&](int r_off, int Rc) -> int {
        const int R2 = (Rc + 1) / 2;
        P.R = Rc;
        P.R2 = R2;
        P.r_off = r_off;
        P.Mg2 = (const float2 *)(Mg2 + (size_t)(r_off / 2) * mstr * 2);
        P.Vg2 = (const float2 *)(Vg2 + (size_t)(r_off / 2) * vstr * 2);
        P.post_mean = post_mean ? post_mean + r_off : nullptr;
        P.post_scale = post_scale ? post_scale + r_off : nullptr;
        P.post_shift = post_shift ? post_shift + r_off : nullptr;
        if (R2 >= PPW) {
            P.SPG = 1;
            P.UPG = (R2 + PPW - 1) / PPW;
        } else {
            P.SPG = PPW / R2;  // several whole samples per unit
            P.UPG = 1;
            long long cap = (B + di.sm_count - 1) / di.sm_count;  // keep every SM busy on small batches
            if (cap < 1) cap = 1;
            if (P.SPG > cap) P.SPG = (int)cap;
        }
        const long long n_tiles = (B + P.SPG - 1) / P.SPG;
        if (n_tiles > 0x7fffffffLL / P.UPG) {
            set_error("%s: batch too large", who);
            return ARMNET_ERR_SHAPE;
        }
        P.n_tiles = (int)n_tiles;
        grid = (unsigned)(n_tiles < di.sm_count ? n_tiles : di.sm_count);
        const long long units_per_cta = (n_tiles + grid - 1) / grid * P.UPG;
        P.NW = (int)(units_per_cta < max_warps ? units_per_cta : max_warps);
        if (tuning().force_nw >= 1 && tuning().force_nw <= max_warps) P.NW = tuning().force_nw;  // tuning experiments only
        P.lockstep = tuning().lockstep ? 1 : 0;  // default: units handed out dynamically

        // TMA eligibility
        P.row_bytes = round_up(E * 4, 16);
        P.tma_gather = ((ld * 4) % 16 == 0 && (uintptr_t)table % 16 == 0 && P.row_bytes <= ld * 4) ? 1 : 0;
        if (tuning().no_tma_gather) P.tma_gather = 0;
        // a unit's output rows are contiguous when pairs never straddle samples (R even); 16-byte size/alignment of each
        // bulk store is re-checked per unit in the kernel
        P.tma_store = (!bwd && !tuning().no_tma_store && Rc % 2 == 0 && (uintptr_t)out_z % 16 == 0) ? 1 : 0;

        // shared-memory budget: ids/values of an epoch of tiles are preloaded; the rest of the space becomes gather
        // slots (the deeper the ring, the further ahead the TMA gathers run)
        const int n_local_max = (int)((n_tiles + grid - 1) / grid);
        const int in_flight = (P.NW + P.UPG - 1) / P.UPG;  // tiles being consumed at once
        const int rows_pad = round_up(P.SPG * F, 4);
        P.TPE = n_local_max;
        if ((long long)P.TPE * rows_pad * 8 > 32 * 1024) P.TPE = (32 * 1024) / (rows_pad * 8);
        if (P.TPE < 1) P.TPE = 1;
        int best_slots = 0;
        for (int ns = kMaxSlots; ns >= in_flight + 2; --ns) {
            P.n_slots = ns;
            const SmemLayout Lt(I->FP, E_lanes, E_stride, ES, P, bwd != nullptr);
            if (Lt.total <= di.smem_optin) {
                best_slots = ns;
                break;
            }
        }
        if (best_slots == 0) {
            P.n_slots = in_flight + 2;
            const SmemLayout Lt(I->FP, E_lanes, E_stride, ES, P, bwd != nullptr);
            smem_total = Lt.total;
            return ARMNET_ERR_UNSUPPORTED;
        }
        if (best_slots > n_local_max + in_flight + 1) best_slots = n_local_max + in_flight + 1;  // no more than useful
        if (best_slots < in_flight + 2) best_slots = in_flight + 2;
        P.n_slots = best_slots;
        P.look = P.n_slots - in_flight - 1;
        if (tuning().force_look >= 1 && tuning().force_look <= P.look) P.look = tuning().force_look;  // experiments only
        const SmemLayout L(I->FP, E_lanes, E_stride, ES, P, bwd != nullptr);
        smem_total = L.total;
        return ARMNET_OK;
    };

    // K*O tiling: when the tables of all K*O rows do not fit one CTA's shared memory (large --h x --nattn_head), the plain
    // forward runs in passes over row chunks (each pass gathers the batch again: 2 KB per sample against 40 B per row of
    // output).  Validation outputs and the backward need the whole row range in one pass.
    int chunk = R;
    rc = tmem_run ? ARMNET_OK : plan(0, R);
    if (rc == ARMNET_ERR_UNSUPPORTED) {
        const bool can_tile = !bwd && !out_tau && !out_p && !out_g && !out_s;
        if (can_tile) {
            chunk = (R / 2) / 64 * 64;
            for (; chunk >= 64; chunk -= 64)
                if (plan(0, chunk) == ARMNET_OK) break;
        }
        if (!can_tile || chunk < 64) {
            set_error("%s: F=%d E=%d K*O=%d needs %d bytes of shared memory per CTA (limit %d)", who, F, E, R, smem_total,
                      di.smem_optin);
            return ARMNET_ERR_UNSUPPORTED;
        }
        kernel = I->kernel;          // the tiled passes always run on armnet_fwd_kernel
        max_warps = kMaxWarps;
    } else if (rc != ARMNET_OK) {
        return rc;
    }

    if (mode != 2) {
        int n_prep = 0;
        if (mode == 1 || !tmem_run) {  // the pair tables (a mode-1 workspace must serve every kernel kind)
            const int total = R2_all * (mstr + vstr) * 2;
            int blocks = (total + 255) / 256;
            if (blocks > di.sm_count * 4) blocks = di.sm_count * 4;
            attn_prepare_kernel<<<blocks, 256, 0, st>>>(bilinear_w, query, att_values, w_is_linear_layout, F, E, D, O, R,
                                                         R2_all, E_lanes, mstr, vstr, scale, P.ep.am1, Mg2, Vg2);
            ARMNET_CUDA_TRY(cudaGetLastError());
            ++n_prep;
        }
        if ((mode == 1 && tmem_ok) || tmem_run) {
            rc = tmem_prepare(bilinear_w, query, att_values, w_is_linear_layout, P.ep.am1, F, E, D, K, O, tmem_ws, st);
            if (rc != ARMNET_OK) return rc;
            ++n_prep;
        }
        if (mode == 1) {
            note_launches(n_prep);
            return ARMNET_OK;
        }
    }
    if (tmem_run) {
        rc = tmem_launch(who, ids, ids_i32, values, table, V, ld, P.ep, B, F, E, K, O, clamp, clamp_lo, clamp_hi,
                         clamp_inplace, post_mean, post_scale, post_shift, out_z, tmem_ws, err_flag, st);
        if (rc != ARMNET_OK) return rc;
        note_launches(mode == 2 ? 1 : 2);
        return ARMNET_OK;
    }
    ARMNET_CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, di.smem_optin));
    int n_launch = 0;
    for (int r_off = 0; r_off < R; r_off += chunk) {
        const int Rc = (R - r_off < chunk) ? R - r_off : chunk;
        rc = plan(r_off, Rc);
        if (rc != ARMNET_OK) {
            if (rc == ARMNET_ERR_UNSUPPORTED)
                set_error("%s: F=%d E=%d rows %d..%d need %d bytes of shared memory per CTA (limit %d)", who, F, E, r_off,
                          r_off + Rc, smem_total, di.smem_optin);
            return rc;
        }
        if (r_off > 0) P.clamp_inplace = 0;   // the values were clamped (and written back) by the first pass
        void *args[] = {(void *)&P};
        ARMNET_CUDA_TRY(cudaLaunchKernel(kernel, dim3(grid), dim3(P.NW * 32), args, (size_t)smem_total, st));
        ++n_launch;
    }
    note_launches((mode == 2 ? 0 : 1) + n_launch);
    return ARMNET_OK;
}

}  // namespace armnet

extern "C" int armnet_fused_fwd_f32(const void *ids, int ids_i32, float *values, const float *table, int64_t V,
                                    int64_t ld, const float *bilinear_w, const float *query, const float *att_values,
                                    int w_is_linear_layout, float alpha, int solver, int n_iter, int64_t B, int F,
                                    int E, int D, int K, int O, int clamp, float clamp_lo, float clamp_hi,
                                    int clamp_inplace, const float *post_mean, const float *post_scale,
                                    const float *post_shift, float *out_z, float *out_tau, float *out_p, float *out_g,
                                    float *out_s, void *workspace, int *err_flag, void *stream) {
    return armnet::launch_fused("fused_fwd", nullptr, ids, ids_i32, values, table, V, ld, bilinear_w, query, att_values,
                                w_is_linear_layout, alpha, solver, n_iter, B, F, E, D, K, O, clamp, clamp_lo, clamp_hi,
                                clamp_inplace, post_mean, post_scale, post_shift, out_z, out_tau, out_p, out_g, out_s,
                                workspace, err_flag, stream);
}

extern "C" int armnet_fused_prepare_f32(const float *bilinear_w, const float *query, const float *att_values,
                                        int w_is_linear_layout, float alpha, int F, int E, int D, int K, int O,
                                        void *workspace, void *stream) {
    // the batch arguments are not touched in mode 1; `workspace` stands in for them to pass the pointer checks
    float *w = static_cast<float *>(workspace);
    return armnet::launch_fused("fused_prepare", nullptr, workspace, 1, w, w, 1, E, bilinear_w, query, att_values,
                                w_is_linear_layout, alpha, ARMNET_SOLVER_AUTO, 50, 1, F, E, D, K, O, 0, 0.f, 0.f, 0, nullptr,
                                nullptr, nullptr, w, nullptr, nullptr, nullptr, nullptr, workspace, nullptr, stream, 1);
}

extern "C" int armnet_fused_fwd_prepared_f32(const void *ids, int ids_i32, float *values, const float *table, int64_t V,
                                             int64_t ld, float alpha, int solver, int n_iter, int64_t B, int F, int E, int D,
                                             int K, int O, int clamp, float clamp_lo, float clamp_hi, int clamp_inplace,
                                             const float *post_mean, const float *post_scale, const float *post_shift,
                                             float *out_z, float *out_tau, float *out_p, float *out_g, float *out_s,
                                             const void *workspace, int *err_flag, void *stream) {
    return armnet::launch_fused("fused_fwd_prepared", nullptr, ids, ids_i32, values, table, V, ld, nullptr, nullptr, nullptr,
                                0, alpha, solver, n_iter, B, F, E, D, K, O, clamp, clamp_lo, clamp_hi,
                                clamp_inplace, post_mean, post_scale, post_shift, out_z, out_tau, out_p, out_g, out_s,
                                const_cast<void *>(workspace), err_flag, stream, 2);
}

extern "C" int armnet_fused_bwd_supported(int F, int E) {
    const armnet::FwdInstance *I = armnet::select_instance(F, E);
    return (I && I->kernel_bwd) ? 1 : 0;
}

extern "C" int armnet_fused_fwd_kernel_kind(int F, int E, int K, int O, float alpha, int solver) {
    using namespace armnet;
    if (F <= 0 || E <= 0 || K <= 0 || O <= 0) return 0;
    const FwdInstance *I = select_instance(F, E);
    EntmaxParams ep;
    if (!I || make_entmax_params(alpha, F, solver, 50, &ep) != ARMNET_OK) return 0;
    if (use_tmem_kernel(F, E, K * O, ep.mode)) return 3;
    return select_mma_instance(I, F, E, K * O, ep.mode) ? 2 : 1;
}

extern "C" int armnet_fused_bwd_f32(const void *ids, int ids_i32, float *values, const float *table, int64_t V,
                                    int64_t ld, const float *bilinear_w, const float *query, const float *att_values,
                                    int w_is_linear_layout, float alpha, int64_t B, int F, int E, int D, int K, int O,
                                    const float *z, const float *dz, const float *tau, float *out_w, float *out_dg,
                                    float *acc_dvalues, float *acc_dm, void *workspace, int *err_flag, void *stream) {
    armnet::BwdArgs b = {z, dz, tau, out_w, out_dg, acc_dvalues, acc_dm};
    // values were clamped by the forward; the solver is not used (tau is an input)
    return armnet::launch_fused("fused_bwd", &b, ids, ids_i32, values, table, V, ld, bilinear_w, query, att_values,
                                w_is_linear_layout, alpha, ARMNET_SOLVER_AUTO, 50, B, F, E, D, K, O, 0, 0.f, 0.f, 0,
                                nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, workspace,
                                err_flag, stream);
}
