// armnet_fwd_tmem_kernel (fused_fwd_tmem.cuh): compiled instances, operand preparation and launch.
#include <math.h>
#include <string.h>

#include "fused_fwd_tmem.cuh"

namespace armnet {

static const TmemInstance kTmemInstances[] = {
    ARMNET_TMEM_INSTANCE(20, 1),       // 39 fields, nemb <= 10: Criteo shape (C2a / C2b)
    ARMNET_TMEM_INSTANCE(20, 0),       // 40 fields
    ARMNET_TMEM_INSTANCE_WIDE(20, 1),  // 39 fields, nemb 11..16 (config 4 shape)
    ARMNET_TMEM_INSTANCE_WIDE(20, 0),
};

static const TmemInstance *select_tmem_instance(int F, int E) {
    for (const TmemInstance &I : kTmemInstances)
        if (F == 2 * I.NP - (I.odd ? 1 : 0) && E <= I.EL && (I.EL == kTmEL || E > kTmEL)) return &I;
    return nullptr;
}

bool tmem_shape_supported(int F, int E, int R) {
    // tensor memory: A operand (R/128 blocks x 32 columns) + 4 D slots of 2 samples x 2 NP fields
    const TmemInstance *I = E >= 1 ? select_tmem_instance(F, E) : nullptr;
    if (I == nullptr || R % 128 != 0 || R < 128 || (R / 128) * tm_a_cols(I->EL) + 4 * 4 * I->NP > 512) return false;
    // shared memory (sm_100: 227 KB opt-in per CTA) with the shallowest gather ring
    TmemParams P;
    memset(&P, 0, sizeof(P));
    P.F = F;
    P.E = E;
    P.R = R;
    P.row_bytes = (E * 4 + 15) / 16 * 16;
    P.tma_store = 1;
    P.look = 1;
    P.n_raw = 2;
    return TmemSmem(I->NP, 2, P).total <= 232448;
}

// Workspace of the TMEM kernel: Apk floats [R/128][128][32 | 48], then Vpk float2 [R][vstr] (16-byte padded).
static size_t tmem_apk_bytes(int R, int EL) { return (size_t)R * tm_a_cols(EL) * 4; }
static size_t tmem_vpk_bytes(int F, int E, int R) {
    const TmemInstance *I = select_tmem_instance(F, E);
    return (((size_t)R * (I->NP + 2) * 8) + 15) / 16 * 16;
}
size_t tmem_workspace_bytes(int F, int E, int R) {
    if (!tmem_shape_supported(F, E, R)) return 0;
    return tmem_apk_bytes(R, select_tmem_instance(F, E)->EL) + tmem_vpk_bytes(F, E, R);
}

// Apk[(kb*128 + i)*KA + k]: TMEM lane i of A block kb is neuron r = 128 kb + i;  EL = 10, KA = 32 (EL = 16, KA = 48):
//   k in [0,EL): M'[x=k][r];  [EL,2EL): M'[x=k-EL][r];  [2EL,3EL): M'[x] - trunc_tf32(M'[x]);  beyond: 0
//   M'[x][r] = (alpha-1) * d_k^-0.5 * sum_y W[k,x,y] Q[k,o,y]   (armnet.py:33-34, entmax.py:42; one-head: W[x,y] = W_lin[y,x])
// Vpk[r*vstr + j] = (V[r][2j], V[r][2j+1])                                                          (armnet.py:36)
__global__ void attn_prepare_tmem_kernel(const float *__restrict__ W, const float *__restrict__ Q,
                                         const float *__restrict__ Vals, int lin_layout, int F, int E, int D, int O, int R,
                                         int NP, int vstr, int EL, float scale, float am1, float *__restrict__ Apk,
                                         float *__restrict__ Vpk) {
    const int KA = tm_a_cols(EL);
    const int nA = R * KA;
    const int nV = R * vstr * 2;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < nA + nV; idx += gridDim.x * blockDim.x) {
        if (idx < nA) {
            const int row = idx / KA, k = idx - row * KA;
            const int r = row;                            // = 128 kb + i
            const int x = k < EL ? k : (k < 2 * EL ? k - EL : k - 2 * EL);
            float a = 0.f;
            if (k < 3 * EL && x < E) {
                const int kh = r / O;
                const float *q = Q + (long long)r * D;
                if (lin_layout) {
                    for (int y = 0; y < D; ++y) a = fmaf(W[y * E + x], q[y], a);
                } else {
                    const float *w = W + ((long long)kh * E + x) * D;
                    for (int y = 0; y < D; ++y) a = fmaf(w[y], q[y], a);
                }
                a = (a * scale) * am1;
                if (k >= 2 * EL) a = a - __uint_as_float(__float_as_uint(a) & 0xffffe000u);
            }
            Apk[idx] = a;
        } else {
            const int i2 = idx - nA;
            const int comp = i2 & 1, q2 = i2 >> 1;       // float2 index
            const int r = q2 / vstr, j = q2 - r * vstr;
            float v = 0.f;
            if (j < NP) {
                const int f = 2 * j + comp;
                if (f < F) v = Vals[(long long)r * F + f];
            }
            Vpk[i2] = v;
        }
    }
}

int tmem_prepare(const float *bilinear_w, const float *query, const float *att_values, int w_is_linear_layout, float am1,
                 int F, int E, int D, int K, int O, void *workspace, cudaStream_t st) {
    const int R = K * O;
    const TmemInstance *I = select_tmem_instance(F, E);
    DeviceInfo di;
    int rc = get_device_info(&di);
    if (rc != ARMNET_OK) return rc;
    float *Apk = (float *)workspace;
    float *Vpk = (float *)((char *)workspace + tmem_apk_bytes(R, I->EL));
    const int vstr = I->NP + 2;
    const int total = R * tm_a_cols(I->EL) + R * vstr * 2;
    int blocks = (total + 255) / 256;
    if (blocks > di.sm_count * 4) blocks = di.sm_count * 4;
    const float scale = (float)pow((double)D, -0.5);  // armnet.py:15
    attn_prepare_tmem_kernel<<<blocks, 256, 0, st>>>(bilinear_w, query, att_values, w_is_linear_layout, F, E, D, O, R, I->NP,
                                                      vstr, I->EL, scale, am1, Apk, Vpk);
    ARMNET_CUDA_TRY(cudaGetLastError());
    return ARMNET_OK;
}

int tmem_launch(const char *who, const void *ids, int ids_i32, float *values, const float *table, int64_t V, int64_t ld,
                const EntmaxParams &ep, int64_t B, int F, int E, int K, int O, int clamp, float clamp_lo, float clamp_hi,
                int clamp_inplace, const float *post_mean, const float *post_scale, const float *post_shift, float *out_z,
                const void *workspace, int *err_flag, cudaStream_t st) {
    const int R = K * O;
    const TmemInstance *I = select_tmem_instance(F, E);
    DeviceInfo di;
    int rc = get_device_info(&di);
    if (rc != ARMNET_OK) return rc;
    TmemParams P;
    memset(&P, 0, sizeof(P));
    P.ids = ids;
    P.values = values;
    P.table = table;
    P.Apk = (const float *)workspace;
    P.Vpk = (const float2 *)((const char *)workspace + tmem_apk_bytes(R, I->EL));
    P.post_mean = post_mean;
    P.post_scale = post_scale;
    P.post_shift = post_shift;
    P.out_z = out_z;
    P.err_flag = err_flag;
    P.V = V;
    P.ld = ld;
    P.B = B;
    P.F = F;
    P.E = E;
    P.R = R;
    P.ids_i32 = ids_i32;
    P.clamp = clamp;
    P.clamp_inplace = clamp_inplace;
    P.clamp_lo = clamp_lo;
    P.clamp_hi = clamp_hi;
    P.ep = ep;
    P.n_tiles = (int)((B + 1) / 2);
    P.row_bytes = (E * 4 + 15) / 16 * 16;
    P.tma_store = ((uintptr_t)out_z % 16 == 0) ? 1 : 0;
    // rows per thread: 1 (logits in registers) by default.  tuning "tmem_rows" == 2 (needs K*O % 256 == 0): two rows per
    // thread, dense rows streamed from tensor memory -- +4 % in the init-weight regime (34.9 M vs 33.5 M samples/s at
    // C2a), -10 % with sparse gates (13.9 M vs 15.5 M), so it is not the default (profiles/r2_v3_summary.md)
    const int NR = (R % 256 == 0 && tuning().tmem_rows == 2 && I->kernel2 != nullptr) ? 2 : 1;
    // gather look-ahead: as deep as the raw ring that fits (<= 16 tiles = 32 samples: the mbarrier block holds 32)
    int look = 16;
    for (; look >= 1; --look) {
        P.look = look;
        P.n_raw = 2 * look;
        const TmemSmem L(I->NP, NR, P);
        if (L.total <= di.smem_optin) break;
    }
    if (look < 1) {
        set_error("%s: F=%d E=%d K*O=%d does not fit the shared memory of the tensor-memory kernel", who, F, E, R);
        return ARMNET_ERR_UNSUPPORTED;
    }
    const TmemSmem L(I->NP, NR, P);
    P.inv_ipt = (unsigned)((0x100000000ull + (unsigned)(R / 128 / NR) - 1) / (unsigned)(R / 128 / NR));
    const void *kernel = NR == 2 ? I->kernel2 : I->kernel;
    const unsigned grid = (unsigned)(P.n_tiles < di.sm_count ? P.n_tiles : di.sm_count);
    ARMNET_CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, di.smem_optin));
    void *args[] = {(void *)&P};
    ARMNET_CUDA_TRY(cudaLaunchKernel(kernel, dim3(grid), dim3(kTmThreads), args, (size_t)L.total, st));
    return ARMNET_OK;
}

}  // namespace armnet

#ifdef ARMNET_TMEM_TRACE
// tuning builds only: copy the timeline trace of the last launch (148 CTAs x 64 clock64() slots) to the host
extern "C" int armnet_debug_tmem_trace(long long *out) {
    return (int)cudaMemcpyFromSymbol(out, armnet::g_tm_trace, sizeof(long long) * 148 * 64);
}
#endif
