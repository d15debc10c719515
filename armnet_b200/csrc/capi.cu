// Version / error / device-info entry points and the host-side helpers shared by all entries.
#include <math.h>
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

namespace armnet {

static thread_local char g_err[512] = "";
static thread_local int g_launches = 0;

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int cuda_fail(cudaError_t e, const char *what) {
    set_error("CUDA error %d (%s) at %s", (int)e, cudaGetErrorString(e), what);
    return ARMNET_ERR_CUDA;
}

void note_launches(int n) { g_launches = n; }

static int env_int(const char *name, int unset) {
    const char *v = getenv(name);
    if (v == nullptr) return unset;
    if (v[0] == '\0') return 1;
    return atoi(v);
}

Tuning &tuning() {
    // function-local static: initialised once (thread-safe), the environment is never read again
    static Tuning t = [] {
        Tuning x;
        x.tmem = env_int("ARMNET_TMEM", -1);
        x.tmem_rows = env_int("ARMNET_TMEM_ROWS", -1);
        x.mma = env_int("ARMNET_MMA", -1);
        const char *sp = getenv("ARMNET_MMA_SPLIT");
        x.mma_split_rna = (sp && sp[0] == 'r') ? 1 : 0;
        x.mma_warps = env_int("ARMNET_MMA_WARPS", -1);
        x.force_nw = env_int("ARMNET_FORCE_NW", -1);
        x.force_look = env_int("ARMNET_FORCE_LOOK", -1);
        x.lockstep = getenv("ARMNET_LOCKSTEP") ? 1 : 0;
        x.no_tma_gather = getenv("ARMNET_NO_TMA_GATHER") ? 1 : 0;
        x.no_tma_store = getenv("ARMNET_NO_TMA_STORE") ? 1 : 0;
        x.gemm_1cta = getenv("ARMNET_GEMM_1CTA") ? 1 : 0;
        return x;
    }();
    return t;
}

int get_device_info(DeviceInfo *out) {
    // Read-only cache of immutable device properties, one slot per device ordinal.
    static int cached_sm[64];
    static int cached_smem[64];
    static bool have[64];
    int dev = 0;
    ARMNET_CUDA_TRY(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64 || !have[dev]) {
        int sm = 0, smem = 0;
        ARMNET_CUDA_TRY(cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, dev));
        ARMNET_CUDA_TRY(cudaDeviceGetAttribute(&smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
        if (dev >= 0 && dev < 64) {
            cached_sm[dev] = sm;
            cached_smem[dev] = smem;
            have[dev] = true;
        } else {
            out->sm_count = sm;
            out->smem_optin = smem;
            return ARMNET_OK;
        }
    }
    out->sm_count = cached_sm[dev];
    out->smem_optin = cached_smem[dev];
    return ARMNET_OK;
}

int make_entmax_params(float alpha, int F, int solver, int n_iter, EntmaxParams *ep) {
    if (!(alpha >= 1.f) || F <= 0) {
        set_error("entmax: alpha must be >= 1 and F > 0 (alpha=%g, F=%d)", (double)alpha, F);
        return ARMNET_ERR_SHAPE;
    }
    memset(ep, 0, sizeof(*ep));
    ep->n_iter = n_iter > 0 ? n_iter : 50;
    ep->inv_F = 1.f / (float)F;
    if (alpha == 1.f) {  // models/armnet.py:12 -- nn.Softmax(dim=-1)
        ep->mode = POW_SOFTMAX;
        ep->am1 = 1.f;
        ep->q = 1.f;
        return ARMNET_OK;
    }
    // entmax.py:31-36: alpha lives in an fp32 tensor, so alpha-1 and 1/(alpha-1) are fp32 roundings.
    const float am1 = alpha - 1.f;
    const float q = 1.f / am1;
    ep->am1 = am1;
    ep->q = q;
    ep->qm1 = q - 1.f;
    // entmax.py:18,47: (1/d) ** (alpha-1) with the scalar base cast to fp32.
    ep->cF = (float)pow((double)(float)(1.0 / (double)F), (double)am1);
    ep->uni_k = ep->qm1 * 0.5f / ep->cF;
    ep->uni_var = 0.04f * ep->cF * ep->cF / (float)F;
    if (solver == ARMNET_SOLVER_BISECT || q < 1.f) {
        ep->mode = POW_BISECT;  // alpha > 2: f is not convex, keep the reference's bracketing search
    } else if (q == 1.f) {
        ep->mode = POW_LINEAR;
    } else if (q == 2.f) {
        ep->mode = POW_SQUARE;
    } else {
        ep->mode = POW_GENERAL;
    }
    return ARMNET_OK;
}

}  // namespace armnet

extern "C" {

int armnet_version(void) { return ARMNET_B200_VERSION; }

const char *armnet_last_error_string(void) { return armnet::g_err; }

int armnet_last_launch_count(void) { return armnet::g_launches; }

int armnet_set_tuning(const char *key, int value) {
    if (key == nullptr) return ARMNET_ERR_NULL;
    armnet::Tuning &t = armnet::tuning();
    struct { const char *name; int *field; } tab[] = {
        {"tmem", &t.tmem}, {"tmem_rows", &t.tmem_rows}, {"mma", &t.mma}, {"mma_split_rna", &t.mma_split_rna}, {"mma_warps", &t.mma_warps},
        {"force_nw", &t.force_nw}, {"force_look", &t.force_look}, {"lockstep", &t.lockstep},
        {"no_tma_gather", &t.no_tma_gather}, {"no_tma_store", &t.no_tma_store}, {"gemm_1cta", &t.gemm_1cta}};
    for (auto &e : tab) {
        if (strcmp(key, e.name) == 0) {
            *e.field = value;
            return ARMNET_OK;
        }
    }
    armnet::set_error("armnet_set_tuning: unknown key '%s'", key);
    return ARMNET_ERR_UNSUPPORTED;
}

int armnet_device_info(int *sm_count, int *smem_optin_bytes) {
    armnet::DeviceInfo di;
    int rc = armnet::get_device_info(&di);
    if (rc != ARMNET_OK) return rc;
    if (sm_count) *sm_count = di.sm_count;
    if (smem_optin_bytes) *smem_optin_bytes = di.smem_optin;
    return ARMNET_OK;
}

}  // extern "C"
