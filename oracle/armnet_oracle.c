/*
 * armnet_oracle.c -- plain-C restatement of the ARM-Net forward hot path.  TEST INFRASTRUCTURE, NOT PRODUCT CODE:
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load liboracle (oracle/_build/).
 *
 * It restates, loop by loop in fp32, what the reference (nusdbsystem/ARM-Net @ 7aeb3a4) computes with ATen calls:
 *   value clamp                      models/armnet.py:82
 *   e = table[id] * value            models/layers.py:20-21
 *   g = (e . W . Q^T) * d_k^-0.5     models/armnet.py:33-34   (one-head: keys = e W_lin^T, g = keys . Q^T, armnet_1h.py:30-32)
 *   p = entmax_alpha(g) by 50 bisection steps, or softmax when alpha == 1      utils/entmax.py:29-68, armnet.py:12
 *   w = p * values                   models/armnet.py:36
 *   z = exp(sum_f w_f e_f)           models/armnet.py:86-87
 * Same operation order as the reference (e.W first, then .Q; scale applied to g; p from the LAST midpoint; final
 * renormalisation).  libm powf/expf differ from ATen's vectorised kernels by <= 1-2 ulp, so this port is pinned to the
 * reference-generated fixtures (tests/golden) within 2e-6 on gates and 1e-6 norm-relative on z, not bit-for-bit -- the
 * bit-exact pin is the torch restatement oracle/armnet_oracle.py (tests/test_oracle_golden.py checks both).
 *
 * Parallel over samples with OpenMP (the CPU baseline uses every host core, like torch does for the reference).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* entmax.py:29-68 for one row of d logits (in place: x -> p). scratch: d floats. */
static void entmax_bisect_row(float *x, int d, float alpha, int n_iter, float *scratch) {
    const float am1 = alpha - 1.0f;           /* alpha is an fp32 tensor in the reference */
    const float inv = 1.0f / am1;
    float mx = -INFINITY;
    for (int i = 0; i < d; ++i) {
        x[i] = x[i] * am1;                    /* entmax.py:42 */
        if (x[i] > mx) mx = x[i];
    }
    float tau_lo = mx - 1.0f;                 /* entmax.py:46: _gp(1, alpha) == 1 */
    const float tau_hi = mx - powf((float)(1.0 / d), am1); /* entmax.py:47 */
    float f_lo = 0.f;
    for (int i = 0; i < d; ++i) {
        float t = x[i] - tau_lo;
        f_lo += powf(t > 0.f ? t : 0.f, inv);
    }
    f_lo -= 1.0f;
    float dm = tau_hi - tau_lo;
    float sum = 1.f;
    for (int it = 0; it < n_iter; ++it) {     /* entmax.py:53-61 */
        dm *= 0.5f;
        const float tau_m = tau_lo + dm;
        sum = 0.f;
        for (int i = 0; i < d; ++i) {
            float t = x[i] - tau_m;
            scratch[i] = powf(t > 0.f ? t : 0.f, inv);
            sum += scratch[i];
        }
        if ((sum - 1.0f) * f_lo >= 0.f) tau_lo = tau_m;
    }
    for (int i = 0; i < d; ++i) x[i] = scratch[i] / sum; /* entmax.py:63-64 */
}

static void softmax_row(float *x, int d) {   /* armnet.py:12 nn.Softmax(dim=-1) */
    float mx = -INFINITY, s = 0.f;
    for (int i = 0; i < d; ++i) if (x[i] > mx) mx = x[i];
    for (int i = 0; i < d; ++i) { x[i] = expf(x[i] - mx); s += x[i]; }
    for (int i = 0; i < d; ++i) x[i] /= s;
}

/*
 * Whole hot path. Layouts as in include/armnet_b200.h: ids [B,F] int64, values [B,F] (clamped IN PLACE),
 * table [V,E]; multi-head (lin_layout == 0): W [K,E,D], Q [K,O,D], vals [K,O,F]; one-head: W = W_lin [D,E], K == 1.
 * Outputs (any may be NULL): e [B,F,E], g [B,K*O,F], p [B,K*O,F], s [B,K*O,E], z [B,K*O,E].
 * Returns 0, or -1 if an id is outside [0, V) (reference: IndexError, layers.py:20).
 */
int armnet_oracle_hot_path(const int64_t *ids, float *values, const float *table, int64_t V, const float *W,
                           const float *Q, const float *vals, int lin_layout, float alpha, int n_iter, int64_t B,
                           int F, int E, int D, int K, int O, float *out_e, float *out_g, float *out_p, float *out_s,
                           float *out_z) {
    const int R = K * O;
    const float scale = (float)pow((double)D, -0.5); /* armnet.py:15 */
    int bad = 0;
#pragma omp parallel for schedule(static)
    for (int64_t b = 0; b < B; ++b) {
        float *e = (float *)malloc(sizeof(float) * ((size_t)F * E + (size_t)F * K * D + 2 * (size_t)F));
        float *keys = e + (size_t)F * E;      /* [F][K][D] = e . W */
        float *row = keys + (size_t)F * K * D;
        float *scratch = row + F;
        for (int f = 0; f < F; ++f) {
            float v = values[b * F + f];
            v = v < 0.001f ? 0.001f : (v > 1.0f ? 1.0f : v);   /* armnet.py:82 clamp_(0.001, 1.) */
            values[b * F + f] = v;
            const int64_t id = ids[b * F + f];
            if (id < 0 || id >= V) {
#pragma omp atomic write
                bad = 1;
                for (int x = 0; x < E; ++x) e[f * E + x] = 0.f;
                continue;
            }
            for (int x = 0; x < E; ++x) e[f * E + x] = table[id * E + x] * v;   /* layers.py:20-21 */
        }
        if (out_e) memcpy(out_e + (size_t)b * F * E, e, sizeof(float) * F * E);
        for (int f = 0; f < F; ++f)           /* first contraction of 'bfx,kxy,koy->bkof': e.W (armnet_1h: nn.Linear) */
            for (int k = 0; k < K; ++k)
                for (int y = 0; y < D; ++y) {
                    float a = 0.f;
                    for (int x = 0; x < E; ++x)
                        a += e[f * E + x] * (lin_layout ? W[y * E + x] : W[((size_t)k * E + x) * D + y]);
                    keys[((size_t)f * K + k) * D + y] = a;
                }
        for (int r = 0; r < R; ++r) {
            const int k = r / O;
            for (int f = 0; f < F; ++f) {     /* second contraction: .Q, then * scale (armnet.py:33-34) */
                float a = 0.f;
                for (int y = 0; y < D; ++y) a += keys[((size_t)f * K + k) * D + y] * Q[(size_t)r * D + y];
                row[f] = a * scale;
            }
            if (out_g) memcpy(out_g + ((size_t)b * R + r) * F, row, sizeof(float) * F);
            if (alpha == 1.0f) softmax_row(row, F);
            else entmax_bisect_row(row, F, alpha, n_iter, scratch);
            if (out_p) memcpy(out_p + ((size_t)b * R + r) * F, row, sizeof(float) * F);
            for (int x = 0; x < E; ++x) {     /* armnet.py:36,86-87 */
                float a = 0.f;
                for (int f = 0; f < F; ++f) a += (row[f] * vals[(size_t)r * F + f]) * e[f * E + x];
                if (out_s) out_s[((size_t)b * R + r) * E + x] = a;
                if (out_z) out_z[((size_t)b * R + r) * E + x] = expf(a);
            }
        }
        free(e);
    }
    return bad ? -1 : 0;
}

int armnet_oracle_num_threads(void) {
#ifdef _OPENMP
    extern int omp_get_max_threads(void);
    return omp_get_max_threads();
#else
    return 1;
#endif
}
