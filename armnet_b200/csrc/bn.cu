// Train-mode arm_bn: BatchNorm1d(K*O) over the interaction output z [B, C = K*O, L = nemb], statistics per channel
// over (B, L) (models/armnet.py:67,89, armnet_1h.py:65,85; torch defaults momentum 0.1, eps 1e-5), forward and backward.
// cuDNN's kernel for this layout (bn_fw_tr_1C11_kernel_NCHW with L = 10..16) takes 0.7 ms at the Criteo shape;
// these are three streaming passes (partial sums -> finalize -> apply), each HBM bound.
//
// Numerics: z = exp(s) sits close to 1 with a tiny spread at initialisation, so E[z^2] - E[z]^2 cancels in fp32.
// The sums are taken around a per-channel pivot (the channel's first element): d = z - pivot, var = E[d^2] - E[d]^2.
// Partial sums are per CTA and combined in a fixed order in double precision: results are run-to-run deterministic.
#include "common.cuh"

namespace armnet {

constexpr int BN_THREADS = 256;

// mode 0 (forward):  s1 += z - pivot[ch]         s2 += (z - pivot[ch])^2
// mode 1 (backward): s1 += dy                    s2 += dy * (z - mean[ch])
// x, y: [B][CL] row-major (CL = C * L); one CTA owns rows [b0, b1); partial: [gridDim.x][CL][2].
// VEC = 4: a thread owns 4 consecutive positions (float4 loads; they may straddle two channels), rows unrolled by 4
// -> 64 bytes in flight per thread, enough memory-level parallelism to stream at HBM rate.
template <int VEC>
__global__ void __launch_bounds__(BN_THREADS) bn_partial_kernel(const float *__restrict__ x, const float *__restrict__ y,
                                                               const float *__restrict__ ref, long long B, int CL, int L,
                                                               int rows_per_cta, int mode, float *__restrict__ partial) {
    const long long b0 = (long long)blockIdx.x * rows_per_cta;
    const long long b1 = min(B, b0 + rows_per_cta);
    for (int j = threadIdx.x * VEC; j < CL; j += BN_THREADS * VEC) {
        float r[VEC], s1[VEC], s2[VEC];
#pragma unroll
        for (int k = 0; k < VEC; ++k) {
            const int ch = (j + k) / L;
            r[k] = mode == 0 ? __ldg(x + (long long)ch * L) : __ldg(ref + ch);  // pivot = z[0, ch, 0] | mean
            s1[k] = s2[k] = 0.f;
        }
#pragma unroll 4
        for (long long b = b0; b < b1; ++b) {
            float v[VEC], g[VEC];
            if (VEC == 4) {
                const float4 t = __ldg(reinterpret_cast<const float4 *>(x + b * CL + j));
                v[0] = t.x, v[1 % VEC] = t.y, v[2 % VEC] = t.z, v[3 % VEC] = t.w;
                if (mode != 0) {
                    const float4 u = __ldg(reinterpret_cast<const float4 *>(y + b * CL + j));
                    g[0] = u.x, g[1 % VEC] = u.y, g[2 % VEC] = u.z, g[3 % VEC] = u.w;
                }
            } else {
                v[0] = __ldg(x + b * CL + j);
                if (mode != 0) g[0] = __ldg(y + b * CL + j);
            }
#pragma unroll
            for (int k = 0; k < VEC; ++k) {
                if (mode == 0) {
                    const float d = v[k] - r[k];
                    s1[k] += d;
                    s2[k] = fmaf(d, d, s2[k]);
                } else {
                    s1[k] += g[k];
                    s2[k] = fmaf(g[k], v[k] - r[k], s2[k]);
                }
            }
        }
        float2 *dst = reinterpret_cast<float2 *>(partial) + (long long)blockIdx.x * CL + j;
#pragma unroll
        for (int k = 0; k < VEC; ++k) dst[k] = make_float2(s1[k], s2[k]);
    }
}

// One CTA per channel: fixed-order sum of the partials of its L positions over all CTAs of the partial pass, in double.
// (One WARP per channel left 512 warps walking 19 MB of partials in chains of L2 latencies: 33 us at config 4; a CTA per
// channel has every load of a thread in flight at once.)  Thread t adds elements t, t + 256, ... in order, the warp sums
// by shuffles and the eight warp sums are added in warp order: run-to-run deterministic.
// mode 0: mean, invstd, running-stat update.   mode 1: sum_dy, sum_dy_xmu (for dweight / dbias and the apply pass).
__global__ void __launch_bounds__(BN_THREADS) bn_finalize_kernel(const float *__restrict__ partial, const float *__restrict__ x,
                                                                int n_cta, int C, int L, long long n, float eps,
                                                                float momentum, int mode, float *__restrict__ out1,
                                                                float *__restrict__ out2, float *__restrict__ running_mean,
                                                                float *__restrict__ running_var,
                                                                const float *__restrict__ invstd_in,
                                                                float *__restrict__ raw_out) {
    __shared__ double red[2][BN_THREADS / 32];
    const int warp = blockIdx.x;   // the channel
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (warp >= C) return;
    const int CL = C * L;
    double s1 = 0.0, s2 = 0.0;
    const int total = n_cta * L;
    constexpr int kU = 8;
    for (int i0 = threadIdx.x; i0 < total; i0 += BN_THREADS * kU) {
        float2 t[kU];
#pragma unroll
        for (int u = 0; u < kU; ++u) {
            const int i = i0 + BN_THREADS * u;
            const int cta = i / L, l = i - cta * L;
            t[u] = i < total ? __ldg(reinterpret_cast<const float2 *>(partial) + (long long)cta * CL + warp * L + l)
                             : make_float2(0.f, 0.f);
        }
#pragma unroll
        for (int u = 0; u < kU; ++u) {
            s1 += (double)t[u].x;
            s2 += (double)t[u].y;
        }
    }
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) {
        s1 += __shfl_xor_sync(0xffffffffu, s1, m);
        s2 += __shfl_xor_sync(0xffffffffu, s2, m);
    }
    if (lane == 0) {
        red[0][wid] = s1;
        red[1][wid] = s2;
    }
    __syncthreads();
    if (threadIdx.x != 0) return;
    s1 = s2 = 0.0;
#pragma unroll
    for (int w = 0; w < BN_THREADS / 32; ++w) {
        s1 += red[0][w];
        s2 += red[1][w];
    }
    if (mode == 0) {
        const double pivot = (double)x[(long long)warp * L];
        const double md = s1 / (double)n;
        const double var = fmax(s2 / (double)n - md * md, 0.0);  // biased, as F.batch_norm normalises with
        const double mean = pivot + md;
        out1[warp] = (float)mean;
        out2[warp] = (float)(1.0 / sqrt(var + (double)eps));
        if (running_mean != nullptr) {  // nn.BatchNorm1d: running stats use the unbiased variance
            const double unb = n > 1 ? var * (double)n / (double)(n - 1) : var;
            running_mean[warp] = (float)((1.0 - momentum) * (double)running_mean[warp] + momentum * mean);
            running_var[warp] = (float)((1.0 - momentum) * (double)running_var[warp] + momentum * unb);
        }
    } else {
        out1[warp] = (float)s1;                       // dbias   = sum dy
        out2[warp] = (float)s2 * invstd_in[warp];     // dweight = sum dy * xhat
        raw_out[warp] = (float)s2;                    // sum dy (x - mean), for the apply pass
    }
}

// forward:  out = (x - mean) * (invstd * weight) + bias
// backward: out = weight * invstd * (dy - sum_dy / n - (x - mean) * invstd^2 * sum_dy_xmu / n)
// Same ownership as bn_partial_kernel: a CTA owns rows [b0, b1), a thread walks its positions with the channel
// constants in registers: out = p0 * x + p1 * dy + p2.
template <int VEC>
__global__ void __launch_bounds__(BN_THREADS) bn_apply_kernel(const float *__restrict__ x, const float *__restrict__ dy,
                                                             const float *__restrict__ mean, const float *__restrict__ invstd,
                                                             const float *__restrict__ weight, const float *__restrict__ bias,
                                                             const float *__restrict__ sum_dy, const float *__restrict__ sum_dy_xmu,
                                                             long long B, int CL, int L, int rows_per_cta, float inv_n,
                                                             int mode, float *__restrict__ out) {
    const long long b0 = (long long)blockIdx.x * rows_per_cta;
    const long long b1 = min(B, b0 + rows_per_cta);
    for (int j = threadIdx.x * VEC; j < CL; j += BN_THREADS * VEC) {
        // fwd: (x - m) * a + c        bwd: (dy - sd - (x - m) * kk) * a      (x - m first: x ~ m when z = exp(s) ~ 1)
        float m[VEC], a[VEC], c[VEC], kk[VEC];
#pragma unroll
        for (int k = 0; k < VEC; ++k) {
            const int ch = (j + k) / L;
            const float is = __ldg(invstd + ch), w = weight ? __ldg(weight + ch) : 1.f;
            m[k] = __ldg(mean + ch);
            a[k] = is * w;
            if (mode == 0) {
                c[k] = bias ? __ldg(bias + ch) : 0.f;
                kk[k] = 0.f;
            } else {
                c[k] = __ldg(sum_dy + ch) * inv_n;
                kk[k] = __ldg(sum_dy_xmu + ch) * is * is * inv_n;
            }
        }
#pragma unroll 4
        for (long long b = b0; b < b1; ++b) {
            float v[VEC], g[VEC], o[VEC];
            if (VEC == 4) {
                const float4 t = __ldg(reinterpret_cast<const float4 *>(x + b * CL + j));
                v[0] = t.x, v[1 % VEC] = t.y, v[2 % VEC] = t.z, v[3 % VEC] = t.w;
                if (mode != 0) {
                    const float4 u = __ldg(reinterpret_cast<const float4 *>(dy + b * CL + j));
                    g[0] = u.x, g[1 % VEC] = u.y, g[2 % VEC] = u.z, g[3 % VEC] = u.w;
                }
            } else {
                v[0] = __ldg(x + b * CL + j);
                if (mode != 0) g[0] = __ldg(dy + b * CL + j);
            }
#pragma unroll
            for (int k = 0; k < VEC; ++k) {
                const float xm = v[k] - m[k];
                o[k] = mode == 0 ? fmaf(xm, a[k], c[k]) : (g[k] - c[k] - xm * kk[k]) * a[k];
            }
            if (VEC == 4)
                *reinterpret_cast<float4 *>(out + b * CL + j) = make_float4(o[0], o[1 % VEC], o[2 % VEC], o[3 % VEC]);
            else
                out[b * CL + j] = o[0];
        }
    }
}

static int bn_grid_rows(long long B, int *rows_per_cta) {
    DeviceInfo di;
    if (get_device_info(&di) != ARMNET_OK) di.sm_count = 148;
    long long target = (long long)di.sm_count * 2;
    long long rpc = (B + target - 1) / target;
    if (rpc < 1) rpc = 1;
    *rows_per_cta = (int)rpc;
    return (int)((B + rpc - 1) / rpc);
}

}  // namespace armnet

using namespace armnet;

extern "C" {

size_t armnet_bn_workspace_floats(int64_t B, int C, int L) {
    if (B <= 0 || C <= 0 || L <= 0) return 0;
    int rpc;
    const int n_cta = bn_grid_rows(B, &rpc);
    return (size_t)n_cta * C * L * 2 + (size_t)C;
}

int armnet_bn_train_fwd_f32(const float *x, int64_t B, int C, int L, const float *weight, const float *bias,
                            float *running_mean, float *running_var, float momentum, float eps, float *out,
                            float *save_mean, float *save_invstd, float *workspace, void *stream) {
    if (!x || !out || !save_mean || !save_invstd || !workspace) {
        set_error("armnet_bn_train_fwd_f32: null pointer");
        return ARMNET_ERR_NULL;
    }
    if (B <= 0 || C <= 0 || L <= 0 || (B * L < 2)) {
        // nn.BatchNorm1d raises "Expected more than 1 value per channel when training"
        set_error("armnet_bn_train_fwd_f32: expected more than 1 value per channel when training (B=%lld, L=%d)",
                  (long long)B, L);
        return ARMNET_ERR_SHAPE;
    }
    cudaStream_t st = (cudaStream_t)stream;
    int rpc;
    const int n_cta = bn_grid_rows(B, &rpc);
    const int CL = C * L;
    const bool v4 = (CL % 4 == 0) && (((uintptr_t)x | (uintptr_t)out) % 16 == 0);
    if (v4) bn_partial_kernel<4><<<n_cta, BN_THREADS, 0, st>>>(x, nullptr, nullptr, B, CL, L, rpc, 0, workspace);
    else bn_partial_kernel<1><<<n_cta, BN_THREADS, 0, st>>>(x, nullptr, nullptr, B, CL, L, rpc, 0, workspace);
    bn_finalize_kernel<<<C, BN_THREADS, 0, st>>>(
        workspace, x, n_cta, C, L, (long long)B * L, eps, momentum, 0, save_mean, save_invstd, running_mean, running_var, nullptr, nullptr);
    if (v4)
        bn_apply_kernel<4><<<n_cta, BN_THREADS, 0, st>>>(x, nullptr, save_mean, save_invstd, weight, bias, nullptr, nullptr,
                                                         B, CL, L, rpc, 0.f, 0, out);
    else
        bn_apply_kernel<1><<<n_cta, BN_THREADS, 0, st>>>(x, nullptr, save_mean, save_invstd, weight, bias, nullptr, nullptr,
                                                         B, CL, L, rpc, 0.f, 0, out);
    ARMNET_CUDA_TRY(cudaGetLastError());
    note_launches(3);
    return ARMNET_OK;
}

int armnet_bn_train_bwd_f32(const float *x, const float *dy, int64_t B, int C, int L, const float *weight,
                            const float *save_mean, const float *save_invstd, float *dx, float *dweight, float *dbias,
                            float *workspace, void *stream) {
    if (!x || !dy || !save_mean || !save_invstd || !dx || !dweight || !dbias || !workspace) {
        set_error("armnet_bn_train_bwd_f32: null pointer");
        return ARMNET_ERR_NULL;
    }
    if (B <= 0 || C <= 0 || L <= 0) {
        set_error("armnet_bn_train_bwd_f32: bad sizes");
        return ARMNET_ERR_SHAPE;
    }
    cudaStream_t st = (cudaStream_t)stream;
    int rpc;
    const int n_cta = bn_grid_rows(B, &rpc);
    const int CL = C * L;
    float *raw = workspace + (size_t)n_cta * CL * 2;  // [C] sum dy (x - mean)
    const bool v4 = (CL % 4 == 0) && (((uintptr_t)x | (uintptr_t)dy | (uintptr_t)dx) % 16 == 0);
    if (v4) bn_partial_kernel<4><<<n_cta, BN_THREADS, 0, st>>>(x, dy, save_mean, B, CL, L, rpc, 1, workspace);
    else bn_partial_kernel<1><<<n_cta, BN_THREADS, 0, st>>>(x, dy, save_mean, B, CL, L, rpc, 1, workspace);
    bn_finalize_kernel<<<C, BN_THREADS, 0, st>>>(
        workspace, x, n_cta, C, L, (long long)B * L, 0.f, 0.f, 1, dbias, dweight, nullptr, nullptr, save_invstd, raw);
    const float inv_n = 1.f / (float)((long long)B * L);
    if (v4)
        bn_apply_kernel<4><<<n_cta, BN_THREADS, 0, st>>>(x, dy, save_mean, save_invstd, weight, nullptr, dbias, raw, B, CL, L,
                                                         rpc, inv_n, 1, dx);
    else
        bn_apply_kernel<1><<<n_cta, BN_THREADS, 0, st>>>(x, dy, save_mean, save_invstd, weight, nullptr, dbias, raw, B, CL, L,
                                                         rpc, inv_n, 1, dx);
    ARMNET_CUDA_TRY(cudaGetLastError());
    note_launches(3);
    return ARMNET_OK;
}

}  // extern "C"
