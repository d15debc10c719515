// armnet_entmax_f32 / armnet_entmax_bwd_f32: row-wise alpha-entmax over the last axis as a standalone op
// (utils/entmax.py:29-68 forward, :71-80 backward; nn.Softmax(dim=-1) when alpha == 1, models/armnet.py:12).
//
// One thread solves one row held in registers (entmax.cuh); a 128-row tile is staged through shared memory so the
// global loads and stores are coalesced.  Algorithmic bytes: 8*F per row forward, 12*F per row backward.
#include "entmax.cuh"

namespace armnet {

constexpr int kRowsPerCta = 128;

__host__ __device__ inline int padded_stride(int F) { return (F & 1) ? F : F + 1; }  // odd stride: conflict-free columns

template <int FP>
__global__ void __launch_bounds__(kRowsPerCta) entmax_fwd_kernel(const float *__restrict__ x, long long rows, int F,
                                                                 const EntmaxParams ep, float *__restrict__ p) {
    extern __shared__ float tile[];
    const int Fs = padded_stride(F);
    const int tid = threadIdx.x;
    const long long n_tiles = (rows + kRowsPerCta - 1) / kRowsPerCta;
    for (long long t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        const long long row0 = t * kRowsPerCta;
        const int nr = (int)min((long long)kRowsPerCta, rows - row0);
        const float *src = x + row0 * F;
        for (int i = tid; i < nr * F; i += kRowsPerCta) tile[(i / F) * Fs + (i % F)] = src[i];
        __syncthreads();
        const bool valid = tid < nr;
        const float *mine = tile + (valid ? tid : 0) * Fs;
        float X[1][FP];
#pragma unroll
        for (int f = 0; f < FP; ++f) X[0][f] = (f < F) ? mine[f] * ep.am1 : neg_inf();  // entmax.py:42
        float tau[1], mx[1], mean[1];
        row_max_mean<FP, false>(X[0], F, ep, mx[0], mean[0]);
        entmax_solve_tau<1, FP, false>(X, F, ep, mx, mean, tau);
        float s = 0.f;
#pragma unroll
        for (int f = 0; f < FP; ++f) {
            X[0][f] = gate_unnorm_rt(X[0][f], tau[0], ep);
            s += X[0][f];
        }
        __syncthreads();  // every thread has read its (or row 0's) logits before any row is overwritten
        if (valid) {
            float *dst = tile + tid * Fs;
#pragma unroll
            for (int f = 0; f < FP; ++f)
                if (f < F) dst[f] = __fdiv_rn(X[0][f], s);  // entmax.py:63-64
        }
        __syncthreads();
        float *out = p + row0 * F;
        for (int i = tid; i < nr * F; i += kRowsPerCta) out[i] = tile[(i / F) * Fs + (i % F)];
        __syncthreads();
    }
}

// dX = dY*gppr - (sum dY*gppr / sum gppr) * gppr,  gppr = p^(2-alpha) on the support (entmax.py:74-80).
// alpha == 1 (softmax): the same expression with gppr = p is the softmax Jacobian.
__global__ void __launch_bounds__(kRowsPerCta) entmax_bwd_kernel(const float *__restrict__ p, const float *__restrict__ dp,
                                                                 long long rows, int F, float two_minus_alpha,
                                                                 float *__restrict__ dx) {
    extern __shared__ float tile[];
    const int Fs = padded_stride(F);
    float *tp = tile;
    float *td = tile + kRowsPerCta * Fs;
    const int tid = threadIdx.x;
    const long long n_tiles = (rows + kRowsPerCta - 1) / kRowsPerCta;
    for (long long t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        const long long row0 = t * kRowsPerCta;
        const int nr = (int)min((long long)kRowsPerCta, rows - row0);
        for (int i = tid; i < nr * F; i += kRowsPerCta) {
            const int o = (i / F) * Fs + (i % F);
            tp[o] = p[row0 * F + i];
            td[o] = dp[row0 * F + i];
        }
        __syncthreads();
        if (tid < nr) {
            float *mp = tp + tid * Fs;
            float *md = td + tid * Fs;
            float sg = 0.f, sdg = 0.f;
            for (int f = 0; f < F; ++f) {
                const float pv = mp[f];
                float g = 0.f;
                if (pv > 0.f) g = (two_minus_alpha == 0.f) ? 1.f : fast_ex2(two_minus_alpha * fast_lg2(pv));
                mp[f] = g;
                sg += g;
                sdg = fmaf(md[f], g, sdg);
            }
            const float qv = __fdiv_rn(sdg, sg);
            for (int f = 0; f < F; ++f) md[f] = (md[f] - qv) * mp[f];
        }
        __syncthreads();
        for (int i = tid; i < nr * F; i += kRowsPerCta) dx[row0 * F + i] = td[(i / F) * Fs + (i % F)];
        __syncthreads();
    }
}

template <int FP>
static int launch_fwd(const float *x, long long rows, int F, const EntmaxParams &ep, float *p, unsigned grid,
                      cudaStream_t st) {
    const size_t smem = (size_t)kRowsPerCta * padded_stride(F) * sizeof(float);
    entmax_fwd_kernel<FP><<<grid, kRowsPerCta, smem, st>>>(x, rows, F, ep, p);
    return ARMNET_OK;
}

}  // namespace armnet

extern "C" int armnet_entmax_f32(const float *x, int64_t rows, int F, float alpha, int solver, int n_iter, float *p,
                                 void *stream) {
    using namespace armnet;
    if (rows == 0 && F > 0) {
        note_launches(0);
        return ARMNET_OK;
    }
    if (!x || !p) {
        set_error("entmax: null pointer");
        return ARMNET_ERR_NULL;
    }
    if (rows < 0 || F <= 0) {
        set_error("entmax: bad shape rows=%lld F=%d", (long long)rows, F);
        return ARMNET_ERR_SHAPE;
    }
    if (F > 64) {
        set_error("entmax: F=%d exceeds the 64-wide register row of this build", F);
        return ARMNET_ERR_UNSUPPORTED;
    }
    EntmaxParams ep;
    int rc = make_entmax_params(alpha, F, solver, n_iter, &ep);
    if (rc != ARMNET_OK) return rc;
    note_launches(0);
    if (rows == 0) return ARMNET_OK;
    DeviceInfo di;
    rc = get_device_info(&di);
    if (rc != ARMNET_OK) return rc;
    const long long n_tiles = (rows + kRowsPerCta - 1) / kRowsPerCta;
    const unsigned grid = (unsigned)(n_tiles < (long long)di.sm_count * 8 ? n_tiles : (long long)di.sm_count * 8);
    cudaStream_t st = (cudaStream_t)stream;
    if (F <= 4) launch_fwd<4>(x, rows, F, ep, p, grid, st);
    else if (F <= 8) launch_fwd<8>(x, rows, F, ep, p, grid, st);
    else if (F <= 12) launch_fwd<12>(x, rows, F, ep, p, grid, st);
    else if (F <= 16) launch_fwd<16>(x, rows, F, ep, p, grid, st);
    else if (F <= 24) launch_fwd<24>(x, rows, F, ep, p, grid, st);
    else if (F <= 32) launch_fwd<32>(x, rows, F, ep, p, grid, st);
    else if (F <= 40) launch_fwd<40>(x, rows, F, ep, p, grid, st);
    else if (F <= 48) launch_fwd<48>(x, rows, F, ep, p, grid, st);
    else launch_fwd<64>(x, rows, F, ep, p, grid, st);
    ARMNET_CUDA_TRY(cudaGetLastError());
    note_launches(1);
    return ARMNET_OK;
}

extern "C" int armnet_entmax_bwd_f32(const float *p, const float *dp, int64_t rows, int F, float alpha, float *dx,
                                     void *stream) {
    using namespace armnet;
    if (rows == 0 && F > 0) {
        note_launches(0);
        return ARMNET_OK;
    }
    if (!p || !dp || !dx) {
        set_error("entmax_bwd: null pointer");
        return ARMNET_ERR_NULL;
    }
    if (rows < 0 || F <= 0 || !(alpha >= 1.f)) {
        set_error("entmax_bwd: bad shape rows=%lld F=%d alpha=%g", (long long)rows, F, (double)alpha);
        return ARMNET_ERR_SHAPE;
    }
    note_launches(0);
    if (rows == 0) return ARMNET_OK;
    const size_t smem = 2 * (size_t)kRowsPerCta * padded_stride(F) * sizeof(float);
    if (smem > 200 * 1024) {
        set_error("entmax_bwd: F=%d needs %zu bytes of shared memory", F, smem);
        return ARMNET_ERR_UNSUPPORTED;
    }
    DeviceInfo di;
    int rc = get_device_info(&di);
    if (rc != ARMNET_OK) return rc;
    if (smem > 48 * 1024)
        ARMNET_CUDA_TRY(cudaFuncSetAttribute(entmax_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const long long n_tiles = (rows + kRowsPerCta - 1) / kRowsPerCta;
    const unsigned grid = (unsigned)(n_tiles < (long long)di.sm_count * 4 ? n_tiles : (long long)di.sm_count * 4);
    // entmax.py:74: Y ** (2 - alpha) with alpha an fp32 tensor
    const float tma = 2.f - alpha;
    entmax_bwd_kernel<<<grid, kRowsPerCta, smem, (cudaStream_t)stream>>>(p, dp, rows, F, tma, dx);
    ARMNET_CUDA_TRY(cudaGetLastError());
    note_launches(1);
    return ARMNET_OK;
}
