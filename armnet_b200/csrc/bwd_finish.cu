// armnet_fused_bwd_finish_f32: from the per-row partials of armnet_fused_bwd_f32 to the parameter gradients, on the
// package's own kernels (round 1 finished with two cuBLAS products and ATen index_add_ / einsum on the host side).
//   de[b,f,:]  = sum_r w[b,f,r] ds[b,r,:] + sum_r dg[b,f,r] Mt[r,:]      ds = dz * z,  Mt[r,x] = d_k^-0.5 sum_y W[k,x,y] Q[k,o,y]
//   dT[id[b,f],:] += de[b,f,:] * value[b,f]                               dense [V,E] gradient (layers.py:12,20-21)
//   dW[k,x,y] = d_k^-0.5 sum_o dm[k,o,x] Q[k,o,y],  dQ[k,o,y] = d_k^-0.5 sum_x dm[k,o,x] W[k,x,y]      (armnet.py:33)
// de is a [F x R] . [R x E] product per sample with its own operands (5.2 GFLOP and 0.92 GB per batch at config 4: HBM
// bound if the math keeps up), so it runs on the tensor cores as warp-level mma.sync.m16n8k8 TF32 with the 3xTF32 split
// (fp32-grade, like every other product of the package): one CTA per sample, the reduction axis r dealt to 8 warps in
// steps of 8 rows, partial tiles summed in shared memory, then one atomic add per element into dT (duplicate ids of a
// batch accumulate like the reference's dense embedding gradient).
#include <math.h>

#include "common.cuh"

namespace armnet {

__device__ __forceinline__ void mma_tf32_16x8x8(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3,
                                                uint32_t b0, uint32_t b1) {
    asm("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
// x = hi + lo: the tensor core reads the top 19 bits of an operand register, so x itself serves as hi
__device__ __forceinline__ uint32_t tf32_lo(float x) {
    return __float_as_uint(x - __uint_as_float(__float_as_uint(x) & 0xffffe000u));
}

constexpr int kFinWarps = 8;

// Mt[r][x] = scale * sum_y W[k,x,y] Q[k,o,y]  (one-head: W[x,y] = W_lin[y,x])
__global__ void bwd_mt_kernel(const float *__restrict__ W, const float *__restrict__ Q, int lin_layout, int E, int D,
                              int O, int R, float scale, float *__restrict__ Mt) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= R * E) return;
    const int r = i / E, x = i - r * E, k = r / O;
    const float *q = Q + (long long)r * D;
    float a = 0.f;
    if (lin_layout) {
        for (int y = 0; y < D; ++y) a = fmaf(W[y * E + x], q[y], a);
    } else {
        const float *w = W + ((long long)k * E + x) * D;
        for (int y = 0; y < D; ++y) a = fmaf(w[y], q[y], a);
    }
    Mt[i] = a * scale;
}

// MT: 16-field tiles (F <= 16 MT).  One CTA per sample.  ds = dz * z and Mt fragments come straight from global memory (each
// element once per CTA; staging them in shared memory first measured slower: 576 vs 315 us at config 4); w / dg stream
// as 16-byte loads along r: the reduction index of
// an MMA step may be permuted freely, so a lane's k-slots (t, t+4) of two consecutive steps are the four consecutive rows
// r0 + 4t .. r0 + 4t + 3 of a 16-row block.
template <int MT>
__global__ void __launch_bounds__(kFinWarps * 32) bwd_embed_grad_kernel(
    const void *__restrict__ ids, int ids_i32, const float *__restrict__ values, const float *__restrict__ wA,
    const float *__restrict__ gA, const float *__restrict__ z, const float *__restrict__ dz,
    const float *__restrict__ Mt, long long V, int F, int E, int R, float *__restrict__ dT) {
    extern __shared__ __align__(16) float fin_smem[];
    const int E_pad = (E + 7) & ~7;
    const int R16 = (R + 15) & ~15;
    float *de_s = fin_smem;                 // [16 MT][E_pad]
    const long long b = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    const float *zb = z + b * (long long)R * E;
    const float *dzb = dz + b * (long long)R * E;
    for (int i = tid; i < 16 * MT * E_pad; i += blockDim.x) de_s[i] = 0.f;
    __syncthreads();
    const float *wb = wA + b * (long long)F * R;
    const float *gb = gA + b * (long long)F * R;
    const int n_blocks = R16 / 16;
    const bool vec = (R & 3) == 0 && ((reinterpret_cast<uintptr_t>(wA) | reinterpret_cast<uintptr_t>(gA)) & 15) == 0;
    // embedding lanes in groups of 16 (two n-tiles): nemb <= 16 is one pass over w / dg
    for (int x0 = 0; x0 < E; x0 += 16) {
        float acc[MT][2][4];
#pragma unroll
        for (int m = 0; m < MT; ++m)
#pragma unroll
            for (int n = 0; n < 2; ++n)
#pragma unroll
                for (int i = 0; i < 4; ++i) acc[m][n][i] = 0.f;
        for (int kb = warp; kb < n_blocks; kb += kFinWarps) {
            const int r4 = kb * 16 + 4 * t;   // this lane's four reduction rows
            // A: w / dg at (field 16m + g | + 8, rows r4 .. r4 + 3)
            float4 aw[MT][2], ag[MT][2];
#pragma unroll
            for (int m = 0; m < MT; ++m)
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int f = 16 * m + g + 8 * h;
                    float4 a = make_float4(0.f, 0.f, 0.f, 0.f), c = a;
                    if (f < F) {
                        const float *pw = wb + (long long)f * R + r4, *pg = gb + (long long)f * R + r4;
                        if (vec && r4 + 3 < R) {
                            a = __ldg(reinterpret_cast<const float4 *>(pw));
                            c = __ldg(reinterpret_cast<const float4 *>(pg));
                        } else {
                            if (r4 + 0 < R) a.x = __ldg(pw + 0), c.x = __ldg(pg + 0);
                            if (r4 + 1 < R) a.y = __ldg(pw + 1), c.y = __ldg(pg + 1);
                            if (r4 + 2 < R) a.z = __ldg(pw + 2), c.z = __ldg(pg + 2);
                            if (r4 + 3 < R) a.w = __ldg(pw + 3), c.w = __ldg(pg + 3);
                        }
                    }
                    aw[m][h] = a;
                    ag[m][h] = c;
                }
#pragma unroll
            for (int st = 0; st < 2; ++st) {   // two MMA steps: rows (r4, r4 + 1) and (r4 + 2, r4 + 3) in k-slots (t, t + 4)
                float bs[2][2], bm[2][2];
#pragma unroll
                for (int n = 0; n < 2; ++n) {   // ds = dz * z and Mt at rows (ra, ra + 1), column x: each element once per CTA
                    const int x = x0 + 8 * n + g;
                    const int ra = r4 + 2 * st;
                    const bool v0 = x < E && ra < R, v1 = x < E && ra + 1 < R;
                    const long long o = (long long)ra * E + x;
                    bs[n][0] = v0 ? __ldg(dzb + o) * __ldg(zb + o) : 0.f;
                    bs[n][1] = v1 ? __ldg(dzb + o + E) * __ldg(zb + o + E) : 0.f;
                    bm[n][0] = v0 ? __ldg(Mt + o) : 0.f;
                    bm[n][1] = v1 ? __ldg(Mt + o + E) : 0.f;
                }
#pragma unroll
                for (int m = 0; m < MT; ++m) {
                    // a0 (row g, k t)  a1 (row g+8, k t)  a2 (row g, k t+4)  a3 (row g+8, k t+4)
                    const float w0 = st ? aw[m][0].z : aw[m][0].x, w1 = st ? aw[m][1].z : aw[m][1].x;
                    const float w2 = st ? aw[m][0].w : aw[m][0].y, w3 = st ? aw[m][1].w : aw[m][1].y;
                    const float g0 = st ? ag[m][0].z : ag[m][0].x, g1 = st ? ag[m][1].z : ag[m][1].x;
                    const float g2 = st ? ag[m][0].w : ag[m][0].y, g3 = st ? ag[m][1].w : ag[m][1].y;
                    const uint32_t awh[4] = {__float_as_uint(w0), __float_as_uint(w1), __float_as_uint(w2), __float_as_uint(w3)};
                    const uint32_t awl[4] = {tf32_lo(w0), tf32_lo(w1), tf32_lo(w2), tf32_lo(w3)};
                    const uint32_t agh[4] = {__float_as_uint(g0), __float_as_uint(g1), __float_as_uint(g2), __float_as_uint(g3)};
                    const uint32_t agl[4] = {tf32_lo(g0), tf32_lo(g1), tf32_lo(g2), tf32_lo(g3)};
#pragma unroll
                    for (int n = 0; n < 2; ++n) {
                        const uint32_t sh0 = __float_as_uint(bs[n][0]), sh1 = __float_as_uint(bs[n][1]);
                        const uint32_t sl0 = tf32_lo(bs[n][0]), sl1 = tf32_lo(bs[n][1]);
                        const uint32_t mh0 = __float_as_uint(bm[n][0]), mh1 = __float_as_uint(bm[n][1]);
                        const uint32_t ml0 = tf32_lo(bm[n][0]), ml1 = tf32_lo(bm[n][1]);
                        // 3xTF32: lo * hi + hi * lo + hi * hi, small terms first
                        mma_tf32_16x8x8(acc[m][n], awl[0], awl[1], awl[2], awl[3], sh0, sh1);
                        mma_tf32_16x8x8(acc[m][n], awh[0], awh[1], awh[2], awh[3], sl0, sl1);
                        mma_tf32_16x8x8(acc[m][n], agl[0], agl[1], agl[2], agl[3], mh0, mh1);
                        mma_tf32_16x8x8(acc[m][n], agh[0], agh[1], agh[2], agh[3], ml0, ml1);
                        mma_tf32_16x8x8(acc[m][n], awh[0], awh[1], awh[2], awh[3], sh0, sh1);
                        mma_tf32_16x8x8(acc[m][n], agh[0], agh[1], agh[2], agh[3], mh0, mh1);
                    }
                }
            }
        }
        // D fragment: d0 (row g, col 2t) d1 (row g, col 2t+1) d2 (row g+8, col 2t) d3 (row g+8, col 2t+1)
#pragma unroll
        for (int m = 0; m < MT; ++m)
#pragma unroll
            for (int n = 0; n < 2; ++n) {
                const int xc = x0 + 8 * n + 2 * t;
                if (xc < E_pad) {
                    atomicAdd(&de_s[(16 * m + g) * E_pad + xc], acc[m][n][0]);
                    atomicAdd(&de_s[(16 * m + g) * E_pad + xc + 1], acc[m][n][1]);
                    atomicAdd(&de_s[(16 * m + g + 8) * E_pad + xc], acc[m][n][2]);
                    atomicAdd(&de_s[(16 * m + g + 8) * E_pad + xc + 1], acc[m][n][3]);
                }
            }
    }
    __syncthreads();
    // dT[id] += de * value: no gradient for ids outside [0, V) (the forward treated them as zero rows)
    for (int i = tid; i < F * E; i += blockDim.x) {
        const int f = i / E, x = i - f * E;
        const long long id = ids_i32 ? (long long)reinterpret_cast<const int *>(ids)[b * F + f]
                                     : reinterpret_cast<const long long *>(ids)[b * F + f];
        if ((unsigned long long)id < (unsigned long long)V)
            atomicAdd(dT + id * E + x, de_s[f * E_pad + x] * __ldg(values + b * F + f));
    }
}

// dW / dQ from the batch-accumulated dm [R][E] (gradient of the pre-contracted matrix, unscaled).  One warp per output
// element: the contraction runs over O (dW) or E (dQ) with the lanes striding it.
__global__ void bwd_attn_grad_kernel(const float *__restrict__ dm, const float *__restrict__ W, const float *__restrict__ Q,
                                     int lin_layout, int E, int D, int K, int O, float scale, float *__restrict__ dW,
                                     float *__restrict__ dQ) {
    const int nW = K * E * D, nQ = K * O * D;
    const int lane = threadIdx.x & 31;
    const int warps = (gridDim.x * blockDim.x) >> 5;
    for (int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; i < nW + nQ; i += warps) {
        float a = 0.f;
        if (i < nW) {
            // multi-head: dW[k][x][y] = sum_o dm[k,o,x] Q[k,o,y];  one-head (K = 1): dW_lin[y][x] = the same sum
            int k, x, y;
            if (lin_layout) {
                k = 0;
                y = i / E;
                x = i - y * E;
            } else {
                k = i / (E * D);
                const int rem = i - k * E * D;
                x = rem / D;
                y = rem - x * D;
            }
            for (int o = lane; o < O; o += 32)
                a = fmaf(dm[((long long)k * O + o) * E + x], Q[((long long)k * O + o) * D + y], a);
        } else {
            const int j = i - nW;
            const int r = j / D, y = j - r * D, k = r / O;
            for (int x = lane; x < E; x += 32) {
                const float w = lin_layout ? W[y * E + x] : W[((long long)k * E + x) * D + y];
                a = fmaf(dm[(long long)r * E + x], w, a);
            }
        }
#pragma unroll
        for (int m = 16; m > 0; m >>= 1) a += __shfl_xor_sync(0xffffffffu, a, m);
        if (lane == 0) {
            if (i < nW)
                dW[i] = a * scale;
            else
                dQ[i - nW] = a * scale;
        }
    }
}

}  // namespace armnet

extern "C" size_t armnet_fused_bwd_finish_workspace_bytes(int E, int K, int O) {
    if (E <= 0 || K <= 0 || O <= 0) return 0;
    return ((size_t)K * O * E * 4 + 15) / 16 * 16;
}

extern "C" int armnet_fused_bwd_finish_f32(const void *ids, int ids_i32, const float *values, int64_t V, int64_t B, int F,
                                           int E, int D, int K, int O, const float *bilinear_w, const float *query,
                                           int w_is_linear_layout, const float *w, const float *dg, const float *z,
                                           const float *dz, const float *acc_dm, float *dT, float *dW, float *dQ,
                                           void *workspace, void *stream) {
    using namespace armnet;
    note_launches(0);
    if (!ids || !values || !bilinear_w || !query || !w || !dg || !z || !dz || !acc_dm || !dT || !dW || !dQ || !workspace) {
        set_error("fused_bwd_finish: null pointer");
        return ARMNET_ERR_NULL;
    }
    if (V <= 0 || B < 0 || F <= 0 || E <= 0 || D <= 0 || K <= 0 || O <= 0 || (w_is_linear_layout && K != 1)) {
        set_error("fused_bwd_finish: bad shape V=%lld B=%lld F=%d E=%d D=%d K=%d O=%d", (long long)V, (long long)B, F, E, D, K,
                  O);
        return ARMNET_ERR_SHAPE;
    }
    if (F > 64) {
        set_error("fused_bwd_finish: %d fields (at most 64)", F);
        return ARMNET_ERR_UNSUPPORTED;
    }
    cudaStream_t st = (cudaStream_t)stream;
    const int R = K * O;
    const float scale = (float)pow((double)D, -0.5);
    float *Mt = (float *)workspace;
    bwd_mt_kernel<<<(R * E + 255) / 256, 256, 0, st>>>(bilinear_w, query, w_is_linear_layout, E, D, O, R, scale, Mt);
    ARMNET_CUDA_TRY(cudaGetLastError());
    int launches = 1;
    if (B > 0) {
        const int MT = (F + 15) / 16;
        const int E_pad = (E + 7) & ~7;
        const size_t smem = (size_t)16 * MT * E_pad * sizeof(float);
        DeviceInfo di;
        int rc = get_device_info(&di);
        if (rc != ARMNET_OK) return rc;
        if (smem > (size_t)di.smem_optin) {
            set_error("fused_bwd_finish: K*O=%d nemb=%d needs %zu bytes of shared memory per CTA (limit %d)", R, E, smem,
                      di.smem_optin);
            return ARMNET_ERR_UNSUPPORTED;
        }
        const void *kern = MT == 1 ? (const void *)bwd_embed_grad_kernel<1> : MT == 2 ? (const void *)bwd_embed_grad_kernel<2>
                         : MT == 3 ? (const void *)bwd_embed_grad_kernel<3> : (const void *)bwd_embed_grad_kernel<4>;
        ARMNET_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        const dim3 grid((unsigned)B), block(kFinWarps * 32);
        switch (MT) {
            case 1: bwd_embed_grad_kernel<1><<<grid, block, smem, st>>>(ids, ids_i32, values, w, dg, z, dz, Mt, V, F, E, R, dT); break;
            case 2: bwd_embed_grad_kernel<2><<<grid, block, smem, st>>>(ids, ids_i32, values, w, dg, z, dz, Mt, V, F, E, R, dT); break;
            case 3: bwd_embed_grad_kernel<3><<<grid, block, smem, st>>>(ids, ids_i32, values, w, dg, z, dz, Mt, V, F, E, R, dT); break;
            default: bwd_embed_grad_kernel<4><<<grid, block, smem, st>>>(ids, ids_i32, values, w, dg, z, dz, Mt, V, F, E, R, dT); break;
        }
        ARMNET_CUDA_TRY(cudaGetLastError());
        ++launches;
    }
    const int total = K * E * D + K * O * D;   // one warp per output element
    bwd_attn_grad_kernel<<<(total + 7) / 8, 256, 0, st>>>(acc_dm, bilinear_w, query, w_is_linear_layout, E, D, K, O, scale,
                                                             dW, dQ);
    ARMNET_CUDA_TRY(cudaGetLastError());
    note_launches(launches + 1);
    return ARMNET_OK;
}
