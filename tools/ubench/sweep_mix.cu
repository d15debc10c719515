// Issue-rate ceiling of the fused entmax + cross-product sweep's instruction mix on one SM (round-2 question: every
// mapping of armnet_fwd_tmem_kernel lands at ~2.0 warp instructions per cycle per SM -- is that the hardware's limit for
// this mix or a scheduling problem?).  Per field pair and row: FADD2, 2 FMNMX, 2 MUFU.LG2, FMUL2, 2 MUFU.EX2, FADD2, FMUL2,
// FADD2, FMUL2 + 10 FFMA2 [+ 6 broadcast LDS + 1 LDS.128].  Variants: packed FFMA2 vs scalar FFMA, with / without the
// shared-memory reads, with / without the MUFU chain, 2 / 3 / 4 warps per SM sub-partition.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o sweep_mix sweep_mix.cu && ./sweep_mix
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
    float2 d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(reinterpret_cast<uint64_t &>(d))
        : "l"(reinterpret_cast<uint64_t &>(a)), "l"(reinterpret_cast<uint64_t &>(b)), "l"(reinterpret_cast<uint64_t &>(c)));
    return d;
}
__device__ __forceinline__ float2 fadd2(float2 a, float2 b) {
    float2 d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(reinterpret_cast<uint64_t &>(d))
        : "l"(reinterpret_cast<uint64_t &>(a)), "l"(reinterpret_cast<uint64_t &>(b)));
    return d;
}
__device__ __forceinline__ float2 fmul2(float2 a, float2 b) {
    float2 d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(reinterpret_cast<uint64_t &>(d))
        : "l"(reinterpret_cast<uint64_t &>(a)), "l"(reinterpret_cast<uint64_t &>(b)));
    return d;
}
__device__ __forceinline__ float lg2f_(float x) { float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float ex2f_(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

constexpr int NP = 20;

// MODE bits: 1 = packed FFMA2 (else scalar FFMA), 2 = e / V from shared memory (else registers), 4 = MUFU chain on
template <int MODE>
__global__ void __launch_bounds__(512, 1) mix_kernel(const float *in, float *out, long long *cycles, int reps) {
    extern __shared__ __align__(16) float sm[];
    for (int i = threadIdx.x; i < 40 * 32 + 32 * 44; i += blockDim.x) sm[i] = 0.001f * (i % 97);
    __syncthreads();
    float2 X[NP];
#pragma unroll
    for (int j = 0; j < NP; ++j) X[j] = make_float2(in[threadIdx.x] + 0.01f * j, in[threadIdx.x] + 0.013f * j);
    float2 acc[5];
    float accs[10];
#pragma unroll
    for (int x = 0; x < 5; ++x) acc[x] = make_float2(0.f, 0.f);
#pragma unroll
    for (int x = 0; x < 10; ++x) accs[x] = 0.f;
    float2 s = make_float2(0.f, 0.f), s1 = s;
    float tau = in[0] * 0.5f;
    const float4 *vr = reinterpret_cast<const float4 *>(sm + 40 * 32 + (threadIdx.x & 31) * 44);
    const float2 qm1 = make_float2(0.4286f, 0.4286f);
    long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
        const float2 nt = make_float2(-tau, -tau);
#pragma unroll
        for (int j = 0; j < NP; ++j) {
            float2 u = fadd2(X[j], nt);
            u.x = fmaxf(u.x, 0.f);
            u.y = fmaxf(u.y, 0.f);
            float2 g;
            if (MODE & 4) {
                const float2 l = fmul2(make_float2(lg2f_(u.x), lg2f_(u.y)), qm1);
                g = make_float2(ex2f_(l.x), ex2f_(l.y));
            } else {
                g = fmul2(u, qm1);
            }
            s1 = fadd2(s1, g);
            const float2 p = fmul2(g, u);
            s = fadd2(s, p);
            float2 v;
            if (MODE & 2) {
                const float4 v4 = vr[j >> 1];
                v = (j & 1) ? make_float2(v4.z, v4.w) : make_float2(v4.x, v4.y);
            } else {
                v = make_float2(1.01f, 0.99f);
            }
            const float2 w = fmul2(p, v);
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const float wf = h ? w.y : w.x;
                float e[10];
                if (MODE & 2) {
                    const float *row = sm + (2 * j + h) * 32;
                    const float4 c0 = *reinterpret_cast<const float4 *>(row);
                    const float4 c1 = *reinterpret_cast<const float4 *>(row + 4);
                    const float2 c2 = *reinterpret_cast<const float2 *>(row + 8);
                    e[0] = c0.x; e[1] = c0.y; e[2] = c0.z; e[3] = c0.w; e[4] = c1.x; e[5] = c1.y; e[6] = c1.z; e[7] = c1.w;
                    e[8] = c2.x; e[9] = c2.y;
                } else {
#pragma unroll
                    for (int x = 0; x < 10; ++x) e[x] = X[(j + x) % NP].x;
                }
                if (MODE & 1) {
                    const float2 w2 = make_float2(wf, wf);
#pragma unroll
                    for (int x = 0; x < 5; ++x) acc[x] = ffma2(w2, make_float2(e[2 * x], e[2 * x + 1]), acc[x]);
                } else {
#pragma unroll
                    for (int x = 0; x < 10; ++x) accs[x] = fmaf(wf, e[x], accs[x]);
                }
            }
        }
        tau += 1e-9f * (s.x + s1.y);
    }
    long long t1 = clock64();
    float r = s.x + s.y + s1.x + s1.y;
#pragma unroll
    for (int x = 0; x < 5; ++x) r += acc[x].x + acc[x].y;
#pragma unroll
    for (int x = 0; x < 10; ++x) r += accs[x];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char *name, int warps, float *in, float *out, long long *cy) {
    const int reps = 200;
    const int smem = (40 * 32 + 32 * 44) * 4;
    mix_kernel<MODE><<<148, warps * 32, smem>>>(in, out, cy, reps);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%s: %s\n", name, cudaGetErrorString(e)); return; }
    long long h[148];
    cudaMemcpy(h, cy, sizeof(h), cudaMemcpyDeviceToHost);
    double mx = 0;
    for (int i = 0; i < 148; ++i) mx = h[i] > mx ? h[i] : mx;
    // per SM sub-partition: warps/4 warps x reps x NP field pairs
    const double per_pair = mx / (reps * NP * (warps / 4.0));
    printf("%-44s %2d warps: %7.1f SMSP-cycles per (row-warp, field pair)\n", name, warps, per_pair);
}

int main() {
    float *in, *out;
    long long *cy;
    cudaMalloc(&in, 4096);
    cudaMalloc(&out, 148 * 512 * 4);
    cudaMalloc(&cy, 148 * 8);
    float h[1024];
    for (int i = 0; i < 1024; ++i) h[i] = 0.5f + 0.0001f * i;
    cudaMemcpy(in, h, 4096, cudaMemcpyHostToDevice);
    for (int warps : {8, 12, 16}) {
        run<1 | 2 | 4>("FFMA2 + LDS + MUFU (the real mix)", warps, in, out, cy);
        run<1 | 4>("FFMA2 + MUFU, no LDS", warps, in, out, cy);
        run<1 | 2>("FFMA2 + LDS, no MUFU", warps, in, out, cy);
        run<1>("FFMA2 only (+ packed entmax ops)", warps, in, out, cy);
        run<2 | 4>("scalar FFMA + LDS + MUFU", warps, in, out, cy);
        run<0>("scalar FFMA only", warps, in, out, cy);
    }
    return 0;
}
