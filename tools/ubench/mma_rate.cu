// Micro-benchmark: issue rate of the warp-level mma.sync.m16n8k8 TF32 instruction on sm_100a (B200), alone and
// interleaved with packed-FP32 work, to decide whether the per-sample E x F products of the fused forward can move to
// the tensor cores without TMEM.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_rate mma_rate.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

template <int NACC, int FMA_PER_MMA>
__global__ void __launch_bounds__(512, 1) rate_kernel(float *out, int iters, float seed) {
    float d[NACC][4];
    uint32_t a[4], b[2];
    for (int i = 0; i < 4; ++i) a[i] = __float_as_uint(seed + threadIdx.x * 1e-3f + i);
    for (int i = 0; i < 2; ++i) b[i] = __float_as_uint(seed * 0.5f + i);
    for (int n = 0; n < NACC; ++n)
        for (int i = 0; i < 4; ++i) d[n][i] = 0.f;
    float2 f[8];
    for (int i = 0; i < 8; ++i) f[i] = make_float2(seed + i, seed - i);
    const float2 k = make_float2(1.0001f, 0.9999f);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int n = 0; n < NACC; ++n) {
            mma_tf32(d[n], a, b);
#pragma unroll
            for (int j = 0; j < FMA_PER_MMA; ++j) {
                float2 &x = f[(n * FMA_PER_MMA + j) % 8];
                asm volatile("fma.rn.f32x2 %0, %1, %2, %0;"
                             : "+l"(reinterpret_cast<uint64_t &>(x))
                             : "l"(reinterpret_cast<const uint64_t &>(x)), "l"(reinterpret_cast<const uint64_t &>(k)));
            }
        }
    }
    float s = 0.f;
    for (int n = 0; n < NACC; ++n)
        for (int i = 0; i < 4; ++i) s += d[n][i];
    for (int i = 0; i < 8; ++i) s += f[i].x + f[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NACC, int FMA>
static void run(const char *name, int warps, float *out) {
    const int iters = 4096;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    rate_kernel<NACC, FMA><<<148, warps * 32>>>(out, 64, 1.f);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    rate_kernel<NACC, FMA><<<148, warps * 32>>>(out, iters, 1.f);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    const double mmas_per_smsp = (double)iters * NACC * warps / 4.0;
    const double cycles = ms * 1e-3 * 1.965e9;
    printf("%-28s warps/SM %2d: %.3f ms, %.2f cycles per MMA per SMSP, %.1f TFLOP/s TF32, FFMA2/MMA %d\n", name, warps, ms,
           cycles / mmas_per_smsp, 148.0 * warps * iters * NACC * 2048.0 / (ms * 1e-3) / 1e12, FMA);
}

int main() {
    float *out;
    cudaMalloc(&out, 148 * 512 * sizeof(float));
    run<8, 0>("mma only, 8 acc", 4, out);
    run<8, 0>("mma only, 8 acc", 8, out);
    run<8, 0>("mma only, 8 acc", 12, out);
    run<8, 0>("mma only, 8 acc", 16, out);
    run<2, 0>("mma only, 2 acc", 12, out);
    run<8, 2>("mma + 2 FFMA2", 12, out);
    run<8, 4>("mma + 4 FFMA2", 12, out);
    run<8, 8>("mma + 8 FFMA2", 12, out);
    run<8, 8>("mma + 8 FFMA2", 16, out);
    cudaError_t e = cudaDeviceSynchronize();
    printf("status: %s\n", cudaGetErrorString(e));
    return e != cudaSuccess;
}
