// Micro-benchmark for the round-2 plan of DESIGN.md 7 (one row per thread through TMEM): how fast are the SMALL
// tcgen05 shapes the fused forward needs, and how fast can warps move rows in and out of TMEM?
//   (1) tcgen05.mma.kind::tf32, M = 128, K = 8, A and B from shared memory (K-major, SWIZZLE_128B), N in {16 .. 256}:
//       cycles per MMA when one thread issues a long chain (the guide's floor is M*N/256 cycles; N = 40 -> 20).
//   (2) the same with the A operand in TMEM (the cross product: A = gates*values written by the threads, N = 16).
//   (3) tcgen05.ld.32x32b.x32 / tcgen05.st.32x32b.x32 by 4 warps (128 rows): cycles per 128 x 32 fp32 block.
//   (4) [added after the first run, NOT RUN YET] the same MMAs rotating over 8 INDEPENDENT accumulator blocks: is the
//       ~119 cycles per small MMA measured in (1)/(2) a latency of dependent accumulation or a fixed issue interval?
// First run (round 1, last GPU seconds; profiles/r1_v7_tcgen05_small_ubench.txt): (1) 119 cycles per MMA for every
// N <= 128 (161 at N = 256), (2) 118.6 cycles, (3) 50 / 52 cycles per 128 x 32 block.
// Build / run:  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tcgen05_small tcgen05_small.cu && ./tcgen05_small
// Every wait is bounded (a stuck barrier traps instead of hanging the box); wrap the run in `timeout 60` anyway.
// The PTX wrappers are the ones armnet_b200/csrc/mlp.cu runs in production, plus tcgen05.st and the TMEM-A MMA form.
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ bool mbar_try(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait_bounded(uint64_t *bar, uint32_t parity) {
    for (long long spin = 0; spin < (1ll << 26); ++spin)
        if (mbar_try(bar, parity)) return;
    asm volatile("trap;");  // never hang the box
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_alloc(uint32_t *slot, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}
// D[tmem] (+)= A[smem] . B[smem]^T
__device__ __forceinline__ void umma_tf32_ss(uint32_t d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n" ::"r"(d),
        "l"(da), "l"(db), "r"(idesc), "r"(acc)
        : "memory");
}
// D[tmem] (+)= A[tmem] . B[smem]^T   (A: 128 lanes x K 32-bit columns at `a`)
__device__ __forceinline__ void umma_tf32_ts(uint32_t d, uint32_t a, uint64_t db, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n}\n" ::"r"(d),
        "r"(a), "l"(db), "r"(idesc), "r"(acc)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_st_32x32(uint32_t taddr, const uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
        "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]),
        "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]),
        "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// K-major operand tile, rows of 128 bytes, SWIZZLE_128B (see mlp.cu: umma_desc_sw128)
__device__ __forceinline__ uint64_t umma_desc_sw128(const void *tile) {
    const uint64_t addr = (uint64_t)((smem_u32(tile) & 0x3FFFFu) >> 4);
    return addr | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
__host__ __device__ constexpr uint32_t umma_idesc_tf32(int m, int n) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

// mode 0: SS MMAs; 1: TS MMAs (A in TMEM); 2: tcgen05.ld loop; 3: tcgen05.st loop; 4 / 5: SS / TS MMAs over 8
// independent accumulator blocks, never accumulating.  out[blockIdx] = cycles of the loop.
__global__ void __launch_bounds__(128, 1) bench_kernel(int mode, int N, int iters, long long *out, float *sink) {
    extern __shared__ __align__(1024) unsigned char smem[];
    float *tileA = reinterpret_cast<float *>(smem);              // 128 rows x 32 floats (K = 32: four K = 8 steps)
    float *tileB = reinterpret_cast<float *>(smem + 16384);      // up to 256 rows x 32 floats
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base_slot;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < (16384 + 32768) / 4; i += blockDim.x) reinterpret_cast<float *>(smem)[i] = 1.0f + (i & 7);
    if (tid == 0) {
        mbar_init(&bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic writes -> visible to the tensor core
    if (warp == 0) tmem_alloc(&tmem_base_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_base_slot;
    long long t0 = 0, t1 = 0;
    if (mode == 4 || mode == 5) {
        if (tid == 0) {
            const uint32_t idesc = umma_idesc_tf32(128, N);
            const uint64_t da = umma_desc_sw128(tileA), db = umma_desc_sw128(tileB);
            const uint32_t stride = (uint32_t)(N <= 32 ? N : 32);  // 8 blocks inside columns [256, 512) when N <= 32
            t0 = clock64();
            for (int i = 0; i < iters; ++i) {
                const uint64_t adv = (uint64_t)((i & 3) * 32 >> 4);
                const uint32_t dcol = tmem + 256 + (uint32_t)(i & 7) * stride;
                if (mode == 4)
                    umma_tf32_ss(dcol, da + adv, db + adv, idesc, 0u);
                else
                    umma_tf32_ts(dcol, tmem + (uint32_t)((i & 3) * 8), db + adv, idesc, 0u);
            }
            umma_commit(&bar);
            mbar_wait_bounded(&bar, 0);
            t1 = clock64();
        }
    } else if (mode <= 1) {
        if (tid == 0) {
            const uint32_t idesc = umma_idesc_tf32(128, N);
            const uint64_t da = umma_desc_sw128(tileA), db = umma_desc_sw128(tileB);
            t0 = clock64();
            for (int i = 0; i < iters; ++i) {
                const uint64_t adv = (uint64_t)((i & 3) * 32 >> 4);       // next K = 8 slice: +32 bytes
                const uint32_t dcol = tmem + 256 + (uint32_t)((i >> 2) & 1) * (uint32_t)(N <= 128 ? N : 0);
                if (mode == 0)
                    umma_tf32_ss(dcol, da + adv, db + adv, idesc, (uint32_t)(i & 3));
                else
                    umma_tf32_ts(dcol, tmem + (uint32_t)((i & 3) * 8), db + adv, idesc, (uint32_t)(i & 3));
            }
            umma_commit(&bar);
            mbar_wait_bounded(&bar, 0);
            t1 = clock64();
        }
    } else {
        uint32_t r[32];
        for (int i = 0; i < 32; ++i) r[i] = tid + i;
        const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16);  // this warp's 32 lanes
        __syncthreads();
        t0 = clock64();
        for (int i = 0; i < iters; ++i) {
            const uint32_t col = (uint32_t)((i & 7) * 32);
            if (mode == 2) {
                tmem_ld_32x32(taddr + col, r);
                if ((i & 3) == 3) tmem_ld_wait();
            } else {
                tmem_st_32x32(taddr + col, r);
                if ((i & 3) == 3) tmem_st_wait();
            }
        }
        tmem_ld_wait();
        tmem_st_wait();
        __syncthreads();
        t1 = clock64();
        float s = 0.f;
        for (int i = 0; i < 32; ++i) s += __uint_as_float(r[i]);
        if (s == 123.456f) sink[tid] = s;
    }
    tc_fence_before();
    __syncthreads();
    if (tid == 0) out[blockIdx.x] = t1 - t0;
    if (warp == 0) tmem_dealloc(tmem, 512);
}

static int run(const char *name, int mode, int N, int iters, long long *d_out, float *d_sink, double per) {
    cudaFuncSetAttribute(bench_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384 + 32768 + 1024);
    bench_kernel<<<148, 128, 16384 + 32768 + 1024>>>(mode, N, iters, d_out, d_sink);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
        printf("%-34s N=%3d: %s\n", name, N, cudaGetErrorString(e));
        return 1;
    }
    long long h[148];
    cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost);
    long long mx = 0;
    for (int i = 0; i < 148; ++i) mx = h[i] > mx ? h[i] : mx;
    printf("%-34s N=%3d: %8.1f cycles per op (floor M*N/256 = %.1f)\n", name, N, (double)mx / iters, per);
    return 0;
}

int main() {
    long long *d_out;
    float *d_sink;
    cudaMalloc(&d_out, 148 * sizeof(long long));
    cudaMalloc(&d_sink, 128 * sizeof(float));
    const int iters = 4096;
    const int Ns[] = {16, 32, 40, 48, 64, 128, 256};
    for (int N : Ns)
        if (run("tcgen05.mma tf32 M=128 K=8, A smem", 0, N, iters, d_out, d_sink, 128.0 * N / 256)) break;  // sticky error
    for (int N : {16, 32, 64})
        if (run("tcgen05.mma tf32 M=128 K=8, A TMEM", 1, N, iters, d_out, d_sink, 128.0 * N / 256)) break;
    for (int N : {16, 32})
        if (run("... independent accumulators, A smem", 4, N, iters, d_out, d_sink, 128.0 * N / 256)) break;
    for (int N : {16, 32})
        if (run("... independent accumulators, A TMEM", 5, N, iters, d_out, d_sink, 128.0 * N / 256)) break;
    run("tcgen05.ld 32x32b.x32, 4 warps", 2, 0, iters, d_out, d_sink, 0);
    run("tcgen05.st 32x32b.x32, 4 warps", 3, 0, iters, d_out, d_sink, 0);
    return 0;
}
