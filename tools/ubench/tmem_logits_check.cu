// Correctness + timing check of the tcgen05 building blocks of armnet_fwd_tmem_kernel (csrc/fused_fwd_tmem.cuh):
//   logits X[r][(b,f)] = sum_x M'[x][r] e[b,f,x] as ONE chain of four K = 8 TF32 MMAs over a "packed 3xTF32" K axis:
//     A row (neuron r)  = [ M0..M9 | M0..M5 | M6..M9 | m0..m9 | 0 0 ]      (M = fp32 value, m = M - trunc_tf32(M))
//     B row (sample b, field f) = [ e0..e9 | l0..l5 | l6..l9 | e0..e9 | 0 0 ]   (l = e - trunc_tf32(e))
//   so that sum_k A[k] B[k] = sum_x (M e + M l + m e): fp32-grade accuracy if the tensor core TRUNCATES its fp32 inputs
//   to TF32 (top 19 bits), which is what this program verifies (a rounding core would leave ~2^-11 relative errors).
//   A comes from shared memory (SS form) and from TMEM (TS form, written by tcgen05.st); B tiles of 80 rows (2 samples x 40
//   fields), rows of 128 bytes, SWIZZLE_128B written BY THREADS (no TMA): chunk c of row r lives at chunk c ^ (r & 7).
//   D (128 lanes x 80 columns fp32) is read back with tcgen05.ld.32x32b.x16.
// Also times: 8 MMAs (2 A blocks x 4 K steps, N = 80) per item, back to back over two D slots.
// Build / run: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_logits_check tmem_logits_check.cu && ./tmem_logits_check
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
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
    asm volatile("trap;");
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
__device__ __forceinline__ void umma_tf32_ss(uint32_t d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n" ::"r"(d),
        "l"(da), "l"(db), "r"(idesc), "r"(acc)
        : "memory");
}
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
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&r)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(r[0]),
                 "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
                 : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ uint64_t umma_desc_sw128(const void *tile) {
    const uint64_t addr = (uint64_t)((smem_u32(tile) & 0x3FFFFu) >> 4);
    return addr | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
__host__ __device__ constexpr uint32_t umma_idesc_tf32(int m, int n) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

constexpr int NB = 80;   // B rows (2 samples x 40 fields)
constexpr int KP = 32;   // packed K

// packed [rows][32] fp32 (row-major, unswizzled) -> SWIZZLE_128B tile in shared memory
__device__ void fill_sw128(float *tile, const float *src, int rows, int tid, int nthr) {
    for (int i = tid; i < rows * 8; i += nthr) {
        const int r = i >> 3, c = i & 7;
        const float4 v = reinterpret_cast<const float4 *>(src)[r * 8 + c];
        reinterpret_cast<float4 *>(tile)[r * 8 + (c ^ (r & 7))] = v;
    }
}

// out_ss / out_ts: [128][80]; cycles[0] = timing loop (iters items of 8 MMAs)
__global__ void __launch_bounds__(128, 1) check_kernel(const float *A, const float *Bm, float *out_ss, float *out_ts,
                                                       long long *cycles, int iters) {
    extern __shared__ __align__(1024) unsigned char smem[];
    float *tileA = reinterpret_cast<float *>(smem);            // 128 x 128 B = 16 KB
    float *tileB = reinterpret_cast<float *>(smem + 16384);    // 80 x 128 B = 10 KB
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    fill_sw128(tileA, A, 128, tid, 128);
    fill_sw128(tileB, Bm, NB, tid, 128);
    if (tid == 0) {
        mbar_init(&bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (warp == 0) tmem_alloc(&tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;
    // A operand into TMEM columns [0, 32): lane = row
    {
        const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16);
        for (int c8 = 0; c8 < 4; ++c8) {
            uint32_t r[8];
            for (int j = 0; j < 8; ++j) r[j] = __float_as_uint(A[tid * KP + c8 * 8 + j]);
            tmem_st8(taddr + c8 * 8, r);
        }
        tmem_st_wait();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t idesc = umma_idesc_tf32(128, NB);
    const uint32_t d_ss = tmem + 64, d_ts = tmem + 192;
    if (tid == 0) {
        const uint64_t da = umma_desc_sw128(tileA), db = umma_desc_sw128(tileB);
        for (int k = 0; k < 4; ++k) umma_tf32_ss(d_ss, da + (uint64_t)(k * 32 >> 4), db + (uint64_t)(k * 32 >> 4), idesc, k);
        for (int k = 0; k < 4; ++k) umma_tf32_ts(d_ts, tmem + (uint32_t)(k * 8), db + (uint64_t)(k * 32 >> 4), idesc, k);
        umma_commit(&bar);
    }
    mbar_wait_bounded(&bar, 0);
    tc_fence_after();
    {
        const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
        for (int c = 0; c < NB; c += 16) {
            uint32_t r[16];
            tmem_ld16(d_ss + lane_base + c, r);
            tmem_ld_wait();
            for (int j = 0; j < 16; ++j) out_ss[tid * NB + c + j] = __uint_as_float(r[j]);
            tmem_ld16(d_ts + lane_base + c, r);
            tmem_ld_wait();
            for (int j = 0; j < 16; ++j) out_ts[tid * NB + c + j] = __uint_as_float(r[j]);
        }
    }
    tc_fence_before();
    __syncthreads();
    // timing: items of 8 MMAs (2 A blocks x 4 K steps) alternating between two D slots of 160 columns
    long long t0 = 0, t1 = 0;
    if (tid == 0) {
        tc_fence_after();
        const uint64_t da = umma_desc_sw128(tileA), db = umma_desc_sw128(tileB);
        t0 = clock64();
        for (int it = 0; it < iters; ++it) {
            const uint32_t d = tmem + 128 + (uint32_t)(it & 1) * 160;
            for (int j = 0; j < 2; ++j)
                for (int k = 0; k < 4; ++k)
                    umma_tf32_ts(d + j * 80, tmem + (uint32_t)(k * 8), db + (uint64_t)(k * 32 >> 4), idesc, k);
        }
        umma_commit(&bar);
        mbar_wait_bounded(&bar, 1);
        t1 = clock64();
        cycles[0] = t1 - t0;
        t0 = clock64();
        for (int it = 0; it < iters; ++it) {
            const uint32_t d = tmem + 128 + (uint32_t)(it & 1) * 160;
            for (int j = 0; j < 2; ++j)
                for (int k = 0; k < 4; ++k)
                    umma_tf32_ss(d + j * 80, da + (uint64_t)(k * 32 >> 4), db + (uint64_t)(k * 32 >> 4), idesc, k);
        }
        umma_commit(&bar);
        mbar_wait_bounded(&bar, 0);
        t1 = clock64();
        cycles[1] = t1 - t0;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512);
    (void)lane;
}

static float trunc_tf32(float x) {
    uint32_t u;
    memcpy(&u, &x, 4);
    u &= 0xffffe000u;
    float y;
    memcpy(&y, &u, 4);
    return y;
}

int main() {
    const int E = 10;
    std::vector<float> M(128 * E), e(NB * E), A(128 * KP, 0.f), Bp(NB * KP, 0.f);
    srand(1);
    auto rnd = [] { return (float)rand() / RAND_MAX * 2.f - 1.f; };
    for (auto &v : M) v = rnd() * 0.3f;
    for (auto &v : e) v = rnd() * 2.f;
    for (int r = 0; r < 128; ++r) {
        float *a = &A[r * KP];
        const float *m = &M[r * E];
        for (int x = 0; x < 10; ++x) a[x] = m[x];
        for (int x = 0; x < 6; ++x) a[10 + x] = m[x];
        for (int x = 6; x < 10; ++x) a[16 + x - 6] = m[x];
        for (int x = 0; x < 10; ++x) a[20 + x] = m[x] - trunc_tf32(m[x]);
    }
    for (int n = 0; n < NB; ++n) {
        float *b = &Bp[n * KP];
        const float *v = &e[n * E];
        for (int x = 0; x < 10; ++x) b[x] = v[x];
        for (int x = 0; x < 6; ++x) b[10 + x] = v[x] - trunc_tf32(v[x]);
        for (int x = 6; x < 10; ++x) b[16 + x - 6] = v[x] - trunc_tf32(v[x]);
        for (int x = 0; x < 10; ++x) b[20 + x] = v[x];
    }
    float *dA, *dB, *dss, *dts;
    long long *dcy;
    cudaMalloc(&dA, A.size() * 4);
    cudaMalloc(&dB, Bp.size() * 4);
    cudaMalloc(&dss, 128 * NB * 4);
    cudaMalloc(&dts, 128 * NB * 4);
    cudaMalloc(&dcy, 16);
    cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, Bp.data(), Bp.size() * 4, cudaMemcpyHostToDevice);
    const int iters = 2048;
    cudaFuncSetAttribute(check_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384 + 10240 + 1024);
    check_kernel<<<1, 128, 16384 + 10240 + 1024>>>(dA, dB, dss, dts, dcy, iters);
    cudaError_t err = cudaDeviceSynchronize();
    if (err != cudaSuccess) {
        printf("kernel failed: %s\n", cudaGetErrorString(err));
        return 1;
    }
    std::vector<float> ss(128 * NB), ts(128 * NB);
    long long cy[2];
    cudaMemcpy(ss.data(), dss, ss.size() * 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(ts.data(), dts, ts.size() * 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(cy, dcy, 16, cudaMemcpyDeviceToHost);
    double mx = 0, ess = 0, ets = 0, e1 = 0;
    for (int r = 0; r < 128; ++r)
        for (int n = 0; n < NB; ++n) {
            double ref = 0, one = 0;
            for (int x = 0; x < E; ++x) {
                ref += (double)M[r * E + x] * e[n * E + x];
                one += (double)trunc_tf32(M[r * E + x]) * trunc_tf32(e[n * E + x]);
            }
            mx = fmax(mx, fabs(ref));
            ess = fmax(ess, fabs(ss[r * NB + n] - ref));
            ets = fmax(ets, fabs(ts[r * NB + n] - ref));
            e1 = fmax(e1, fabs(one - ref));
        }
    printf("max|ref| %.4f   SS: max err %.3e (norm-rel %.3e)   TS: max err %.3e (norm-rel %.3e)   plain 1xTF32 would be %.3e\n",
           mx, ess, ess / mx, ets, ets / mx, e1 / mx);
    printf("sample: ref-ish ss[5][7]=%.7f ts[5][7]=%.7f\n", ss[5 * NB + 7], ts[5 * NB + 7]);
    printf("timing: TS item (8 MMAs, N=80): %.1f cycles;  SS item: %.1f cycles\n", (double)cy[0] / iters, (double)cy[1] / iters);
    const bool ok = ess / mx < 2e-6 && ets / mx < 2e-6;
    printf(ok ? "PASS\n" : "FAIL\n");
    return ok ? 0 : 2;
}
