// The trailing dense MLP of ARMNetModel.forward in eval mode (models/layers.py:68-88 called from models/armnet.py:92,
// models/armnet_1h.py:89): [Linear -> BatchNorm1d -> ReLU -> Dropout] x n -> Linear(., noutput).
//
// The first Linear ([B, K*O*E] x [K*O*E, mlp_nhid], 10.7 GFLOP at the Criteo shape) is the one true GEMM of the hot
// path and runs on the 5th-generation tensor cores (tcgen05.mma, accumulators in TMEM, operands staged by TMA).
// fp32 parity (the reference multiplies in fp32, layers.py:73) is kept with the 3xTF32 split:
//     x = x_hi + x_lo,  w = w_hi + w_lo   (hi = fp32 rounded to the 10-bit TF32 mantissa, lo = exact remainder)
//     x.w ~= x_hi.w_hi + x_lo.w_hi + x_hi.w_lo          (dropped x_lo.w_lo ~ 2^-22 relative)
// three kind::tf32 MMAs per K step, the dominant term and the two corrections in separate fp32 TMEM accumulators
// (the tensor core rounds each accumulation towards zero).  w_hi / w_lo are split once per weight
// version (armnet_mlp_split_weight_f32); x is split in shared memory by the converter warps of the GEMM kernel.
//
// Two GEMM kernels: mlp_gemm_tf32x3_pair_kernel (CTA pair, cta_group::2: the default) and the single-CTA
// mlp_gemm_tf32x3_kernel it grew out of (kept for A/B runs behind ARMNET_GEMM_1CTA).
//
// Everything after the first Linear (its bias + BatchNorm + ReLU, the narrow hidden layers, the output Linear) is
// one CUDA-core kernel (mlp_tail_kernel): < 3 % of the FLOPs, so one launch instead of ~8.
#include <cuda.h>
#include <stdlib.h>

#include "entmax_pair.cuh"  // ffma2 / splat2 (packed fp32)

namespace armnet {

// ------------------------------------------------------------------ tile configuration of the tensor-core GEMM
constexpr int GM = 128;      // rows of x per CTA (UMMA M, one TMEM lane per row)
constexpr int GN = 256;      // output features per CTA (UMMA N, one TMEM column per feature)
constexpr int GK = 32;       // floats of the reduction axis per stage = 128 bytes = one SWIZZLE_128B row
constexpr int UK = 8;        // UMMA K for kind::tf32 (32 bytes)
constexpr int G_STAGES = 2;
constexpr int G_A_BYTES = GM * GK * 4;                              // 16 KB: x tile (hi after conversion)
constexpr int G_B_BYTES = GN * GK * 4;                              // 32 KB: w_hi or w_lo tile
constexpr int G_STAGE_BYTES = 2 * G_A_BYTES + 2 * G_B_BYTES;        // 96 KB: x_hi, x_lo, w_hi, w_lo
constexpr int G_SMEM_BYTES = G_STAGES * G_STAGE_BYTES + 1024 /*alignment*/ + 256 /*barriers*/;
constexpr int G_THREADS = 192;  // warp 0 TMA producer, warp 1 MMA issuer / TMEM owner, warps 2-5 converter + epilogue
constexpr int G_CONV_THREADS = 128;

// ------------------------------------------------------------------ PTX wrappers (tcgen05 / TMA tensor copies)
__device__ __forceinline__ void tma_load_2d(void *smem_dst, const CUtensorMap *map, int c0, int c1, uint64_t *bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::
            "r"(smem_u32(smem_dst)),
        "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void prefetch_tensormap(const CUtensorMap *map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_alloc(uint32_t *smem_slot, uint32_t cols) {  // one full warp
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_slot)),
                 "r"(cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t cols) {  // the same warp that allocated
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}
// D[tmem] (+)= A[smem] . B[smem]^T, both operands K-major, fp32 accumulate. Issued by ONE thread.
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                          uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// mbarrier arrive once every tcgen05.mma issued so far by this thread has completed (implies fence::before_thread_sync)
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}
// 32 lanes x 32 consecutive columns of fp32 accumulators -> 32 registers per thread (thread = TMEM lane = row)
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
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// Shared-memory matrix descriptor of a K-major operand tile whose rows are 128 bytes, SWIZZLE_128B (what TMA wrote):
// start address >> 4 in bits [0,14); leading byte offset (ignored for swizzled K-major, 1) in [16,30); stride byte
// offset = 8 rows x 128 B = 1024 B >> 4 in [32,46); descriptor version 1 (sm_100) in [46,48); layout type 2 in [61,64).
__device__ __forceinline__ uint64_t umma_desc_sw128(const void *smem_tile) {
    const uint64_t addr = (uint64_t)((smem_u32(smem_tile) & 0x3FFFFu) >> 4);
    return addr | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
// Instruction descriptor, kind::tf32: D fp32 (bits 4-5 = 1), A/B TF32 (bits 7-9, 10-12 = 2), both K-major (bits 15,16 = 0),
// N >> 3 in bits [17,23), M >> 4 in bits [24,29).
__host__ __device__ constexpr uint32_t umma_idesc_tf32(int m, int n) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

// fp32 -> nearest value with a 10-bit mantissa (ties away from zero in magnitude; the remainder x - hi is exact)
__device__ __forceinline__ float tf32_hi(float x) {
    return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u);
}

__global__ void split_tf32_kernel(const float *__restrict__ w, float *__restrict__ hi, float *__restrict__ lo,
                                  long long n) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        const float x = w[i];
        const float h = tf32_hi(x);
        hi[i] = h;
        lo[i] = x - h;
    }
}

struct GemmParams {
    float *out;          // [splits][MB][N][32] partial sums over each split's slice of the reduction axis
    float *dense;        // or (splits == 1 only): the product itself, row-major [M][N] (+ bias[n] when bias != null)
    const float *bias;
    int M, N, K;
    int MB;              // ceil(M / 32) sample blocks
    int kb_total;        // ceil(K / GK)
    int splits;
};

// grid = (ceil(M/GM), ceil(N/GN), splits), 192 threads, one output tile per CTA.
__global__ void __launch_bounds__(G_THREADS, 1)
    mlp_gemm_tf32x3_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_wh,
                           const __grid_constant__ CUtensorMap map_wl, const GemmParams P) {
    extern __shared__ unsigned char smem_raw[];
    // SWIZZLE_128B tiles need 1024-byte alignment
    unsigned char *smem = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    // stage s: x_hi | x_lo | w_hi | w_lo
    auto s_xh = [&](int s) { return smem + s * G_STAGE_BYTES; };
    auto s_xl = [&](int s) { return smem + s * G_STAGE_BYTES + G_A_BYTES; };
    auto s_wh = [&](int s) { return smem + s * G_STAGE_BYTES + 2 * G_A_BYTES; };
    auto s_wl = [&](int s) { return smem + s * G_STAGE_BYTES + 2 * G_A_BYTES + G_B_BYTES; };
    uint64_t *bar_full = reinterpret_cast<uint64_t *>(smem + G_STAGES * G_STAGE_BYTES);  // TMA landed (x, w_hi, w_lo)
    uint64_t *bar_conv = bar_full + G_STAGES;                                            // x split into hi / lo
    uint64_t *bar_empty = bar_conv + G_STAGES;                                           // MMAs of the stage retired
    uint64_t *bar_acc = bar_empty + G_STAGES;                                            // accumulator complete
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bar_acc + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m0 = blockIdx.x * GM, n0 = blockIdx.y * GN;
    const int kb0 = (int)((long long)P.kb_total * blockIdx.z / P.splits);
    const int kb1 = (int)((long long)P.kb_total * (blockIdx.z + 1) / P.splits);
    const int n_kb = kb1 - kb0;

    if (warp == 0 && lane == 0) {
        prefetch_tensormap(&map_x);
        prefetch_tensormap(&map_wh);
        prefetch_tensormap(&map_wl);
        for (int s = 0; s < G_STAGES; ++s) {
            mbar_init(&bar_full[s], 1);
            mbar_init(&bar_conv[s], G_CONV_THREADS / 32);
            mbar_init(&bar_empty[s], 1);
        }
        mbar_init(bar_acc, 1);
        mbar_fence_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, 2 * GN);  // columns [0,GN): x_hi.w_hi, [GN,2GN): the two correction terms
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_acc = *tmem_slot;

    if (warp == 0) {
        // ===== TMA producer: x tile [GM x GK], w_hi / w_lo tiles [GN x GK]; out-of-range rows / columns arrive as zeros
        if (lane == 0) {
            for (int i = 0; i < n_kb; ++i) {
                const int s = i % G_STAGES;
                const uint32_t ph = (uint32_t)(i / G_STAGES) & 1u;
                mbar_wait(&bar_empty[s], ph ^ 1u);
                mbar_arrive_expect_tx(&bar_full[s], (uint32_t)(G_A_BYTES + 2 * G_B_BYTES));
                const int kc = (kb0 + i) * GK;
                tma_load_2d(s_xh(s), &map_x, kc, m0, &bar_full[s]);
                tma_load_2d(s_wh(s), &map_wh, kc, n0, &bar_full[s]);
                tma_load_2d(s_wl(s), &map_wl, kc, n0, &bar_full[s]);
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ===== MMA issuer: per K step of 8, acc += x_hi.w_hi + x_lo.w_hi + x_hi.w_lo
        if (lane == 0) {
            constexpr uint32_t idesc = umma_idesc_tf32(GM, GN);
            for (int i = 0; i < n_kb; ++i) {
                const int s = i % G_STAGES;
                const uint32_t ph = (uint32_t)(i / G_STAGES) & 1u;
                mbar_wait(&bar_full[s], ph);
                mbar_wait(&bar_conv[s], ph);
                tc_fence_after();
                const uint64_t d_xh = umma_desc_sw128(s_xh(s)), d_xl = umma_desc_sw128(s_xl(s));
                const uint64_t d_wh = umma_desc_sw128(s_wh(s)), d_wl = umma_desc_sw128(s_wl(s));
#pragma unroll
                for (int k = 0; k < GK / UK; ++k) {
                    const uint64_t adv = (uint64_t)((k * UK * 4) >> 4);  // 32 bytes along K inside the swizzle atom
                    // The tensor core rounds each accumulation towards zero; keeping the small correction terms in
                    // their own accumulator shortens the rounding chain of the dominant term threefold.
                    const uint32_t first = (i > 0 || k > 0) ? 1u : 0u;
                    umma_tf32(tmem_acc, d_xh + adv, d_wh + adv, idesc, first);
                    umma_tf32(tmem_acc + GN, d_xl + adv, d_wh + adv, idesc, first);
                    umma_tf32(tmem_acc + GN, d_xh + adv, d_wl + adv, idesc, 1u);
                }
                umma_commit(&bar_empty[s]);  // frees the stage for the producer once these MMAs have read it
            }
            umma_commit(bar_acc);
        }
        __syncwarp();
    } else {
        // ===== converter warps: split the x tile TMA delivered into hi (in place) and lo, element-wise -- the
        // swizzled placement is untouched, so the two tiles share the descriptor geometry
        const int ct = threadIdx.x - 64;
        for (int i = 0; i < n_kb; ++i) {
            const int s = i % G_STAGES;
            const uint32_t ph = (uint32_t)(i / G_STAGES) & 1u;
            mbar_wait(&bar_full[s], ph);
            float4 *xh = reinterpret_cast<float4 *>(s_xh(s));
            float4 *xl = reinterpret_cast<float4 *>(s_xl(s));
#pragma unroll
            for (int j = 0; j < G_A_BYTES / 16 / G_CONV_THREADS; ++j) {
                const int q = j * G_CONV_THREADS + ct;
                const float4 v = xh[q];
                float4 h, l;
                h.x = tf32_hi(v.x);
                h.y = tf32_hi(v.y);
                h.z = tf32_hi(v.z);
                h.w = tf32_hi(v.w);
                l.x = v.x - h.x;
                l.y = v.y - h.y;
                l.z = v.z - h.z;
                l.w = v.w - h.w;
                xh[q] = h;
                xl[q] = l;
            }
            fence_proxy_async_smem();  // generic-proxy writes -> visible to the tensor core's async-proxy reads
            __syncwarp();
            if (lane == 0) mbar_arrive(&bar_conv[s]);
        }
        // ===== epilogue: TMEM -> registers -> global partial sums. Warp w may touch TMEM lanes [32 (w%4), +32).
        mbar_wait(bar_acc, 0);
        tc_fence_after();
        const int q4 = warp & 3;
        const int row = m0 + q4 * 32 + lane;
        // partial sums are stored in blocks of 32 samples, [split][m / 32][n][m % 32]: the 32 rows x 32 columns a warp
        // holds are 4 KB of contiguous memory (coalesced, no power-of-two stride), and one CTA of the tail kernel
        // (32 samples) reads one contiguous [N][32] block per split.
        float *ocol = P.out + ((((long long)blockIdx.z * P.MB + (m0 >> 5) + q4) * P.N + n0) << 5) + lane;
#pragma unroll 1
        for (int c = 0; c < GN / 32; ++c) {
            if (n0 + c * 32 >= P.N) break;  // warp-uniform
            uint32_t r[32], r2[32];
            tmem_ld_32x32(tmem_acc + ((uint32_t)(q4 * 32) << 16) + (uint32_t)(c * 32), r);
            tmem_ld_32x32(tmem_acc + ((uint32_t)(q4 * 32) << 16) + (uint32_t)(GN + c * 32), r2);
            tmem_ld_wait();
            if (P.dense != nullptr) {
                // row-major output: transpose the warp's 32 x 32 block through shared memory (the operand stages are idle
                // now) so that every store instruction writes 128 contiguous bytes of one output row
                float *tile = reinterpret_cast<float *>(smem) + q4 * (32 * 33);
#pragma unroll
                for (int j = 0; j < 32; ++j) tile[lane * 33 + j] = __uint_as_float(r[j]) + __uint_as_float(r2[j]);
                __syncwarp();
                const int col = n0 + c * 32 + lane;
                const float bv = (P.bias != nullptr && col < P.N) ? __ldg(P.bias + col) : 0.f;
                const int row0 = m0 + q4 * 32;
#pragma unroll 8
                for (int rr = 0; rr < 32; ++rr)
                    if (row0 + rr < P.M && col < P.N) P.dense[(long long)(row0 + rr) * P.N + col] = tile[rr * 33 + lane] + bv;
                __syncwarp();
            } else if (row < P.M) {
#pragma unroll
                for (int j = 0; j < 32; ++j)
                    if (n0 + c * 32 + j < P.N)
                        ocol[(c * 32 + j) << 5] = __uint_as_float(r[j]) + __uint_as_float(r2[j]);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_acc, 2 * GN);
}

// ------------------------------------------------------------------ the same GEMM on a CTA PAIR (cta_group::2)
// Two CTAs of a cluster (the two SMs of a TPC) share one 256 x 256 output tile: each holds 128 rows of x and HALF of
// the W tile (128 of the 256 output features); tcgen05.mma.cta_group::2 (issued by the leader CTA only) multiplies
// the pair's 256 rows with all 256 features, each SM's tensor core reading only its own shared memory.  Per SM and
// K step that is 4 KB (x) + 4 KB (W half) of operand reads instead of 4 + 8, 48 KB instead of 80 KB of TMA traffic per
// stage, and 64 KB stages -> a 3-deep ring.
//   barriers, per stage:  bar_x    (own CTA)   own x tile landed             -> own converter warps
//                         bar_w    (leader)    both CTAs' W halves landed    -> MMA issuer   (TMA ...cta_group::2 signals
//                                                                               the leader's barrier from either CTA)
//                         bar_conv (leader)    8 converter warps (2 CTAs) done, rank 1 arrives through the cluster
//                         bar_empty(each CTA)  tcgen05.commit ...multicast::cluster from the leader frees the stage in both
constexpr int G2_STAGES = 3;
constexpr int G2_BH_BYTES = (GN / 2) * GK * 4;                        // 16 KB: this CTA's half of a W tile
constexpr int G2_STAGE_BYTES = 2 * G_A_BYTES + 2 * G2_BH_BYTES;       // 64 KB
constexpr int G2_SMEM_BYTES = G2_STAGES * G2_STAGE_BYTES + 1024 + 256;

__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive on the barrier at the same shared-memory offset in CTA 0 of the cluster
__device__ __forceinline__ void mbar_arrive_leader(uint64_t *bar) {
    asm volatile(
        "{\n"
        ".reg .b32 ra;\n"
        "mapa.shared::cluster.u32 ra, %0, 0;\n"
        "mbarrier.arrive.release.cluster.shared::cluster.b64 _, [ra];\n"
        "}\n" ::"r"(smem_u32(bar))
        : "memory");
}
// TMA tile load whose completion bytes are counted on the LEADER CTA's barrier (bit 24 of a shared::cluster address is
// the CTA rank inside the pair)
__device__ __forceinline__ void tma_load_2d_pair(void *smem_dst, const CUtensorMap *map, int c0, int c1, uint64_t *bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], "
        "[%2];" ::"r"(smem_u32(smem_dst)),
        "l"(map), "r"(smem_u32(bar) & 0xFEFFFFFFu), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tmem_alloc_pair(uint32_t *smem_slot, uint32_t cols) {  // one full warp in EACH CTA
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_slot)),
                 "r"(cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t addr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}
__device__ __forceinline__ void umma_tf32_pair(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                               uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit_pair(uint64_t *bar) {  // arrives on `bar` in BOTH CTAs
    asm volatile(
        "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
            smem_u32(bar)),
        "h"((uint16_t)3)
        : "memory");
}

// grid = (2 * ceil(M/256), ceil(N/GN), splits), clusters of 2 along x, 192 threads per CTA.
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(G_THREADS, 1)
    mlp_gemm_tf32x3_pair_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_wh,
                                const __grid_constant__ CUtensorMap map_wl, const GemmParams P) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char *smem = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    auto s_xh = [&](int s) { return smem + s * G2_STAGE_BYTES; };
    auto s_xl = [&](int s) { return smem + s * G2_STAGE_BYTES + G_A_BYTES; };
    auto s_wh = [&](int s) { return smem + s * G2_STAGE_BYTES + 2 * G_A_BYTES; };
    auto s_wl = [&](int s) { return smem + s * G2_STAGE_BYTES + 2 * G_A_BYTES + G2_BH_BYTES; };
    uint64_t *bar_x = reinterpret_cast<uint64_t *>(smem + G2_STAGES * G2_STAGE_BYTES);
    uint64_t *bar_w = bar_x + G2_STAGES;
    uint64_t *bar_conv = bar_w + G2_STAGES;
    uint64_t *bar_empty = bar_conv + G2_STAGES;
    uint64_t *bar_acc = bar_empty + G2_STAGES;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bar_acc + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    const int m0 = blockIdx.x * GM;                       // blockIdx.x = 2 * pair + rank
    const int n0 = blockIdx.y * GN, nh = n0 + (int)rank * (GN / 2);
    const int kb0 = (int)((long long)P.kb_total * blockIdx.z / P.splits);
    const int kb1 = (int)((long long)P.kb_total * (blockIdx.z + 1) / P.splits);
    const int n_kb = kb1 - kb0;

    if (warp == 0 && lane == 0) {
        prefetch_tensormap(&map_x);
        prefetch_tensormap(&map_wh);
        prefetch_tensormap(&map_wl);
        for (int s = 0; s < G2_STAGES; ++s) {
            mbar_init(&bar_x[s], 1);
            mbar_init(&bar_w[s], 1);
            mbar_init(&bar_conv[s], 2 * (G_CONV_THREADS / 32));
            mbar_init(&bar_empty[s], 1);
        }
        mbar_init(bar_acc, 1);
        mbar_fence_init();
    }
    if (warp == 1) tmem_alloc_pair(tmem_slot, 2 * GN);
    tc_fence_before();
    cluster_sync_all();  // barriers of both CTAs are initialised before anyone signals across the pair
    __syncthreads();     // (redundant after the cluster barrier; compute-sanitizer's racecheck only models CTA barriers)
    tc_fence_after();
    const uint32_t tmem_acc = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            for (int i = 0; i < n_kb; ++i) {
                const int s = i % G2_STAGES;
                const uint32_t ph = (uint32_t)(i / G2_STAGES) & 1u;
                mbar_wait(&bar_empty[s], ph ^ 1u);
                const int kc = (kb0 + i) * GK;
                mbar_arrive_expect_tx(&bar_x[s], (uint32_t)G_A_BYTES);
                tma_load_2d(s_xh(s), &map_x, kc, m0, &bar_x[s]);
                if (leader) mbar_arrive_expect_tx(&bar_w[s], (uint32_t)(4 * G2_BH_BYTES));  // both halves, both CTAs
                tma_load_2d_pair(s_wh(s), &map_wh, kc, nh, &bar_w[s]);
                tma_load_2d_pair(s_wl(s), &map_wl, kc, nh, &bar_w[s]);
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        if (leader && lane == 0) {
            constexpr uint32_t idesc = umma_idesc_tf32(2 * GM, GN);
            for (int i = 0; i < n_kb; ++i) {
                const int s = i % G2_STAGES;
                const uint32_t ph = (uint32_t)(i / G2_STAGES) & 1u;
                mbar_wait(&bar_w[s], ph);
                mbar_wait(&bar_conv[s], ph);
                tc_fence_after();
                const uint64_t d_xh = umma_desc_sw128(s_xh(s)), d_xl = umma_desc_sw128(s_xl(s));
                const uint64_t d_wh = umma_desc_sw128(s_wh(s)), d_wl = umma_desc_sw128(s_wl(s));
#pragma unroll
                for (int k = 0; k < GK / UK; ++k) {
                    const uint64_t adv = (uint64_t)((k * UK * 4) >> 4);
                    const uint32_t first = (i > 0 || k > 0) ? 1u : 0u;
                    umma_tf32_pair(tmem_acc, d_xh + adv, d_wh + adv, idesc, first);
                    umma_tf32_pair(tmem_acc + GN, d_xl + adv, d_wh + adv, idesc, first);
                    umma_tf32_pair(tmem_acc + GN, d_xh + adv, d_wl + adv, idesc, 1u);
                }
                umma_commit_pair(&bar_empty[s]);
            }
            umma_commit_pair(bar_acc);
        }
        __syncwarp();
    } else {
        const int ct = threadIdx.x - 64;
        for (int i = 0; i < n_kb; ++i) {
            const int s = i % G2_STAGES;
            const uint32_t ph = (uint32_t)(i / G2_STAGES) & 1u;
            mbar_wait(&bar_x[s], ph);
            float4 *xh = reinterpret_cast<float4 *>(s_xh(s));
            float4 *xl = reinterpret_cast<float4 *>(s_xl(s));
#pragma unroll
            for (int j = 0; j < G_A_BYTES / 16 / G_CONV_THREADS; ++j) {
                const int q = j * G_CONV_THREADS + ct;
                const float4 v = xh[q];
                float4 h, l;
                h.x = tf32_hi(v.x);
                h.y = tf32_hi(v.y);
                h.z = tf32_hi(v.z);
                h.w = tf32_hi(v.w);
                l.x = v.x - h.x;
                l.y = v.y - h.y;
                l.z = v.z - h.z;
                l.w = v.w - h.w;
                xh[q] = h;
                xl[q] = l;
            }
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive_leader(&bar_conv[s]);
        }
        mbar_wait(bar_acc, 0);
        tc_fence_after();
        const int q4 = warp & 3;
        const int row = m0 + q4 * 32 + lane;
        float *ocol = P.out + ((((long long)blockIdx.z * P.MB + (m0 >> 5) + q4) * P.N + n0) << 5) + lane;
#pragma unroll 1
        for (int c = 0; c < GN / 32; ++c) {
            if (n0 + c * 32 >= P.N) break;  // warp-uniform
            uint32_t r[32], r2[32];
            tmem_ld_32x32(tmem_acc + ((uint32_t)(q4 * 32) << 16) + (uint32_t)(c * 32), r);
            tmem_ld_32x32(tmem_acc + ((uint32_t)(q4 * 32) << 16) + (uint32_t)(GN + c * 32), r2);
            tmem_ld_wait();
            if (P.dense != nullptr) {
                // row-major output: transpose the warp's 32 x 32 block through shared memory (the operand stages are idle
                // now) so that every store instruction writes 128 contiguous bytes of one output row
                float *tile = reinterpret_cast<float *>(smem) + q4 * (32 * 33);
#pragma unroll
                for (int j = 0; j < 32; ++j) tile[lane * 33 + j] = __uint_as_float(r[j]) + __uint_as_float(r2[j]);
                __syncwarp();
                const int col = n0 + c * 32 + lane;
                const float bv = (P.bias != nullptr && col < P.N) ? __ldg(P.bias + col) : 0.f;
                const int row0 = m0 + q4 * 32;
#pragma unroll 8
                for (int rr = 0; rr < 32; ++rr)
                    if (row0 + rr < P.M && col < P.N) P.dense[(long long)(row0 + rr) * P.N + col] = tile[rr * 33 + lane] + bv;
                __syncwarp();
            } else if (row < P.M) {
#pragma unroll
                for (int j = 0; j < 32; ++j)
                    if (n0 + c * 32 + j < P.N)
                        ocol[(c * 32 + j) << 5] = __uint_as_float(r[j]) + __uint_as_float(r2[j]);
            }
        }
    }
    tc_fence_before();
    cluster_sync_all();  // nobody leaves (or frees TMEM) while the peer can still signal / multicast into this CTA
    if (warp == 1) tmem_dealloc_pair(tmem_acc, 2 * GN);
}

// ------------------------------------------------------------------ everything after the first Linear
// packed parameters (floats): a1[H] c1[H] | per further hidden layer: Wt[H_in=H][H] a[H] c[H] | Wf[NO][H] bf[NO]
//   h1 = relu(sum_s partial[s] * a1 + c1)          a = bn.weight / sqrt(running_var + eps)  (layers.py:75, eval)
//   h_{l+1} = relu((h_l . W_l^T) * a_l + c_l)       c = (linear.bias - running_mean) * a + bn.bias
//   y = h_last . Wf^T + bf                                                                      (layers.py:81)
constexpr int T_COLS = 256;      // output features per pass
constexpr int T_THREADS = 256;   // 8 warps: 2 halves of the reduction axis x 4 column blocks of 64
constexpr int TS = 32;           // samples per CTA = one block of the partial-sum layout
constexpr int TP = TS + 4;       // pitch (floats) of an activation row hT[feature][sample]: conflict-free float4 access
constexpr int T_CH = 32;         // rows of W^T staged per chunk ([T_CH][T_COLS] floats = 32 KB), ring of n_buf chunks
constexpr int T_MAXH = 512;
constexpr int T_RED_WARPS = 8;

struct TailParams {
    const float *partials;  // [splits][MB][H][32]  (what the GEMM kernel writes)
    const float *packed;
    float *y;               // [B][NO]
    long long B;
    int splits, H, n_rest, NO;
    int n_buf_log2;         // log2 of the number of W^T chunk buffers (4 or 2, as shared memory allows)
};

// One CTA = 32 samples through every layer after the first GEMM.  Activations live in shared memory feature-major
// (hT[k][s]).  A hidden layer is a register-tiled SGEMM: a warp covers 32 samples x 64 output features with 8 x 8
// accumulators per thread (packed FFMA2), so one k step costs 4 LDS.128 per 32 FFMA2 -- shared-memory return bandwidth
// (512 B per LDS.128, broadcast or not) is the limiter of a CUDA-core GEMM; W^T chunks arrive by TMA bulk copies (one
// 1 KB row segment per lane of warp 0) into a ring of buffers, and the two halves of each chunk's rows go to warps
// 0-3 / 4-7, whose partial sums meet in shared memory at the end of the layer.
__global__ void __launch_bounds__(T_THREADS, 1) mlp_tail_kernel(const TailParams P) {
    extern __shared__ __align__(128) unsigned char tsm_raw[];
    uint64_t *bar = reinterpret_cast<uint64_t *>(tsm_raw);
    float *wbuf = reinterpret_cast<float *>(tsm_raw + 128);
    const int H = P.H;
    const int NBL = P.n_buf_log2, NB = 1 << NBL;
    float *hA = wbuf + NB * T_CH * T_COLS;
    float *hB = hA + H * TP;
    float *red = hB + H * TP;  // [T_RED_WARPS][NO][TS]
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const long long b0 = (long long)blockIdx.x * TS;
    const int ns = (int)min((long long)TS, P.B - b0);
    const long long MB = (P.B + TS - 1) / TS;
    const float *pk = P.packed;
    if (tid == 0) {
        for (int i = 0; i < NB; ++i) mbar_init(&bar[i], 1);
        mbar_fence_init();
    }
    // ---- first layer's epilogue: reduce the split-K partial sums, bias + BatchNorm (folded), ReLU.
    // float4 t of this CTA's [H][32] block = feature t / 8, samples 4 (t % 8) .. +3: coalesced reads, conflict-free writes.
    const int n_f4 = H * (TS / 4);
    for (int t0 = tid; t0 < n_f4; t0 += 4 * T_THREADS) {
        float4 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int sp = 0; sp < P.splits; ++sp) {
            const float4 *src = reinterpret_cast<const float4 *>(P.partials + (((long long)sp * MB + blockIdx.x) * H << 5));
            float4 w[4];
#pragma unroll
            for (int u = 0; u < 4; ++u)
                w[u] = (t0 + u * T_THREADS < n_f4) ? __ldg(src + t0 + u * T_THREADS) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                v[u].x += w[u].x;
                v[u].y += w[u].y;
                v[u].z += w[u].z;
                v[u].w += w[u].w;
            }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int t = t0 + u * T_THREADS;
            if (t < n_f4) {
                const int n = t >> 3, s4 = (t & 7) * 4;
                const float a1 = pk[n], c1 = pk[H + n];
                float4 o;
                o.x = (s4 + 0 < ns) ? fmaxf(fmaf(v[u].x, a1, c1), 0.f) : 0.f;
                o.y = (s4 + 1 < ns) ? fmaxf(fmaf(v[u].y, a1, c1), 0.f) : 0.f;
                o.z = (s4 + 2 < ns) ? fmaxf(fmaf(v[u].z, a1, c1), 0.f) : 0.f;
                o.w = (s4 + 3 < ns) ? fmaxf(fmaf(v[u].w, a1, c1), 0.f) : 0.f;
                *reinterpret_cast<float4 *>(hA + n * TP + s4) = o;
            }
        }
    }
    pk += 2 * H;
    __syncthreads();
    // ---- further hidden layers
    uint32_t it = 0;  // chunks consumed so far: buffer it % NB, mbarrier parity (it / NB) & 1 (NB a power of two)
    const int n_chunks = (H + T_CH - 1) / T_CH;
    for (int l = 0; l < P.n_rest; ++l) {
        const float *Wt = pk;  // [H][H], input feature major
        const float *a = pk + (long long)H * H, *c = a + H;
        for (int g0 = 0; g0 < H; g0 += T_COLS) {
            const int ncols = min(T_COLS, H - g0);
            auto issue = [&](int cidx, int buf) {  // whole warp 0
                const int rows = min(T_CH, H - cidx * T_CH);
                if (lane == 0) mbar_arrive_expect_tx(&bar[buf], (uint32_t)(rows * ncols * 4));
                __syncwarp();
                if (lane < rows)
                    tma_load_bulk(wbuf + (buf * T_CH + lane) * T_COLS, Wt + (long long)(cidx * T_CH + lane) * H + g0,
                                  (uint32_t)(ncols * 4), &bar[buf]);
            };
            if (warp == 0) {
                for (int i = 0; i < NB && i < n_chunks; ++i) issue(i, (int)((it + i) & (NB - 1)));
            }
            // thread tile: samples {4 sg .. +3, 16 + 4 sg .. +3} x features {4 cg .. +3, 32 + 4 cg .. +3} of the warp's 64
            const int kh = warp >> 2, wc = (warp & 3) * 64, sg = lane >> 3, cg = lane & 7;
            float2 acc[4][8];  // [sample pair][feature]
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = make_float2(0.f, 0.f);
            for (int ci = 0; ci < n_chunks; ++ci, ++it) {
                const int buf = (int)(it & (NB - 1));
                mbar_wait(&bar[buf], (it >> NBL) & 1u);
                const int rows = min(T_CH, H - ci * T_CH);
                const int r0 = kh ? rows / 2 : 0, r1 = kh ? rows : rows / 2;
                const float *wb = wbuf + buf * T_CH * T_COLS + wc + 4 * cg;
                const float *hrow = hA + (ci * T_CH) * TP + 4 * sg;
#pragma unroll 4
                for (int kk = r0; kk < r1; ++kk) {
                    const float4 w0 = *reinterpret_cast<const float4 *>(wb + kk * T_COLS);
                    const float4 w1 = *reinterpret_cast<const float4 *>(wb + kk * T_COLS + 32);
                    const float4 h0 = *reinterpret_cast<const float4 *>(hrow + kk * TP);
                    const float4 h1 = *reinterpret_cast<const float4 *>(hrow + kk * TP + 16);
                    const float2 hp[4] = {make_float2(h0.x, h0.y), make_float2(h0.z, h0.w), make_float2(h1.x, h1.y),
                                          make_float2(h1.z, h1.w)};
                    const float wv[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const float2 w2 = splat2(wv[j]);
#pragma unroll
                        for (int i = 0; i < 4; ++i) acc[i][j] = ffma2(hp[i], w2, acc[i][j]);
                    }
                }
                __syncthreads();  // everyone is done with this buffer
                if (warp == 0 && ci + NB < n_chunks) issue(ci + NB, buf);
            }
            // the upper half of the reduction axis hands its sums to the lower half through the (now idle) chunk ring
            float2 *xch = reinterpret_cast<float2 *>(wbuf);
            if (kh == 1) {
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 8; ++j) xch[(i * 8 + j) * 128 + (tid - 128)] = acc[i][j];
            }
            __syncthreads();
            if (kh == 0) {
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int cl = wc + (j < 4 ? 4 * cg + j : 32 + 4 * cg + (j - 4));  // feature inside this pass
                    if (cl < ncols) {
                        const int n = g0 + cl;
                        const float an = a[n], cn = c[n];
                        float o[8];
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const float2 t = xch[(i * 8 + j) * 128 + tid];
                            o[2 * i] = fmaxf(fmaf(acc[i][j].x + t.x, an, cn), 0.f);
                            o[2 * i + 1] = fmaxf(fmaf(acc[i][j].y + t.y, an, cn), 0.f);
                        }
                        *reinterpret_cast<float4 *>(hB + n * TP + 4 * sg) = make_float4(o[0], o[1], o[2], o[3]);
                        *reinterpret_cast<float4 *>(hB + n * TP + 16 + 4 * sg) = make_float4(o[4], o[5], o[6], o[7]);
                    }
                }
            }
            fence_proxy_async_smem();  // generic accesses to the exchange area before the next TMA writes into it
            __syncthreads();           // the exchange area is a chunk buffer again
        }
        pk += (long long)H * H + 2 * H;
        __syncthreads();
        float *t = hA;
        hA = hB;
        hB = t;
    }
    // ---- output Linear (layers.py:81): lane = sample, 8 warps split the feature axis, then a shared-memory sum
    const float *Wf = pk, *bf = pk + (long long)P.NO * H;
    if (warp < T_RED_WARPS) {
        const int kw0 = (int)((long long)H * warp / T_RED_WARPS), kw1 = (int)((long long)H * (warp + 1) / T_RED_WARPS);
        for (int o = 0; o < P.NO; ++o) {
            float accs = 0.f;
            for (int k = kw0; k < kw1; ++k) accs = fmaf(hA[k * TP + lane], __ldg(Wf + (long long)o * H + k), accs);
            red[(warp * P.NO + o) * TS + lane] = accs;
        }
    }
    __syncthreads();
    for (int t = tid; t < ns * P.NO; t += T_THREADS) {
        const int sidx = t / P.NO, o = t - sidx * P.NO;
        float v = bf[o];
#pragma unroll
        for (int w = 0; w < T_RED_WARPS; ++w) v += red[(w * P.NO + o) * TS + sidx];
        P.y[(b0 + sidx) * P.NO + o] = v;
    }
}

// ------------------------------------------------------------------ a hidden layer after the first, on the tensor cores
// One launch = one hidden Linear(H_in -> H_out <= 256) of layers.py:73-77 in eval mode, for 128 samples per CTA:
//   x[m][k]  = relu(sum_s in[s][m/32][k][m%32] * a_in[k] + c_in[k])      the previous layer's activation, built by the
//                                                                         producer warps straight into the swizzled A tile
//   acc[m][n] = sum_k x[m][k] w[n][k]                                     tcgen05.mma kind::tf32, 3xTF32 like the first GEMM
//   final:     y[m][o] = sum_n relu(acc[m][n] a_out[n] + c_out[n]) wf[o][n] + bf[o]      (this layer's BN/ReLU + layers.py:81)
//   otherwise: out[0][m/32][n][m%32] = acc[m][n]                          (the next launch applies a_out / c_out)
// so "split-K sums of layer 1 -> y" is ONE kernel on ceil(B/128) SMs (10 us at B = 4096) instead of the CUDA-core tail
// kernel on B/32 SMs (27 us).  Roles as in mlp_gemm_tf32x3_kernel: warp 0 = TMA producer of the W tiles, warp 1 = TMEM owner
// + MMA issuer, warps 2-5 = producers of the A tile (thread = sample row: coalesced reads along the sample axis of the
// partial-sum layout) and epilogue (thread = TMEM lane = sample).
constexpr int HT_MAX_NO = 4;
struct HiddenParams {
    const float *in;        // [in_splits][MB][H_in][32]
    const float *a_in, *c_in;
    const float *a_out, *c_out, *wf, *bf;   // final mode (y != null)
    float *y;               // [B][NO]
    float *out;             // [1][MB][H_out][32] when y == null
    long long B;
    int MB, in_splits, H_in, H_out, NO, kb_total;
};

constexpr int HT_THREADS = 64 + 2 * G_CONV_THREADS;   // TMA warp, MMA warp, 8 producer / epilogue warps
__global__ void __launch_bounds__(HT_THREADS, 1)
    mlp_hidden_tc_kernel(const __grid_constant__ CUtensorMap map_wh, const __grid_constant__ CUtensorMap map_wl,
                         const HiddenParams P) {
    extern __shared__ __align__(1024) unsigned char smem_hid[];   // SWIZZLE_128B tiles need 1024-byte alignment
    unsigned char *smem = smem_hid;
    auto s_xh = [&](int s) { return smem + s * G_STAGE_BYTES; };
    auto s_xl = [&](int s) { return smem + s * G_STAGE_BYTES + G_A_BYTES; };
    auto s_wh = [&](int s) { return smem + s * G_STAGE_BYTES + 2 * G_A_BYTES; };
    auto s_wl = [&](int s) { return smem + s * G_STAGE_BYTES + 2 * G_A_BYTES + G_B_BYTES; };
    uint64_t *bar_w = reinterpret_cast<uint64_t *>(smem + G_STAGES * G_STAGE_BYTES);   // W tiles landed (TMA)
    uint64_t *bar_a = bar_w + G_STAGES;                                                 // A tile produced (4 warps)
    uint64_t *bar_empty = bar_a + G_STAGES;                                             // MMAs of the stage retired
    uint64_t *bar_acc = bar_empty + G_STAGES;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bar_acc + 1);
    float *prm = reinterpret_cast<float *>(smem + G_STAGES * G_STAGE_BYTES + 256);      // a_in c_in [H_in] | a_out c_out [H_out] | wf [NO][H_out]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m0 = blockIdx.x * GM;
    const int n_kb = P.kb_total;
    const int Hi = P.H_in, Ho = P.H_out;
    float *p_ain = prm, *p_cin = prm + Hi, *p_aout = prm + 2 * Hi, *p_cout = p_aout + Ho, *p_wf = p_cout + Ho;

    if (warp == 0 && lane == 0) {
        prefetch_tensormap(&map_wh);
        prefetch_tensormap(&map_wl);
        for (int s = 0; s < G_STAGES; ++s) {
            mbar_init(&bar_w[s], 1);
            mbar_init(&bar_a[s], 2 * G_CONV_THREADS / 32);
            mbar_init(&bar_empty[s], 1);
        }
        mbar_init(bar_acc, 1);
        mbar_fence_init();
    }
    for (int i = threadIdx.x; i < Hi; i += HT_THREADS) {
        p_ain[i] = P.a_in[i];
        p_cin[i] = P.c_in[i];
    }
    if (P.y != nullptr) {
        for (int i = threadIdx.x; i < Ho; i += HT_THREADS) {
            p_aout[i] = P.a_out[i];
            p_cout[i] = P.c_out[i];
        }
        for (int i = threadIdx.x; i < P.NO * Ho; i += HT_THREADS) p_wf[i] = P.wf[i];
    }
    if (warp == 1) tmem_alloc(tmem_slot, 2 * GN);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_acc = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            for (int i = 0; i < n_kb; ++i) {
                const int s = i % G_STAGES;
                const uint32_t ph = (uint32_t)(i / G_STAGES) & 1u;
                mbar_wait(&bar_empty[s], ph ^ 1u);
                mbar_arrive_expect_tx(&bar_w[s], (uint32_t)(2 * G_B_BYTES));
                tma_load_2d(s_wh(s), &map_wh, i * GK, 0, &bar_w[s]);
                tma_load_2d(s_wl(s), &map_wl, i * GK, 0, &bar_w[s]);
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = umma_idesc_tf32(GM, GN);
            for (int i = 0; i < n_kb; ++i) {
                const int s = i % G_STAGES;
                const uint32_t ph = (uint32_t)(i / G_STAGES) & 1u;
                mbar_wait(&bar_w[s], ph);
                mbar_wait(&bar_a[s], ph);
                tc_fence_after();
                const uint64_t d_xh = umma_desc_sw128(s_xh(s)), d_xl = umma_desc_sw128(s_xl(s));
                const uint64_t d_wh = umma_desc_sw128(s_wh(s)), d_wl = umma_desc_sw128(s_wl(s));
#pragma unroll
                for (int k = 0; k < GK / UK; ++k) {
                    const uint64_t adv = (uint64_t)((k * UK * 4) >> 4);
                    const uint32_t first = (i > 0 || k > 0) ? 1u : 0u;
                    umma_tf32(tmem_acc, d_xh + adv, d_wh + adv, idesc, first);
                    umma_tf32(tmem_acc + GN, d_xl + adv, d_wh + adv, idesc, first);
                    umma_tf32(tmem_acc + GN, d_xh + adv, d_wl + adv, idesc, 1u);
                }
                umma_commit(&bar_empty[s]);
            }
            umma_commit(bar_acc);
        }
        __syncwarp();
    } else {
        // ===== A producers (8 warps): thread = (sample row m of the CTA's 128, half of the K block's 32 features).  The
        // loads of K block i + 1 (all splits, 64 registers) are in flight while block i is activated, split and stored.
        const int pt = threadIdx.x - 64;
        const int m = pt & (GM - 1), half = pt >> 7;
        const long long blk = (long long)(m0 >> 5) + (m >> 5);
        const bool row_ok = (long long)m0 + m < P.B;   // rows past the batch are never read (the GEMM does not write them): zeros
        const float *src = P.in + ((blk * Hi) << 5) + lane;
        const long long split_stride = ((long long)P.MB * Hi) << 5;
        constexpr int HK = GK / 2, MAXS = 4;
        // the next K block is in flight (registers) while the current one is split and stored.  (Two blocks in flight, 128
        // more registers, measured no faster: the ring of two 96 KB operand stages sets the pace, not the load latency.)
        float raw[MAXS][HK];
        auto issue = [&](int i) {
#pragma unroll
            for (int sp = 0; sp < MAXS; ++sp)
#pragma unroll
                for (int j = 0; j < HK; ++j) {
                    const int k = i * GK + half * HK + j;
                    raw[sp][j] = (row_ok && sp < P.in_splits && k < Hi) ? __ldg(src + sp * split_stride + ((long long)k << 5)) : 0.f;
                }
        };
        issue(0);
#pragma unroll 1
        for (int i = 0; i < n_kb; ++i) {
            const int s = i % G_STAGES;
            const uint32_t ph = (uint32_t)(i / G_STAGES) & 1u;
            float v[HK];
#pragma unroll
            for (int j = 0; j < HK; ++j) {
                float acc = raw[0][j];
#pragma unroll
                for (int sp = 1; sp < MAXS; ++sp) acc += raw[sp][j];
                const int k = i * GK + half * HK + j;
                v[j] = (row_ok && k < Hi) ? fmaxf(fmaf(acc, p_ain[k], p_cin[k]), 0.f) : 0.f;
            }
            if (i + 1 < n_kb) issue(i + 1);
            mbar_wait(&bar_empty[s], ph ^ 1u);   // the MMAs that read this stage's previous tile have retired
            unsigned char *xh = s_xh(s) + m * 128, *xl = s_xl(s) + m * 128;
#pragma unroll
            for (int cc = 0; cc < HK / 4; ++cc) {
                const int c = half * (HK / 4) + cc;
                float4 h, l;
                h.x = tf32_hi(v[4 * cc + 0]);
                h.y = tf32_hi(v[4 * cc + 1]);
                h.z = tf32_hi(v[4 * cc + 2]);
                h.w = tf32_hi(v[4 * cc + 3]);
                l.x = v[4 * cc + 0] - h.x;
                l.y = v[4 * cc + 1] - h.y;
                l.z = v[4 * cc + 2] - h.z;
                l.w = v[4 * cc + 3] - h.w;
                const int pos = (c ^ (m & 7)) << 4;   // SWIZZLE_128B: 16-byte chunk c of row m sits at c ^ (m % 8)
                *reinterpret_cast<float4 *>(xh + pos) = h;
                *reinterpret_cast<float4 *>(xl + pos) = l;
            }
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bar_a[s]);
        }
        // ===== epilogue: thread = TMEM lane = sample row (warp w reads lanes [32 (w % 4), +32)); the two warps of a lane
        // quadrant take alternate blocks of 32 columns and meet in shared memory for the output Linear
        mbar_wait(bar_acc, 0);
        tc_fence_after();
        const int q4 = warp & 3, cpar = (warp - 2) >> 2;
        const int er = q4 * 32 + lane;            // row of the accumulator tile this thread reads
        const long long b = (long long)m0 + er;
        float yo[HT_MAX_NO];
#pragma unroll
        for (int o = 0; o < HT_MAX_NO; ++o) yo[o] = 0.f;
        float *ocol = P.out ? P.out + (((((long long)(m0 >> 5) + q4) * Ho)) << 5) + lane : nullptr;
#pragma unroll 1
        for (int c = cpar; c < GN / 32; c += 2) {
            if (c * 32 >= Ho) break;
            uint32_t r[32], r2[32];
            tmem_ld_32x32(tmem_acc + ((uint32_t)(q4 * 32) << 16) + (uint32_t)(c * 32), r);
            tmem_ld_32x32(tmem_acc + ((uint32_t)(q4 * 32) << 16) + (uint32_t)(GN + c * 32), r2);
            tmem_ld_wait();
            if (P.y != nullptr) {
                // this layer's BatchNorm + ReLU (layers.py:75-76), then the output Linear's dot products (layers.py:81)
                float h2[32];
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    const int n = c * 32 + j;
                    h2[j] = n < Ho ? fmaxf(fmaf(__uint_as_float(r[j]) + __uint_as_float(r2[j]), p_aout[n], p_cout[n]), 0.f) : 0.f;
                }
#pragma unroll
                for (int o = 0; o < HT_MAX_NO; ++o) {
                    if (o < P.NO) {   // warp-uniform
                        const float *wrow = p_wf + o * Ho + c * 32;
#pragma unroll
                        for (int j = 0; j < 32; ++j)
                            if (c * 32 + j < Ho) yo[o] = fmaf(h2[j], wrow[j], yo[o]);
                    }
                }
            } else if ((long long)(m0 >> 5) + q4 < P.MB) {
#pragma unroll
                for (int j = 0; j < 32; ++j)
                    if (c * 32 + j < Ho) ocol[(c * 32 + j) << 5] = __uint_as_float(r[j]) + __uint_as_float(r2[j]);
            }
        }
        if (P.y != nullptr) {
            // the operand stages are idle now: odd-column warps hand their sums to the even-column warps
            float *xch = reinterpret_cast<float *>(smem);   // [HT_MAX_NO][128]
            if (cpar == 1) {
#pragma unroll
                for (int o = 0; o < HT_MAX_NO; ++o) xch[o * GM + er] = yo[o];
            }
            named_bar_sync(1, 2 * G_CONV_THREADS);
            if (cpar == 0 && b < P.B) {
#pragma unroll
                for (int o = 0; o < HT_MAX_NO; ++o)
                    if (o < P.NO) P.y[b * P.NO + o] = (yo[o] + xch[o * GM + er]) + __ldg(P.bf + o);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_acc, 2 * GN);
}

// ------------------------------------------------------------------ out-of-place transpose (training GEMM operands)
// out[c][r] = in[r][c]; 64 x 64 tiles through shared memory, 16-byte global accesses on both sides when the shapes allow.
// The backward of the first Linear needs x^T (134 MB at config 4); the generic strided copy reaches ~1.8 TB/s.
__global__ void __launch_bounds__(256) transpose_f32_kernel(const float *__restrict__ in, long long rows, long long cols,
                                                            float *__restrict__ out) {
    __shared__ float tile[64][65];
    const long long r0 = (long long)blockIdx.y * 64, c0 = (long long)blockIdx.x * 64;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;  // 16 x 16 threads, each a 4-wide strip, 4 passes
    const bool vec = (rows % 4 == 0) && (cols % 4 == 0);
#pragma unroll
    for (int p = 0; p < 4; ++p) {
        const int lr = ty + 16 * p;
        const long long r = r0 + lr, c = c0 + 4 * tx;
        if (r < rows) {
            if (vec && c + 3 < cols) {
                const float4 v = __ldg(reinterpret_cast<const float4 *>(in + r * cols + c));
                tile[lr][4 * tx + 0] = v.x;
                tile[lr][4 * tx + 1] = v.y;
                tile[lr][4 * tx + 2] = v.z;
                tile[lr][4 * tx + 3] = v.w;
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (c + j < cols) tile[lr][4 * tx + j] = __ldg(in + r * cols + c + j);
            }
        }
    }
    __syncthreads();
#pragma unroll
    for (int p = 0; p < 4; ++p) {
        const int lc = ty + 16 * p;                    // column of the input tile = row of the output
        const long long c = c0 + lc, r = r0 + 4 * tx;  // output row c, output columns r .. r+3
        if (c < cols) {
            if (vec && r + 3 < rows) {
                *reinterpret_cast<float4 *>(out + c * rows + r) =
                    make_float4(tile[4 * tx + 0][lc], tile[4 * tx + 1][lc], tile[4 * tx + 2][lc], tile[4 * tx + 3][lc]);
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (r + j < rows) out[c * rows + r + j] = tile[4 * tx + j][lc];
            }
        }
    }
}

// ------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
    // The driver entry point is resolved through the runtime, so the library has no link-time libcuda dependency.
    static EncodeTiledFn fn = nullptr;
    if (fn == nullptr) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess ||
            qres != cudaDriverEntryPointSuccess)
            return nullptr;
        fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

// 2-D fp32 row-major [rows][cols] (pitch = cols), box [box_rows][GK], 128-byte swizzle, zeros outside the tensor.
static int make_map(CUtensorMap *map, const float *base, long long rows, long long cols, int box_rows) {
    EncodeTiledFn enc = get_encode_fn();
    if (enc == nullptr) {
        set_error("cuTensorMapEncodeTiled is not available from this driver");
        return ARMNET_ERR_CUDA;
    }
    const cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    const cuuint64_t gstride[1] = {(cuuint64_t)cols * 4};
    const cuuint32_t box[2] = {(cuuint32_t)GK, (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float *>(base), gdim, gstride, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed with CUresult %d (rows=%lld cols=%lld)", (int)r, rows, cols);
        return ARMNET_ERR_CUDA;
    }
    return ARMNET_OK;
}

}  // namespace armnet

using namespace armnet;

extern "C" {

int armnet_mlp_split_weight_f32(const float *w, int64_t n, float *w_hi, float *w_lo, void *stream) {
    if (!w || !w_hi || !w_lo) {
        set_error("armnet_mlp_split_weight_f32: null pointer");
        return ARMNET_ERR_NULL;
    }
    if (n <= 0) return ARMNET_OK;
    split_tf32_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(w, w_hi, w_lo, (long long)n);
    ARMNET_CUDA_TRY(cudaGetLastError());
    note_launches(1);
    return ARMNET_OK;
}

int armnet_mlp_linear_splits(int64_t B, int K, int N) {
    if (B <= 0 || K <= 0 || N <= 0) return 0;
    DeviceInfo di;
    if (get_device_info(&di) != ARMNET_OK) return 0;
    const long long tiles = ((B + GM - 1) / GM) * ((N + GN - 1) / GN);
    const int kb_total = (K + GK - 1) / GK;
    // One wave of CTAs.  The tensor core rounds every accumulation towards zero, so the error grows with the length of
    // the accumulation chain (K = 5120: 1.3e-5 unsplit, 3e-6 at 4 splits -- cuBLAS fp32 SGEMM: 3e-6); never run unsplit
    // chains longer than 64 K-blocks.
    long long s = di.sm_count / (tiles > 0 ? tiles : 1);
    if (s < (kb_total + 63) / 64) s = (kb_total + 63) / 64;
    if (s < 1) s = 1;
    if (s > 16) s = 16;
    if (s > kb_total) s = kb_total;
    return (int)s;
}

static int gemm_tf32x3_impl(const float *x, int64_t B, int K, const float *w_hi, const float *w_lo, int N, int splits,
                           float *partials, float *dense, const float *bias, void *stream) {
    if (!x || !w_hi || !w_lo || (!partials && !dense)) {
        set_error("armnet_mlp_linear_tf32x3: null pointer");
        return ARMNET_ERR_NULL;
    }
    if (B <= 0 || K <= 0 || N <= 0 || splits <= 0 || splits > (K + GK - 1) / GK) {
        set_error("armnet_mlp_linear_tf32x3: bad sizes B=%lld K=%d N=%d splits=%d", (long long)B, K, N, splits);
        return ARMNET_ERR_SHAPE;
    }
    if (K % 4 != 0 || ((uintptr_t)x & 15) || ((uintptr_t)w_hi & 15) || ((uintptr_t)w_lo & 15) ||
        ((uintptr_t)partials & 15) || ((uintptr_t)dense & 15)) {
        set_error("armnet_mlp_linear_tf32x3: TMA needs 16-byte aligned rows (K %% 4 == 0) and base pointers");
        return ARMNET_ERR_ALIGN;
    }
    CUtensorMap mx, mh, ml;
    int rc;
    if ((rc = make_map(&mx, x, B, K, GM)) != ARMNET_OK) return rc;
    const int w_box = tuning().gemm_1cta != 0 ? GN : GN / 2;  // the pair kernel loads W in halves
    if ((rc = make_map(&mh, w_hi, N, K, w_box)) != ARMNET_OK) return rc;
    if ((rc = make_map(&ml, w_lo, N, K, w_box)) != ARMNET_OK) return rc;
    GemmParams P;
    P.out = partials;
    P.dense = dense;
    P.bias = bias;
    P.M = (int)B;
    P.MB = (int)((B + 31) / 32);
    P.N = N;
    P.K = K;
    P.kb_total = (K + GK - 1) / GK;
    P.splits = splits;
    static bool attr_set[64];
    int dev = 0;
    ARMNET_CUDA_TRY(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64 || !attr_set[dev]) {
        ARMNET_CUDA_TRY(cudaFuncSetAttribute(mlp_gemm_tf32x3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             G_SMEM_BYTES));
        ARMNET_CUDA_TRY(cudaFuncSetAttribute(mlp_gemm_tf32x3_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             G2_SMEM_BYTES));
        if (dev >= 0 && dev < 64) attr_set[dev] = true;
    }
    if (tuning().gemm_1cta != 0) {  // tuning experiments only: the single-CTA (cta_group::1) kernel
        dim3 grid((unsigned)((B + GM - 1) / GM), (unsigned)((N + GN - 1) / GN), (unsigned)splits);
        mlp_gemm_tf32x3_kernel<<<grid, G_THREADS, G_SMEM_BYTES, (cudaStream_t)stream>>>(mx, mh, ml, P);
    } else {
        dim3 grid((unsigned)(2 * ((B + 2 * GM - 1) / (2 * GM))), (unsigned)((N + GN - 1) / GN), (unsigned)splits);
        mlp_gemm_tf32x3_pair_kernel<<<grid, G_THREADS, G2_SMEM_BYTES, (cudaStream_t)stream>>>(mx, mh, ml, P);
    }
    ARMNET_CUDA_TRY(cudaGetLastError());
    note_launches(1);
    return ARMNET_OK;
}

int armnet_mlp_linear_tf32x3(const float *x, int64_t B, int K, const float *w_hi, const float *w_lo, int N, int splits,
                             float *partials, void *stream) {
    if (!partials) {
        set_error("armnet_mlp_linear_tf32x3: null pointer");
        return ARMNET_ERR_NULL;
    }
    return gemm_tf32x3_impl(x, B, K, w_hi, w_lo, N, splits, partials, nullptr, nullptr, stream);
}

int armnet_linear_tf32x3_dense(const float *x, int64_t B, int K, const float *w_hi, const float *w_lo, int N,
                               const float *bias, float *y, void *stream) {
    if (!y) {
        set_error("armnet_linear_tf32x3_dense: null pointer");
        return ARMNET_ERR_NULL;
    }
    return gemm_tf32x3_impl(x, B, K, w_hi, w_lo, N, 1, nullptr, y, bias, stream);
}

int armnet_mlp_hidden_tc_f32(const float *in_partials, int in_splits, int64_t B, int H_in, const float *a_in,
                             const float *c_in, const float *w_hi, const float *w_lo, int H_out, const float *a_out,
                             const float *c_out, const float *wf, const float *bf, int NO, float *y, float *out_partials,
                             void *stream) {
    if (!in_partials || !a_in || !c_in || !w_hi || !w_lo || (!y && !out_partials) ||
        (y && (!a_out || !c_out || !wf || !bf))) {
        set_error("armnet_mlp_hidden_tc_f32: null pointer");
        return ARMNET_ERR_NULL;
    }
    if (B <= 0 || H_in <= 0 || H_out <= 0 || in_splits <= 0 || (y && NO <= 0)) {
        set_error("armnet_mlp_hidden_tc_f32: bad sizes B=%lld H_in=%d H_out=%d splits=%d NO=%d", (long long)B, H_in, H_out,
                  in_splits, NO);
        return ARMNET_ERR_SHAPE;
    }
    if (H_out > GN || H_in > 1024 || (y && NO > HT_MAX_NO) || in_splits > 4) {
        set_error("armnet_mlp_hidden_tc_f32: H_out <= %d, H_in <= 1024, outputs <= %d, splits <= 4 (got %d, %d, %d, %d)", GN,
                  HT_MAX_NO, H_out, H_in, NO, in_splits);
        return ARMNET_ERR_UNSUPPORTED;
    }
    if (H_in % 4 != 0 || ((uintptr_t)w_hi & 15) || ((uintptr_t)w_lo & 15)) {
        set_error("armnet_mlp_hidden_tc_f32: TMA needs 16-byte aligned weight rows (H_in %% 4 == 0) and base pointers");
        return ARMNET_ERR_ALIGN;
    }
    CUtensorMap mh, ml;
    int rc;
    if ((rc = make_map(&mh, w_hi, H_out, H_in, GN)) != ARMNET_OK) return rc;
    if ((rc = make_map(&ml, w_lo, H_out, H_in, GN)) != ARMNET_OK) return rc;
    HiddenParams P;
    P.in = in_partials;
    P.a_in = a_in;
    P.c_in = c_in;
    P.a_out = a_out;
    P.c_out = c_out;
    P.wf = wf;
    P.bf = bf;
    P.y = y;
    P.out = y ? nullptr : out_partials;
    P.B = B;
    P.MB = (int)((B + 31) / 32);
    P.in_splits = in_splits;
    P.H_in = H_in;
    P.H_out = H_out;
    P.NO = y ? NO : 0;
    P.kb_total = (H_in + GK - 1) / GK;
    const int smem = G_STAGES * G_STAGE_BYTES + 1024 + 256 + (2 * H_in + 2 * H_out + HT_MAX_NO * H_out) * (int)sizeof(float);
    DeviceInfo di;
    if ((rc = get_device_info(&di)) != ARMNET_OK) return rc;
    if (smem > di.smem_optin) {
        set_error("armnet_mlp_hidden_tc_f32: %d bytes of shared memory needed (> %d)", smem, di.smem_optin);
        return ARMNET_ERR_UNSUPPORTED;
    }
    ARMNET_CUDA_TRY(cudaFuncSetAttribute(mlp_hidden_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    mlp_hidden_tc_kernel<<<(unsigned)((B + GM - 1) / GM), HT_THREADS, smem, (cudaStream_t)stream>>>(mh, ml, P);
    ARMNET_CUDA_TRY(cudaGetLastError());
    note_launches(1);
    return ARMNET_OK;
}

size_t armnet_mlp_tail_packed_floats(int H, int n_rest, int NO) {
    if (H <= 0 || n_rest < 0 || NO <= 0) return 0;
    return (size_t)2 * H + (size_t)n_rest * ((size_t)H * H + 2 * (size_t)H) + (size_t)NO * H + NO;
}

int armnet_mlp_tail_f32(const float *partials, int splits, int64_t B, int H, int n_rest, int NO, const float *packed,
                        float *y, void *stream) {
    if (!partials || !packed || !y) {
        set_error("armnet_mlp_tail_f32: null pointer");
        return ARMNET_ERR_NULL;
    }
    if (B <= 0 || H <= 0 || n_rest < 0 || NO <= 0 || splits <= 0) {
        set_error("armnet_mlp_tail_f32: bad sizes B=%lld H=%d n_rest=%d NO=%d splits=%d", (long long)B, H, n_rest, NO,
                  splits);
        return ARMNET_ERR_SHAPE;
    }
    if (H % 4 != 0 || H > T_MAXH || NO > 64 || ((uintptr_t)packed & 15) || ((uintptr_t)partials & 15)) {
        set_error("armnet_mlp_tail_f32: hidden width %d / outputs %d not supported (H %% 4 == 0, H <= %d, NO <= 64)", H,
                  NO, T_MAXH);
        return ARMNET_ERR_UNSUPPORTED;
    }
    DeviceInfo di;
    int rc = get_device_info(&di);
    if (rc != ARMNET_OK) return rc;
    int n_buf = 4;
    size_t smem = 0;
    for (; n_buf >= 2; n_buf -= 2) {
        smem = 128 + ((size_t)n_buf * T_CH * T_COLS + (size_t)2 * H * TP + (size_t)T_RED_WARPS * NO * TS) * sizeof(float);
        if (smem <= (size_t)di.smem_optin) break;
    }
    if (n_buf < 2) {
        set_error("armnet_mlp_tail_f32: hidden width %d needs %zu bytes of shared memory (> %d)", H, smem, di.smem_optin);
        return ARMNET_ERR_UNSUPPORTED;
    }
    ARMNET_CUDA_TRY(cudaFuncSetAttribute(mlp_tail_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    TailParams P;
    P.partials = partials;
    P.packed = packed;
    P.y = y;
    P.B = B;
    P.splits = splits;
    P.H = H;
    P.n_rest = n_rest;
    P.NO = NO;
    P.n_buf_log2 = n_buf == 4 ? 2 : 1;
    mlp_tail_kernel<<<(unsigned)((B + TS - 1) / TS), T_THREADS, smem, (cudaStream_t)stream>>>(P);
    ARMNET_CUDA_TRY(cudaGetLastError());
    note_launches(1);
    return ARMNET_OK;
}

}  // extern "C"

extern "C" int armnet_transpose_f32(const float *in, int64_t rows, int64_t cols, float *out, void *stream) {
    if (!in || !out) {
        set_error("armnet_transpose_f32: null pointer");
        return ARMNET_ERR_NULL;
    }
    if (rows < 0 || cols < 0) {
        set_error("armnet_transpose_f32: bad sizes");
        return ARMNET_ERR_SHAPE;
    }
    if (rows == 0 || cols == 0) return ARMNET_OK;
    if (((uintptr_t)in & 15) || ((uintptr_t)out & 15)) {
        set_error("armnet_transpose_f32: buffers must be 16-byte aligned");
        return ARMNET_ERR_ALIGN;
    }
    dim3 grid((unsigned)((cols + 63) / 64), (unsigned)((rows + 63) / 64));
    if (grid.y > 65535) {
        set_error("armnet_transpose_f32: more than 65535 row tiles");
        return ARMNET_ERR_UNSUPPORTED;
    }
    transpose_f32_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(in, rows, cols, out);
    ARMNET_CUDA_TRY(cudaGetLastError());
    note_launches(1);
    return ARMNET_OK;
}
