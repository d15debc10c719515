// Fused ARM-Net forward with the attention logits on the 5th-generation tensor cores (tcgen05 + TMEM), sm_100a.
//   value clamp -> embedding row gather -> attention logits -> alpha-entmax gates -> gates*values
//   -> log-space cross-feature product -> exp [-> eval-mode arm_bn]   (models/armnet.py:82-89, 26-36)
//
// Same function and the same C ABI entry as armnet_fwd_kernel (fused_fwd.cuh); what changes is who computes the logits
//   X[b, r, f] = sum_x M'[x, r] e[b, f, x]       (armnet.py:33-34 with entmax.py:42's (alpha-1) folded into M'),
// half of the FP32-pipe work of armnet_fwd_kernel.  Here they are a tcgen05.mma.kind::tf32 per (128 neurons, 2 samples):
//   * A = rows of M'^T (M = 128 neurons per MMA), resident in TMEM for the whole kernel (written once with tcgen05.st);
//     B = the gathered embedding rows of a TILE of two samples (N = 2 x 40 rows), in shared memory, K-major, 128-byte
//     swizzle; D (128 lanes x 80 columns fp32) in TMEM.  Four K = 8 steps cover a "packed 3xTF32" K axis of 32:
//         A row = [ M0..M9 | M0..M9 | m0..m9 | 0 0 ],  B row = [ e0..e9 | l0..l9 | e0..e9 | 0 0 ]
//     with m = M - trunc_tf32(M), l = e - trunc_tf32(e): the tensor core reads the top 19 bits of an fp32 operand, so
//     sum_k A_k B_k = sum_x (M e + M l + m e) -- fp32-grade logits (5.6e-7 norm-relative on B200,
//     tools/ubench/tmem_logits_check.cu) from ONE accumulation chain of 4 MMAs (~45 cycles each).  nemb 11..16 use the
//     wide layout (EL = 16): B row = [ e0..e15 | l0..l15 ], A row = [ M | M | m ] in 48 columns, six K = 8 steps of which
//     the last two re-read B floats 0..15 (tm_a_cols / tm_k_steps / tm_b_float below).
//   * TMEM lane i of the D block of A-block kb is neuron 128 kb + i, so a thread reads ITS row's 40 logits with
//     tcgen05.ld (fields packed in (f, f+1) register pairs -- the layout entmax_rows.cuh works on) and the D slot is
//     released as soon as the registers are loaded (four slots: the MMAs of the next items run under the current entmax).
//     NR = 2 (opt-in): a thread owns the same lane of two A blocks, see the kernel template.
//   * Everything after the logits is thread-private like in armnet_fwd_kernel: entmax (entmax_rows.cuh: MUFU-free
//     pre-solve, q-norm Newton, last sweep fused with the cross product), s[x] = sum_f p_f V_f e[f,x] on FFMA2 with the
//     e rows read (warp-broadcast) from the B tile itself -- its first 10 floats are the exact fp32 e --, exp, optional
//     eval-mode arm_bn, one TMA bulk store per warp-unit (32 consecutive neurons x E of one sample).
// Roles: 16 consumer warps (4 per TMEM lane quadrant; units = (item, quadrant, sample) taken in order from one counter
// per quadrant) + three producer warps that only meet through mbarriers: the GATHER warp requests the ids / values of
// tile t + 1 while it issues tile t (clamp in place and range check at use: it never waits for a load inside its loop) and
// issues the TMA bulk row gathers up to 16 tiles ahead into a raw ring; the CONVERT warp turns
// landed rows into B-tile rows (scale by the value: e = T[id] v exactly as layers.py:21; split; swizzled store); the MMA
// warp issues the tcgen05.mma of an item the moment its tile is converted and a D slot is free (a single producer warp
// doing all three in sequence starved the consumers: 28 % issue utilisation, profiles/r2_v3_summary.md).  No CTA-wide
// barrier in steady state; every mbarrier wait is bounded (a protocol bug traps instead of hanging the device).
// Start-up (tools/trace_tmem.py, -DARMNET_TMEM_TRACE): consumer warps 0-3 put M'^T into tensor memory block by block
// (the first MMA needs block 0), the first logits appear ~7 us after CTA entry (ids 1 us, 78 bulk copies 2.3 us to issue,
// rows 0.8 us, convert 0.7 us); the ring indices and the item -> tile index carry no runtime division.
// Requirements (host-checked, everything else runs on armnet_fwd_kernel / armnet_fwd_mma_kernel): F in {2NP-1, 2NP} for a
// compiled NP, E <= 16 (E > 10: opt-in, tuning tmem = 1), K*O a multiple of 128 with (K*O / 128) * (32 | 48) + 320 <= 512
// tensor-memory columns, 16-byte-aligned table rows (the module's padded shadow table), no validation outputs,
// solver != literal bisection.
#pragma once

#include "entmax_rows.cuh"
#include "entmax_stream.cuh"
#include "fused_fwd.cuh"

namespace armnet {

#ifndef ARMNET_TMEM_CONSUMERS
#define ARMNET_TMEM_CONSUMERS 16
#endif
constexpr int kTmConsumers = ARMNET_TMEM_CONSUMERS;   // consumer warps (a multiple of 4: TMEM lane quadrants)
#ifndef ARMNET_TMEM_SETMAXNREG
#define ARMNET_TMEM_SETMAXNREG 0
#endif
// 20 warps = 16 consumers (4 per TMEM lane quadrant) + the producer warpgroup: gather warp, convert warp, MMA warp
// (+ 1 idle).  5 warps per SM sub-partition -> 96 registers per thread: 40 logit + 10 accumulator registers leave room
// for the software pipelining of the sweeps.  Measured at C2a (profiles/r2_v3_summary.md): 12 consumers x 128 registers
// 33.5 M / 15.5 M samples/s (init / trained-like), 16 x 96: 35.2 M / 16.1 M, 20 x 80: 34.3 M / 16.2 M.
// ARMNET_TMEM_SETMAXNREG (experiment, no gain: ptxas keeps the launch-bound register cap): the producer warpgroup gives
// registers back and the consumers grow.
constexpr int kTmWarpGather = kTmConsumers, kTmWarpConvert = kTmConsumers + 1, kTmWarpMma = kTmConsumers + 2;
constexpr int kTmThreads = (kTmConsumers + 4) * 32;
constexpr int kTmTiles = 4;                           // B-tile ring (tiles of 2 samples)
constexpr int kTmKP = 32;                             // packed K (floats per operand row = 128 bytes)
constexpr int kTmEL = 10;                             // embedding lanes of the narrow packed layout (nemb <= 10)
constexpr int kTmELWide = 16;                         // ... of the wide one (nemb <= 16)
// Packed operands.  B row (128 bytes): EL = 10: [e0..9 | l0..9 | e0..9 | 0 0], EL = 16: [e0..15 | l0..15]  (l = e - trunc_tf32(e)).
// A row: EL = 10: [M | M | m | 0 0] = 32 columns, four K = 8 steps against B floats 0, 8, 16, 24;
//        EL = 16: [M | M | m]       = 48 columns, six steps against B floats 0, 8 (M e), 16, 24 (M l), 0, 8 (m e).
__host__ __device__ constexpr int tm_a_cols(int el) { return el <= kTmEL ? 32 : 48; }
__host__ __device__ constexpr int tm_k_steps(int el) { return tm_a_cols(el) / 8; }
__host__ __device__ constexpr int tm_b_float(int el, int k) { return el <= kTmEL ? 8 * k : 8 * (k < 4 ? k : k - 4); }

// Timeline trace (tuning builds only, -DARMNET_TMEM_TRACE): clock64() of CTA-local events, 64 slots per CTA.
#ifdef ARMNET_TMEM_TRACE
__device__ long long g_tm_trace[148 * 64];
#define TM_TRACE(slot)                                                                        \
    do {                                                                                      \
        if ((threadIdx.x & 31) == 0 && blockIdx.x < 148) g_tm_trace[blockIdx.x * 64 + (slot)] = clock64(); \
    } while (0)
#define TM_TRACE_ONCE(slot, cond)                                                             \
    do {                                                                                      \
        if (cond) TM_TRACE(slot);                                                             \
    } while (0)
#else
#define TM_TRACE(slot) do { } while (0)
#define TM_TRACE_ONCE(slot, cond) do { } while (0)
#endif

#ifndef ARMNET_TM_UMULHI
#define ARMNET_TM_UMULHI 1     // item -> tile index by multiply-high instead of a runtime division
#endif

struct TmemParams {
    const void *ids;
    float *values;
    const float *table;
    const float *Apk;         // [R/128][128][32]  packed rows of M'^T (attn_prepare_tmem_kernel)
    const float2 *Vpk;        // [R][vstr]         att_values, field-packed (f, f+1) pairs per row
    const float *post_mean;   // [R] or null: eval-mode arm_bn, out = (z - mean) * scale + shift
    const float *post_scale;
    const float *post_shift;
    float *out_z;             // [B][R][E]
    int *err_flag;
    long long V, ld, B;
    int F, E, R;
    int ids_i32;
    int clamp, clamp_inplace;
    float clamp_lo, clamp_hi;
    EntmaxParams ep;
    int n_tiles;     // ceil(B / 2)
    int n_raw;       // raw ring depth in samples (= 2 * look)
    int look;        // tiles of gather look-ahead
    int row_bytes;   // bytes fetched per embedding row (multiple of 16)
    int tma_store;
    unsigned inv_ipt;   // ceil(2^32 / items per tile): item -> tile index by __umulhi
};

// Shared-memory carve-up, identical on host (sizing) and device (pointers). Offsets in bytes.
struct TmemSmem {
    int off_bar, off_tiles, off_V, off_raw, off_vals, off_out, total;
    int tile_bytes, v_bytes, vstr, raw_sample_bytes, fpad, out_floats;
    __host__ __device__ static int up(int x, int a) { return (x + a - 1) / a * a; }
    __host__ __device__ TmemSmem(int NP, int NR, const TmemParams &P) {
        const int NFP = 2 * NP;
        fpad = NFP;
        vstr = NP + 2;                     // float2 per row; (NP + 2) * 8 bytes = 16 * odd for NP = 20: conflict-free LDS.128
        tile_bytes = 2 * NFP * kTmKP * 4;  // 2 samples x NFP rows x 128 bytes (a multiple of 1024 when NFP % 8 == 0)
        v_bytes = up(P.R * vstr * 8, 16);
        raw_sample_bytes = up(P.F * P.row_bytes, 16);
        out_floats = 64 * P.E;             // per warp: two pieces of 32 rows (NR = 2: one unit; NR = 1: two units in flight)
        off_bar = 0;                       // mbarriers + counters: 1024 bytes
        off_tiles = 1024;
        off_V = off_tiles + kTmTiles * tile_bytes;
        off_raw = up(off_V + v_bytes, 128);
        off_vals = up(off_raw + P.n_raw * raw_sample_bytes, 16);
        off_out = up(off_vals + P.n_raw * fpad * 4, 128);
        total = off_out + (P.tma_store ? kTmConsumers * out_floats * 4 : 0);
    }
};

// ------------------------------------------------------------------ tcgen05 wrappers (same PTX as csrc/mlp.cu)
__device__ __forceinline__ void tm_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tm_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tm_alloc(uint32_t *slot, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tm_dealloc(uint32_t addr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}
// D[tmem] (+)= A[tmem] . B[smem]^T   (A: 128 lanes x 8 32-bit columns at `a`)
__device__ __forceinline__ void tm_mma_tf32_ts(uint32_t d, uint32_t a, uint64_t db, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n}\n" ::"r"(d),
        "r"(a), "l"(db), "r"(idesc), "r"(acc)
        : "memory");
}
__device__ __forceinline__ void tm_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void tm_st8(uint32_t taddr, const float (&r)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "f"(r[0]),
                 "f"(r[1]), "f"(r[2]), "f"(r[3]), "f"(r[4]), "f"(r[5]), "f"(r[6]), "f"(r[7])
                 : "memory");
}
__device__ __forceinline__ void tm_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tm_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// 32 lanes x 16 pairs of adjacent columns -> X[0..15] of each lane
__device__ __forceinline__ void tm_ld32(uint32_t taddr, float2 *X) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=f"(X[0].x), "=f"(X[0].y), "=f"(X[1].x), "=f"(X[1].y), "=f"(X[2].x), "=f"(X[2].y), "=f"(X[3].x), "=f"(X[3].y),
          "=f"(X[4].x), "=f"(X[4].y), "=f"(X[5].x), "=f"(X[5].y), "=f"(X[6].x), "=f"(X[6].y), "=f"(X[7].x), "=f"(X[7].y),
          "=f"(X[8].x), "=f"(X[8].y), "=f"(X[9].x), "=f"(X[9].y), "=f"(X[10].x), "=f"(X[10].y), "=f"(X[11].x),
          "=f"(X[11].y), "=f"(X[12].x), "=f"(X[12].y), "=f"(X[13].x), "=f"(X[13].y), "=f"(X[14].x), "=f"(X[14].y),
          "=f"(X[15].x), "=f"(X[15].y)
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tm_ld8(uint32_t taddr, float2 *X) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=f"(X[0].x), "=f"(X[0].y), "=f"(X[1].x), "=f"(X[1].y), "=f"(X[2].x), "=f"(X[2].y), "=f"(X[3].x),
                   "=f"(X[3].y)
                 : "r"(taddr)
                 : "memory");
}
// K-major operand tile, rows of 128 bytes, SWIZZLE_128B, 8-row atoms 1024 bytes apart (csrc/mlp.cu: umma_desc_sw128)
__device__ __forceinline__ uint64_t tm_desc_sw128(const void *tile) {
    const uint64_t addr = (uint64_t)((smem_u32(tile) & 0x3FFFFu) >> 4);
    return addr | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
__host__ __device__ constexpr uint32_t tm_idesc_tf32(int m, int n) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
// Bounded mbarrier wait: a protocol bug traps (the launch fails with an error) instead of hanging the device.
__device__ __forceinline__ void tm_wait(uint64_t *bar, uint32_t parity) {
#pragma unroll 1
    for (uint32_t spin = 0; spin < (1u << 24); ++spin) {
        uint32_t ok;
        asm volatile(
            "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
            : "=r"(ok)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
        if (ok) return;
    }
    asm volatile("trap;");
}
__device__ __forceinline__ float trunc_tf32(float x) { return __uint_as_float(__float_as_uint(x) & 0xffffe000u); }

// NP field pairs per row: F = 2 NP (ODD = false) or 2 NP - 1 (ODD = true).  NFP = 2 NP rows per sample in the B tile.
// NR = 1: a thread owns one row, its logits in registers (4 D slots of one A block, freed as soon as they are loaded).
// NR = 2: a thread owns two rows (same TMEM lane of two A blocks) whose logits are streamed from tensor memory on every
//         pass (entmax_stream.cuh): every shared-memory read of an embedding row feeds two rows -- the one-row mapping is
//         bound by the shared-memory -> register return path -- and a D slot (2 blocks) lives until its units are done.
template <int NP, bool ODD, int NR, int EL>
__global__ void __launch_bounds__(kTmThreads, 1) armnet_fwd_tmem_kernel(const __grid_constant__ TmemParams P) {
    static_assert(EL == kTmEL || (EL == kTmELWide && NR == 1), "embedding lanes: 10 (both thread mappings) or 16 (one row per thread)");
    constexpr int KA = tm_a_cols(EL), KSTEPS = tm_k_steps(EL);
    constexpr int NFP = 2 * NP;
    static_assert(NFP % 8 == 0 && 2 * NFP <= 256 && (2 * NFP) % 16 == 0, "two samples of NFP rows are the N of one MMA");
    static_assert(NP % 4 == 0 && (NP < 16 || NP >= 16), "tcgen05.ld shapes: one .x32 for 16 pairs, .x8 for every 4 more");
    constexpr int DSLOT = NR * 2 * NFP;  // TMEM columns of a D slot: NR A blocks (128 neurons each) x 2 samples x NFP fields
    constexpr int NSLOT = 4 / NR;        // D slots (A operand: <= 192 columns, D: 320)
    extern __shared__ __align__(1024) unsigned char smem_tm[];
    unsigned char *smem = smem_tm;
    const TmemSmem L(NP, NR, P);
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem + L.off_bar);
    uint64_t *v_ready = bar;                        // V table landed
    uint64_t *d_full = bar + 1;                     // [4]  MMAs of an item done (tcgen05.commit)
    uint64_t *d_empty = bar + 5;                    // [4]  all 8 units of an item hold their logits in registers
    uint64_t *tile_full = bar + 9;                  // [kTmTiles]  tile converted (e rows visible to the consumers)
    uint64_t *tile_empty = tile_full + kTmTiles;    // [kTmTiles]  every unit of the tile is done reading e
    uint64_t *raw_full = tile_empty + kTmTiles;     // [n_raw <= 32]  gathered rows of a sample landed
    uint64_t *raw_empty = raw_full + 32;            // [n_raw <= 32]  the convert warp is done with the sample's raw rows
    uint64_t *a_ready = bar + 90;                   // [NBLK <= 8]  A block kb (128 rows of M'^T) sits in tensor memory
    int *next_unit = reinterpret_cast<int *>(smem + L.off_bar + 960);   // [4] one counter per lane quadrant
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + L.off_bar + 992);
    unsigned char *tiles = smem + L.off_tiles;
    const float2 *Vs = reinterpret_cast<const float2 *>(smem + L.off_V);
    unsigned char *raw = smem + L.off_raw;
    float *vals = reinterpret_cast<float *>(smem + L.off_vals);
    float *outs = reinterpret_cast<float *>(smem + L.off_out);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int F = P.F, E = P.E, R = P.R;
    const int NBLK = R >> 7;         // A blocks of 128 neurons
    const int IPT = NBLK / NR;       // items per tile (an item = NR blocks x 2 samples)
    const int A_COLS = NBLK * KA;
    const int G = (int)gridDim.x;
    const int n_local = (P.n_tiles - (int)blockIdx.x + G - 1) / G;   // tiles blockIdx.x + t * G
    const EntmaxParams ep = P.ep;
    TM_TRACE_ONCE(0, warp == 0);

    const int rows_tile = 2 * F;          // ids / values of a tile are contiguous in the [B, F] arrays
    constexpr int KR = (2 * NFP + 31) / 32;

    if (tid == 0) {
        mbar_init(v_ready, 1);
        for (int kb = 0; kb < NBLK; ++kb) mbar_init(&a_ready[kb], 4);
        for (int s = 0; s < NSLOT; ++s) {
            mbar_init(&d_full[s], 1);
            mbar_init(&d_empty[s], 8);
        }
        for (int s = 0; s < kTmTiles; ++s) {
            mbar_init(&tile_full[s], 1);
            mbar_init(&tile_empty[s], IPT * 8);
        }
        for (int s = 0; s < P.n_raw; ++s) {
            mbar_init(&raw_full[s], 1);
            mbar_init(&raw_empty[s], 1);
        }
        for (int qd = 0; qd < 4; ++qd) next_unit[qd] = 0;
        mbar_fence_init();
        mbar_arrive_expect_tx(v_ready, (uint32_t)L.v_bytes);
        tma_load_bulk(smem + L.off_V, P.Vpk, (uint32_t)L.v_bytes, v_ready);
    }
    if (warp == kTmWarpMma) tm_alloc(tmem_slot, 512);
    // B tiles start as zeros: the pad rows (field NFP-1 for odd F, samples past the batch) and the two pad floats of
    // every row are never written again
    for (int i = tid; i < kTmTiles * L.tile_bytes / 16; i += kTmThreads)
        reinterpret_cast<float4 *>(tiles)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    fence_proxy_async_smem();
    tm_fence_before();
    __syncthreads();
    tm_fence_after();
    TM_TRACE_ONCE(1, warp == 0);
    const uint32_t tmem = *tmem_slot;
    // A operand: TMEM columns [0, A_COLS), lane i of block kb = packed row (kb, i).  A warp can only write the 32 lanes of
    // its own quadrant (warp % 4): consumer warps 0-3 load one quadrant each, published block by block (the first MMA needs
    // block 0 only).  (Letting the four producer-group warps do it instead delays the first MMA: 116.7 vs 110.6 us.)
    auto load_a_quadrant = [&]() {
        const int qa = warp & 3;
        const uint32_t taddr = tmem + ((uint32_t)(qa * 32) << 16);
#pragma unroll 1
        for (int kb = 0; kb < NBLK; ++kb) {
            const float4 *src = reinterpret_cast<const float4 *>(P.Apk + ((long long)kb * 128 + qa * 32 + lane) * KA);
            float4 v[KA / 4];
#pragma unroll
            for (int c = 0; c < KA / 4; ++c) v[c] = __ldg(src + c);
#pragma unroll
            for (int c8 = 0; c8 < KA / 8; ++c8) {
                const float4 a = v[2 * c8], b = v[2 * c8 + 1];
                const float r[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
                tm_st8(taddr + (uint32_t)(kb * KA + c8 * 8), r);
            }
            tm_st_wait();
            tm_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&a_ready[kb]);   // only the MMA warp waits for it, block by block
            TM_TRACE_ONCE(2, qa == 0 && kb == 0);
        }
    };
    if (warp < 4) load_a_quadrant();
    const uint32_t d_base = tmem + (uint32_t)A_COLS;

#if ARMNET_TMEM_SETMAXNREG
    if (warp >= kTmConsumers) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
    } else {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 152;");
    }
#endif
    auto sample_valid = [&](int t, int s) { return 2 * ((long long)blockIdx.x + (long long)t * G) + s < P.B; };

    if (warp == kTmWarpGather) {
        // =========================================================== gather warp: ids / values -> TMA bulk row gathers
        // `prefetch` only ISSUES the loads of a tile's ids / values (one tile ahead); `finish` clamps / range-checks at use
        long long pid[KR];
        float pv[KR];
        auto prefetch = [&](int t) {
            const long long row0 = ((long long)blockIdx.x + (long long)t * G) * rows_tile;
#pragma unroll
            for (int k = 0; k < KR; ++k) {
                const int idx = lane + 32 * k;
                const long long row = row0 + idx;
                long long id = 0;
                float v = 0.f;
                if (idx < rows_tile && t < n_local && row < P.B * F) {
                    id = P.ids_i32 ? (long long)__ldg(reinterpret_cast<const int *>(P.ids) + row)
                                   : __ldg(reinterpret_cast<const long long *>(P.ids) + row);
                    v = P.values[row];
                }
                pid[k] = id;
                pv[k] = v;
            }
        };
        auto finish = [&](int t) {
            const long long row0 = ((long long)blockIdx.x + (long long)t * G) * rows_tile;
#pragma unroll
            for (int k = 0; k < KR; ++k) {
                const int idx = lane + 32 * k;
                const long long row = row0 + idx;
                long long id = pid[k];
                float v = pv[k];
                if (idx < rows_tile && t < n_local && row < P.B * F) {
                    if (P.clamp) {  // armnet.py:82 -- in place on the caller's tensor
                        const float vc = fminf(fmaxf(v, P.clamp_lo), P.clamp_hi);
                        if (P.clamp_inplace && vc != v) P.values[row] = vc;
                        v = vc;
                    }
                    if ((unsigned long long)id >= (unsigned long long)P.V) {  // reference: IndexError (layers.py:20)
                        if (P.err_flag) atomicOr(P.err_flag, 1);
                        id = 0;
                        v = 0.f;
                    }
                }
                pid[k] = id;
                pv[k] = v;
            }
        };
        prefetch(0);
        TM_TRACE(53);
        int rs0 = 0, use = 0;   // raw-ring slot of the tile's first sample (n_raw is even: 2 t % n_raw) and its lap, no division
        for (int t = 0; t < n_local; ++t) {
            finish(t);
            // raw slots of tile t were last used by tile t - look: wait until the convert warp has read them
            if (lane < 2 && sample_valid(t, lane) && use > 0) tm_wait(&raw_empty[rs0 + lane], (uint32_t)(use - 1) & 1u);
            __syncwarp();
            TM_TRACE_ONCE(55, t == 0);
#pragma unroll
            for (int k = 0; k < KR; ++k) {   // the values first: they are published by the arrive below
                const int idx = lane + 32 * k;
                if (idx < rows_tile) {
                    const int s = idx >= F ? 1 : 0, f = idx - s * F;
                    if (sample_valid(t, s)) vals[(rs0 + s) * L.fpad + f] = pv[k];
                }
            }
            __syncwarp();
            TM_TRACE_ONCE(56, t == 0);
            if (lane < 2 && sample_valid(t, lane))
                mbar_arrive_expect_tx(&raw_full[rs0 + lane], (uint32_t)(F * P.row_bytes));
            __syncwarp();
            TM_TRACE_ONCE(57, t == 0);
#pragma unroll
            for (int k = 0; k < KR; ++k) {
                const int idx = lane + 32 * k;
                if (idx < rows_tile) {
                    const int s = idx >= F ? 1 : 0, f = idx - s * F;
                    if (sample_valid(t, s)) {
                        const int rs = rs0 + s;
                        tma_load_bulk(raw + rs * L.raw_sample_bytes + f * P.row_bytes, P.table + pid[k] * P.ld,
                                      (uint32_t)P.row_bytes, &raw_full[rs]);
                    }
                }
            }
            TM_TRACE_ONCE(3, t == 0);
            prefetch(t + 1);
            rs0 += 2;
            if (rs0 >= P.n_raw) {
                rs0 = 0;
                ++use;
            }
        }
        TM_TRACE(12);
    } else if (warp == kTmWarpConvert) {
        // =========================================================== convert warp
        // landed rows of tile t -> B tile: e = row * v (layers.py:21), packed [e | lo(e) | e | 0 0], 128-byte swizzle
        int rs0 = 0, use = 0;   // as in the gather warp
        for (int t = 0; t < n_local; ++t) {
            const int bt = t % kTmTiles;
            if (t >= kTmTiles) tm_wait(&tile_empty[bt], (uint32_t)(t / kTmTiles - 1) & 1u);
            for (int s = 0; s < 2; ++s)
                if (sample_valid(t, s)) tm_wait(&raw_full[rs0 + s], (uint32_t)use & 1u);
            unsigned char *tile = tiles + bt * L.tile_bytes;
            TM_TRACE_ONCE(4, t == 0);
#pragma unroll
            for (int k = 0; k < KR; ++k) {
                const int idx = lane + 32 * k;
                if (idx < rows_tile) {
                    const int s = idx >= F ? 1 : 0, f = idx - s * F;
                    if (sample_valid(t, s)) {
                        const int rs = rs0 + s;
                        const float4 *src = reinterpret_cast<const float4 *>(raw + rs * L.raw_sample_bytes + f * P.row_bytes);
                        const float v = vals[rs * L.fpad + f];
                        constexpr int EC4 = (EL + 3) / 4;   // 16-byte chunks of a gathered row
                        float e[4 * EC4], l[4 * EC4];
#pragma unroll
                        for (int c = 0; c < EC4; ++c) {
                            float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
                            if (c * 16 < P.row_bytes) q = src[c];
                            e[4 * c + 0] = q.x;
                            e[4 * c + 1] = q.y;
                            e[4 * c + 2] = q.z;
                            e[4 * c + 3] = q.w;
                        }
#pragma unroll
                        for (int x = 0; x < 4 * EC4; ++x) {
                            e[x] = (x < EL && x < E) ? __fmul_rn(e[x], v) : 0.f;
                            l[x] = e[x] - trunc_tf32(e[x]);
                        }
                        const int n = s * NFP + f;
                        float4 *dst = reinterpret_cast<float4 *>(tile + n * 128);
                        const int sw = n & 7;
                        if constexpr (EL == kTmEL) {
                            dst[0 ^ sw] = make_float4(e[0], e[1], e[2], e[3]);
                            dst[1 ^ sw] = make_float4(e[4], e[5], e[6], e[7]);
                            dst[2 ^ sw] = make_float4(e[8], e[9], l[0], l[1]);
                            dst[3 ^ sw] = make_float4(l[2], l[3], l[4], l[5]);
                            dst[4 ^ sw] = make_float4(l[6], l[7], l[8], l[9]);
                            dst[5 ^ sw] = make_float4(e[0], e[1], e[2], e[3]);
                            dst[6 ^ sw] = make_float4(e[4], e[5], e[6], e[7]);
                            dst[7 ^ sw] = make_float4(e[8], e[9], 0.f, 0.f);
                        } else {
#pragma unroll
                            for (int c = 0; c < 4; ++c) {
                                dst[c ^ sw] = make_float4(e[4 * c], e[4 * c + 1], e[4 * c + 2], e[4 * c + 3]);
                                dst[(4 + c) ^ sw] = make_float4(l[4 * c], l[4 * c + 1], l[4 * c + 2], l[4 * c + 3]);
                            }
                        }
                    }
                }
            }
            fence_proxy_async_smem();  // generic-proxy writes -> visible to the tensor core's reads
            __syncwarp();
            if (lane == 0) mbar_arrive(&tile_full[bt]);
            if (lane < 2 && sample_valid(t, lane)) mbar_arrive(&raw_empty[rs0 + lane]);
            TM_TRACE_ONCE(5, t == 0);
            rs0 += 2;
            if (rs0 >= P.n_raw) {
                rs0 = 0;
                ++use;
            }
        }
        TM_TRACE(13);
    } else if (warp == kTmWarpMma) {
        // =========================================================== MMA warp: one elected lane issues tcgen05.mma
        const uint32_t idesc = tm_idesc_tf32(128, 2 * NFP);
        for (int t = 0; t < n_local; ++t) {
            const int bt = t % kTmTiles;
            tm_wait(&tile_full[bt], (uint32_t)(t / kTmTiles) & 1u);
            const uint64_t db = tm_desc_sw128(tiles + bt * L.tile_bytes);
            for (int h = 0; h < IPT; ++h) {
                const int i = t * IPT + h, slot = i % NSLOT;
                if (i >= NSLOT) tm_wait(&d_empty[slot], (uint32_t)(i / NSLOT - 1) & 1u);
                if (t == 0)
                    for (int j = 0; j < NR; ++j) tm_wait(&a_ready[h * NR + j], 0);
                tm_fence_after();
                TM_TRACE_ONCE(6, i == 0);
                if (lane == 0) {
                    const uint32_t d = d_base + (uint32_t)(slot * DSLOT);
#pragma unroll
                    for (int j = 0; j < NR; ++j)
#pragma unroll
                        for (int k = 0; k < KSTEPS; ++k)
                            tm_mma_tf32_ts(d + (uint32_t)(j * 2 * NFP), tmem + (uint32_t)((h * NR + j) * KA + k * 8),
                                           db + (uint64_t)(tm_b_float(EL, k) * 4 >> 4), idesc, (uint32_t)k);
                    tm_commit(&d_full[slot]);
                }
                __syncwarp();
                TM_TRACE_ONCE(7, i == 0);
            }
        }
        TM_TRACE(14);
    } else if (warp > kTmWarpMma) {
        // idle warp of the producer warpgroup
    } else {
        // =========================================================== consumer warps
        const int qd = warp & 3;
        const int n_units = n_local * IPT * 2;
        float *ost_warp = outs + warp * L.out_floats;
        int obuf = 0;
        tm_wait(v_ready, 0);
#if ARMNET_TM_UMULHI
#define TM_DIV_IPT(i) ((int)__umulhi((uint32_t)(i), P.inv_ipt))   // host-computed ceil(2^32 / IPT): no division, no register
#else
#define TM_DIV_IPT(i) ((i) / IPT)
#endif
        int u_next = 0;
        bool first_unit = true;
        (void)first_unit;
        if (lane == 0) u_next = atomicAdd(&next_unit[qd], 1);
        if constexpr (NR == 2) {
        for (;;) {
            const int u = __shfl_sync(0xffffffffu, u_next, 0);
            if (u >= n_units) break;
            const int i = u >> 1, s = u & 1;       // item, sample inside the tile
            const int t = TM_DIV_IPT(i), h = i - t * IPT;
            const int slot = i % NSLOT;
            const int bt = t % kTmTiles;
            const long long b = 2 * ((long long)blockIdx.x + (long long)t * G) + s;
            const bool valid = b < P.B;
            const int r0 = h * 256 + qd * 32 + lane;   // this lane's two neurons: r0 and r0 + 128

            tm_wait(&d_full[slot], (uint32_t)(i / NSLOT) & 1u);
            tm_fence_after();
            if (lane == 0) u_next = atomicAdd(&next_unit[qd], 1);   // claim the next unit: the latency hides under this one
            tm_wait(&tile_full[bt], (uint32_t)(t / kTmTiles) & 1u);

            float2 acc[2][EL / 2];
            float tau[2], S[2] = {1.f, 1.f};
            if (valid) {
                const uint32_t t0 = d_base + ((uint32_t)(qd * 32) << 16) + (uint32_t)(slot * DSLOT + s * NFP);
                const unsigned char *eb = tiles + bt * L.tile_bytes + s * (NFP * 128);
                const float4 *vr0 = reinterpret_cast<const float4 *>(Vs + (size_t)r0 * L.vstr);
                const float4 *vr1 = reinterpret_cast<const float4 *>(Vs + (size_t)(r0 + 128) * L.vstr);
                // field pair j = 4c + k (c: runtime chunk, k: unrolled): V pairs (2j, 2j+1) sit in float4 2c + (k >> 1) of the
                // row; e rows f = 8c + 2k, 8c + 2k + 1 are 1024 c bytes into the sample's rows, swizzle f & 7 = 2k, 2k + 1
                auto vrow = [&](int n, int c, int k) -> float2 {
                    const float4 v4 = (n ? vr1 : vr0)[2 * c + (k >> 1)];
                    return (k & 1) ? make_float2(v4.z, v4.w) : make_float2(v4.x, v4.y);
                };
                auto fma_field = [&](const unsigned char *row, int sw, float w0, float w1) {
                    const float4 c0 = *reinterpret_cast<const float4 *>(row + ((0 ^ sw) << 4));
                    const float4 c1 = *reinterpret_cast<const float4 *>(row + ((1 ^ sw) << 4));
                    const float2 c2 = *reinterpret_cast<const float2 *>(row + ((2 ^ sw) << 4));
                    const float2 e2[EL / 2] = {make_float2(c0.x, c0.y), make_float2(c0.z, c0.w), make_float2(c1.x, c1.y),
                                               make_float2(c1.z, c1.w), c2};
                    const float2 a = splat2(w0), bw = splat2(w1);
#pragma unroll
                    for (int x = 0; x < EL / 2; ++x) {
                        acc[0][x] = ffma2(a, e2[x], acc[0][x]);     // s[x] += w_f e[f,x]  (armnet.py:86-87)
                        acc[1][x] = ffma2(bw, e2[x], acc[1][x]);
                    }
                };
                auto cross = [&](int c, int k, float2 w0, float2 w1) {
                    const unsigned char *rows8 = eb + c * (8 * 128);
                    fma_field(rows8 + (2 * k) * 128, 2 * k, w0.x, w1.x);
                    if (!ODD || k < 3 || c < NP / 4 - 1) fma_field(rows8 + (2 * k + 1) * 128, 2 * k + 1, w0.y, w1.y);
                };
                auto reset = [&]() {
#pragma unroll
                    for (int x = 0; x < EL / 2; ++x) acc[0][x] = acc[1][x] = make_float2(0.f, 0.f);
                };
                if (!stream_dense_entmax_cross<NP, ODD>(t0, t0 + 2 * NFP, ep, tau, S, vrow, cross, reset)) {
                    // not near-uniform (or alpha in {1, 1.5, 2}): one row at a time with its logits in registers
#pragma unroll 1
                    for (int n = 0; n < 2; ++n) {
                        float2 X[1][NP];
                        const uint32_t tn = t0 + (uint32_t)(n * 2 * NFP);
                        if (NP >= 16) tm_ld32(tn, &X[0][0]);
#pragma unroll
                        for (int j = (NP >= 16 ? 16 : 0); j < NP; j += 4) tm_ld8(tn + 2 * j, &X[0][j]);
                        tm_ld_wait();
                        if (ODD) X[0][NP - 1].y = neg_inf();
                        const float4 *vr = n ? vr1 : vr0;
                        float2 a1[EL / 2];
                        float tau1[1], S1r[1] = {1.f};
                        auto vrow1 = [&](int, int j) -> float2 {
                            const float4 v4 = vr[j >> 1];
                            return (j & 1) ? make_float2(v4.z, v4.w) : make_float2(v4.x, v4.y);
                        };
                        auto fma1 = [&](int f, float w) {
                            const unsigned char *row = eb + f * 128;
                            const int sw = f & 7;
                            const float4 c0 = *reinterpret_cast<const float4 *>(row + ((0 ^ sw) << 4));
                            const float4 c1 = *reinterpret_cast<const float4 *>(row + ((1 ^ sw) << 4));
                            const float2 c2 = *reinterpret_cast<const float2 *>(row + ((2 ^ sw) << 4));
                            const float2 w2 = splat2(w);
                            a1[0] = ffma2(w2, make_float2(c0.x, c0.y), a1[0]);
                            a1[1] = ffma2(w2, make_float2(c0.z, c0.w), a1[1]);
                            a1[2] = ffma2(w2, make_float2(c1.x, c1.y), a1[2]);
                            a1[3] = ffma2(w2, make_float2(c1.z, c1.w), a1[3]);
                            a1[4] = ffma2(w2, c2, a1[4]);
                        };
                        auto cross1 = [&](int j, const float2 (&w)[1]) {
                            fma1(2 * j, w[0].x);
                            if (!ODD || j < NP - 1) fma1(2 * j + 1, w[0].y);
                        };
                        auto reset1 = [&]() {
#pragma unroll
                            for (int x = 0; x < EL / 2; ++x) a1[x] = make_float2(0.f, 0.f);
                        };
                        rows_entmax_cross<1, NP, ODD>(X, ep, tau1, S1r, vrow1, cross1, reset1);
                        if (n == 0) {
#pragma unroll
                            for (int x = 0; x < EL / 2; ++x) acc[0][x] = a1[x];
                            S[0] = S1r[0];
                        } else {
#pragma unroll
                            for (int x = 0; x < EL / 2; ++x) acc[1][x] = a1[x];
                            S[1] = S1r[0];
                        }
                    }
                }
            }
            // logits and e rows are not read any more: hand the D slot and the tile back (one arrival per unit each)
            tm_fence_before();
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(&d_empty[slot]);
                mbar_arrive(&tile_empty[bt]);
            }
            if (!valid) continue;

            // ---- s = acc / S (entmax.py:63-64 renormalisation), z = exp(s) (armnet.py:86) [, eval-mode arm_bn, :89]
            float *ost_base = ost_warp;
            if (P.tma_store) {
                if (lane == 0) tma_store_wait_read<0>();  // the previous unit's bulk stores (a whole unit ago) have drained
                __syncwarp();
            }
#pragma unroll
            for (int n = 0; n < 2; ++n) {
                const int r = r0 + n * 128;
                float z[EL];
                const float2 k2 = splat2(__frcp_rn(S[n]) * 1.4426950408889634f);
#pragma unroll
                for (int x = 0; x < EL / 2; ++x) {
                    const float2 t2 = fmul2(acc[n][x], k2);
                    z[2 * x] = fast_ex2(t2.x);
                    z[2 * x + 1] = fast_ex2(t2.y);
                }
                if (P.post_scale != nullptr) {
                    const float m = __ldg(P.post_mean + r), a = __ldg(P.post_scale + r), sh = __ldg(P.post_shift + r);
#pragma unroll
                    for (int x = 0; x < EL; ++x) z[x] = fmaf(z[x] - m, a, sh);
                }
                float *dst = P.tma_store ? ost_base + n * 32 * E + lane * E
                                         : P.out_z + ((long long)b * R + r) * (long long)E;
                if ((E & 1) == 0 && P.tma_store) {
#pragma unroll
                    for (int x = 0; x < EL; x += 2)
                        if (x < E) *reinterpret_cast<float2 *>(dst + x) = make_float2(z[x], z[x + 1]);
                } else {
#pragma unroll
                    for (int x = 0; x < EL; ++x)
                        if (x < E) dst[x] = z[x];
                }
            }
            if (P.tma_store) {
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) {   // two pieces of 32 consecutive rows (neurons 256h + 32qd .. and + 128)
                    float *g0 = P.out_z + ((long long)b * R + h * 256 + qd * 32) * (long long)E;
                    tma_store_bulk(g0, ost_base, (uint32_t)(32 * E * 4));
                    tma_store_bulk(g0 + 128 * E, ost_base + 32 * E, (uint32_t)(32 * E * 4));
                    tma_store_commit();
                }
            }
        }
        } else {
        for (;;) {
            const int u = __shfl_sync(0xffffffffu, u_next, 0);
            if (u >= n_units) break;
            const int i = u >> 1, s = u & 1;       // item, sample inside the tile
            const int t = TM_DIV_IPT(i), kb = i - t * NBLK;   // NR == 1: IPT == NBLK
            const int slot = i % NSLOT;
            const int bt = t % kTmTiles;
            const long long b = 2 * ((long long)blockIdx.x + (long long)t * G) + s;
            const bool valid = b < P.B;
            const int r = kb * 128 + qd * 32 + lane;   // this lane's neuron

            // ---- logits of this lane's row of sample s: TMEM -> registers, fields packed in (f, f+1) pairs
            float2 X[1][NP];
            tm_wait(&d_full[slot], (uint32_t)(i / NSLOT) & 1u);
            tm_fence_after();
            TM_TRACE_ONCE(36 + warp, first_unit);
            first_unit = false;
            {
                const uint32_t tn = d_base + ((uint32_t)(qd * 32) << 16) + (uint32_t)(slot * DSLOT + s * NFP);
                if (NP >= 16) tm_ld32(tn, &X[0][0]);
#pragma unroll
                for (int j = (NP >= 16 ? 16 : 0); j < NP; j += 4) tm_ld8(tn + 2 * j, &X[0][j]);
                tm_ld_wait();
            }
            tm_fence_before();
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(&d_empty[slot]);               // the MMAs of item i + NSLOT may overwrite the slot
                u_next = atomicAdd(&next_unit[qd], 1);     // claim the next unit now: the latency hides under this one
            }
            if (ODD) X[0][NP - 1].y = neg_inf();
            tm_wait(&tile_full[bt], (uint32_t)(t / kTmTiles) & 1u);


            float2 acc[EL / 2];
            float tau[1], S[1] = {1.f};
            if (valid) {
                // e rows of sample s: row n = s * NFP + f of the tile, chunk c at (c ^ (n & 7)); NFP % 8 == 0 -> n & 7 == f & 7
                const unsigned char *eb = tiles + bt * L.tile_bytes + s * (NFP * 128);
                const float4 *vr = reinterpret_cast<const float4 *>(Vs + (size_t)r * L.vstr);
                auto vrow = [&](int, int j) -> float2 {
                    const float4 v4 = vr[j >> 1];
                    return (j & 1) ? make_float2(v4.z, v4.w) : make_float2(v4.x, v4.y);
                };
                auto fma_field = [&](int f, float w) {
                    const unsigned char *row = eb + f * 128;
                    const int sw = f & 7;
                    const float4 c0 = *reinterpret_cast<const float4 *>(row + ((0 ^ sw) << 4));
                    const float4 c1 = *reinterpret_cast<const float4 *>(row + ((1 ^ sw) << 4));
                    const float2 w2 = splat2(w);
                    acc[0] = ffma2(w2, make_float2(c0.x, c0.y), acc[0]);     // s[x] += w_f e[f,x]  (armnet.py:86-87)
                    acc[1] = ffma2(w2, make_float2(c0.z, c0.w), acc[1]);
                    acc[2] = ffma2(w2, make_float2(c1.x, c1.y), acc[2]);
                    acc[3] = ffma2(w2, make_float2(c1.z, c1.w), acc[3]);
                    if constexpr (EL == kTmEL) {
                        const float2 c2 = *reinterpret_cast<const float2 *>(row + ((2 ^ sw) << 4));
                        acc[4] = ffma2(w2, c2, acc[4]);
                    } else {
                        const float4 c2 = *reinterpret_cast<const float4 *>(row + ((2 ^ sw) << 4));
                        const float4 c3 = *reinterpret_cast<const float4 *>(row + ((3 ^ sw) << 4));
                        acc[4] = ffma2(w2, make_float2(c2.x, c2.y), acc[4]);
                        acc[5] = ffma2(w2, make_float2(c2.z, c2.w), acc[5]);
                        acc[6] = ffma2(w2, make_float2(c3.x, c3.y), acc[6]);
                        acc[7] = ffma2(w2, make_float2(c3.z, c3.w), acc[7]);
                    }
                };
                auto cross = [&](int j, const float2 (&w)[1]) {
                    fma_field(2 * j, w[0].x);
                    if (!ODD || j < NP - 1) fma_field(2 * j + 1, w[0].y);
                };
                auto reset = [&]() {
#pragma unroll
                    for (int x = 0; x < EL / 2; ++x) acc[x] = make_float2(0.f, 0.f);
                };
                rows_entmax_cross<1, NP, ODD>(X, ep, tau, S, vrow, cross, reset);
            }
            TM_TRACE_ONCE(10, warp == 0 && u == 0);
            // every lane is done reading the tile: hand it back (one arrival per unit)
            __syncwarp();
            if (lane == 0) mbar_arrive(&tile_empty[bt]);
            if (!valid) continue;

            // ---- s = acc / S (entmax.py:63-64 renormalisation), z = exp(s) (armnet.py:86) [, eval-mode arm_bn, :89]
            float z[EL];
            {
                const float2 k2 = splat2(__frcp_rn(S[0]) * 1.4426950408889634f);
#pragma unroll
                for (int x = 0; x < EL / 2; ++x) {
                    const float2 t2 = fmul2(acc[x], k2);
                    z[2 * x] = fast_ex2(t2.x);
                    z[2 * x + 1] = fast_ex2(t2.y);
                }
            }
            if (P.post_scale != nullptr) {
                const float m = __ldg(P.post_mean + r), a = __ldg(P.post_scale + r), sh = __ldg(P.post_shift + r);
#pragma unroll
                for (int x = 0; x < EL; ++x) z[x] = fmaf(z[x] - m, a, sh);
            }
            // the unit's 32 rows are contiguous in out_z: stage them, one TMA bulk store
            float *gdst = P.out_z + ((long long)b * R + kb * 128 + qd * 32) * (long long)E;
            if (P.tma_store) {
                float *ost_base = ost_warp + obuf * (L.out_floats / 2);
                obuf ^= 1;
                if (lane == 0) tma_store_wait_read<1>();  // the bulk store before the previous one has drained this buffer
                __syncwarp();
                float *ost = ost_base + lane * E;
                if ((E & 1) == 0) {
#pragma unroll
                    for (int x = 0; x < EL; x += 2)
                        if (x < E) *reinterpret_cast<float2 *>(ost + x) = make_float2(z[x], z[x + 1]);
                } else {
#pragma unroll
                    for (int x = 0; x < EL; ++x)
                        if (x < E) ost[x] = z[x];
                }
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) {
                    tma_store_bulk(gdst, ost_base, (uint32_t)(32 * E * 4));
                    tma_store_commit();
                }
            } else {
                float *dst = gdst + lane * E;
#pragma unroll
                for (int x = 0; x < EL; ++x)
                    if (x < E) dst[x] = z[x];
            }
        }
        }
        TM_TRACE(16 + warp);
        if (P.tma_store && lane == 0) tma_store_wait_all<0>();
    }
    tm_fence_before();
    __syncthreads();
    TM_TRACE_ONCE(15, warp == 0);
    if (warp == kTmWarpMma) tm_dealloc(tmem, 512);
}

// One compiled shape of the kernel.
struct TmemInstance {
    int NP;
    int odd;
    int EL;                // embedding lanes of the packed layout: 10 (nemb <= 10) or 16 (nemb <= 16)
    const void *kernel;    // one row per thread, logits in registers
    const void *kernel2;   // two rows per thread, logits streamed from tensor memory (needs K*O % 256 == 0; EL = 10 only)
};
#define ARMNET_TMEM_INSTANCE(NP, ODD)                                                      \
    { NP, ODD, kTmEL, (const void *)&armnet_fwd_tmem_kernel<NP, (ODD) != 0, 1, kTmEL>,     \
      (const void *)&armnet_fwd_tmem_kernel<NP, (ODD) != 0, 2, kTmEL> }
#define ARMNET_TMEM_INSTANCE_WIDE(NP, ODD)                                                          \
    { NP, ODD, kTmELWide, (const void *)&armnet_fwd_tmem_kernel<NP, (ODD) != 0, 1, kTmELWide>, nullptr }

}  // namespace armnet
