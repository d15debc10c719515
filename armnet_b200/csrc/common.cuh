// Shared device/host helpers for libarmnet_b200.so (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/armnet_b200.h"

namespace armnet {

// ------------------------------------------------------------------ host-side error plumbing
void set_error(const char *fmt, ...);          // capi.cu (thread-local text)
int cuda_fail(cudaError_t e, const char *what);  // records text, returns ARMNET_ERR_CUDA
void note_launches(int n);

#define ARMNET_CUDA_TRY(expr)                                        \
    do {                                                             \
        cudaError_t _e = (expr);                                     \
        if (_e != cudaSuccess) return ::armnet::cuda_fail(_e, #expr); \
    } while (0)

struct DeviceInfo {
    int sm_count;
    int smem_optin;
};
int get_device_info(DeviceInfo *out);  // cached per device (read-only after first use)

// Tuning / experiment switches.  Read ONCE per process from the ARMNET_* environment variables (never in a launch
// path); armnet_set_tuning() changes them afterwards (tests, A/B runs).  -1 = the library's own default.
struct Tuning {
    int tmem;           // ARMNET_TMEM         1 / 0: force armnet_fwd_tmem_kernel on / off where the shape allows it
    int tmem_rows;      // ARMNET_TMEM_ROWS    1 / 2: rows per thread of armnet_fwd_tmem_kernel (default 1; 2 needs K*O % 256 == 0)
    int mma;            // ARMNET_MMA          1 / 0: force armnet_fwd_mma_kernel on / off
    int mma_split_rna;  // ARMNET_MMA_SPLIT=rna
    int mma_warps;      // ARMNET_MMA_WARPS    12: the 12-warp instance
    int force_nw;       // ARMNET_FORCE_NW     CTA size of armnet_fwd_kernel
    int force_look;     // ARMNET_FORCE_LOOK   gather look-ahead (tiles)
    int lockstep;       // ARMNET_LOCKSTEP
    int no_tma_gather;  // ARMNET_NO_TMA_GATHER
    int no_tma_store;   // ARMNET_NO_TMA_STORE
    int gemm_1cta;      // ARMNET_GEMM_1CTA    first MLP Linear on the single-CTA kernel
};
Tuning &tuning();

// ------------------------------------------------------------------ entmax solver parameters
enum PowMode : int {
    POW_SOFTMAX = 0,    // alpha == 1   : p = exp(g - max)
    POW_GENERAL = 1,    // 1 < alpha < 2, q = 1/(alpha-1) > 1 : Newton on tau, p = u^q via lg2/ex2
    POW_SQUARE = 2,     // alpha == 1.5 : p = u*u, Newton without MUFU
    POW_LINEAR = 3,     // alpha == 2   : p = u, Michelot
    POW_BISECT = 4,     // reference algorithm (any alpha > 1), p = u^q via lg2/ex2
};

struct EntmaxParams {
    int mode;        // PowMode
    int n_iter;      // bisection halvings (POW_BISECT)
    float am1;       // fp32(alpha) - 1            (entmax.py:31-36,42: alpha is an fp32 tensor)
    float q;         // 1 / am1 in fp32            (entmax.py:22)
    float qm1;       // q - 1
    float cF;        // F^-(alpha-1) = (1/d)^(alpha-1), the offset of tau_hi (entmax.py:47)
    float inv_F;     // 1 / F
    float uni_k;     // (q-1) / (2 cF): second-order term of the near-uniform closed-form start
    float uni_var;   // largest variance of X for which that start is tried: (0.2 cF)^2 / F
};

// Fills *ep for (alpha, F, solver). Returns ARMNET_OK or ARMNET_ERR_SHAPE.
int make_entmax_params(float alpha, int F, int solver, int n_iter, EntmaxParams *ep);

#ifdef __CUDACC__
// ------------------------------------------------------------------ device PTX wrappers

__device__ __forceinline__ float fast_lg2(float x) {
    float y;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float fast_ex2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}

// TMA (non-tensor bulk) global -> shared, completion counted in bytes on an mbarrier. SASS: UBLKCP.
__device__ __forceinline__ void tma_load_bulk(void *smem_dst, const void *gmem_src, uint32_t bytes,
                                              uint64_t *bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_u32(smem_dst)),
        "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}
// TMA (non-tensor bulk) shared -> global, tracked by the bulk async-group of the issuing thread.
__device__ __forceinline__ void tma_store_bulk(void *gmem_dst, const void *smem_src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gmem_dst),
                 "r"(smem_u32(smem_src)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void tma_store_wait_read() {
    asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
template <int N>
__device__ __forceinline__ void tma_store_wait_all() {
    asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory");
}
// generic-proxy smem writes -> visible to the async proxy (TMA store) that follows
__device__ __forceinline__ void fence_proxy_async_smem() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
#endif  // __CUDACC__

}  // namespace armnet
