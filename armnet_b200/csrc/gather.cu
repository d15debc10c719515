// armnet_embed_gather_f32: layers.Embedding.forward (models/layers.py:15-21) + value clamp (models/armnet.py:82).
//
// HBM-bound gather.  Algorithmic bytes per sample: F*(8 id + 4 value + 4E row read + 4E row write).
// One thread owns VEC consecutive floats of the flattened [B*F, E] output, so stores are fully coalesced and each
// gathered row is read by E/VEC adjacent lanes (all sectors of the row are consumed by one warp instruction).
#include "common.cuh"

namespace armnet {

template <int VEC>
struct VecT;
template <>
struct VecT<1> { using type = float; };
template <>
struct VecT<2> { using type = float2; };
template <>
struct VecT<4> { using type = float4; };

__device__ __forceinline__ float scale1(float a, float v) { return __fmul_rn(a, v); }
__device__ __forceinline__ float2 scale1(float2 a, float v) { return make_float2(__fmul_rn(a.x, v), __fmul_rn(a.y, v)); }
__device__ __forceinline__ float4 scale1(float4 a, float v) {
    return make_float4(__fmul_rn(a.x, v), __fmul_rn(a.y, v), __fmul_rn(a.z, v), __fmul_rn(a.w, v));
}

template <int VEC, bool I32>
__global__ void __launch_bounds__(256) embed_gather_kernel(const void *__restrict__ ids_, float *__restrict__ values,
                                                           const float *__restrict__ table, long long V, long long ld,
                                                           long long n_rows, int E, float *__restrict__ out, int clamp,
                                                           float lo, float hi, int inplace, int *__restrict__ err_flag) {
    using T = typename VecT<VEC>::type;
    const int epv = E / VEC;  // vector chunks per row
    const long long total = n_rows * epv;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x; j < total; j += stride) {
        const long long row = j / epv;
        const int c = (int)(j - row * epv);
        long long id = I32 ? (long long)reinterpret_cast<const int *>(ids_)[row]
                           : reinterpret_cast<const long long *>(ids_)[row];
        float v = values[row];
        if (clamp) {
            const float vc = fminf(fmaxf(v, lo), hi);
            if (inplace && c == 0 && vc != v) values[row] = vc;
            v = vc;
        }
        T r;
        if ((unsigned long long)id >= (unsigned long long)V) {  // reference raises IndexError (layers.py:20)
            if (err_flag && c == 0) atomicOr(err_flag, 1);
            r = T{};
        } else {
            r = scale1(__ldg(reinterpret_cast<const T *>(table + id * ld) + c), v);
        }
        reinterpret_cast<T *>(out)[j] = r;
    }
}

}  // namespace armnet

extern "C" int armnet_embed_gather_f32(const void *ids, int ids_i32, float *values, const float *table, int64_t V,
                                       int64_t ld, int64_t B, int F, int E, float *out, int clamp, float clamp_lo,
                                       float clamp_hi, int clamp_inplace, int *err_flag, void *stream) {
    using namespace armnet;
    if (B == 0) {
        note_launches(0);
        return ARMNET_OK;
    }
    if (!ids || !values || !table || !out) {
        set_error("embed_gather: null pointer");
        return ARMNET_ERR_NULL;
    }
    if (V <= 0 || B < 0 || F <= 0 || E <= 0 || ld < E) {
        set_error("embed_gather: bad shape V=%lld ld=%lld B=%lld F=%d E=%d", (long long)V, (long long)ld, (long long)B, F, E);
        return ARMNET_ERR_SHAPE;
    }
    note_launches(0);
    if (B == 0) return ARMNET_OK;
    int vec = 1;
    const uintptr_t al = (uintptr_t)table | (uintptr_t)out;
    if (E % 4 == 0 && ld % 4 == 0 && al % 16 == 0) vec = 4;
    else if (E % 2 == 0 && ld % 2 == 0 && al % 8 == 0) vec = 2;
    if ((uintptr_t)table % 4 || (uintptr_t)out % 4 || (uintptr_t)values % 4 || (uintptr_t)ids % (ids_i32 ? 4 : 8)) {
        set_error("embed_gather: misaligned pointer");
        return ARMNET_ERR_ALIGN;
    }
    DeviceInfo di;
    int rc = get_device_info(&di);
    if (rc != ARMNET_OK) return rc;
    const long long n_rows = (long long)B * F;
    const long long total = n_rows * (E / vec);
    long long blocks = (total + 255) / 256;
    const long long cap = (long long)di.sm_count * 8;  // 8 resident 256-thread CTAs per SM, grid-stride beyond that
    if (blocks > cap) blocks = cap;
    cudaStream_t st = (cudaStream_t)stream;
#define LAUNCH(VEC, I32)                                                                                       \
    embed_gather_kernel<VEC, I32><<<(unsigned)blocks, 256, 0, st>>>(ids, values, table, V, ld, n_rows, E, out, \
                                                                    clamp, clamp_lo, clamp_hi, clamp_inplace, err_flag)
    if (ids_i32) {
        if (vec == 4) LAUNCH(4, true); else if (vec == 2) LAUNCH(2, true); else LAUNCH(1, true);
    } else {
        if (vec == 4) LAUNCH(4, false); else if (vec == 2) LAUNCH(2, false); else LAUNCH(1, false);
    }
#undef LAUNCH
    ARMNET_CUDA_TRY(cudaGetLastError());
    note_launches(1);
    return ARMNET_OK;
}

// armnet_linear_gather_f32: layers.Linear.forward (models/layers.py:31-37), the 1-wide embedding of the LR term used by
// the zoo models (afm.py:40, xdfm.py:46,67):  y[b] = sum_f weight[ids[b,f]] * values[b,f] + bias.
// One warp per sample: lanes stride over the fields (coalesced id / value reads), shuffle tree for the sum.
namespace armnet {
template <bool I32>
__global__ void __launch_bounds__(256) linear_gather_kernel(const void *__restrict__ ids_, const float *__restrict__ values,
                                                            const float *__restrict__ weight, long long V, long long B,
                                                            int F, const float *__restrict__ bias, float *__restrict__ y,
                                                            int *__restrict__ err_flag) {
    const int lane = threadIdx.x & 31;
    const long long warps = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long b = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5; b < B; b += warps) {
        float acc = 0.f;
        for (int f = lane; f < F; f += 32) {
            const long long id = I32 ? (long long)reinterpret_cast<const int *>(ids_)[b * F + f]
                                     : reinterpret_cast<const long long *>(ids_)[b * F + f];
            if ((unsigned long long)id >= (unsigned long long)V) {
                if (err_flag) atomicOr(err_flag, 1);
            } else {
                acc = fmaf(__ldg(weight + id), __ldg(values + b * F + f), acc);
            }
        }
#pragma unroll
        for (int m = 16; m > 0; m >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, m);
        if (lane == 0) y[b] = acc + (bias ? __ldg(bias) : 0.f);
    }
}
}  // namespace armnet

extern "C" int armnet_linear_gather_f32(const void *ids, int ids_i32, const float *values, const float *weight, int64_t V,
                                        int64_t B, int F, const float *bias, float *y, int *err_flag, void *stream) {
    using namespace armnet;
    note_launches(0);
    if (B == 0) return ARMNET_OK;
    if (!ids || !values || !weight || !y) {
        set_error("linear_gather: null pointer");
        return ARMNET_ERR_NULL;
    }
    if (V <= 0 || B < 0 || F <= 0) {
        set_error("linear_gather: bad shape V=%lld B=%lld F=%d", (long long)V, (long long)B, F);
        return ARMNET_ERR_SHAPE;
    }
    DeviceInfo di;
    int rc = get_device_info(&di);
    if (rc != ARMNET_OK) return rc;
    long long blocks = (B * 32 + 255) / 256;
    const long long cap = (long long)di.sm_count * 8;
    if (blocks > cap) blocks = cap;
    if (ids_i32)
        linear_gather_kernel<true><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(ids, values, weight, V, B, F, bias, y, err_flag);
    else
        linear_gather_kernel<false><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(ids, values, weight, V, B, F, bias, y, err_flag);
    ARMNET_CUDA_TRY(cudaGetLastError());
    note_launches(1);
    return ARMNET_OK;
}
