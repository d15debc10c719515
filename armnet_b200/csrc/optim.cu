// Dense Adam over the flat parameter / gradient bucket of data-parallel training, fused with the averaging and the
// [-c, c] gradient clamp the reference applies through per-parameter hooks (train.py:62-65: optim.Adam(lr) over every
// parameter, p.register_hook(lambda g: g.clamp(-1, 1))).  One pass over (p, g, m, v): 16 bytes read + 12 written per
// element, HBM bound; the stock path is ~6 elementwise passes (div, clamp, lerp, mul/addcmul, sqrt/div/add, addcdiv).
#include "common.cuh"

namespace armnet {

struct AdamParams {
    float *p, *m, *v;
    const float *g;
    long long n;
    float grad_scale, clamp;              // g <- clamp(g * grad_scale, -clamp, clamp); clamp <= 0: no clamp
    float beta1, beta2, eps;
    float step_size, inv_bc2_sqrt;        // lr / (1 - beta1^t),  1 / sqrt(1 - beta2^t)
};

__device__ __forceinline__ void adam_one(float &p, float g, float &m, float &v, const AdamParams &A) {
    g *= A.grad_scale;
    if (A.clamp > 0.f) g = fminf(fmaxf(g, -A.clamp), A.clamp);
    m = fmaf(g - m, 1.f - A.beta1, m);                       // exp_avg.lerp_(grad, 1 - beta1)
    v = fmaf(v, A.beta2, (1.f - A.beta2) * g * g);           // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, 1 - beta2)
    const float denom = fmaf(sqrtf(v), A.inv_bc2_sqrt, A.eps);
    p = fmaf(-A.step_size, __fdiv_rn(m, denom), p);          // param.addcdiv_(exp_avg, denom, value=-step_size)
}

__global__ void __launch_bounds__(256) clamp_adam_kernel(const AdamParams A) {
    const long long n4 = A.n >> 2;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        float4 p = reinterpret_cast<float4 *>(A.p)[i];
        const float4 g = __ldg(reinterpret_cast<const float4 *>(A.g) + i);
        float4 m = reinterpret_cast<float4 *>(A.m)[i];
        float4 v = reinterpret_cast<float4 *>(A.v)[i];
        adam_one(p.x, g.x, m.x, v.x, A);
        adam_one(p.y, g.y, m.y, v.y, A);
        adam_one(p.z, g.z, m.z, v.z, A);
        adam_one(p.w, g.w, m.w, v.w, A);
        reinterpret_cast<float4 *>(A.p)[i] = p;
        reinterpret_cast<float4 *>(A.m)[i] = m;
        reinterpret_cast<float4 *>(A.v)[i] = v;
    }
    // tail (n % 4 elements)
    const long long i = (n4 << 2) + (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < A.n) adam_one(A.p[i], A.g[i], A.m[i], A.v[i], A);
}

}  // namespace armnet

using namespace armnet;

extern "C" int armnet_clamp_adam_f32(float *param, const float *grad, float *exp_avg, float *exp_avg_sq, int64_t n,
                                     float grad_scale, float clamp, float lr, float beta1, float beta2, float eps,
                                     int64_t step, void *stream) {
    if (!param || !grad || !exp_avg || !exp_avg_sq) {
        set_error("armnet_clamp_adam_f32: null pointer");
        return ARMNET_ERR_NULL;
    }
    if (n < 0 || step < 1) {
        set_error("armnet_clamp_adam_f32: n >= 0 and step >= 1 required (n=%lld step=%lld)", (long long)n, (long long)step);
        return ARMNET_ERR_SHAPE;
    }
    if (((uintptr_t)param | (uintptr_t)grad | (uintptr_t)exp_avg | (uintptr_t)exp_avg_sq) & 15) {
        set_error("armnet_clamp_adam_f32: buffers must be 16-byte aligned");
        return ARMNET_ERR_ALIGN;
    }
    if (n == 0) return ARMNET_OK;
    AdamParams A;
    A.p = param;
    A.g = grad;
    A.m = exp_avg;
    A.v = exp_avg_sq;
    A.n = n;
    A.grad_scale = grad_scale;
    A.clamp = clamp;
    A.beta1 = beta1;
    A.beta2 = beta2;
    A.eps = eps;
    // torch.optim.Adam (_single_tensor_adam): python-double bias corrections, rounded to fp32 at use
    const double bc1 = 1.0 - pow((double)beta1, (double)step);
    const double bc2 = 1.0 - pow((double)beta2, (double)step);
    A.step_size = (float)((double)lr / bc1);
    A.inv_bc2_sqrt = (float)(1.0 / sqrt(bc2));
    DeviceInfo di;
    int rc = get_device_info(&di);
    if (rc != ARMNET_OK) return rc;
    long long blocks = ((n >> 2) + 255) / 256;
    const long long cap = (long long)di.sm_count * 8;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    clamp_adam_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(A);
    ARMNET_CUDA_TRY(cudaGetLastError());
    note_launches(1);
    return ARMNET_OK;
}
