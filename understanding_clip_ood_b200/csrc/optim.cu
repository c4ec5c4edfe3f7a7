// Fused multi-tensor AdamW: ONE launch updates every parameter of a group (torch.optim.AdamW semantics — decoupled weight
// decay, bias-corrected moments — as the reference's training loop configures it, deps/open_clip/src/training/main.py:299-326).
// The host hands over a device table of (param, grad, exp_avg, exp_avg_sq, count, dtypes) entries and a device list of
// (entry, offset) chunks, one chunk of kChunk elements per thread block.  Moments are fp32; parameters / gradients may be
// fp32, bf16 or fp16 (precision='bf16' keeps 16-bit weights like the reference does).
#include "../../include/b200clip.h"
#include "common.cuh"
#include "internal.h"

namespace b200clip {

namespace {

constexpr int kChunk = 4096;

__device__ __forceinline__ float ld_any(const void* p, int dtype, int64_t i) {
    if (dtype == 0) return static_cast<const float*>(p)[i];
    if (dtype == 1) return __bfloat162float(static_cast<const __nv_bfloat16*>(p)[i]);
    return __half2float(static_cast<const __half*>(p)[i]);
}
__device__ __forceinline__ void st_any(void* p, int dtype, int64_t i, float v) {
    if (dtype == 0) static_cast<float*>(p)[i] = v;
    else if (dtype == 1) static_cast<__nv_bfloat16*>(p)[i] = __float2bfloat16_rn(v);
    else static_cast<__half*>(p)[i] = __float2half_rn(v);
}

__global__ void __launch_bounds__(256)
adamw_kernel(const b200clip_adamw_tensor* __restrict__ items, const int32_t* __restrict__ chunk_item, const int64_t* __restrict__ chunk_off,
             float lr, float beta1, float beta2, float eps, float weight_decay, float bias_c1, float bias_c2_sqrt, float grad_scale) {
    const b200clip_adamw_tensor it = items[chunk_item[blockIdx.x]];
    const int64_t begin = chunk_off[blockIdx.x];
    const int64_t end = min(it.count, begin + kChunk);
    const float decay = 1.0f - lr * weight_decay;
    const float step = lr / bias_c1;
    for (int64_t i = begin + threadIdx.x; i < end; i += 256) {
        const float g = ld_any(it.grad, it.grad_dtype, i) * grad_scale;
        float p = ld_any(it.param, it.param_dtype, i) * decay;
        const float m = fmaf(beta1, it.exp_avg[i], (1.0f - beta1) * g);
        const float v = fmaf(beta2, it.exp_avg_sq[i], (1.0f - beta2) * g * g);
        it.exp_avg[i] = m;
        it.exp_avg_sq[i] = v;
        p -= step * m / (sqrtf(v) / bias_c2_sqrt + eps);
        st_any(it.param, it.param_dtype, i, p);
    }
}

}  // namespace

int adamw_step(const b200clip_adamw_tensor* items, const int32_t* chunk_item, const int64_t* chunk_off, int chunks, float lr, float beta1,
               float beta2, float eps, float weight_decay, int step, float grad_scale, cudaStream_t stream) {
    B2C_CHECK_ARG(items && chunk_item && chunk_off && chunks > 0 && step >= 1, "adamw_step: bad arguments");
    const float bc1 = 1.0f - powf(beta1, static_cast<float>(step));
    const float bc2 = sqrtf(1.0f - powf(beta2, static_cast<float>(step)));
    adamw_kernel<<<chunks, 256, 0, stream>>>(items, chunk_item, chunk_off, lr, beta1, beta2, eps, weight_decay, bc1, bc2, grad_scale);
    B2C_LAUNCH_CHECK("adamw_kernel");
    return 0;
}

}  // namespace b200clip
