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

// four consecutive elements of a parameter / gradient of any of the three dtypes (16-byte or 8-byte access)
__device__ __forceinline__ void ld4_any(const void* p, int dtype, int64_t i, float (&v)[4]) {
    if (dtype == 0) {
        const float4 f = *reinterpret_cast<const float4*>(static_cast<const float*>(p) + i);
        v[0] = f.x; v[1] = f.y; v[2] = f.z; v[3] = f.w;
    } else {
        const uint2 u = *reinterpret_cast<const uint2*>(static_cast<const uint16_t*>(p) + i);
        if (dtype == 1) {
            v[0] = __uint_as_float(u.x << 16); v[1] = __uint_as_float(u.x & 0xFFFF0000u);
            v[2] = __uint_as_float(u.y << 16); v[3] = __uint_as_float(u.y & 0xFFFF0000u);
        } else {
            const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&u.x)), b = __half22float2(*reinterpret_cast<const __half2*>(&u.y));
            v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
        }
    }
}
__device__ __forceinline__ void st4_any(void* p, int dtype, int64_t i, const float (&v)[4]) {
    if (dtype == 0) {
        *reinterpret_cast<float4*>(static_cast<float*>(p) + i) = make_float4(v[0], v[1], v[2], v[3]);
    } else if (dtype == 1) {
        const __nv_bfloat162 a = __floats2bfloat162_rn(v[0], v[1]), b = __floats2bfloat162_rn(v[2], v[3]);
        *reinterpret_cast<uint2*>(static_cast<uint16_t*>(p) + i) = make_uint2(*reinterpret_cast<const uint32_t*>(&a), *reinterpret_cast<const uint32_t*>(&b));
    } else {
        const __half2 a = __floats2half2_rn(v[0], v[1]), b = __floats2half2_rn(v[2], v[3]);
        *reinterpret_cast<uint2*>(static_cast<uint16_t*>(p) + i) = make_uint2(*reinterpret_cast<const uint32_t*>(&a), *reinterpret_cast<const uint32_t*>(&b));
    }
}

__device__ __forceinline__ void adamw_update(float& p, float& m, float& v, float g, float decay, float step, float beta1, float beta2, float eps,
                                             float bias_c2_sqrt) {
    p *= decay;
    m = fmaf(beta1, m, (1.0f - beta1) * g);
    v = fmaf(beta2, v, (1.0f - beta2) * g * g);
    p -= step * m / (sqrtf(v) / bias_c2_sqrt + eps);
}

__global__ void __launch_bounds__(256)
adamw_kernel(const b200clip_adamw_tensor* __restrict__ items, const int32_t* __restrict__ chunk_item, const int64_t* __restrict__ chunk_off,
             float lr, float beta1, float beta2, float eps, float weight_decay, float bias_c1, float bias_c2_sqrt, float grad_scale) {
    const b200clip_adamw_tensor it = items[chunk_item[blockIdx.x]];
    const int64_t begin = chunk_off[blockIdx.x];
    const int64_t end = min(it.count, begin + kChunk);
    const float decay = 1.0f - lr * weight_decay;
    const float step = lr / bias_c1;
    // 16-byte accesses where the tensors allow them (chunks start at multiples of 4096 elements; the bases must be aligned)
    const bool vec = ((reinterpret_cast<uintptr_t>(it.param) | reinterpret_cast<uintptr_t>(it.grad) | reinterpret_cast<uintptr_t>(it.exp_avg) |
                       reinterpret_cast<uintptr_t>(it.exp_avg_sq)) & 15) == 0;
    int64_t i = begin;
    if (vec) {
        const int64_t end4 = begin + (end - begin) / 4 * 4;
        for (i = begin + threadIdx.x * 4; i < end4; i += 256 * 4) {
            float g[4], p[4];
            ld4_any(it.grad, it.grad_dtype, i, g);
            ld4_any(it.param, it.param_dtype, i, p);
            float4 m4 = *reinterpret_cast<const float4*>(it.exp_avg + i), v4 = *reinterpret_cast<const float4*>(it.exp_avg_sq + i);
            float m[4] = {m4.x, m4.y, m4.z, m4.w}, v[4] = {v4.x, v4.y, v4.z, v4.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) adamw_update(p[j], m[j], v[j], g[j] * grad_scale, decay, step, beta1, beta2, eps, bias_c2_sqrt);
            *reinterpret_cast<float4*>(it.exp_avg + i) = make_float4(m[0], m[1], m[2], m[3]);
            *reinterpret_cast<float4*>(it.exp_avg_sq + i) = make_float4(v[0], v[1], v[2], v[3]);
            st4_any(it.param, it.param_dtype, i, p);
        }
        i = end4;
    }
    for (i += threadIdx.x; i < end; i += 256) {
        const float g = ld_any(it.grad, it.grad_dtype, i) * grad_scale;
        float p = ld_any(it.param, it.param_dtype, i);
        float m = it.exp_avg[i], v = it.exp_avg_sq[i];
        adamw_update(p, m, v, g, decay, step, beta1, beta2, eps, bias_c2_sqrt);
        it.exp_avg[i] = m;
        it.exp_avg_sq[i] = v;
        st_any(it.param, it.param_dtype, i, p);
    }
}

// multi-tensor conversion: tensors of a list converted (fp32 <-> 16-bit, or copied) by ONE launch; chunks as in adamw_kernel
__global__ void __launch_bounds__(256)
multi_cast_kernel(const b200clip_cast_tensor* __restrict__ items, const int32_t* __restrict__ chunk_item, const int64_t* __restrict__ chunk_off) {
    const b200clip_cast_tensor it = items[chunk_item[blockIdx.x]];
    const int64_t begin = chunk_off[blockIdx.x];
    const int64_t end = min(it.count, begin + kChunk);
    const bool vec = ((reinterpret_cast<uintptr_t>(it.src) | reinterpret_cast<uintptr_t>(it.dst)) & 15) == 0;
    int64_t i = begin;
    if (vec) {
        const int64_t end4 = begin + (end - begin) / 4 * 4;
        for (i = begin + threadIdx.x * 4; i < end4; i += 256 * 4) {
            float v[4];
            ld4_any(it.src, it.src_dtype, i, v);
            st4_any(it.dst, it.dst_dtype, i, v);
        }
        i = end4;
    }
    for (i += threadIdx.x; i < end; i += 256) st_any(it.dst, it.dst_dtype, i, ld_any(it.src, it.src_dtype, i));
}

}  // namespace

int multi_cast(const b200clip_cast_tensor* items, const int32_t* chunk_item, const int64_t* chunk_off, int chunks, cudaStream_t stream) {
    B2C_CHECK_ARG(items && chunk_item && chunk_off && chunks > 0, "multi_cast: bad arguments");
    multi_cast_kernel<<<chunks, 256, 0, stream>>>(items, chunk_item, chunk_off);
    B2C_LAUNCH_CHECK("multi_cast_kernel");
    return 0;
}

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
