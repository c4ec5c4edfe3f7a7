// Input-side gather kernels (HBM-bound, coalesced on the write side):
//   patchify   : NCHW image -> [B*g*g, kpad] patch matrix (K order = channel, ky, kx — the memory order of
//                conv1.weight.view(W, 3*P*P), transformer.py:461,602-604) + class-token rows (transformer.py:607-609)
//   text_embed : token-embedding gather + positional add (model.py:272-274) and the EOT argmax of
//                text_global_pool (transformer.py:654)
#include "common.cuh"
#include "internal.h"

namespace b200clip {

namespace {

template <typename T> __device__ __forceinline__ T cast_from_f(float v);
template <> __device__ __forceinline__ float cast_from_f<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 cast_from_f<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }
template <> __device__ __forceinline__ __half cast_from_f<__half>(float v) { return __float2half_rn(v); }
template <typename T> __device__ __forceinline__ float cast_to_f(T v);
template <> __device__ __forceinline__ float cast_to_f<float>(float v) { return v; }
template <> __device__ __forceinline__ float cast_to_f<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <> __device__ __forceinline__ float cast_to_f<__half>(__half v) { return __half2float(v); }
// value after a cast to the activation dtype (fp32 tables are cast at use in the 16-bit modes)
template <typename T> __device__ __forceinline__ float rt(float v) { return cast_to_f<T>(cast_from_f<T>(v)); }

// VEC elements (16 bytes) per thread when the patch size allows it, else 1.  PC / GC / KC > 0: patch size, grid and padded row
// length known at compile time (the CLIP geometries), so that the five integer divisions per 16-byte vector become
// multiply-shifts: the generic form was issue-bound (75 % of the issue slots at 51 % of the DRAM peak).
template <typename T, int VEC, int PC = 0, int GC = 0, int KC = 0>
__global__ void __launch_bounds__(256)
patchify_kernel(const T* __restrict__ image, T* __restrict__ patches, int batch, int S_, int P_, int g_, int kpad_, int cls_slot) {
    // cls_slot = 1: token layout, g*g+1 rows per image with an all-zero row in the class-token slot (row 0)
    const int P = PC > 0 ? PC : P_;
    const int g = GC > 0 ? GC : g_;
    const int kpad = KC > 0 ? KC : kpad_;
    const int S = PC > 0 ? PC * GC : S_;
    const int kreal = 3 * P * P;
    const int vec_per_row = kpad / VEC;
    const int rows_per_img = g * g + cls_slot;
    const int64_t total = static_cast<int64_t>(batch) * rows_per_img * vec_per_row;
    for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const int64_t row = i / vec_per_row;
        const int col = static_cast<int>(i - row * vec_per_row) * VEC;
        T* dst = patches + row * kpad + col;
        const int b = static_cast<int>(row / rows_per_img);
        const int pr = static_cast<int>(row - static_cast<int64_t>(b) * rows_per_img) - cls_slot;
        if (col >= kreal || pr < 0) {
#pragma unroll
            for (int e = 0; e < VEC; ++e) dst[e] = cast_from_f<T>(0.f);
            continue;
        }
        const int gy = pr / g, gx = pr - gy * g;
        const int c = col / (P * P);
        const int rem = col - c * P * P;
        const int ky = rem / P, kx = rem - ky * P;
        const T* src = image + ((static_cast<int64_t>(b) * 3 + c) * S + (gy * P + ky)) * S + gx * P + kx;
        if constexpr (VEC * sizeof(T) == 16) {
            *reinterpret_cast<uint4*>(dst) = __ldg(reinterpret_cast<const uint4*>(src));
        } else if constexpr (VEC * sizeof(T) == 4 && VEC == 2) {
            *reinterpret_cast<uint32_t*>(dst) = __ldg(reinterpret_cast<const uint32_t*>(src));
        } else {
            dst[0] = src[0];
        }
    }
}

// uint8 pixels -> normalised activations inside the im2col: ToTensor (x / 255) and Normalize ((x - mean) / std) of the
// reference's preprocessing (deps/open_clip/src/open_clip/transform.py:274-392, constants.py:1-2) with the same operation
// order in fp32, then ONE rounding to the activation dtype (= `.half()` / `.to(bfloat16)` of the evaluation scripts).  The
// host then uploads 1 byte per pixel instead of 2 or 4 and the normalised image never exists in memory.
struct Norm3 {
    float mean[3], std[3];
};
template <typename T, int VEC, int PC = 0, int GC = 0, int KC = 0>
__global__ void __launch_bounds__(256)
patchify_u8_kernel(const uint8_t* __restrict__ image, T* __restrict__ patches, int batch, int S_, int P_, int g_, int kpad_, int cls_slot,
                   const Norm3 nrm) {
    // The result depends on (pixel value, channel) only: every block tabulates the 3 x 256 values once with the reference's
    // operation order and IEEE divisions (two fp32 divisions per pixel made the straightforward form compute-bound).
    __shared__ T lut[3][256];
    for (int i = threadIdx.x; i < 3 * 256; i += blockDim.x) {
        const int c = i >> 8, v = i & 255;
        lut[c][v] = cast_from_f<T>(__fdiv_rn(__fsub_rn(__fdiv_rn(static_cast<float>(v), 255.0f), nrm.mean[c]), nrm.std[c]));
    }
    __syncthreads();
    const int P = PC > 0 ? PC : P_;
    const int g = GC > 0 ? GC : g_;
    const int kpad = KC > 0 ? KC : kpad_;
    const int S = PC > 0 ? PC * GC : S_;
    const int kreal = 3 * P * P;
    const int vec_per_row = kpad / VEC;
    const int rows_per_img = g * g + cls_slot;
    const int64_t total = static_cast<int64_t>(batch) * rows_per_img * vec_per_row;
    for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const int64_t row = i / vec_per_row;
        const int col = static_cast<int>(i - row * vec_per_row) * VEC;
        T* dst = patches + row * kpad + col;
        const int b = static_cast<int>(row / rows_per_img);
        const int pr = static_cast<int>(row - static_cast<int64_t>(b) * rows_per_img) - cls_slot;
        T vals[VEC];
        if (col >= kreal || pr < 0) {
#pragma unroll
            for (int e = 0; e < VEC; ++e) vals[e] = cast_from_f<T>(0.f);
        } else {
            const int gy = pr / g, gx = pr - gy * g;
            const int c = col / (P * P);
            const int rem = col - c * P * P;
            const int ky = rem / P, kx = rem - ky * P;
            const uint8_t* src = image + ((static_cast<int64_t>(b) * 3 + c) * S + (gy * P + ky)) * S + gx * P + kx;
            uint8_t px[VEC];
            if constexpr (VEC == 8) {
                const uint2 u = __ldg(reinterpret_cast<const uint2*>(src));
                const uint32_t w2[2] = {u.x, u.y};
#pragma unroll
                for (int e = 0; e < 8; ++e) px[e] = static_cast<uint8_t>(w2[e >> 2] >> (8 * (e & 3)));
            } else if constexpr (VEC == 2) {
                const unsigned short u = __ldg(reinterpret_cast<const unsigned short*>(src));
                px[0] = static_cast<uint8_t>(u & 0xff);
                px[1] = static_cast<uint8_t>(u >> 8);
            } else {
#pragma unroll
                for (int e = 0; e < VEC; ++e) px[e] = src[e];
            }
#pragma unroll
            for (int e = 0; e < VEC; ++e) vals[e] = lut[c][px[e]];
        }
        if constexpr (VEC * sizeof(T) == 16) {
            *reinterpret_cast<uint4*>(dst) = *reinterpret_cast<const uint4*>(vals);
        } else if constexpr (VEC * sizeof(T) == 4) {
            *reinterpret_cast<uint32_t*>(dst) = *reinterpret_cast<const uint32_t*>(vals);
        } else {
#pragma unroll
            for (int e = 0; e < VEC; ++e) dst[e] = vals[e];
        }
    }
}

template <typename T>
__global__ void cls_rows_kernel(const float* __restrict__ class_emb, const float* __restrict__ pos, T* __restrict__ x,
                                int batch, int L, int width) {
    const int64_t total = static_cast<int64_t>(batch) * width;
    for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const int b = static_cast<int>(i / width);
        const int c = static_cast<int>(i - static_cast<int64_t>(b) * width);
        x[static_cast<int64_t>(b) * L * width + c] = cast_from_f<T>(rt<T>(class_emb[c]) + rt<T>(pos[c]));
    }
}

// one warp per output row (t, l)
template <typename T>
__global__ void __launch_bounds__(256)
text_embed_kernel(const int64_t* __restrict__ text, int ctx, const float* __restrict__ tok_emb,
                  const float* __restrict__ pos_emb, T* __restrict__ x, int T_, int L, int width, int vocab) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t r = static_cast<int64_t>(blockIdx.x) * 8 + warp;
    if (r >= static_cast<int64_t>(T_) * L) return;
    const int t = static_cast<int>(r / L);
    const int l = static_cast<int>(r - static_cast<int64_t>(t) * L);
    int64_t id = text[static_cast<int64_t>(t) * ctx + l];
    id = id < 0 ? 0 : (id >= vocab ? vocab - 1 : id);
    const float* e = tok_emb + id * width;
    const float* p = pos_emb + static_cast<int64_t>(l) * width;
    T* o = x + r * width;
    for (int c = lane * 4; c < width; c += 128) {
        const float4 ev = __ldg(reinterpret_cast<const float4*>(e + c));
        const float4 pv = __ldg(reinterpret_cast<const float4*>(p + c));
        o[c + 0] = cast_from_f<T>(rt<T>(ev.x) + rt<T>(pv.x));
        o[c + 1] = cast_from_f<T>(rt<T>(ev.y) + rt<T>(pv.y));
        o[c + 2] = cast_from_f<T>(rt<T>(ev.z) + rt<T>(pv.z));
        o[c + 3] = cast_from_f<T>(rt<T>(ev.w) + rt<T>(pv.w));
    }
}

// eot[t] = first index of the maximum token id over the FULL context (text.argmax(dim=-1))
__global__ void eot_kernel(const int64_t* __restrict__ text, int ctx, int32_t* __restrict__ eot, int T_, int L) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int t = blockIdx.x * 8 + warp;
    if (t >= T_) return;
    long long best = -0x7fffffffffffffffLL - 1;
    int best_i = 0x7fffffff;
    for (int l = lane; l < ctx; l += 32) {
        const long long v = text[static_cast<int64_t>(t) * ctx + l];
        if (v > best) { best = v; best_i = l; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const long long ov = __shfl_xor_sync(0xffffffffu, best, o);
        const int oi = __shfl_xor_sync(0xffffffffu, best_i, o);
        if (ov > best || (ov == best && oi < best_i)) { best = ov; best_i = oi; }
    }
    if (lane == 0) eot[t] = best_i < L ? best_i : L - 1;  // host guarantees eot < L; clamp keeps a bad call in bounds
}

template <typename T>
int patchify_t(const void* image, void* patches, int batch, int S, int P, int kpad, const float* class_emb, const float* pos,
               void* x, int width, int cls_slot, cudaStream_t stream) {
    const int g = S / P;
    constexpr int V = 16 / sizeof(T);
    const bool vec_ok = (P % V == 0) && (S % V == 0) && (kpad % V == 0) &&
                        (reinterpret_cast<uintptr_t>(image) % 16 == 0) && (reinterpret_cast<uintptr_t>(patches) % 16 == 0);
    // 16-bit types with an even patch size that is not a multiple of 8 (ViT-L/14): two elements (4 bytes) per thread
    const bool vec2_ok = !vec_ok && sizeof(T) == 2 && (P % 2 == 0) && (S % 2 == 0) && (kpad % 2 == 0) &&
                         (reinterpret_cast<uintptr_t>(image) % 4 == 0) && (reinterpret_cast<uintptr_t>(patches) % 4 == 0);
    const int64_t total = static_cast<int64_t>(batch) * (g * g + cls_slot) * (vec_ok ? kpad / V : (vec2_ok ? kpad / 2 : kpad));
    int64_t blocks = (total + 255) / 256;
    if (blocks > static_cast<int64_t>(num_sms()) * 32) blocks = static_cast<int64_t>(num_sms()) * 32;
    const T* img = static_cast<const T*>(image);
    T* pat = static_cast<T*>(patches);
    const int nb = static_cast<int>(blocks);
    if (vec_ok && S == 224 && P == 32 && kpad == 3072)        // ViT-B/32
        patchify_kernel<T, V, 32, 7, 3072><<<nb, 256, 0, stream>>>(img, pat, batch, S, P, g, kpad, cls_slot);
    else if (vec_ok && S == 224 && P == 16 && kpad == 768)    // ViT-B/16
        patchify_kernel<T, V, 16, 14, 768><<<nb, 256, 0, stream>>>(img, pat, batch, S, P, g, kpad, cls_slot);
    else if (vec_ok)
        patchify_kernel<T, V><<<static_cast<int>(blocks), 256, 0, stream>>>(static_cast<const T*>(image), static_cast<T*>(patches),
                                                                            batch, S, P, g, kpad, cls_slot);
    else if (vec2_ok && S == 224 && P == 14 && kpad == 640)   // ViT-L/14
        patchify_kernel<T, 2, 14, 16, 640><<<nb, 256, 0, stream>>>(img, pat, batch, S, P, g, kpad, cls_slot);
    else if (vec2_ok)
        patchify_kernel<T, 2><<<nb, 256, 0, stream>>>(img, pat, batch, S, P, g, kpad, cls_slot);
    else
        patchify_kernel<T, 1><<<static_cast<int>(blocks), 256, 0, stream>>>(static_cast<const T*>(image), static_cast<T*>(patches),
                                                                            batch, S, P, g, kpad, cls_slot);
    B2C_LAUNCH_CHECK("patchify_kernel");
    if (x != nullptr) {
        const int64_t tot2 = static_cast<int64_t>(batch) * width;
        int blocks2 = static_cast<int>((tot2 + 255) / 256);
        if (blocks2 > num_sms() * 8) blocks2 = num_sms() * 8;
        cls_rows_kernel<T><<<blocks2, 256, 0, stream>>>(class_emb, pos, static_cast<T*>(x), batch, g * g + 1, width);
        B2C_LAUNCH_CHECK("cls_rows_kernel");
    }
    return 0;
}

template <typename T>
int patchify_u8_t(const uint8_t* image, const Norm3& nrm, void* patches, int batch, int S, int P, int kpad, const float* class_emb,
                  const float* pos, void* x, int width, int cls_slot, cudaStream_t stream) {
    const int g = S / P;
    // 8 pixels (one 8-byte load, one 16-byte store for the 16-bit dtypes) per thread when a patch row allows it, else 2
    const bool vec_ok = sizeof(T) == 2 && (P % 8 == 0) && (S % 8 == 0) && (kpad % 8 == 0) && (reinterpret_cast<uintptr_t>(image) % 8 == 0) &&
                        (reinterpret_cast<uintptr_t>(patches) % 16 == 0);
    const bool vec2_ok = !vec_ok && sizeof(T) == 2 && (P % 2 == 0) && (S % 2 == 0) && (kpad % 2 == 0) &&
                         (reinterpret_cast<uintptr_t>(image) % 2 == 0) && (reinterpret_cast<uintptr_t>(patches) % 4 == 0);
    const int64_t total = static_cast<int64_t>(batch) * (g * g + cls_slot) * (vec_ok ? kpad / 8 : (vec2_ok ? kpad / 2 : kpad));
    int64_t blocks = (total + 255) / 256;
    if (blocks > static_cast<int64_t>(num_sms()) * 16) blocks = static_cast<int64_t>(num_sms()) * 16;
    const int nb = static_cast<int>(blocks);
    T* pat = static_cast<T*>(patches);
    if (vec_ok && S == 224 && P == 32 && kpad == 3072)
        patchify_u8_kernel<T, 8, 32, 7, 3072><<<nb, 256, 0, stream>>>(image, pat, batch, S, P, g, kpad, cls_slot, nrm);
    else if (vec_ok && S == 224 && P == 16 && kpad == 768)
        patchify_u8_kernel<T, 8, 16, 14, 768><<<nb, 256, 0, stream>>>(image, pat, batch, S, P, g, kpad, cls_slot, nrm);
    else if (vec_ok)
        patchify_u8_kernel<T, 8><<<nb, 256, 0, stream>>>(image, pat, batch, S, P, g, kpad, cls_slot, nrm);
    else if (vec2_ok && S == 224 && P == 14 && kpad == 640)
        patchify_u8_kernel<T, 2, 14, 16, 640><<<nb, 256, 0, stream>>>(image, pat, batch, S, P, g, kpad, cls_slot, nrm);
    else if (vec2_ok)
        patchify_u8_kernel<T, 2><<<nb, 256, 0, stream>>>(image, pat, batch, S, P, g, kpad, cls_slot, nrm);
    else
        patchify_u8_kernel<T, 1><<<nb, 256, 0, stream>>>(image, pat, batch, S, P, g, kpad, cls_slot, nrm);
    B2C_LAUNCH_CHECK("patchify_u8_kernel");
    if (x != nullptr) {
        const int64_t tot2 = static_cast<int64_t>(batch) * width;
        int blocks2 = static_cast<int>((tot2 + 255) / 256);
        if (blocks2 > num_sms() * 8) blocks2 = num_sms() * 8;
        cls_rows_kernel<T><<<blocks2, 256, 0, stream>>>(class_emb, pos, static_cast<T*>(x), batch, g * g + 1, width);
        B2C_LAUNCH_CHECK("cls_rows_kernel");
    }
    return 0;
}

template <typename T>
int text_embed_t(const int64_t* text, int ctx, const float* tok_emb, const float* pos_emb, void* x, int32_t* eot, int T_,
                 int L, int width, int vocab, cudaStream_t stream) {
    const int64_t rows = static_cast<int64_t>(T_) * L;
    text_embed_kernel<T><<<static_cast<int>((rows + 7) / 8), 256, 0, stream>>>(text, ctx, tok_emb, pos_emb, static_cast<T*>(x),
                                                                               T_, L, width, vocab);
    B2C_LAUNCH_CHECK("text_embed_kernel");
    if (eot != nullptr) {
        eot_kernel<<<(T_ + 7) / 8, 256, 0, stream>>>(text, ctx, eot, T_, L);
        B2C_LAUNCH_CHECK("eot_kernel");
    }
    return 0;
}

}  // namespace

int patchify(int dtype, const void* image, void* patches, int batch, int image_size, int patch, int kpad,
             const float* class_emb, const float* pos, void* x, int width, cudaStream_t stream, int cls_slot) {
    B2C_CHECK_ARG(batch > 0 && image_size > 0 && patch > 0 && image_size % patch == 0,
                  "patchify: bad geometry batch=%d image=%d patch=%d", batch, image_size, patch);
    B2C_CHECK_ARG(kpad >= 3 * patch * patch, "patchify: kpad=%d smaller than 3*P*P=%d", kpad, 3 * patch * patch);
    B2C_CHECK_ARG(x == nullptr || (class_emb != nullptr && pos != nullptr), "patchify: class rows need class_emb and pos");
    switch (dtype) {
        case 0: return patchify_t<float>(image, patches, batch, image_size, patch, kpad, class_emb, pos, x, width, cls_slot, stream);
        case 1: return patchify_t<__nv_bfloat16>(image, patches, batch, image_size, patch, kpad, class_emb, pos, x, width, cls_slot, stream);
        case 2: return patchify_t<__half>(image, patches, batch, image_size, patch, kpad, class_emb, pos, x, width, cls_slot, stream);
    }
    set_last_error("patchify: unknown dtype %d", dtype);
    return -1;
}

int patchify_u8(int dtype, const uint8_t* image, const float* mean, const float* std, void* patches, int batch, int image_size, int patch,
                int kpad, const float* class_emb, const float* pos, void* x, int width, cudaStream_t stream, int cls_slot) {
    B2C_CHECK_ARG(image != nullptr && mean != nullptr && std != nullptr && patches != nullptr, "patchify_u8: null pointer");
    B2C_CHECK_ARG(batch > 0 && image_size > 0 && patch > 0 && image_size % patch == 0,
                  "patchify: bad geometry batch=%d image=%d patch=%d", batch, image_size, patch);
    B2C_CHECK_ARG(kpad >= 3 * patch * patch, "patchify: kpad=%d smaller than 3*P*P=%d", kpad, 3 * patch * patch);
    B2C_CHECK_ARG(x == nullptr || (class_emb != nullptr && pos != nullptr), "patchify: class rows need class_emb and pos");
    Norm3 nrm;
    for (int c = 0; c < 3; ++c) {
        B2C_CHECK_ARG(std[c] > 0.f, "patchify_u8: std[%d] must be positive", c);
        nrm.mean[c] = mean[c];
        nrm.std[c] = std[c];
    }
    switch (dtype) {
        case 0: return patchify_u8_t<float>(image, nrm, patches, batch, image_size, patch, kpad, class_emb, pos, x, width, cls_slot, stream);
        case 1: return patchify_u8_t<__nv_bfloat16>(image, nrm, patches, batch, image_size, patch, kpad, class_emb, pos, x, width, cls_slot, stream);
        case 2: return patchify_u8_t<__half>(image, nrm, patches, batch, image_size, patch, kpad, class_emb, pos, x, width, cls_slot, stream);
    }
    set_last_error("patchify_u8: unknown dtype %d", dtype);
    return -1;
}

int text_embed_v(int dtype, const int64_t* text, int ctx, const float* tok_emb, const float* pos_emb, void* x, int32_t* eot,
                 int T, int L, int width, int vocab, cudaStream_t stream) {
    B2C_CHECK_ARG(T > 0 && L > 0 && L <= ctx && width % 4 == 0, "text_embed: bad shape T=%d L=%d ctx=%d width=%d", T, L, ctx, width);
    switch (dtype) {
        case 0: return text_embed_t<float>(text, ctx, tok_emb, pos_emb, x, eot, T, L, width, vocab, stream);
        case 1: return text_embed_t<__nv_bfloat16>(text, ctx, tok_emb, pos_emb, x, eot, T, L, width, vocab, stream);
        case 2: return text_embed_t<__half>(text, ctx, tok_emb, pos_emb, x, eot, T, L, width, vocab, stream);
    }
    set_last_error("text_embed: unknown dtype %d", dtype);
    return -1;
}

int eot_argmax(const int64_t* text, int ctx, int32_t* eot, int T, cudaStream_t stream) {
    B2C_CHECK_ARG(T > 0 && ctx > 0, "eot_argmax: bad shape T=%d ctx=%d", T, ctx);
    eot_kernel<<<(T + 7) / 8, 256, 0, stream>>>(text, ctx, eot, T, ctx);
    B2C_LAUNCH_CHECK("eot_kernel");
    return 0;
}

int text_embed(int dtype, const int64_t* text, int ctx, const float* tok_emb, const float* pos_emb, void* x, int32_t* eot,
               int T, int L, int width, cudaStream_t stream) {
    return text_embed_v(dtype, text, ctx, tok_emb, pos_emb, x, eot, T, L, width, 0x7fffffff, stream);
}

}  // namespace b200clip
