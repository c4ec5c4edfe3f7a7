// Zero-shot similarity stage: L2-normalise -> image x class-prompt logits -> top-k, fused in one kernel.
//
// Replaces xclip/zero_shot.py:42-60,103-109 (F.normalize of the image features, the tensordot with
// prompt_feat, argmax) and training/zero_shot.py:11-14 (topk(5)).  The contraction is 2*D*C = 0.35 MFLOP per
// image (D=512, C=345): HBM/latency-bound, so it runs on the FP32 pipe with fp32 accumulation — which
// is also what the fp32 index-parity gate needs.  Eight image rows per CTA are normalised into shared
// memory once; each warp then streams class rows of prompt_feat (L2-resident, 128-bit loads) against
// all eight images; logits land in shared memory, from where they are written out (coalesced) and
// scanned by a per-row warp-level top-k (k <= 8) that breaks ties towards the lower class index, like
// torch.argmax.  16-bit modes round the normalised features and the logits to the storage dtype at
// the same points the reference's bf16/fp16 tensors do.
//
// class_mean: prompt_feat[c] = normalize(mean_t normalize(txt_feat[c, t]))  (xclip/zero_shot.py:231-234).
#include "common.cuh"
#include "internal.h"

namespace b200clip {

namespace {

constexpr int kRows = 8;  // image rows per CTA == warps per CTA

template <typename T> __device__ __forceinline__ float ld_as_f(const T* p);
template <> __device__ __forceinline__ float ld_as_f<float>(const float* p) { return *p; }
template <> __device__ __forceinline__ float ld_as_f<__nv_bfloat16>(const __nv_bfloat16* p) { return __bfloat162float(*p); }
template <> __device__ __forceinline__ float ld_as_f<__half>(const __half* p) { return __half2float(*p); }
template <typename T> __device__ __forceinline__ float rnd(float v) {
    if constexpr (sizeof(T) == 2) return round16<T>(v);
    else return v;
}

template <typename T>
__global__ void __launch_bounds__(kRows * 32)
zeroshot_kernel(const T* __restrict__ img, const T* __restrict__ prompt, float* __restrict__ logits,
                int64_t* __restrict__ topk_idx, float* __restrict__ topk_val, int B, int C, int D, int k, int normalize_img,
                float logit_scale) {
    extern __shared__ float zs_smem[];
    float* s_img = zs_smem;               // [kRows][D]
    float* s_log = zs_smem + kRows * D;   // [kRows][C]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int row0 = blockIdx.x * kRows;

    // 1) load + normalise image row `warp`
    {
        const int r = row0 + warp;
        float sq = 0.f;
        if (r < B) {
            for (int d = lane; d < D; d += 32) {
                const float v = ld_as_f<T>(img + static_cast<int64_t>(r) * D + d);
                s_img[warp * D + d] = v;
                sq += v * v;
            }
        } else {
            for (int d = lane; d < D; d += 32) s_img[warp * D + d] = 0.f;
        }
        if (normalize_img) {
            float nrm = rnd<T>(sqrtf(warp_sum(sq)));
            const float denom = fmaxf(nrm, 1e-12f);
            __syncwarp();
            for (int d = lane; d < D; d += 32) s_img[warp * D + d] = rnd<T>(s_img[warp * D + d] / denom);
        }
    }
    __syncthreads();

    // 2) logits: warp-strided over classes, lanes over D
    for (int c = warp; c < C; c += kRows) {
        float acc[kRows];
#pragma unroll
        for (int r = 0; r < kRows; ++r) acc[r] = 0.f;
        const T* pr = prompt + static_cast<int64_t>(c) * D;
        for (int d = lane * 4; d < D; d += 128) {
            float p[4];
            if constexpr (sizeof(T) == 4) {
                const float4 v = __ldg(reinterpret_cast<const float4*>(pr + d));
                p[0] = v.x; p[1] = v.y; p[2] = v.z; p[3] = v.w;
            } else {
                const uint2 v = __ldg(reinterpret_cast<const uint2*>(pr + d));
                const float2 a = Half16<T>::unpack(v.x), b = Half16<T>::unpack(v.y);
                p[0] = a.x; p[1] = a.y; p[2] = b.x; p[3] = b.y;
            }
#pragma unroll
            for (int r = 0; r < kRows; ++r) {
                const float4 x = *reinterpret_cast<const float4*>(s_img + r * D + d);
                acc[r] = fmaf(x.x, p[0], acc[r]);
                acc[r] = fmaf(x.y, p[1], acc[r]);
                acc[r] = fmaf(x.z, p[2], acc[r]);
                acc[r] = fmaf(x.w, p[3], acc[r]);
            }
        }
#pragma unroll
        for (int r = 0; r < kRows; ++r) acc[r] = warp_sum(acc[r]);
        if (lane == 0) {
#pragma unroll
            for (int r = 0; r < kRows; ++r) s_log[r * C + c] = logit_scale * rnd<T>(acc[r]);
        }
    }
    __syncthreads();

    // 3) write logits (coalesced) and run the per-row warp top-k
    const int r = row0 + warp;
    if (r >= B) return;
    float* lr = s_log + warp * C;
    if (logits != nullptr)
        for (int c = lane; c < C; c += 32) logits[static_cast<int64_t>(r) * C + c] = lr[c];
    if (topk_idx == nullptr) return;
    for (int i = 0; i < k; ++i) {
        float best = -INFINITY;
        int best_c = 0x7fffffff;
        for (int c = lane; c < C; c += 32) {
            const float v = lr[c];
            if (v > best || (v == best && c < best_c)) { best = v; best_c = c; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, best, o);
            const int oc = __shfl_xor_sync(0xffffffffu, best_c, o);
            if (ov > best || (ov == best && oc < best_c)) { best = ov; best_c = oc; }
        }
        if (best_c == 0x7fffffff) best_c = 0;  // all remaining entries are -inf / NaN
        if (lane == 0) {
            topk_idx[static_cast<int64_t>(r) * k + i] = best_c;
            if (topk_val != nullptr) topk_val[static_cast<int64_t>(r) * k + i] = best;
            lr[best_c] = -INFINITY;
        }
        __syncwarp();
    }
}

// Per-row warp-level top-k over a materialised logit matrix x[B, C] (row pitch ldx): the 16-bit zero-shot path computes
// the logits with the tcgen05 GEMM (csrc/gemm_pair.cu) and finishes here.  One warp per row; the row (<= 1024 classes:
// 32 per lane) lives in registers, each of the k rounds is a lane-local scan + a 5-step shuffle arg-max; ties go to the
// lower class index like torch.argmax / torch.topk.  Optionally also emits the row as fp32 (`logits_out`, pitch C).
constexpr int kTopkWarps = 8;
template <typename T, int VPL>
__global__ void __launch_bounds__(kTopkWarps * 32)
topk_rows_kernel(const T* __restrict__ x, int64_t ldx, int B, int C, int k, int64_t* __restrict__ topk_idx,
                 float* __restrict__ topk_val, float* __restrict__ logits_out) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int r = blockIdx.x * kTopkWarps + warp;
    if (r >= B) return;
    const T* xr = x + static_cast<int64_t>(r) * ldx;
    float v[VPL];
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
        const int c = lane + i * 32;
        v[i] = c < C ? ld_as_f<T>(xr + c) : -INFINITY;
        if (logits_out != nullptr && c < C) logits_out[static_cast<int64_t>(r) * C + c] = v[i];
    }
    if (topk_idx == nullptr) return;
    for (int it = 0; it < k; ++it) {
        float best = -INFINITY;
        int best_c = 0x7fffffff;
#pragma unroll
        for (int i = 0; i < VPL; ++i) {
            const int c = lane + i * 32;
            if (c < C && (v[i] > best || (v[i] == best && c < best_c))) {
                best = v[i];
                best_c = c;
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, best, o);
            const int oc = __shfl_xor_sync(0xffffffffu, best_c, o);
            if (ov > best || (ov == best && oc < best_c)) {
                best = ov;
                best_c = oc;
            }
        }
        if (best_c == 0x7fffffff) best_c = 0;  // all remaining entries are -inf / NaN
        if (lane == 0) {
            topk_idx[static_cast<int64_t>(r) * k + it] = best_c;
            if (topk_val != nullptr) topk_val[static_cast<int64_t>(r) * k + it] = best;
        }
        // the winning lane retires its element (c = lane + i*32  ->  i = c / 32)
#pragma unroll
        for (int i = 0; i < VPL; ++i)
            if (lane + i * 32 == best_c) v[i] = -INFINITY;
    }
}

template <typename T>
int topk_rows_t(const void* x, int64_t ldx, int B, int C, int k, int64_t* idx, float* val, float* logits_out, cudaStream_t stream) {
    const int vpl = (C + 31) / 32;
    const int grid = (B + kTopkWarps - 1) / kTopkWarps;
    const T* xp = static_cast<const T*>(x);
#define TK_CASE(V) \
    topk_rows_kernel<T, V><<<grid, kTopkWarps * 32, 0, stream>>>(xp, ldx, B, C, k, idx, val, logits_out)
    if (vpl <= 4) TK_CASE(4);
    else if (vpl <= 8) TK_CASE(8);
    else if (vpl <= 12) TK_CASE(12);
    else if (vpl <= 16) TK_CASE(16);
    else if (vpl <= 32) TK_CASE(32);
    else {
        set_last_error("topk_rows: C=%d exceeds the 1024 classes of the register-resident path", C);
        return -1;
    }
#undef TK_CASE
    B2C_LAUNCH_CHECK("topk_rows_kernel");
    return 0;
}

// one CTA (128 threads) per class
template <typename T>
__global__ void __launch_bounds__(128)
class_mean_kernel(const T* __restrict__ txt, T* __restrict__ out, int templates, int D) {
    extern __shared__ float cm_smem[];
    float* s_den = cm_smem;  // [templates]
    __shared__ float s_red[4];
    const int c = blockIdx.x;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const T* base = txt + static_cast<int64_t>(c) * templates * D;
    for (int t = warp; t < templates; t += 4) {
        float sq = 0.f;
        for (int d = lane; d < D; d += 32) {
            const float v = ld_as_f<T>(base + static_cast<int64_t>(t) * D + d);
            sq += v * v;
        }
        sq = warp_sum(sq);
        if (lane == 0) s_den[t] = fmaxf(rnd<T>(sqrtf(sq)), 1e-12f);
    }
    __syncthreads();
    float sq_local = 0.f;
    // each thread owns columns d = tid, tid+128, ...; keep the means in registers (D <= 128*8)
    float mean[8];
    int nd = 0;
    for (int d = threadIdx.x; d < D; d += 128, ++nd) {
        float s = 0.f;
        for (int t = 0; t < templates; ++t) s += rnd<T>(ld_as_f<T>(base + static_cast<int64_t>(t) * D + d) / s_den[t]);
        const float m = rnd<T>(s / static_cast<float>(templates));
        mean[nd] = m;
        sq_local += m * m;
    }
    sq_local = warp_sum(sq_local);
    if (lane == 0) s_red[warp] = sq_local;
    __syncthreads();
    const float denom = fmaxf(rnd<T>(sqrtf(s_red[0] + s_red[1] + s_red[2] + s_red[3])), 1e-12f);
    nd = 0;
    for (int d = threadIdx.x; d < D; d += 128, ++nd) {
        const float v = mean[nd] / denom;
        if constexpr (sizeof(T) == 4) out[static_cast<int64_t>(c) * D + d] = v;
        else out[static_cast<int64_t>(c) * D + d] = Half16<T>::from_f(v);
    }
}

template <typename T>
int zeroshot_t(const void* img, const void* prompt, float* logits, int64_t* topk_idx, float* topk_val, int B, int C, int D, int k,
               int normalize_img, float logit_scale, cudaStream_t stream) {
    const size_t smem = static_cast<size_t>(kRows) * (D + C) * sizeof(float);
    B2C_CHECK_ARG(smem <= 200 * 1024, "zeroshot: C=%d, D=%d need %zu B of shared memory (> 200 KiB)", C, D, smem);
    auto kern = zeroshot_kernel<T>;
    if (smem > 48 * 1024) B2C_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    kern<<<(B + kRows - 1) / kRows, kRows * 32, smem, stream>>>(static_cast<const T*>(img), static_cast<const T*>(prompt), logits,
                                                                topk_idx, topk_val, B, C, D, k, normalize_img, logit_scale);
    B2C_LAUNCH_CHECK("zeroshot_kernel");
    return 0;
}

template <typename T>
int class_mean_t(const void* txt, void* out, int classes, int templates, int D, cudaStream_t stream) {
    class_mean_kernel<T><<<classes, 128, templates * sizeof(float), stream>>>(static_cast<const T*>(txt), static_cast<T*>(out),
                                                                             templates, D);
    B2C_LAUNCH_CHECK("class_mean_kernel");
    return 0;
}

}  // namespace

int topk_rows(int dtype, const void* x, int64_t ldx, int B, int C, int k, int64_t* topk_idx, float* topk_val, float* logits_out,
              cudaStream_t stream) {
    B2C_CHECK_ARG(x != nullptr && B > 0 && C > 0 && ldx >= C, "topk_rows: bad input B=%d C=%d ld=%lld", B, C, (long long)ldx);
    B2C_CHECK_ARG(k >= 0 && k <= 8 && k <= C, "topk_rows: k=%d must be in [0, min(8, C)]", k);
    B2C_CHECK_ARG(k == 0 || topk_idx != nullptr, "topk_rows: k > 0 needs topk_idx");
    B2C_CHECK_ARG(topk_idx != nullptr || logits_out != nullptr, "topk_rows: nothing to write");
    if (k == 0) topk_idx = nullptr;
    switch (dtype) {
        case 0: return topk_rows_t<float>(x, ldx, B, C, k, topk_idx, topk_val, logits_out, stream);
        case 1: return topk_rows_t<__nv_bfloat16>(x, ldx, B, C, k, topk_idx, topk_val, logits_out, stream);
        case 2: return topk_rows_t<__half>(x, ldx, B, C, k, topk_idx, topk_val, logits_out, stream);
    }
    set_last_error("topk_rows: unknown dtype %d", dtype);
    return -1;
}

int zeroshot(int dtype, const void* img_feat, const void* prompt_feat, float* logits, int64_t* topk_idx, float* topk_val, int B,
             int C, int D, int k, int normalize_img, float logit_scale, cudaStream_t stream) {
    B2C_CHECK_ARG(B > 0 && C > 0 && D > 0, "zeroshot: empty input B=%d C=%d D=%d", B, C, D);
    B2C_CHECK_ARG(D % 4 == 0, "zeroshot: D=%d must be a multiple of 4", D);
    B2C_CHECK_ARG(k >= 0 && k <= 8 && k <= C, "zeroshot: k=%d must be in [0, min(8, C)]", k);
    B2C_CHECK_ARG(topk_idx != nullptr || k == 0 || logits != nullptr, "zeroshot: nothing to write");
    B2C_CHECK_ARG((reinterpret_cast<uintptr_t>(img_feat) | reinterpret_cast<uintptr_t>(prompt_feat)) % 16 == 0,
                  "zeroshot: pointers must be 16-byte aligned");
    if (k == 0) topk_idx = nullptr;
    switch (dtype) {
        case 0: return zeroshot_t<float>(img_feat, prompt_feat, logits, topk_idx, topk_val, B, C, D, k, normalize_img, logit_scale, stream);
        case 1: return zeroshot_t<__nv_bfloat16>(img_feat, prompt_feat, logits, topk_idx, topk_val, B, C, D, k, normalize_img, logit_scale, stream);
        case 2: return zeroshot_t<__half>(img_feat, prompt_feat, logits, topk_idx, topk_val, B, C, D, k, normalize_img, logit_scale, stream);
    }
    set_last_error("zeroshot: unknown dtype %d", dtype);
    return -1;
}

int class_mean(int dtype, const void* txt_feat, void* prompt_feat, int classes, int templates, int D, cudaStream_t stream) {
    B2C_CHECK_ARG(classes > 0 && templates > 0 && D > 0 && D <= 1024, "class_mean: bad shape classes=%d templates=%d D=%d", classes,
                  templates, D);
    switch (dtype) {
        case 0: return class_mean_t<float>(txt_feat, prompt_feat, classes, templates, D, stream);
        case 1: return class_mean_t<__nv_bfloat16>(txt_feat, prompt_feat, classes, templates, D, stream);
        case 2: return class_mean_t<__half>(txt_feat, prompt_feat, classes, templates, D, stream);
    }
    set_last_error("class_mean: unknown dtype %d", dtype);
    return -1;
}

}  // namespace b200clip
