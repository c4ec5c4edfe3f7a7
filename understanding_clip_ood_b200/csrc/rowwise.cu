// HBM-bound row-wise kernels: LayerNorm (with row gather for CLS / EOT pooling) and L2 normalise.
// One warp per row, 128-bit loads/stores, the row stays in registers between the statistics pass
// and the write (algorithmic bytes = one read + one write of the row).
//   LayerNorm / LayerNormFp32: deps/open_clip/src/open_clip/transformer.py:15-30 (fp32 statistics on the
//   up-cast input, fp32 gamma/beta, cast back to the activation dtype).
//   F.normalize: model.py:267,284 and xclip/zero_shot.py:34,50 (x / max(||x||, eps)).
#include "common.cuh"
#include "internal.h"

namespace b200clip {

namespace {

template <typename T> struct Vec16 {
    static constexpr int kElems = 16 / sizeof(T);
};

template <typename T> __device__ __forceinline__ void load_vec(const T* p, float (&f)[Vec16<T>::kElems]);
template <> __device__ __forceinline__ void load_vec<float>(const float* p, float (&f)[4]) {
    const float4 v = *reinterpret_cast<const float4*>(p);
    f[0] = v.x; f[1] = v.y; f[2] = v.z; f[3] = v.w;
}
template <> __device__ __forceinline__ void load_vec<__nv_bfloat16>(const __nv_bfloat16* p, float (&f)[8]) {
    const uint4 v = *reinterpret_cast<const uint4*>(p);
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float2 t = Half16<__nv_bfloat16>::unpack(w[i]);
        f[2 * i] = t.x; f[2 * i + 1] = t.y;
    }
}
template <> __device__ __forceinline__ void load_vec<__half>(const __half* p, float (&f)[8]) {
    const uint4 v = *reinterpret_cast<const uint4*>(p);
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float2 t = Half16<__half>::unpack(w[i]);
        f[2 * i] = t.x; f[2 * i + 1] = t.y;
    }
}
template <typename T> __device__ __forceinline__ void store_vec(T* p, const float (&f)[Vec16<T>::kElems]);
template <> __device__ __forceinline__ void store_vec<float>(float* p, const float (&f)[4]) {
    *reinterpret_cast<float4*>(p) = make_float4(f[0], f[1], f[2], f[3]);
}
template <> __device__ __forceinline__ void store_vec<__nv_bfloat16>(__nv_bfloat16* p, const float (&f)[8]) {
    using H = Half16<__nv_bfloat16>;
    *reinterpret_cast<uint4*>(p) = make_uint4(H::pack(f[0], f[1]), H::pack(f[2], f[3]), H::pack(f[4], f[5]), H::pack(f[6], f[7]));
}
template <> __device__ __forceinline__ void store_vec<__half>(__half* p, const float (&f)[8]) {
    using H = Half16<__half>;
    *reinterpret_cast<uint4*>(p) = make_uint4(H::pack(f[0], f[1]), H::pack(f[2], f[3]), H::pack(f[4], f[5]), H::pack(f[6], f[7]));
}

constexpr int kWarpsPerBlock = 8;

// VPL = 16-byte vectors per lane held in registers
template <typename T, int VPL>
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
layernorm_kernel(const T* x, int64_t ldx, const float* __restrict__ gamma, const float* __restrict__ beta,
                 T* y, int64_t ldy, int rows, int width, float eps, int row_stride_rows,
                 const int32_t* __restrict__ row_idx) {
    constexpr int E = Vec16<T>::kElems;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int r = blockIdx.x * kWarpsPerBlock + warp;
    if (r >= rows) return;
    int64_t in_row = static_cast<int64_t>(r) * row_stride_rows;
    if (row_idx != nullptr) in_row += row_idx[r];
    const T* xr = x + in_row * ldx;
    T* yr = y + static_cast<int64_t>(r) * ldy;
    const int nvec = width / E;

    float v[VPL][E];
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
        const int vi = lane + i * 32;
        if (vi < nvec) {
            load_vec<T>(xr + vi * E, v[i]);
#pragma unroll
            for (int e = 0; e < E; ++e) sum += v[i][e];
        } else {
#pragma unroll
            for (int e = 0; e < E; ++e) v[i][e] = 0.f;
        }
    }
    const float mean = warp_sum(sum) / static_cast<float>(width);
    float sq = 0.f;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
        const int vi = lane + i * 32;
        if (vi < nvec) {
#pragma unroll
            for (int e = 0; e < E; ++e) {
                const float d = v[i][e] - mean;
                sq += d * d;
            }
        }
    }
    const float rstd = rsqrtf(warp_sum(sq) / static_cast<float>(width) + eps);
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
        const int vi = lane + i * 32;
        if (vi < nvec) {
            float o[E];
#pragma unroll
            for (int e4 = 0; e4 < E; e4 += 4) {
                const float4 g = __ldg(reinterpret_cast<const float4*>(gamma + vi * E + e4));
                const float4 b = __ldg(reinterpret_cast<const float4*>(beta + vi * E + e4));
                o[e4 + 0] = (v[i][e4 + 0] - mean) * rstd * g.x + b.x;
                o[e4 + 1] = (v[i][e4 + 1] - mean) * rstd * g.y + b.y;
                o[e4 + 2] = (v[i][e4 + 2] - mean) * rstd * g.z + b.z;
                o[e4 + 3] = (v[i][e4 + 3] - mean) * rstd * g.w + b.w;
            }
            store_vec<T>(yr + vi * E, o);
        }
    }
}

// Per-row LayerNorm statistics only: stats[r] = (mean, rstd) in fp32.  Used when the LayerNorm itself is folded into the
// following GEMM (csrc/gemm_pair.cu, "LN-fold" epilogues): the GEMM then reads the un-normalised residual stream directly
// and the normalised activations are never written to / re-read from HBM.  Read-only stream: one pass over x.
template <typename T, int VPL>
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
row_stats_kernel(const T* __restrict__ x, int64_t ldx, float2* __restrict__ stats, int rows, int width, float eps) {
    constexpr int E = Vec16<T>::kElems;
    constexpr int R = VPL <= 4 ? 2 : 1;  // rows per warp, all of their loads in flight before the first reduction
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int r0 = (blockIdx.x * kWarpsPerBlock + warp) * R;
    pdl_launch_dependents();
    pdl_wait();
    if (r0 >= rows) return;
    const int nvec = width / E;
    float v[R][VPL][E];
#pragma unroll
    for (int j = 0; j < R; ++j) {
        const int r = min(r0 + j, rows - 1);
        const T* xr = x + static_cast<int64_t>(r) * ldx;
#pragma unroll
        for (int i = 0; i < VPL; ++i) {
            const int vi = lane + i * 32;
            if (vi < nvec) {
                load_vec<T>(xr + vi * E, v[j][i]);
            } else {
#pragma unroll
                for (int e = 0; e < E; ++e) v[j][i][e] = 0.f;
            }
        }
    }
    const float inv_w = 1.0f / static_cast<float>(width);
    // packed fp32x2 arithmetic (FADD2 / FFMA2): the kernel was issue-bound (68 % issue slots for 54 % of the DRAM peak)
#pragma unroll
    for (int j = 0; j < R; ++j) {
        uint64_t s2 = 0ull;
#pragma unroll
        for (int i = 0; i < VPL; ++i)
#pragma unroll
            for (int e = 0; e < E; e += 2) s2 = add_f2(s2, pack_f2(v[j][i][e], v[j][i][e + 1]));
        float s_lo, s_hi;
        unpack_f2(s2, s_lo, s_hi);
        const float mean = warp_sum(s_lo + s_hi) * inv_w;
        const uint64_t nmean2 = pack_f2(-mean, -mean);
        uint64_t q2 = 0ull;
#pragma unroll
        for (int i = 0; i < VPL; ++i) {
            if (lane + i * 32 < nvec) {
#pragma unroll
                for (int e = 0; e < E; e += 2) {
                    const uint64_t d2 = add_f2(pack_f2(v[j][i][e], v[j][i][e + 1]), nmean2);
                    q2 = fma_f2(d2, d2, q2);
                }
            }
        }
        float q_lo, q_hi;
        unpack_f2(q2, q_lo, q_hi);
        const float sq = q_lo + q_hi;
        const float rstd = rsqrtf(warp_sum(sq) * inv_w + eps);
        if (lane == 0 && r0 + j < rows) stats[r0 + j] = make_float2(mean, rstd);
    }
}

template <typename T, int VPL>
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
normalize_kernel(const T* x, int64_t ldx, T* y, int64_t ldy, int rows, int dim, float eps) {
    constexpr int E = Vec16<T>::kElems;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int r = blockIdx.x * kWarpsPerBlock + warp;
    if (r >= rows) return;
    const T* xr = x + static_cast<int64_t>(r) * ldx;
    T* yr = y + static_cast<int64_t>(r) * ldy;
    const int nvec = dim / E;
    float v[VPL][E];
    float sq = 0.f;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
        const int vi = lane + i * 32;
        if (vi < nvec) {
            load_vec<T>(xr + vi * E, v[i]);
#pragma unroll
            for (int e = 0; e < E; ++e) sq += v[i][e] * v[i][e];
        }
    }
    // F.normalize = x / x.norm().clamp_min(eps): the norm is accumulated in fp32 and, in the 16-bit modes,
    // rounded to the activation dtype before the division (it is a bf16/fp16 tensor in the reference)
    float nrm = sqrtf(warp_sum(sq));
    if constexpr (sizeof(T) == 2) nrm = round16<T>(nrm);
    const float denom = fmaxf(nrm, eps);
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
        const int vi = lane + i * 32;
        if (vi < nvec) {
            float o[E];
#pragma unroll
            for (int e = 0; e < E; ++e) o[e] = v[i][e] / denom;
            store_vec<T>(yr + vi * E, o);
        }
    }
}

template <typename T>
int launch_ln(const void* x, int64_t ldx, const float* gamma, const float* beta, void* y, int64_t ldy, int rows, int width,
              float eps, int rs, const int32_t* row_idx, cudaStream_t stream) {
    constexpr int E = Vec16<T>::kElems;
    const int nvec = width / E;
    const int vpl = (nvec + 31) / 32;
    const int grid = (rows + kWarpsPerBlock - 1) / kWarpsPerBlock;
    const T* xp = static_cast<const T*>(x);
    T* yp = static_cast<T*>(y);
#define LN_CASE(V)                                                                                                        \
    case V:                                                                                                               \
        layernorm_kernel<T, V><<<grid, kWarpsPerBlock * 32, 0, stream>>>(xp, ldx, gamma, beta, yp, ldy, rows, width, eps, \
                                                                          rs, row_idx);                                   \
        break;
    switch (vpl) {
        LN_CASE(1) LN_CASE(2) LN_CASE(3) LN_CASE(4) LN_CASE(5) LN_CASE(6) LN_CASE(7) LN_CASE(8)
        default: set_last_error("layernorm: width %d too large", width); return -1;
    }
#undef LN_CASE
    B2C_LAUNCH_CHECK("layernorm_kernel");
    return 0;
}

template <typename T>
int launch_norm(const void* x, int64_t ldx, void* y, int64_t ldy, int rows, int dim, float eps, cudaStream_t stream) {
    constexpr int E = Vec16<T>::kElems;
    const int nvec = dim / E;
    const int vpl = (nvec + 31) / 32;
    const int grid = (rows + kWarpsPerBlock - 1) / kWarpsPerBlock;
    const T* xp = static_cast<const T*>(x);
    T* yp = static_cast<T*>(y);
#define NORM_CASE(V) \
    case V: normalize_kernel<T, V><<<grid, kWarpsPerBlock * 32, 0, stream>>>(xp, ldx, yp, ldy, rows, dim, eps); break;
    switch (vpl) {
        NORM_CASE(1) NORM_CASE(2) NORM_CASE(3) NORM_CASE(4) NORM_CASE(5) NORM_CASE(6) NORM_CASE(7) NORM_CASE(8)
        default: set_last_error("normalize: dim %d too large", dim); return -1;
    }
#undef NORM_CASE
    B2C_LAUNCH_CHECK("normalize_kernel");
    return 0;
}

template <typename T>
int launch_stats(const void* x, int64_t ldx, float* stats, int rows, int width, float eps, cudaStream_t stream) {
    constexpr int E = Vec16<T>::kElems;
    const int vpl = (width / E + 31) / 32;
    const int rows_per_warp = vpl <= 4 ? 2 : 1;   // must match R in row_stats_kernel
    const int warps = (rows + rows_per_warp - 1) / rows_per_warp;
    const int grid = (warps + kWarpsPerBlock - 1) / kWarpsPerBlock;
    const T* xp = static_cast<const T*>(x);
    float2* sp = reinterpret_cast<float2*>(stats);
    cudaLaunchConfig_t cfg = {};
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(kWarpsPerBlock * 32);
    cfg.stream = stream;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    cudaError_t le = cudaSuccess;
#define ST_CASE(V) \
    case V: le = cudaLaunchKernelEx(&cfg, row_stats_kernel<T, V>, xp, ldx, sp, rows, width, eps); break;
    switch (vpl) {
        ST_CASE(1) ST_CASE(2) ST_CASE(3) ST_CASE(4) ST_CASE(5) ST_CASE(6) ST_CASE(7) ST_CASE(8)
        default: set_last_error("row_stats: width %d too large", width); return -1;
    }
#undef ST_CASE
    if (le != cudaSuccess) return cuda_fail(le, "cudaLaunchKernelEx(row_stats_kernel)");
    B2C_LAUNCH_CHECK("row_stats_kernel");
    return 0;
}

}  // namespace

int row_stats(int dtype, const void* x, int64_t ldx, float* stats, int rows, int width, float eps, cudaStream_t stream) {
    B2C_CHECK_ARG(rows > 0 && width > 0, "row_stats: empty input rows=%d width=%d", rows, width);
    const int e = 16 / dtype_size(dtype);
    B2C_CHECK_ARG(width % e == 0 && ldx % e == 0, "row_stats: width/ld must be multiples of %d", e);
    B2C_CHECK_ARG(reinterpret_cast<uintptr_t>(x) % 16 == 0 && reinterpret_cast<uintptr_t>(stats) % 8 == 0,
                  "row_stats: x must be 16-byte and stats 8-byte aligned");
    switch (dtype) {
        case 0: return launch_stats<float>(x, ldx, stats, rows, width, eps, stream);
        case 1: return launch_stats<__nv_bfloat16>(x, ldx, stats, rows, width, eps, stream);
        case 2: return launch_stats<__half>(x, ldx, stats, rows, width, eps, stream);
    }
    set_last_error("row_stats: unknown dtype %d", dtype);
    return -1;
}

int layernorm(int dtype, const void* x, int64_t ldx, const float* gamma, const float* beta, void* y, int64_t ldy, int rows,
              int width, float eps, int row_stride_rows, const int32_t* row_idx, cudaStream_t stream) {
    B2C_CHECK_ARG(rows > 0 && width > 0, "layernorm: empty input rows=%d width=%d", rows, width);
    const int e = 16 / dtype_size(dtype);
    B2C_CHECK_ARG(width % e == 0 && ldx % e == 0 && ldy % e == 0, "layernorm: width/ld must be multiples of %d", e);
    B2C_CHECK_ARG(width % 4 == 0, "layernorm: width must be a multiple of 4");
    B2C_CHECK_ARG((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y) | reinterpret_cast<uintptr_t>(gamma) |
                   reinterpret_cast<uintptr_t>(beta)) % 16 == 0, "layernorm: pointers must be 16-byte aligned");
    if (row_stride_rows <= 0) row_stride_rows = 1;
    switch (dtype) {
        case 0: return launch_ln<float>(x, ldx, gamma, beta, y, ldy, rows, width, eps, row_stride_rows, row_idx, stream);
        case 1: return launch_ln<__nv_bfloat16>(x, ldx, gamma, beta, y, ldy, rows, width, eps, row_stride_rows, row_idx, stream);
        case 2: return launch_ln<__half>(x, ldx, gamma, beta, y, ldy, rows, width, eps, row_stride_rows, row_idx, stream);
    }
    set_last_error("layernorm: unknown dtype %d", dtype);
    return -1;
}

int normalize_rows(int dtype, const void* x, int64_t ldx, void* y, int64_t ldy, int rows, int dim, float eps,
                   cudaStream_t stream) {
    B2C_CHECK_ARG(rows > 0 && dim > 0, "normalize: empty input rows=%d dim=%d", rows, dim);
    const int e = 16 / dtype_size(dtype);
    B2C_CHECK_ARG(dim % e == 0 && ldx % e == 0 && ldy % e == 0, "normalize: dim/ld must be multiples of %d", e);
    B2C_CHECK_ARG((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) % 16 == 0,
                  "normalize: pointers must be 16-byte aligned");
    switch (dtype) {
        case 0: return launch_norm<float>(x, ldx, y, ldy, rows, dim, eps, stream);
        case 1: return launch_norm<__nv_bfloat16>(x, ldx, y, ldy, rows, dim, eps, stream);
        case 2: return launch_norm<__half>(x, ldx, y, ldy, rows, dim, eps, stream);
    }
    set_last_error("normalize: unknown dtype %d", dtype);
    return -1;
}

}  // namespace b200clip
