// Backward kernels of the tower blocks (training path): everything of the backward that is not a GEMM.
//   reference semantics: autograd through ResidualAttentionBlock (deps/open_clip/src/open_clip/transformer.py:253-264),
//   F.layer_norm (:15-30), nn.GELU / QuickGELU (:33-36), nn.MultiheadAttention's softmax attention (:224,249-251),
//   nn.Embedding + positional add (model.py:272-274), F.normalize (model.py:267,284), as driven by
//   training/train.py:115-183.  fp32 math everywhere; storage type T = the tower dtype.
//
//   transpose16        out[c][r] = in[r][c] for 16-bit matrices: operand preparation of the dgrad / wgrad GEMMs, which run on
//                      gemm_pair_kernel (C = A B^T with both operands K-major)
//   col_sum            bias gradients: column sums of a gradient matrix, deterministic two-stage reduction
//   ln_backward        d(LayerNorm input) (+ the residual-branch gradient that by-passes the LayerNorm) and per-block partial
//                      sums of d(gamma), d(beta); optional row gather / scatter (CLS / EOT pooled rows)
//   act_backward       dz = da * gelu'(z)  (erf GELU or QuickGELU)
//   attention_backward softmax attention backward per (image, head) with the probabilities recomputed from q, k
//   normalize_backward d(x) of y = x / max(||x||, eps)
//   embedding backward: scatter-add of d(x) rows into d(token_embedding) and the per-position sum for d(positional_embedding)
#include "common.cuh"
#include "internal.h"

#include <mma.h>

#include <cstdlib>

namespace b200clip {

namespace {

template <typename T> __device__ __forceinline__ float to_f(T v);
template <> __device__ __forceinline__ float to_f<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <> __device__ __forceinline__ float to_f<__half>(__half v) { return __half2float(v); }
template <typename T> __device__ __forceinline__ T from_f(float v);
template <> __device__ __forceinline__ float from_f<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }
template <> __device__ __forceinline__ __half from_f<__half>(float v) { return __float2half_rn(v); }

// ------------------------------------------------------------------------------------------------------------------------
// transpose (16-bit): 64 x 64 tiles through shared memory, 2-element vector accesses on both sides
// ------------------------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) transpose_kernel(const T* __restrict__ in, int64_t ldi, T* __restrict__ out, int64_t ldo, int R, int C, int Rpad,
                                                        int act) {
    __shared__ T tile[64][72];
    const int r0 = blockIdx.y * 64, c0 = blockIdx.x * 64;
    const bool vec_in = sizeof(T) == 2 && (ldi % 8 == 0) && (reinterpret_cast<uintptr_t>(in) % 16 == 0) && c0 + 64 <= C;
    if (vec_in) {
        // 64 rows x 8 vectors of 8 elements: 512 vector loads over 256 threads
        for (int i = threadIdx.x; i < 512; i += 256) {
            const int rr = i >> 3, v = (i & 7) * 8;
            uint4 val = make_uint4(0, 0, 0, 0);
            if (r0 + rr < R) val = *reinterpret_cast<const uint4*>(in + static_cast<int64_t>(r0 + rr) * ldi + c0 + v);
            *reinterpret_cast<uint4*>(&tile[rr][v]) = val;
        }
    } else {
        const int tx = threadIdx.x & 63, ty = threadIdx.x >> 6;   // 64 x 4
        for (int i = ty; i < 64; i += 4) {
            const int r = r0 + i, c = c0 + tx;
            tile[i][tx] = (r < R && c < C) ? in[static_cast<int64_t>(r) * ldi + c] : from_f<T>(0.f);
        }
    }
    __syncthreads();
    if (act != 0) {
        // out = act(in)^T: the activation of the MLP is applied on the way through (the recompute keeps only the pre-activation)
        for (int i = threadIdx.x; i < 64 * 64; i += 256) {
            const int rr = i >> 6, cc = i & 63;
            const float v = to_f<T>(tile[rr][cc]);
            tile[rr][cc] = from_f<T>(act == 2 ? quick_gelu(v) : gelu_erf(v));
        }
        __syncthreads();
    }
    const bool vec_out = sizeof(T) == 2 && (ldo % 8 == 0) && (reinterpret_cast<uintptr_t>(out) % 16 == 0) && r0 + 64 <= Rpad;
    if (vec_out) {
        // output row c (64 of them) x 8 vectors of 8 consecutive r
        for (int i = threadIdx.x; i < 512; i += 256) {
            const int cc = i >> 3, v = (i & 7) * 8;
            if (c0 + cc < C) {
                alignas(16) T o[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) o[j] = tile[v + j][cc];
                *reinterpret_cast<uint4*>(out + static_cast<int64_t>(c0 + cc) * ldo + r0 + v) = *reinterpret_cast<const uint4*>(o);
            }
        }
    } else {
        const int tx = threadIdx.x & 63, ty = threadIdx.x >> 6;
        for (int i = ty; i < 64; i += 4) {
            const int c = c0 + i, r = r0 + tx;
            if (c < C && r < Rpad) out[static_cast<int64_t>(c) * ldo + r] = tile[tx][i];   // rows >= R were loaded as zero
        }
    }
}

// ------------------------------------------------------------------------------------------------------------------------
// column sums: partial[chunk][col] over row chunks, then a fixed-order sum over the chunks
// ------------------------------------------------------------------------------------------------------------------------
constexpr int kColChunkRows = 64;
constexpr int kColVecChunkRows = 128;   // rows per block of the vectorised column sum
constexpr int kColSumCounters = 256;   // column blocks (64 columns each) the single-launch column sum has counters for

// block = 32 column pairs x 8 row lanes: 64 columns x kColChunkRows rows per block, rows strided over the 8 lanes, then a
// shared-memory reduction over the lanes (fixed order)
// `counters` != nullptr (one zero-initialised word per 64-column block, left at zero): the row chunk that finishes LAST for a
// column block also sums that block's partials (fixed order: deterministic) and writes out[c] -- no second launch.
template <typename T, typename TO>
__global__ void __launch_bounds__(256) colsum_partial_kernel(const T* __restrict__ g, int64_t ld, int rows, int cols, float* __restrict__ part,
                                                             unsigned int* __restrict__ counters, TO* __restrict__ out, int accumulate) {
    __shared__ float red[8][65];
    __shared__ bool is_last;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int c = blockIdx.x * 64 + tx * 2;
    const int r0 = blockIdx.y * kColChunkRows;
    const int r1 = min(rows, r0 + kColChunkRows);
    float s0 = 0.f, s1 = 0.f;
    if (c + 1 < cols) {
        for (int r = r0 + ty; r < r1; r += 8) {
            const T* p = g + static_cast<int64_t>(r) * ld + c;
            s0 += to_f<T>(p[0]);
            s1 += to_f<T>(p[1]);
        }
    } else if (c < cols) {
        for (int r = r0 + ty; r < r1; r += 8) s0 += to_f<T>(g[static_cast<int64_t>(r) * ld + c]);
    }
    red[ty][tx * 2] = s0;
    red[ty][tx * 2 + 1] = s1;
    __syncthreads();
    if (threadIdx.x < 64) {
        const int cc = blockIdx.x * 64 + threadIdx.x;
        if (cc < cols) {
            float s = 0.f;
#pragma unroll
            for (int k = 0; k < 8; ++k) s += red[k][threadIdx.x];
            part[static_cast<int64_t>(blockIdx.y) * cols + cc] = s;
        }
    }
    if (counters == nullptr) return;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int prev = atomicAdd(counters + blockIdx.x, 1u);
        is_last = prev == gridDim.y - 1;
        if (is_last) counters[blockIdx.x] = 0u;
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    // 64 columns x 4 chunk lanes
    const int col = threadIdx.x & 63, lane4 = threadIdx.x >> 6;
    const int cc = blockIdx.x * 64 + col;
    float s = 0.f;
    if (cc < cols)
        for (int k = lane4; k < static_cast<int>(gridDim.y); k += 4) s += __ldcg(part + static_cast<int64_t>(k) * cols + cc);
    red[lane4][col] = s;
    __syncthreads();
    if (threadIdx.x < 64 && cc < cols) {
        float t = (red[0][col] + red[1][col]) + (red[2][col] + red[3][col]);
        if (accumulate) t += to_f<TO>(out[cc]);
        out[cc] = from_f<TO>(t);
    }
}
// out[c] (+)= sum over chunks of part[chunk][c]: 32 columns x 8 chunk lanes per block (fixed summation order)
template <typename TO>
__global__ void __launch_bounds__(256) colsum_final_kernel(const float* __restrict__ part, int chunks, int cols, TO* __restrict__ out, int accumulate) {
    __shared__ float red[8][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int c = blockIdx.x * 32 + tx;
    float s = 0.f;
    if (c < cols)
        for (int k = ty; k < chunks; k += 8) s += part[static_cast<int64_t>(k) * cols + c];
    red[ty][tx] = s;
    __syncthreads();
    if (ty == 0 && c < cols) {
        float t = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) t += red[k][tx];
        if (accumulate) t += to_f<TO>(out[c]);
        out[c] = from_f<TO>(t);
    }
}

// ------------------------------------------------------------------------------------------------------------------------
// LayerNorm backward.  One warp per row; the block's rows are consecutive, its d(gamma) / d(beta) contributions are summed in
// shared memory and written as one partial row per block (reduced by colsum_final_kernel).
//   xhat = (x - mean) * rstd,  gy = g * gamma,  dx = rstd * (gy - mean_j(gy) - xhat * mean_j(gy * xhat)) (+ dres)
// Row gather: logical row r reads x row r * row_stride + (row_idx ? row_idx[r] : 0) and writes dx to the same physical row.
// ------------------------------------------------------------------------------------------------------------------------
constexpr int kLnWarps = 8;
constexpr int kLnRowsPerBlock = 32;

template <typename T>
__global__ void __launch_bounds__(kLnWarps * 32)
ln_backward_kernel(const T* __restrict__ g, int64_t ldg, const T* __restrict__ x, int64_t ldx, const float* __restrict__ gamma,
                   const T* __restrict__ dres, int64_t ldr, T* __restrict__ dx, int64_t ldd, float* __restrict__ dgamma_part,
                   float* __restrict__ dbeta_part, int rows, int width, float eps, int row_stride, const int32_t* __restrict__ row_idx) {
    extern __shared__ float sm[];   // [2][kLnWarps][width]
    float* sg = sm;
    float* sb = sm + kLnWarps * width;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int j = lane; j < width; j += 32) {
        sg[warp * width + j] = 0.f;
        sb[warp * width + j] = 0.f;
    }
    const int row_begin = blockIdx.x * kLnRowsPerBlock;
    const int row_end = min(rows, row_begin + kLnRowsPerBlock);
    const float inv_w = 1.0f / static_cast<float>(width);
    for (int r = row_begin + warp; r < row_end; r += kLnWarps) {
        const int64_t pr = static_cast<int64_t>(r) * row_stride + (row_idx != nullptr ? row_idx[r] : 0);
        const T* xr = x + pr * ldx;
        const T* gr = g + static_cast<int64_t>(r) * ldg;
        float s = 0.f, q = 0.f;
        for (int j = lane; j < width; j += 32) {
            const float v = to_f<T>(xr[j]);
            s += v;
        }
        const float mean = warp_sum(s) * inv_w;
        for (int j = lane; j < width; j += 32) {
            const float d = to_f<T>(xr[j]) - mean;
            q = fmaf(d, d, q);
        }
        const float rstd = rsqrtf(warp_sum(q) * inv_w + eps);
        float a = 0.f, b = 0.f;
        for (int j = lane; j < width; j += 32) {
            const float xh = (to_f<T>(xr[j]) - mean) * rstd;
            const float gv = to_f<T>(gr[j]);
            const float gy = gv * gamma[j];
            a += gy;
            b = fmaf(gy, xh, b);
            sg[warp * width + j] = fmaf(gv, xh, sg[warp * width + j]);
            sb[warp * width + j] += gv;
        }
        a = warp_sum(a) * inv_w;
        b = warp_sum(b) * inv_w;
        T* dr = dx + pr * ldd;
        for (int j = lane; j < width; j += 32) {
            const float xh = (to_f<T>(xr[j]) - mean) * rstd;
            float v = rstd * (to_f<T>(gr[j]) * gamma[j] - a - xh * b);
            if (dres != nullptr) v += to_f<T>(dres[pr * ldr + j]);
            dr[j] = from_f<T>(v);
        }
    }
    __syncthreads();
    for (int j = threadIdx.x; j < width; j += kLnWarps * 32) {
        float tg = 0.f, tb = 0.f;
#pragma unroll
        for (int w = 0; w < kLnWarps; ++w) {
            tg += sg[w * width + j];
            tb += sb[w * width + j];
        }
        dgamma_part[static_cast<int64_t>(blockIdx.x) * width + j] = tg;
        dbeta_part[static_cast<int64_t>(blockIdx.x) * width + j] = tb;
    }
}

// Vectorised form for widths that are a multiple of 32 lanes x 16 bytes: every lane keeps its VPL vectors of the row in
// registers (ONE read of g, x and dres), and its d(gamma) / d(beta) contributions for its own columns across all rows of the
// warp; the warps of the block are combined through shared memory at the end.
template <typename T> struct VecOf { static constexpr int n = 16 / sizeof(T); };
template <typename T> __device__ __forceinline__ void load16(const T* p, float* f) {
    if constexpr (sizeof(T) == 4) {
        const float4 v = *reinterpret_cast<const float4*>(p);
        f[0] = v.x; f[1] = v.y; f[2] = v.z; f[3] = v.w;
    } else {
        const uint4 v = *reinterpret_cast<const uint4*>(p);
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float2 t = Half16<T>::unpack(w[i]);
            f[2 * i] = t.x;
            f[2 * i + 1] = t.y;
        }
    }
}
template <typename T> __device__ __forceinline__ void store16(T* p, const float* f) {
    if constexpr (sizeof(T) == 4) {
        *reinterpret_cast<float4*>(p) = make_float4(f[0], f[1], f[2], f[3]);
    } else {
        using H = Half16<T>;
        *reinterpret_cast<uint4*>(p) = make_uint4(H::pack(f[0], f[1]), H::pack(f[2], f[3]), H::pack(f[4], f[5]), H::pack(f[6], f[7]));
    }
}

template <typename T, int VPL>
__global__ void __launch_bounds__(kLnWarps * 32)
ln_backward_vec_kernel(const T* __restrict__ g, int64_t ldg, const T* __restrict__ x, int64_t ldx, const float* __restrict__ gamma,
                       const T* __restrict__ dres, int64_t ldr, T* __restrict__ dx, int64_t ldd, float* __restrict__ dgamma_part,
                       float* __restrict__ dbeta_part, int rows, int width, float eps, int row_stride, const int32_t* __restrict__ row_idx) {
    constexpr int V = VecOf<T>::n;
    extern __shared__ float sm[];   // [2][kLnWarps][width]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float gam[VPL * V], ag[VPL * V], ab[VPL * V];
#pragma unroll
    for (int k = 0; k < VPL; ++k)
#pragma unroll
        for (int i = 0; i < V; ++i) {
            gam[k * V + i] = gamma[(k * 32 + lane) * V + i];
            ag[k * V + i] = 0.f;
            ab[k * V + i] = 0.f;
        }
    const int row_begin = blockIdx.x * kLnRowsPerBlock;
    const int row_end = min(rows, row_begin + kLnRowsPerBlock);
    const float inv_w = 1.0f / static_cast<float>(width);
    for (int r = row_begin + warp; r < row_end; r += kLnWarps) {
        const int64_t pr = static_cast<int64_t>(r) * row_stride + (row_idx != nullptr ? row_idx[r] : 0);
        float xv[VPL * V], gv[VPL * V];
        float s = 0.f;
#pragma unroll
        for (int k = 0; k < VPL; ++k) {
            load16<T>(x + pr * ldx + (k * 32 + lane) * V, xv + k * V);
            load16<T>(g + static_cast<int64_t>(r) * ldg + (k * 32 + lane) * V, gv + k * V);
#pragma unroll
            for (int i = 0; i < V; ++i) s += xv[k * V + i];
        }
        const float mean = warp_sum(s) * inv_w;
        float q = 0.f;
#pragma unroll
        for (int i = 0; i < VPL * V; ++i) {
            xv[i] -= mean;
            q = fmaf(xv[i], xv[i], q);
        }
        const float rstd = rsqrtf(warp_sum(q) * inv_w + eps);
        float a = 0.f, b = 0.f;
#pragma unroll
        for (int i = 0; i < VPL * V; ++i) {
            xv[i] *= rstd;                       // xhat
            const float gy = gv[i] * gam[i];
            a += gy;
            b = fmaf(gy, xv[i], b);
            ag[i] = fmaf(gv[i], xv[i], ag[i]);
            ab[i] += gv[i];
        }
        a = warp_sum(a) * inv_w;
        b = warp_sum(b) * inv_w;
#pragma unroll
        for (int k = 0; k < VPL; ++k) {
            float o[V];
            if (dres != nullptr) load16<T>(dres + pr * ldr + (k * 32 + lane) * V, o);
#pragma unroll
            for (int i = 0; i < V; ++i) {
                const float v = rstd * (gv[k * V + i] * gam[k * V + i] - a - xv[k * V + i] * b);
                o[i] = dres != nullptr ? o[i] + v : v;
            }
            store16<T>(dx + pr * ldd + (k * 32 + lane) * V, o);
        }
    }
    float* sg = sm;
    float* sb = sm + kLnWarps * width;
#pragma unroll
    for (int k = 0; k < VPL; ++k)
#pragma unroll
        for (int i = 0; i < V; ++i) {
            sg[warp * width + (k * 32 + lane) * V + i] = ag[k * V + i];
            sb[warp * width + (k * 32 + lane) * V + i] = ab[k * V + i];
        }
    __syncthreads();
    for (int j = threadIdx.x; j < width; j += kLnWarps * 32) {
        float tg = 0.f, tb = 0.f;
#pragma unroll
        for (int w = 0; w < kLnWarps; ++w) {
            tg += sg[w * width + j];
            tb += sb[w * width + j];
        }
        dgamma_part[static_cast<int64_t>(blockIdx.x) * width + j] = tg;
        dbeta_part[static_cast<int64_t>(blockIdx.x) * width + j] = tb;
    }
}

template <typename T>
bool launch_ln_backward_vec(const T* g, int64_t ldg, const T* x, int64_t ldx, const float* gamma, const T* dres, int64_t ldr, T* dx, int64_t ldd,
                            float* pg, float* pb, int rows, int width, float eps, int row_stride, const int32_t* row_idx, int blocks, size_t smem,
                            cudaStream_t stream) {
    constexpr int V = VecOf<T>::n;
    if (width % (32 * V) != 0 || ldg % V != 0 || ldx % V != 0 || ldd % V != 0 || (dres != nullptr && ldr % V != 0)) return false;
    if ((reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(dx) | reinterpret_cast<uintptr_t>(dres) |
         reinterpret_cast<uintptr_t>(gamma)) % 16 != 0)
        return false;
    const int vpl = width / (32 * V);
#define B2C_LN_VEC(N)                                                                                                                       \
    case N:                                                                                                                                 \
        ln_backward_vec_kernel<T, N><<<blocks, kLnWarps * 32, smem, stream>>>(g, ldg, x, ldx, gamma, dres, ldr, dx, ldd, pg, pb, rows, width, eps, \
                                                                              row_stride, row_idx);                                         \
        return true;
    switch (vpl) {
        B2C_LN_VEC(1)
        B2C_LN_VEC(2)
        B2C_LN_VEC(3)
        B2C_LN_VEC(4)
    }   // wider rows (fp32 at W >= 768: the parity mode) keep the scalar kernel
#undef B2C_LN_VEC
    return false;
}

// ------------------------------------------------------------------------------------------------------------------------
// activation backward
// ------------------------------------------------------------------------------------------------------------------------
// Column sums with 16-byte loads: a block covers 32 lanes x V columns (256 for 16-bit types) and kColVecChunkRows rows (its 8 warps
// stride over the rows), writes one partial row, and the block that finishes LAST for its column block (counter per column block,
// zero on entry, left at zero) sums the partials in a fixed order and writes out[c]: one launch, deterministic.
template <typename T, typename TO>
__global__ void __launch_bounds__(256) colsum_vec_kernel(const T* __restrict__ g, int64_t ld, int rows, int cols, float* __restrict__ part,
                                                         unsigned int* __restrict__ counters, TO* __restrict__ out, int accumulate) {
    constexpr int V = VecOf<T>::n;
    constexpr int CB = 32 * V;
    __shared__ float red[8][CB];
    __shared__ bool is_last;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int c = blockIdx.x * CB + lane * V;
    const int r0 = blockIdx.y * kColVecChunkRows;
    const int r1 = min(rows, r0 + kColVecChunkRows);
    float acc[V];
#pragma unroll
    for (int j = 0; j < V; ++j) acc[j] = 0.f;
    if (c < cols) {
        for (int r = r0 + warp; r < r1; r += 8) {
            float v[V];
            load16<T>(g + static_cast<int64_t>(r) * ld + c, v);
#pragma unroll
            for (int j = 0; j < V; ++j) acc[j] += v[j];
        }
    }
#pragma unroll
    for (int j = 0; j < V; ++j) red[warp][lane * V + j] = acc[j];
    __syncthreads();
    for (int t = threadIdx.x; t < CB; t += 256) {
        const int cc = blockIdx.x * CB + t;
        if (cc < cols) {
            float s2 = 0.f;
#pragma unroll
            for (int k = 0; k < 8; ++k) s2 += red[k][t];
            part[static_cast<int64_t>(blockIdx.y) * cols + cc] = s2;
        }
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int prev = atomicAdd(counters + blockIdx.x, 1u);
        is_last = prev == gridDim.y - 1;
        if (is_last) counters[blockIdx.x] = 0u;
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    // the partial rows of this column block: warp w takes chunks w, w + 8, ... (short dependent chains), then the warps are summed
#pragma unroll
    for (int j = 0; j < V; ++j) acc[j] = 0.f;
    if (c < cols) {
        for (int k = warp; k < static_cast<int>(gridDim.y); k += 8) {
            const float4* src = reinterpret_cast<const float4*>(part + static_cast<int64_t>(k) * cols + c);
#pragma unroll
            for (int q4 = 0; q4 < V / 4; ++q4) {
                const float4 v4 = __ldcg(src + q4);
                acc[q4 * 4 + 0] += v4.x; acc[q4 * 4 + 1] += v4.y; acc[q4 * 4 + 2] += v4.z; acc[q4 * 4 + 3] += v4.w;
            }
        }
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < V; ++j) red[warp][lane * V + j] = acc[j];
    __syncthreads();
    for (int t = threadIdx.x; t < CB; t += 256) {
        const int cc = blockIdx.x * CB + t;
        if (cc < cols) {
            float s2 = 0.f;
#pragma unroll
            for (int k = 0; k < 8; ++k) s2 += red[k][t];
            if (accumulate) s2 += to_f<TO>(out[cc]);
            out[cc] = from_f<TO>(s2);
        }
    }
}

__device__ __forceinline__ float act_grad(float x, int quick) {
    if (quick) {
        const float sg = 1.0f / (1.0f + __expf(-1.702f * x));
        return sg * (1.0f + 1.702f * x * (1.0f - sg));
    }
    const float cdf = 0.5f * (1.0f + erff(x * 0.70710678118654752440f));
    const float pdf = 0.3989422804014327f * __expf(-0.5f * x * x);
    return cdf + x * pdf;
}
// 16-byte accesses (8 / 4 elements per thread per step) when `vec`; the scalar loop handles unaligned buffers and the tail
template <typename T>
__global__ void __launch_bounds__(256) act_backward_kernel(const T* __restrict__ da, const T* __restrict__ z, T* __restrict__ dz, int64_t n, int quick,
                                                           int vec) {
    constexpr int V = VecOf<T>::n;
    const int64_t tid = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x, nthr = static_cast<int64_t>(gridDim.x) * blockDim.x;
    const int64_t nv = vec ? n / V : 0;
    for (int64_t i = tid; i < nv; i += nthr) {
        float zv[V], dv[V];
        load16<T>(z + i * V, zv);
        load16<T>(da + i * V, dv);
        if (quick || sizeof(T) == 4) {   // fp32 (parity mode) keeps erff / expf
#pragma unroll
            for (int j = 0; j < V; ++j) dv[j] *= act_grad(zv[j], quick);
        } else {
#pragma unroll
            for (int j = 0; j < V; j += 2) {
                float d0, d1;
                gelu_grad_pair_fast(zv[j], zv[j + 1], d0, d1);
                dv[j] *= d0;
                dv[j + 1] *= d1;
            }
        }
        store16<T>(dz + i * V, dv);
    }
    for (int64_t i = nv * V + tid; i < n; i += nthr) dz[i] = from_f<T>(to_f<T>(da[i]) * act_grad(to_f<T>(z[i]), quick));
}
// forward activation as a separate pass (the training recompute keeps the pre-activation z for the backward)
template <typename T>
__global__ void __launch_bounds__(256) act_forward_kernel(const T* __restrict__ z, T* __restrict__ a, int64_t n, int quick, int vec) {
    constexpr int V = VecOf<T>::n;
    const int64_t tid = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x, nthr = static_cast<int64_t>(gridDim.x) * blockDim.x;
    const int64_t nv = vec ? n / V : 0;
    for (int64_t i = tid; i < nv; i += nthr) {
        float zv[V];
        load16<T>(z + i * V, zv);
        if (quick || sizeof(T) == 4) {
#pragma unroll
            for (int j = 0; j < V; ++j) zv[j] = quick ? quick_gelu(zv[j]) : gelu_erf(zv[j]);
        } else {   // the fitted form of the fused forward epilogue (gemm_pair.cu): the recompute reproduces the forward's values
#pragma unroll
            for (int j = 0; j < V; j += 2) gelu_pair_fast(zv[j], zv[j + 1]);
        }
        store16<T>(a + i * V, zv);
    }
    for (int64_t i = nv * V + tid; i < n; i += nthr) {
        const float x = to_f<T>(z[i]);
        a[i] = from_f<T>(quick ? quick_gelu(x) : gelu_erf(x));
    }
}

// ------------------------------------------------------------------------------------------------------------------------
// attention backward: one CTA per (image, head); K, V of the head in shared memory (storage type), dK / dV accumulated in
// shared memory (fp32), queries in tiles of kTQ rows.  Probabilities are recomputed (no attention matrix is saved).
//   S = scale q k^T (+ causal mask), P = softmax(S), dV += P^T dO, dP = dO V^T, dS = P * (dP - rowsum(P * dP)),
//   dq = scale dS k, dk += scale dS^T q
// ------------------------------------------------------------------------------------------------------------------------
constexpr int kHd = 64;
constexpr int kTQ = 16;
constexpr int kAttnThreads = 256;

template <typename T>
__global__ void __launch_bounds__(kAttnThreads)
attention_backward_kernel(const T* __restrict__ qkv, const T* __restrict__ d_out, T* __restrict__ d_qkv, int L, int heads, int causal) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int W = heads * kHd;
    const int b = blockIdx.x / heads, h = blockIdx.x % heads;
    const int Lp = L + 1;                         // padded probability row
    float* dK = reinterpret_cast<float*>(smem_raw);            // [L][65]
    float* dV = dK + L * (kHd + 1);                            // [L][65]
    float* Qs = dV + L * (kHd + 1);                            // [kTQ][65]
    float* dOs = Qs + kTQ * (kHd + 1);                         // [kTQ][65]
    float* P = dOs + kTQ * (kHd + 1);                          // [kTQ][Lp]
    float* dS = P + kTQ * Lp;                                  // [kTQ][Lp]
    T* Ks = reinterpret_cast<T*>(dS + kTQ * Lp);               // [L][66]
    T* Vs = Ks + L * (kHd + 2);                                // [L][66]
    const int tid = threadIdx.x;
    const float scale = 0.125f;                  // 1 / sqrt(64)
    const T* base = qkv + static_cast<int64_t>(b) * L * 3 * W + h * kHd;
    for (int i = tid; i < L * kHd; i += kAttnThreads) {
        const int l = i / kHd, d = i % kHd;
        Ks[l * (kHd + 2) + d] = base[static_cast<int64_t>(l) * 3 * W + W + d];
        Vs[l * (kHd + 2) + d] = base[static_cast<int64_t>(l) * 3 * W + 2 * W + d];
        dK[l * (kHd + 1) + d] = 0.f;
        dV[l * (kHd + 1) + d] = 0.f;
    }
    __syncthreads();
    for (int q0 = 0; q0 < L; q0 += kTQ) {
        const int nq = min(kTQ, L - q0);
        for (int i = tid; i < kTQ * kHd; i += kAttnThreads) {
            const int r = i / kHd, d = i % kHd;
            const bool ok = r < nq;
            Qs[r * (kHd + 1) + d] = ok ? to_f<T>(base[static_cast<int64_t>(q0 + r) * 3 * W + d]) : 0.f;
            dOs[r * (kHd + 1) + d] = ok ? to_f<T>(d_out[(static_cast<int64_t>(b) * L + q0 + r) * W + h * kHd + d]) : 0.f;
        }
        __syncthreads();
        // S and dP for every (query row, key) pair of the tile
        for (int i = tid; i < kTQ * L; i += kAttnThreads) {
            const int r = i / L, j = i % L;
            float s = 0.f, dp = 0.f;
#pragma unroll 8
            for (int d = 0; d < kHd; ++d) {
                s = fmaf(Qs[r * (kHd + 1) + d], to_f<T>(Ks[j * (kHd + 2) + d]), s);
                dp = fmaf(dOs[r * (kHd + 1) + d], to_f<T>(Vs[j * (kHd + 2) + d]), dp);
            }
            const bool masked = (causal && j > q0 + r) || r >= nq;
            P[r * Lp + j] = masked ? -INFINITY : s * scale;
            dS[r * Lp + j] = dp;
        }
        __syncthreads();
        // row softmax and dS = P * (dP - sum_j P dP): one warp per row
        {
            const int warp = tid >> 5, lane = tid & 31;
            for (int r = warp; r < kTQ; r += kAttnThreads / 32) {
                if (r >= nq) {
                    for (int j = lane; j < L; j += 32) {
                        P[r * Lp + j] = 0.f;
                        dS[r * Lp + j] = 0.f;
                    }
                    continue;
                }
                float m = -INFINITY;
                for (int j = lane; j < L; j += 32) m = fmaxf(m, P[r * Lp + j]);
                m = warp_max(m);
                float sum = 0.f;
                for (int j = lane; j < L; j += 32) {
                    const float e = __expf(P[r * Lp + j] - m);
                    P[r * Lp + j] = e;
                    sum += e;
                }
                const float inv = 1.0f / warp_sum(sum);
                float dot = 0.f;
                for (int j = lane; j < L; j += 32) {
                    const float pv = P[r * Lp + j] * inv;
                    P[r * Lp + j] = pv;
                    dot = fmaf(pv, dS[r * Lp + j], dot);
                }
                dot = warp_sum(dot);
                for (int j = lane; j < L; j += 32) dS[r * Lp + j] = P[r * Lp + j] * (dS[r * Lp + j] - dot);
            }
        }
        __syncthreads();
        // dV[j][d] += sum_r P[r][j] dO[r][d];  dK[j][d] += scale sum_r dS[r][j] q[r][d]
        for (int i = tid; i < L * kHd; i += kAttnThreads) {
            const int j = i / kHd, d = i % kHd;
            float av = 0.f, ak = 0.f;
#pragma unroll
            for (int r = 0; r < kTQ; ++r) {
                av = fmaf(P[r * Lp + j], dOs[r * (kHd + 1) + d], av);
                ak = fmaf(dS[r * Lp + j], Qs[r * (kHd + 1) + d], ak);
            }
            dV[j * (kHd + 1) + d] += av;
            dK[j * (kHd + 1) + d] = fmaf(scale, ak, dK[j * (kHd + 1) + d]);
        }
        // dq[r][d] = scale sum_j dS[r][j] k[j][d]
        for (int i = tid; i < nq * kHd; i += kAttnThreads) {
            const int r = i / kHd, d = i % kHd;
            float a = 0.f;
            for (int j = 0; j < L; ++j) a = fmaf(dS[r * Lp + j], to_f<T>(Ks[j * (kHd + 2) + d]), a);
            d_qkv[(static_cast<int64_t>(b) * L + q0 + r) * 3 * W + h * kHd + d] = from_f<T>(a * scale);
        }
        __syncthreads();
    }
    for (int i = tid; i < L * kHd; i += kAttnThreads) {
        const int l = i / kHd, d = i % kHd;
        T* dst = d_qkv + (static_cast<int64_t>(b) * L + l) * 3 * W + h * kHd + d;
        dst[W] = from_f<T>(dK[l * (kHd + 1) + d]);
        dst[2 * W] = from_f<T>(dV[l * (kHd + 1) + d]);
    }
}

// ------------------------------------------------------------------------------------------------------------------------
// Tensor-core attention backward for the 16-bit modes and short sequences (L <= LP <= 80: ViT-B/32's 50 tokens, the 77-token
// text context): the whole (image, head) problem lives in shared memory, the five small products run on mma.sync through the
// wmma API (16 x 16 x 16 tiles, fp32 accumulation), the softmax algebra in fp32 in between:
//   S = Q K^T, dP = dO V^T -> P = softmax(scale S), dS = scale P (dP - rowsum(P dP)) -> dV = P^T dO, dK = dS^T Q, dQ = dS K
// P and dS are rounded to the storage type for the second round of products (as flash-attention style kernels do).
// ------------------------------------------------------------------------------------------------------------------------
template <typename T> struct WmmaT;
template <> struct WmmaT<__nv_bfloat16> { using type = __nv_bfloat16; };
template <> struct WmmaT<__half> { using type = __half; };

template <typename T, int LP>
__global__ void __launch_bounds__(256) attention_backward_tc_kernel(const T* __restrict__ qkv, const T* __restrict__ d_out, T* __restrict__ d_qkv, int L,
                                                                     int heads, int causal) {
    namespace wmma = nvcuda::wmma;
    using WT = typename WmmaT<T>::type;
    constexpr int LDT = kHd + 8;        // 72: operand tiles [LP][64] in the storage type
    constexpr int LDS = LP + 8;         // score matrices [LP][LP], fp32
    constexpr int LDP = 2 * LDS;        // the same rows seen as storage-type rows (P / dS overwrite the fp32 rows they were made from)
    constexpr int NT = LP / 16;
    constexpr int NJ = (LP + 31) / 32;  // row elements per lane
    constexpr int LDW = 20;             // per-warp output staging tile [16][LDW] fp32
    extern __shared__ __align__(128) unsigned char smem_raw[];
    T* Qs = reinterpret_cast<T*>(smem_raw);
    T* Ks = Qs + LP * LDT;
    T* Vs = Ks + LP * LDT;
    T* dOs = Vs + LP * LDT;
    float* Sf = reinterpret_cast<float*>(dOs + LP * LDT);   // [LP][LDS] fp32: S; after the softmax its rows hold P in the storage type
    float* dPf = Sf + LP * LDS;                              // [LP][LDS] fp32: dP; afterwards scale * dS in the storage type
    float* wstage = dPf + LP * LDS;                          // [8 warps][16][LDW]
    T* Pb = reinterpret_cast<T*>(Sf);
    T* dSb = reinterpret_cast<T*>(dPf);
    const int W = heads * kHd;
    const int b = blockIdx.x / heads, h = blockIdx.x % heads;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const T* base = qkv + static_cast<int64_t>(b) * L * 3 * W + h * kHd;
    const T* dbase = d_out + static_cast<int64_t>(b) * L * W + h * kHd;
    // ---- operand tiles (rows >= L are zero), 16-byte accesses: 8 vectors per 64-element row
    for (int i = tid; i < LP * 8; i += 256) {
        const int l = i >> 3, c = (i & 7) * 8;
        uint4 q = make_uint4(0, 0, 0, 0), k = q, v = q, g = q;
        if (l < L) {
            const T* row = base + static_cast<int64_t>(l) * 3 * W + c;
            q = *reinterpret_cast<const uint4*>(row);
            k = *reinterpret_cast<const uint4*>(row + W);
            v = *reinterpret_cast<const uint4*>(row + 2 * W);
            g = *reinterpret_cast<const uint4*>(dbase + static_cast<int64_t>(l) * W + c);
        }
        *reinterpret_cast<uint4*>(Qs + l * LDT + c) = q;
        *reinterpret_cast<uint4*>(Ks + l * LDT + c) = k;
        *reinterpret_cast<uint4*>(Vs + l * LDT + c) = v;
        *reinterpret_cast<uint4*>(dOs + l * LDT + c) = g;
    }
    __syncthreads();
    // ---- S = Q K^T and dP = dO V^T: 2 * NT * NT output tiles over the 8 warps
    for (int t = warp; t < 2 * NT * NT; t += 8) {
        const int which = t / (NT * NT), tt = t % (NT * NT);
        const int mi = tt / NT, ni = tt % NT;
        const T* A = (which == 0 ? Qs : dOs) + mi * 16 * LDT;
        const T* Bm = (which == 0 ? Ks : Vs) + ni * 16 * LDT;
        wmma::fragment<wmma::accumulator, 16, 16, 16, float> acc;
        wmma::fill_fragment(acc, 0.f);
#pragma unroll
        for (int k = 0; k < kHd; k += 16) {
            wmma::fragment<wmma::matrix_a, 16, 16, 16, WT, wmma::row_major> fa;
            wmma::fragment<wmma::matrix_b, 16, 16, 16, WT, wmma::col_major> fb;
            wmma::load_matrix_sync(fa, reinterpret_cast<const WT*>(A + k), LDT);
            wmma::load_matrix_sync(fb, reinterpret_cast<const WT*>(Bm + k), LDT);
            wmma::mma_sync(acc, fa, fb, acc);
        }
        wmma::store_matrix_sync((which == 0 ? Sf : dPf) + mi * 16 * LDS + ni * 16, acc, LDS, wmma::mem_row_major);
    }
    __syncthreads();
    // ---- softmax rows, dS: one warp per row.  The row is read into registers, then P and scale * dS (storage type) are written
    // over the first half of the fp32 rows they were computed from -- no separate P / dS buffers (the 80-row text variant fits two
    // CTAs per SM that way: 110 KB instead of 143 KB).
    const float scale = 0.125f;
    for (int r = warp; r < LP; r += 8) {
        float pv[NJ], dv[NJ];
        float dot = 0.f;
        const int jmax = r < L ? (causal ? r + 1 : L) : 0;        // keys [0, jmax) take part (none for the padding rows)
        if (r < L) {
            float m = -INFINITY;
#pragma unroll
            for (int t = 0; t < NJ; ++t) {
                const int j = lane + 32 * t;
                pv[t] = j < jmax ? Sf[r * LDS + j] * scale : -INFINITY;
                dv[t] = j < jmax ? dPf[r * LDS + j] : 0.f;
                m = fmaxf(m, pv[t]);
            }
            m = warp_max(m);
            float sum = 0.f;
#pragma unroll
            for (int t = 0; t < NJ; ++t) {
                pv[t] = lane + 32 * t < jmax ? __expf(pv[t] - m) : 0.f;
                sum += pv[t];
            }
            const float inv = 1.0f / warp_sum(sum);
#pragma unroll
            for (int t = 0; t < NJ; ++t) {
                pv[t] *= inv;
                dot = fmaf(pv[t], dv[t], dot);
            }
            dot = warp_sum(dot);
        } else {
#pragma unroll
            for (int t = 0; t < NJ; ++t) pv[t] = dv[t] = 0.f;
        }
        __syncwarp();      // every lane has read its part of the two fp32 rows before they are overwritten
#pragma unroll
        for (int t = 0; t < NJ; ++t) {
            const int j = lane + 32 * t;
            if (j < LP) {
                const bool on = j < jmax;
                Pb[r * LDP + j] = from_f<T>(on ? pv[t] : 0.f);
                dSb[r * LDP + j] = from_f<T>(on ? scale * pv[t] * (dv[t] - dot) : 0.f);
            }
        }
    }
    __syncthreads();
    // ---- dV = P^T dO, dK = dS^T Q, dQ = dS K: 3 * NT * 4 output tiles [16 x 16] of [LP x 64] matrices; every warp converts its tile
    // through a private fp32 staging tile and writes the rows (< L) in the storage type: 32 bytes per row and tile
    float* mine = wstage + warp * 16 * LDW;
    for (int t = warp; t < 3 * NT * 4; t += 8) {
        const int which = t / (NT * 4), tt = t % (NT * 4);
        const int mi = tt / 4, ni = tt % 4;
        wmma::fragment<wmma::accumulator, 16, 16, 16, float> acc;
        wmma::fill_fragment(acc, 0.f);
        const T* Bm = which == 0 ? dOs : (which == 1 ? Qs : Ks);
        if (which < 2) {
            const T* A = which == 0 ? Pb : dSb;        // A^T: element (m, k) = A[k][m]
#pragma unroll
            for (int k = 0; k < LP; k += 16) {
                wmma::fragment<wmma::matrix_a, 16, 16, 16, WT, wmma::col_major> fa;
                wmma::fragment<wmma::matrix_b, 16, 16, 16, WT, wmma::row_major> fb;
                wmma::load_matrix_sync(fa, reinterpret_cast<const WT*>(A + k * LDP + mi * 16), LDP);
                wmma::load_matrix_sync(fb, reinterpret_cast<const WT*>(Bm + k * LDT + ni * 16), LDT);
                wmma::mma_sync(acc, fa, fb, acc);
            }
        } else {
#pragma unroll
            for (int k = 0; k < LP; k += 16) {
                wmma::fragment<wmma::matrix_a, 16, 16, 16, WT, wmma::row_major> fa;
                wmma::fragment<wmma::matrix_b, 16, 16, 16, WT, wmma::row_major> fb;
                wmma::load_matrix_sync(fa, reinterpret_cast<const WT*>(dSb + mi * 16 * LDP + k), LDP);
                wmma::load_matrix_sync(fb, reinterpret_cast<const WT*>(Bm + k * LDT + ni * 16), LDT);
                wmma::mma_sync(acc, fa, fb, acc);
            }
        }
        wmma::store_matrix_sync(mine, acc, LDW, wmma::mem_row_major);
        __syncwarp();
        const int row = lane >> 1, half = lane & 1;
        const int l = mi * 16 + row;
        if (l < L) {
            const float* src = mine + row * LDW + half * 8;
            alignas(16) T o[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) o[j] = from_f<T>(src[j]);
            const int slot = which == 0 ? 2 : (which == 1 ? 1 : 0);      // output order q, k, v <- dQ (2), dK (1), dV (0)
            *reinterpret_cast<uint4*>(d_qkv + (static_cast<int64_t>(b) * L + l) * 3 * W + slot * W + h * kHd + ni * 16 + half * 8) =
                *reinterpret_cast<const uint4*>(o);
        }
        __syncwarp();      // the staging tile is free for this warp's next output tile
    }
}

template <typename T, int LP> constexpr size_t attn_bwd_tc_smem() {
    constexpr size_t LDT = kHd + 8, LDS = LP + 8;
    return 4 * LP * LDT * sizeof(T) + 2 * LP * LDS * sizeof(float) + 8 * 16 * 20 * sizeof(float);
}

template <typename T, int LP>
int launch_attn_bwd_tc(const void* qkv, const void* d_out, void* d_qkv, int batch, int L, int heads, int causal, cudaStream_t stream) {
    constexpr size_t smem = attn_bwd_tc_smem<T, LP>();
    static cudaError_t attr = cudaFuncSetAttribute(attention_backward_tc_kernel<T, LP>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    B2C_CUDA(attr);
    attention_backward_tc_kernel<T, LP><<<batch * heads, 256, smem, stream>>>(static_cast<const T*>(qkv), static_cast<const T*>(d_out), static_cast<T*>(d_qkv), L,
                                                                              heads, causal);
    B2C_LAUNCH_CHECK("attention_backward_tc_kernel");
    return 0;
}

// ------------------------------------------------------------------------------------------------------------------------
// y = x / max(||x||, eps):  dx = (g - y (y . g)) / max(||x||, eps); one warp per row
// ------------------------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) normalize_backward_kernel(const T* __restrict__ x, const T* __restrict__ g, T* __restrict__ dx, int rows, int dim, float eps) {
    const int r = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (r >= rows) return;
    const T* xr = x + static_cast<int64_t>(r) * dim;
    const T* gr = g + static_cast<int64_t>(r) * dim;
    float ss = 0.f, dot = 0.f;
    for (int j = lane; j < dim; j += 32) {
        const float v = to_f<T>(xr[j]);
        ss = fmaf(v, v, ss);
        dot = fmaf(v, to_f<T>(gr[j]), dot);
    }
    const float nrm = fmaxf(sqrtf(warp_sum(ss)), eps);
    dot = warp_sum(dot) / (nrm * nrm);            // (y . g) / ||x||
    for (int j = lane; j < dim; j += 32) dx[static_cast<int64_t>(r) * dim + j] = from_f<T>((to_f<T>(gr[j]) - to_f<T>(xr[j]) * dot) / nrm);
}

// ------------------------------------------------------------------------------------------------------------------------
// embedding backward
// ------------------------------------------------------------------------------------------------------------------------
// d_pos[l][c] (+)= sum_t dx[(t * L + l)][c]: deterministic (one thread per (l, c))
template <typename T>
__global__ void __launch_bounds__(256) period_sum_kernel(const T* __restrict__ dx, int64_t ld, int groups, int L, int width, float* __restrict__ out,
                                                         int64_t ldo, int accumulate) {
    const int c = blockIdx.x * 256 + threadIdx.x;
    const int l = blockIdx.y;
    if (c >= width) return;
    float s = 0.f;
    for (int t = 0; t < groups; ++t) s += to_f<T>(dx[(static_cast<int64_t>(t) * L + l) * ld + c]);
    float* o = out + static_cast<int64_t>(l) * ldo + c;
    *o = accumulate ? *o + s : s;
}
// d_tok[text[t][l]][c] += dx[(t * L + l)][c]  (fp32 atomics: the only non-deterministic sum of the backward)
template <typename T>
__global__ void __launch_bounds__(256) token_scatter_kernel(const T* __restrict__ dx, int64_t ld, const int64_t* __restrict__ text, int ctx, int T_, int L,
                                                            int width, int vocab, float* __restrict__ d_tok) {
    const int row = blockIdx.x;                  // t * L + l
    const int t = row / L, l = row % L;
    const int64_t id = text[static_cast<int64_t>(t) * ctx + l];
    if (id < 0 || id >= vocab) return;
    for (int c = threadIdx.x; c < width; c += 256) atomicAdd(d_tok + id * width + c, to_f<T>(dx[static_cast<int64_t>(row) * ld + c]));
}

template <typename TI, typename TO>
__global__ void __launch_bounds__(256) convert_kernel(const TI* __restrict__ in, TO* __restrict__ out, int64_t n) {
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += static_cast<int64_t>(gridDim.x) * blockDim.x)
        out[i] = from_f<TO>(to_f<TI>(in[i]));
}

inline int grid_for(int64_t n, int per_block = 256, int max_blocks = 148 * 16) {
    int64_t b = (n + per_block - 1) / per_block;
    if (b > max_blocks) b = max_blocks;
    if (b < 1) b = 1;
    return static_cast<int>(b);
}

#define B2C_DISPATCH_T(dtype, ...)                                           \
    do {                                                                     \
        if ((dtype) == 0) { using T = float; __VA_ARGS__; }                  \
        else if ((dtype) == 1) { using T = __nv_bfloat16; __VA_ARGS__; }     \
        else if ((dtype) == 2) { using T = __half; __VA_ARGS__; }            \
        else { B2C_CHECK_ARG(false, "unknown dtype %d", (dtype)); }          \
    } while (0)

}  // namespace

// out[c][r] = in[r][c] for r < R, 0 for R <= r < Rpad (Rpad <= ldo: zero padding of the contraction dimension of a GEMM operand)
// act: 0 = plain transpose, 1 = erf GELU, 2 = QuickGELU applied to every element on the way
int transpose16(int dtype, const void* in, int64_t ldi, void* out, int64_t ldo, int R, int C, int Rpad, cudaStream_t stream, int act) {
    B2C_CHECK_ARG(in && out && R > 0 && C > 0 && Rpad >= R && Rpad <= ldo && act >= 0 && act <= 2, "transpose: bad arguments");
    dim3 grid((C + 63) / 64, (Rpad + 63) / 64);
    B2C_DISPATCH_T(dtype, (transpose_kernel<T><<<grid, 256, 0, stream>>>(static_cast<const T*>(in), ldi, static_cast<T*>(out), ldo, R, C, Rpad, act)));
    B2C_LAUNCH_CHECK("transpose_kernel");
    return 0;
}

int64_t col_sum_scratch_floats(int rows, int cols) { return static_cast<int64_t>((rows + kColChunkRows - 1) / kColChunkRows) * cols; }

// out[c] (+)= sum_r g[r][c]; out_dtype: 0 = fp32, otherwise the storage type `dtype`
int col_sum(int dtype, const void* g, int64_t ld, int rows, int cols, void* out, int out_f32, int accumulate, float* scratch, cudaStream_t stream,
            unsigned int* counters) {
    B2C_CHECK_ARG(g && out && scratch && rows > 0 && cols > 0, "col_sum: bad arguments");
    const int chunks = (rows + kColChunkRows - 1) / kColChunkRows;
    dim3 grid((cols + 63) / 64, chunks);
    const int V = dtype == 0 ? 4 : 8;
    if (counters != nullptr && cols % V == 0 && ld % V == 0 && reinterpret_cast<uintptr_t>(g) % 16 == 0 && (cols + 32 * V - 1) / (32 * V) <= kColSumCounters) {
        // one launch, 16-byte loads: the last row chunk of every column block finishes the sum
        const dim3 vgrid((cols + 32 * V - 1) / (32 * V), (rows + kColVecChunkRows - 1) / kColVecChunkRows);
        if (out_f32 || dtype == 0) {
            B2C_DISPATCH_T(dtype, (colsum_vec_kernel<T, float><<<vgrid, 256, 0, stream>>>(static_cast<const T*>(g), ld, rows, cols, scratch, counters,
                                                                                            static_cast<float*>(out), accumulate)));
        } else {
            B2C_DISPATCH_T(dtype, (colsum_vec_kernel<T, T><<<vgrid, 256, 0, stream>>>(static_cast<const T*>(g), ld, rows, cols, scratch, counters,
                                                                                        static_cast<T*>(out), accumulate)));
        }
        B2C_LAUNCH_CHECK("colsum_vec_kernel");
        return 0;
    }
    B2C_DISPATCH_T(dtype, (colsum_partial_kernel<T, float><<<grid, 256, 0, stream>>>(static_cast<const T*>(g), ld, rows, cols, scratch, nullptr,
                                                                                       static_cast<float*>(nullptr), 0)));
    B2C_LAUNCH_CHECK("colsum_partial_kernel");
    if (out_f32 || dtype == 0) {
        colsum_final_kernel<float><<<(cols + 31) / 32, 256, 0, stream>>>(scratch, chunks, cols, static_cast<float*>(out), accumulate);
    } else {
        B2C_DISPATCH_T(dtype, (colsum_final_kernel<T><<<(cols + 31) / 32, 256, 0, stream>>>(scratch, chunks, cols, static_cast<T*>(out), accumulate)));
    }
    B2C_LAUNCH_CHECK("colsum_final_kernel");
    return 0;
}

int64_t ln_backward_scratch_floats(int rows, int width) { return 2 * static_cast<int64_t>((rows + kLnRowsPerBlock - 1) / kLnRowsPerBlock) * width; }

// dx = LayerNorm backward of g through x (+ dres); d_gamma, d_beta (fp32, += when `accumulate`)
int ln_backward(int dtype, const void* g, int64_t ldg, const void* x, int64_t ldx, const float* gamma, const void* dres, int64_t ldr, void* dx,
                int64_t ldd, float* d_gamma, float* d_beta, int rows, int width, float eps, int row_stride, const int32_t* row_idx, int accumulate,
                float* scratch, cudaStream_t stream) {
    B2C_CHECK_ARG(g && x && gamma && dx && d_gamma && d_beta && scratch && rows > 0 && width > 0, "ln_backward: bad arguments");
    const int blocks = (rows + kLnRowsPerBlock - 1) / kLnRowsPerBlock;
    const size_t smem = 2 * static_cast<size_t>(kLnWarps) * width * sizeof(float);
    B2C_CHECK_ARG(smem <= 48 * 1024, "ln_backward: width %d too large", width);
    float* pg = scratch;
    float* pb = scratch + static_cast<int64_t>(blocks) * width;
    const int rs = row_stride > 0 ? row_stride : 1;
    B2C_DISPATCH_T(dtype, {
        if (!launch_ln_backward_vec<T>(static_cast<const T*>(g), ldg, static_cast<const T*>(x), ldx, gamma, static_cast<const T*>(dres), ldr,
                                       static_cast<T*>(dx), ldd, pg, pb, rows, width, eps, rs, row_idx, blocks, smem, stream))
            ln_backward_kernel<T><<<blocks, kLnWarps * 32, smem, stream>>>(static_cast<const T*>(g), ldg, static_cast<const T*>(x), ldx, gamma,
                                                                           static_cast<const T*>(dres), ldr, static_cast<T*>(dx), ldd, pg, pb, rows,
                                                                           width, eps, rs, row_idx);
    });
    B2C_LAUNCH_CHECK("ln_backward_kernel");
    colsum_final_kernel<float><<<(width + 31) / 32, 256, 0, stream>>>(pg, blocks, width, d_gamma, accumulate);
    B2C_LAUNCH_CHECK("colsum_final_kernel");
    colsum_final_kernel<float><<<(width + 31) / 32, 256, 0, stream>>>(pb, blocks, width, d_beta, accumulate);
    B2C_LAUNCH_CHECK("colsum_final_kernel");
    return 0;
}

int act_backward(int dtype, const void* da, const void* z, void* dz, int64_t n, int quick, cudaStream_t stream) {
    B2C_CHECK_ARG(da && z && dz && n > 0, "act_backward: bad arguments");
    const int vec = ((reinterpret_cast<uintptr_t>(da) | reinterpret_cast<uintptr_t>(z) | reinterpret_cast<uintptr_t>(dz)) & 15) == 0;
    B2C_DISPATCH_T(dtype, (act_backward_kernel<T><<<grid_for(n / 4), 256, 0, stream>>>(static_cast<const T*>(da), static_cast<const T*>(z), static_cast<T*>(dz), n, quick, vec)));
    B2C_LAUNCH_CHECK("act_backward_kernel");
    return 0;
}

int act_forward(int dtype, const void* z, void* a, int64_t n, int quick, cudaStream_t stream) {
    B2C_CHECK_ARG(z && a && n > 0, "act_forward: bad arguments");
    const int vec = ((reinterpret_cast<uintptr_t>(z) | reinterpret_cast<uintptr_t>(a)) & 15) == 0;
    B2C_DISPATCH_T(dtype, (act_forward_kernel<T><<<grid_for(n / 4), 256, 0, stream>>>(static_cast<const T*>(z), static_cast<T*>(a), n, quick, vec)));
    B2C_LAUNCH_CHECK("act_forward_kernel");
    return 0;
}

int attention_backward(int dtype, const void* qkv, const void* d_out, void* d_qkv, int batch, int seq_len, int heads, int causal, cudaStream_t stream) {
    B2C_CHECK_ARG(qkv && d_out && d_qkv && batch > 0 && seq_len > 0 && heads > 0, "attention_backward: bad arguments");
    const int L = seq_len;
    // 16-bit modes, short sequences: the tensor-core kernel (B200CLIP_ATTN_BWD_SIMT=1 keeps the fp32 SIMT kernel for A/B checks)
    static const bool force_simt = [] {
        const char* e = getenv("B200CLIP_ATTN_BWD_SIMT");
        return e != nullptr && e[0] == '1';
    }();
    if (dtype != 0 && L <= 80 && !force_simt) {
        if (dtype == 1) return L <= 64 ? launch_attn_bwd_tc<__nv_bfloat16, 64>(qkv, d_out, d_qkv, batch, L, heads, causal, stream)
                                       : launch_attn_bwd_tc<__nv_bfloat16, 80>(qkv, d_out, d_qkv, batch, L, heads, causal, stream);
        return L <= 64 ? launch_attn_bwd_tc<__half, 64>(qkv, d_out, d_qkv, batch, L, heads, causal, stream)
                       : launch_attn_bwd_tc<__half, 80>(qkv, d_out, d_qkv, batch, L, heads, causal, stream);
    }
    const size_t es = dtype == 0 ? 4 : 2;
    const size_t smem = (2 * static_cast<size_t>(L) * (kHd + 1) + 2 * kTQ * (kHd + 1) + 2 * kTQ * (L + 1)) * sizeof(float) + 2 * static_cast<size_t>(L) * (kHd + 2) * es;
    B2C_CHECK_ARG(smem <= 227 * 1024, "attention_backward: sequence length %d too long for the shared-memory kernel (%zu bytes)", L, smem);
    const int grid = batch * heads;
    if (dtype == 0) {
        static cudaError_t e0 = cudaFuncSetAttribute(attention_backward_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        B2C_CUDA(e0);
        attention_backward_kernel<float><<<grid, kAttnThreads, smem, stream>>>(static_cast<const float*>(qkv), static_cast<const float*>(d_out),
                                                                               static_cast<float*>(d_qkv), L, heads, causal);
    } else if (dtype == 1) {
        static cudaError_t e1 = cudaFuncSetAttribute(attention_backward_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        B2C_CUDA(e1);
        attention_backward_kernel<__nv_bfloat16><<<grid, kAttnThreads, smem, stream>>>(
            static_cast<const __nv_bfloat16*>(qkv), static_cast<const __nv_bfloat16*>(d_out), static_cast<__nv_bfloat16*>(d_qkv), L, heads, causal);
    } else if (dtype == 2) {
        static cudaError_t e2 = cudaFuncSetAttribute(attention_backward_kernel<__half>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        B2C_CUDA(e2);
        attention_backward_kernel<__half><<<grid, kAttnThreads, smem, stream>>>(static_cast<const __half*>(qkv), static_cast<const __half*>(d_out),
                                                                                static_cast<__half*>(d_qkv), L, heads, causal);
    } else {
        B2C_CHECK_ARG(false, "attention_backward: unknown dtype %d", dtype);
    }
    B2C_LAUNCH_CHECK("attention_backward_kernel");
    return 0;
}

int normalize_backward(int dtype, const void* x, const void* g, void* dx, int rows, int dim, float eps, cudaStream_t stream) {
    B2C_CHECK_ARG(x && g && dx && rows > 0 && dim > 0, "normalize_backward: bad arguments");
    B2C_DISPATCH_T(dtype, (normalize_backward_kernel<T><<<(rows + 7) / 8, 256, 0, stream>>>(static_cast<const T*>(x), static_cast<const T*>(g), static_cast<T*>(dx), rows, dim, eps)));
    B2C_LAUNCH_CHECK("normalize_backward_kernel");
    return 0;
}

int period_sum(int dtype, const void* dx, int64_t ld, int groups, int L, int width, float* out, int64_t ldo, int accumulate, cudaStream_t stream) {
    B2C_CHECK_ARG(dx && out && groups > 0 && L > 0 && width > 0, "period_sum: bad arguments");
    dim3 grid((width + 255) / 256, L);
    B2C_DISPATCH_T(dtype, (period_sum_kernel<T><<<grid, 256, 0, stream>>>(static_cast<const T*>(dx), ld, groups, L, width, out, ldo, accumulate)));
    B2C_LAUNCH_CHECK("period_sum_kernel");
    return 0;
}

int token_scatter(int dtype, const void* dx, int64_t ld, const int64_t* text, int ctx, int T_, int L, int width, int vocab, float* d_tok, cudaStream_t stream) {
    B2C_CHECK_ARG(dx && text && d_tok && T_ > 0 && L > 0 && width > 0, "token_scatter: bad arguments");
    B2C_DISPATCH_T(dtype, (token_scatter_kernel<T><<<T_ * L, 256, 0, stream>>>(static_cast<const T*>(dx), ld, text, ctx, T_, L, width, vocab, d_tok)));
    B2C_LAUNCH_CHECK("token_scatter_kernel");
    return 0;
}

// fp32 <-> storage type element-wise copy (gradient hand-over in the mixed-precision modes)
int convert(int dtype_in, const void* in, int dtype_out, void* out, int64_t n, cudaStream_t stream) {
    B2C_CHECK_ARG(in && out && n > 0, "convert: bad arguments");
    const int g = grid_for(n);
    if (dtype_in == 0 && dtype_out == 1) convert_kernel<float, __nv_bfloat16><<<g, 256, 0, stream>>>(static_cast<const float*>(in), static_cast<__nv_bfloat16*>(out), n);
    else if (dtype_in == 0 && dtype_out == 2) convert_kernel<float, __half><<<g, 256, 0, stream>>>(static_cast<const float*>(in), static_cast<__half*>(out), n);
    else if (dtype_in == 1 && dtype_out == 0) convert_kernel<__nv_bfloat16, float><<<g, 256, 0, stream>>>(static_cast<const __nv_bfloat16*>(in), static_cast<float*>(out), n);
    else if (dtype_in == 2 && dtype_out == 0) convert_kernel<__half, float><<<g, 256, 0, stream>>>(static_cast<const __half*>(in), static_cast<float*>(out), n);
    else if (dtype_in == 0 && dtype_out == 0) convert_kernel<float, float><<<g, 256, 0, stream>>>(static_cast<const float*>(in), static_cast<float*>(out), n);
    else B2C_CHECK_ARG(false, "convert: unsupported dtype pair %d -> %d", dtype_in, dtype_out);
    B2C_LAUNCH_CHECK("convert_kernel");
    return 0;
}

}  // namespace b200clip
