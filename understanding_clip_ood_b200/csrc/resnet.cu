// ModifiedResNet image tower (the RN50 family of CLIP), eval mode: one C-ABI call per CLIP.encode_image.
//
//   resnet_forward = ModifiedResNet.forward   (deps/open_clip/src/open_clip/modified_resnet.py:95-181)
//                    Bottleneck.forward       (modified_resnet.py:42-56)
//                    AttentionPool2d.forward  (modified_resnet.py:69-92)
//
// Activations live in HBM as NHWC rows ([B*H*W, C], C contiguous), so every 1x1 convolution is the tensor-core GEMM of the
// transformer towers (gemm_pair_kernel: tcgen05, TMA-fed) over those rows, and a 3x3 convolution is the same GEMM over an
// im2col of the rows with K order (ky, kx, cin).  BatchNorm (inference statistics) is folded into the convolution weights
// and a per-channel shift by the caller; shift, ReLU and the bottleneck's residual add run in the GEMM epilogues
// (B200CLIP_EPI_RELU / B200CLIP_EPI_RESIDUAL_RELU), so a bottleneck is conv1 GEMM -> im2col -> conv2 GEMM -> [avgpool] ->
// conv3 GEMM (+ identity + ReLU), plus [avgpool ->] downsample GEMM on the first block of a stage.  The attention pool is the
// packed-QKV attention of the ViT tower with L = HW + 1 tokens; only the class (mean) token's query is projected, and the
// output projection reads that token's row of every image through the GEMM's row pitch.
//
// The driver only enqueues kernels on the caller's stream into a caller-provided workspace (CUDA-graph capturable).
#include "../../include/b200clip.h"
#include "common.cuh"
#include "internal.h"

#include <algorithm>
#include <cstdlib>

namespace b200clip {

namespace {

inline int64_t align_up(int64_t v, int64_t a) { return (v + a - 1) / a * a; }

template <typename T> __device__ __forceinline__ float to_float(T v);
template <> __device__ __forceinline__ float to_float<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_float<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <> __device__ __forceinline__ float to_float<__half>(__half v) { return __half2float(v); }
template <typename T> __device__ __forceinline__ T from_float(float v);
template <> __device__ __forceinline__ float from_float<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_float<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }
template <> __device__ __forceinline__ __half from_float<__half>(float v) { return __float2half_rn(v); }

// 16-byte vector of T (4 floats or 8 16-bit values)
template <typename T> struct Vec16 {
    static constexpr int kN = 16 / sizeof(T);
    T v[kN];
};
template <typename T> __device__ __forceinline__ Vec16<T> load16(const T* p) {
    Vec16<T> r;
    *reinterpret_cast<uint4*>(r.v) = __ldg(reinterpret_cast<const uint4*>(p));
    return r;
}
template <typename T> __device__ __forceinline__ void store16(T* p, const Vec16<T>& r) {
    *reinterpret_cast<uint4*>(p) = *reinterpret_cast<const uint4*>(r.v);
}

// ---- stem conv1 (3x3, stride 2, padding 1) im2col straight from the NCHW batch -------------------------------------------
// out[(b*Ho + oy)*Wo + ox, (ky*3 + kx)*3 + c] = image[b, c, 2*oy + ky - 1, 2*ox + kx - 1] (0 outside), columns 27..kpad-1 = 0.
// One thread per 16-byte vector of an output row.
template <typename T>
__global__ void __launch_bounds__(256) stem_im2col_kernel(const T* __restrict__ image, T* __restrict__ out, int64_t total_vec, int S,
                                                          int Ho, int kpad) {
    constexpr int kN = Vec16<T>::kN;
    const int vpr = kpad / kN;
    for (int64_t idx = blockIdx.x * 256ll + threadIdx.x; idx < total_vec; idx += gridDim.x * 256ll) {
        const int v = static_cast<int>(idx % vpr);
        const int64_t row = idx / vpr;
        const int ox = static_cast<int>(row % Ho);
        const int oy = static_cast<int>((row / Ho) % Ho);
        const int64_t b = row / (static_cast<int64_t>(Ho) * Ho);
        const T* img = image + b * 3 * S * S;
        Vec16<T> r;
#pragma unroll
        for (int e = 0; e < kN; ++e) {
            const int k = v * kN + e;
            T val = from_float<T>(0.f);
            if (k < 27) {
                const int tap = k / 3, c = k - tap * 3;
                const int iy = 2 * oy + tap / 3 - 1, ix = 2 * ox + tap % 3 - 1;
                if (iy >= 0 && iy < S && ix >= 0 && ix < S) val = __ldg(img + (static_cast<int64_t>(c) * S + iy) * S + ix);
            }
            r.v[e] = val;
        }
        store16(out + idx * kN, r);
    }
}

// ---- 3x3 / stride 1 / padding 1 im2col over NHWC rows ------------------------------------------------------------------------
// in [B*H*W, C] -> out [B*H*W, 9*C], out[row, tap*C + c] = in[row shifted by (tap/3 - 1, tap%3 - 1), c] (0 outside the image).
// One thread per (row, 16-byte vector of the C channels): it derives (x, y) once and moves all nine taps (nine independent loads in
// flight).  Pure data movement, so it is typed by vector only.  (The first version did one thread per OUTPUT vector with four 64-bit
// divisions each: issue-bound at 2.9 TB/s of traffic, 44 % of the RN50 forward.)
__global__ void __launch_bounds__(256) im2col3x3_kernel(const uint4* __restrict__ in, uint4* __restrict__ out, int64_t n_rowvec, int H, int W,
                                                         int cv, int cv_shift) {
    for (int64_t i = blockIdx.x * 256ll + threadIdx.x; i < n_rowvec; i += gridDim.x * 256ll) {
        int64_t row;
        int v;
        if (cv_shift >= 0) {
            row = i >> cv_shift;
            v = static_cast<int>(i & (cv - 1));
        } else {
            row = i / cv;
            v = static_cast<int>(i - row * cv);
        }
        const uint32_t r32 = static_cast<uint32_t>(row);      // rows = B * H * W < 2^32 (checked by the caller)
        const uint32_t q = r32 / static_cast<uint32_t>(W);
        const int x = static_cast<int>(r32 - q * static_cast<uint32_t>(W));
        const int y = static_cast<int>(q % static_cast<uint32_t>(H));
        const uint4* src = in + row * cv + v;
        uint4* dst = out + row * 9 * cv + v;
        uint4 val[9];
#pragma unroll
        for (int tap = 0; tap < 9; ++tap) {
            const int dy = tap / 3 - 1, dx = tap % 3 - 1;
            val[tap] = make_uint4(0u, 0u, 0u, 0u);
            if (static_cast<unsigned>(y + dy) < static_cast<unsigned>(H) && static_cast<unsigned>(x + dx) < static_cast<unsigned>(W))
                val[tap] = __ldg(src + (dy * W + dx) * cv);
        }
#pragma unroll
        for (int tap = 0; tap < 9; ++tap) dst[tap * cv] = val[tap];
    }
}

// ---- AvgPool2d(2) over NHWC rows: [B, H, W, C] -> [B, H/2, W/2, C], fp32 sum of the four taps, one rounding -----------------
template <typename T>
__global__ void __launch_bounds__(256) avgpool2_kernel(const T* __restrict__ in, T* __restrict__ out, int64_t total_vec, int Ho, int Wo,
                                                       int C) {
    constexpr int kN = Vec16<T>::kN;
    const int cv = C / kN;
    const int W = 2 * Wo;
    for (int64_t idx = blockIdx.x * 256ll + threadIdx.x; idx < total_vec; idx += gridDim.x * 256ll) {
        const int v = static_cast<int>(idx % cv);
        const int64_t orow = idx / cv;
        const int ox = static_cast<int>(orow % Wo);
        const int oy = static_cast<int>((orow / Wo) % Ho);
        const int64_t b = orow / (static_cast<int64_t>(Ho) * Wo);
        const int64_t irow = (b * 2 * Ho + 2 * oy) * W + 2 * ox;
        const T* p = in + irow * C + v * kN;
        const Vec16<T> a = load16(p), b2 = load16(p + C), c2 = load16(p + static_cast<int64_t>(W) * C),
                       d = load16(p + static_cast<int64_t>(W) * C + C);
        Vec16<T> r;
#pragma unroll
        for (int e = 0; e < kN; ++e)
            r.v[e] = from_float<T>((to_float(a.v[e]) + to_float(b2.v[e]) + to_float(c2.v[e]) + to_float(d.v[e])) * 0.25f);
        store16(out + idx * kN, r);
    }
}

// ---- AttentionPool2d token matrix (modified_resnet.py:70-72) -------------------------------------------------------------------
// tok[b*(HW+1) + 0, :]     = T(mean_i x[b, i, :]) + T(pos[0, :])
// tok[b*(HW+1) + 1 + i, :] = x[b, i, :] + T(pos[1 + i, :])           (sums rounded once to the tower dtype T, as the reference's
// elementwise ops on T tensors do; the mean is accumulated in fp32).  grid = (channel-vector blocks, B).
template <typename T>
__global__ void __launch_bounds__(128) attnpool_tokens_kernel(const T* __restrict__ x, const float* __restrict__ pos, T* __restrict__ tok,
                                                              int HW, int C) {
    constexpr int kN = Vec16<T>::kN;
    const int v = blockIdx.x * 128 + threadIdx.x;
    if (v * kN >= C) return;
    const int64_t b = blockIdx.y;
    const T* xb = x + b * HW * C + v * kN;
    T* tb = tok + b * (HW + 1) * C + v * kN;
    float acc[kN];
#pragma unroll
    for (int e = 0; e < kN; ++e) acc[e] = 0.f;
    for (int i = 0; i < HW; ++i) {
        const Vec16<T> a = load16(xb + static_cast<int64_t>(i) * C);
        const float* pr = pos + static_cast<int64_t>(1 + i) * C + v * kN;
        Vec16<T> r;
#pragma unroll
        for (int e = 0; e < kN; ++e) {
            const float xv = to_float(a.v[e]);
            acc[e] += xv;
            r.v[e] = from_float<T>(xv + to_float(from_float<T>(__ldg(pr + e))));
        }
        store16(tb + static_cast<int64_t>(1 + i) * C, r);
    }
    Vec16<T> r;
    const float inv = 1.f / static_cast<float>(HW);
#pragma unroll
    for (int e = 0; e < kN; ++e) {
        const float m = to_float(from_float<T>(acc[e] * inv));
        r.v[e] = from_float<T>(m + to_float(from_float<T>(__ldg(pos + v * kN + e))));
    }
    store16(tb, r);
}

// B200CLIP_CONV_IM2COL=1: every 3x3 convolution through im2col + GEMM (the first version of the tower; A/B measurements)
bool conv_via_im2col() {
    static const bool v = [] {
        const char* e = getenv("B200CLIP_CONV_IM2COL");
        return e != nullptr && e[0] == '1';
    }();
    return v;
}

inline int grid_for(int64_t total) {
    int64_t blocks = (total + 255) / 256;
    const int64_t cap = static_cast<int64_t>(num_sms()) * 16;
    return static_cast<int>(blocks < cap ? (blocks > 0 ? blocks : 1) : cap);
}

}  // namespace

int stem_im2col(int dtype, const void* image, void* out, int batch, int image_size, int kpad, cudaStream_t s) {
    B2C_CHECK_ARG(image != nullptr && out != nullptr && batch > 0 && image_size > 0 && image_size % 2 == 0,
                  "stem_im2col: bad arguments (batch=%d image=%d)", batch, image_size);
    const int es = dtype_size(dtype);
    B2C_CHECK_ARG(kpad >= 27 && (kpad * es) % 16 == 0, "stem_im2col: kpad=%d must cover 27 taps in whole 16-byte vectors", kpad);
    B2C_CHECK_ARG(reinterpret_cast<uintptr_t>(out) % 16 == 0, "stem_im2col: output must be 16-byte aligned");
    const int Ho = image_size / 2;
    const int64_t total = static_cast<int64_t>(batch) * Ho * Ho * (kpad * es / 16);
    switch (dtype) {
        case 0: stem_im2col_kernel<float><<<grid_for(total), 256, 0, s>>>(static_cast<const float*>(image), static_cast<float*>(out), total, image_size, Ho, kpad); break;
        case 1: stem_im2col_kernel<__nv_bfloat16><<<grid_for(total), 256, 0, s>>>(static_cast<const __nv_bfloat16*>(image), static_cast<__nv_bfloat16*>(out), total, image_size, Ho, kpad); break;
        case 2: stem_im2col_kernel<__half><<<grid_for(total), 256, 0, s>>>(static_cast<const __half*>(image), static_cast<__half*>(out), total, image_size, Ho, kpad); break;
        default: set_last_error("stem_im2col: unknown dtype %d", dtype); return -1;
    }
    B2C_LAUNCH_CHECK("stem_im2col_kernel");
    return 0;
}

int im2col3x3(int dtype, const void* in, void* out, int batch, int H, int W, int C, cudaStream_t s) {
    B2C_CHECK_ARG(in != nullptr && out != nullptr && batch > 0 && H > 0 && W > 0 && C > 0, "im2col3x3: bad arguments");
    const int es = dtype_size(dtype);
    B2C_CHECK_ARG(dtype >= 0 && dtype <= 2 && (C * es) % 16 == 0, "im2col3x3: C=%d must give whole 16-byte vectors", C);
    B2C_CHECK_ARG((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out)) % 16 == 0, "im2col3x3: buffers must be 16-byte aligned");
    const int cv = C * es / 16;
    const int64_t rows = static_cast<int64_t>(batch) * H * W;
    B2C_CHECK_ARG(rows < (1ll << 32), "im2col3x3: %lld rows do not fit the 32-bit pixel arithmetic", static_cast<long long>(rows));
    int cv_shift = -1;
    for (int sh = 0; sh < 16; ++sh)
        if ((1 << sh) == cv) cv_shift = sh;
    const int64_t total = rows * cv;
    im2col3x3_kernel<<<grid_for(total), 256, 0, s>>>(static_cast<const uint4*>(in), static_cast<uint4*>(out), total, H, W, cv, cv_shift);
    B2C_LAUNCH_CHECK("im2col3x3_kernel");
    return 0;
}

int avgpool2(int dtype, const void* in, void* out, int batch, int H, int W, int C, cudaStream_t s) {
    B2C_CHECK_ARG(in != nullptr && out != nullptr && batch > 0 && H > 0 && W > 0 && H % 2 == 0 && W % 2 == 0 && C > 0,
                  "avgpool2: bad arguments (H=%d W=%d C=%d)", H, W, C);
    const int es = dtype_size(dtype);
    B2C_CHECK_ARG(dtype >= 0 && dtype <= 2 && (C * es) % 16 == 0, "avgpool2: C=%d must give whole 16-byte vectors", C);
    B2C_CHECK_ARG((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out)) % 16 == 0, "avgpool2: buffers must be 16-byte aligned");
    const int Ho = H / 2, Wo = W / 2;
    const int64_t total = static_cast<int64_t>(batch) * Ho * Wo * (C * es / 16);
    switch (dtype) {
        case 0: avgpool2_kernel<float><<<grid_for(total), 256, 0, s>>>(static_cast<const float*>(in), static_cast<float*>(out), total, Ho, Wo, C); break;
        case 1: avgpool2_kernel<__nv_bfloat16><<<grid_for(total), 256, 0, s>>>(static_cast<const __nv_bfloat16*>(in), static_cast<__nv_bfloat16*>(out), total, Ho, Wo, C); break;
        case 2: avgpool2_kernel<__half><<<grid_for(total), 256, 0, s>>>(static_cast<const __half*>(in), static_cast<__half*>(out), total, Ho, Wo, C); break;
    }
    B2C_LAUNCH_CHECK("avgpool2_kernel");
    return 0;
}

int attnpool_tokens(int dtype, const void* x, const float* pos, void* tok, int batch, int HW, int C, cudaStream_t s) {
    B2C_CHECK_ARG(x != nullptr && pos != nullptr && tok != nullptr && batch > 0 && HW > 0 && C > 0, "attnpool_tokens: bad arguments");
    const int es = dtype_size(dtype);
    B2C_CHECK_ARG(dtype >= 0 && dtype <= 2 && (C * es) % 16 == 0, "attnpool_tokens: C=%d must give whole 16-byte vectors", C);
    B2C_CHECK_ARG((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(tok)) % 16 == 0, "attnpool_tokens: buffers must be 16-byte aligned");
    const int vecs = C * es / 16;
    dim3 grid((vecs + 127) / 128, batch);
    switch (dtype) {
        case 0: attnpool_tokens_kernel<float><<<grid, 128, 0, s>>>(static_cast<const float*>(x), pos, static_cast<float*>(tok), HW, C); break;
        case 1: attnpool_tokens_kernel<__nv_bfloat16><<<grid, 128, 0, s>>>(static_cast<const __nv_bfloat16*>(x), pos, static_cast<__nv_bfloat16*>(tok), HW, C); break;
        case 2: attnpool_tokens_kernel<__half><<<grid, 128, 0, s>>>(static_cast<const __half*>(x), pos, static_cast<__half*>(tok), HW, C); break;
    }
    B2C_LAUNCH_CHECK("attnpool_tokens_kernel");
    return 0;
}

// ---- whole tower ------------------------------------------------------------------------------------------------------------------
namespace {

struct RnBuffers {
    char* col;   // im2col rows of the 3x3 convolutions (largest: the stem's 112 x 112 x 9*32, layer2.0's 56 x 56 x 9*128)
    char* x[2];  // block input / output, ping-pong
    char* t[2];  // conv1 / conv2 outputs (stem: conv outputs)
    char* p[2];  // average-pooled conv2 output / average-pooled block input (first block of a strided stage)
    char* id;    // downsample branch output
    char* tok;   // attention-pool tokens [B*L, C]
    char* qkv;   // [B*L, 3C]
    char* att;   // [B*L, C]
    void* sk;    // stream-K workspace of the CTA-pair GEMM
    int64_t total;
};

struct RnSizes {   // elements per image
    int64_t col = 0, x = 0, t = 0, p = 0, id = 0;
};

int check_rn(const b200clip_resnet_cfg* c, const b200clip_resnet_weights* w) {
    B2C_CHECK_ARG(c != nullptr && w != nullptr && w->blocks_host != nullptr, "resnet: null cfg / weights");
    B2C_CHECK_ARG(c->dtype >= 0 && c->dtype <= 2, "resnet: unknown dtype %d", c->dtype);
    const int vec = 16 / dtype_size(c->dtype);
    B2C_CHECK_ARG(c->width > 0 && (c->width / 2) % vec == 0, "resnet: width %d must make width/2 a whole number of 16-byte vectors", c->width);
    B2C_CHECK_ARG(c->image_size > 0 && c->image_size % 32 == 0, "resnet: image_size %d must be a multiple of 32", c->image_size);
    B2C_CHECK_ARG(c->heads * 64 == c->width * 32, "resnet: attention pool needs head width 64 (heads=%d, embed=%d)", c->heads, c->width * 32);
    B2C_CHECK_ARG(c->n_blocks > 0 && c->embed_dim > 0 && c->embed_dim % 8 == 0, "resnet: bad n_blocks / embed_dim");
    B2C_CHECK_ARG(c->stem_kpad >= 27 && (c->stem_kpad * dtype_size(c->dtype)) % 16 == 0, "resnet: bad stem_kpad %d", c->stem_kpad);
    return 0;
}

// per-image buffer sizes, by walking the stage schedule
int plan_rn(const b200clip_resnet_cfg& c, const b200clip_resnet_weights& w, RnSizes& z) {
    const int s1 = c.image_size / 2;
    const int64_t r1 = static_cast<int64_t>(s1) * s1;
    const int half = c.width / 2;
    auto up = [](int64_t& a, int64_t b) { if (b > a) a = b; };
    up(z.col, r1 * c.stem_kpad);
    up(z.col, r1 * 9 * half);
    up(z.t, r1 * c.width);
    int H = s1 / 2;
    int cin = c.width;
    up(z.x, static_cast<int64_t>(H) * H * cin);
    for (int i = 0; i < c.n_blocks; ++i) {
        const b200clip_resnet_block& b = w.blocks_host[i];
        B2C_CHECK_ARG(b.cin == cin && b.planes > 0 && (b.stride == 1 || b.stride == 2), "resnet: block %d has cin=%d (expected %d), planes=%d, stride=%d",
                      i, b.cin, cin, b.planes, b.stride);
        B2C_CHECK_ARG(b.stride == 1 || H % 2 == 0, "resnet: block %d halves an odd resolution %d", i, H);
        B2C_CHECK_ARG((b.down_w != nullptr) == (b.stride > 1 || cin != 4 * b.planes), "resnet: block %d downsample weights do not match its shape", i);
        B2C_CHECK_ARG(b.conv1_w && b.conv1_b && b.conv2_w && b.conv2_b && b.conv3_w && b.conv3_b && (b.down_w == nullptr || b.down_b != nullptr),
                      "resnet: block %d has null weights", i);
        const int64_t rows = static_cast<int64_t>(H) * H;
        up(z.t, rows * b.planes);
        up(z.col, rows * 9 * b.planes);
        const int Ho = H / b.stride;
        const int64_t orows = static_cast<int64_t>(Ho) * Ho;
        if (b.stride > 1) up(z.p, std::max(orows * b.planes, orows * cin));
        if (b.down_w != nullptr) up(z.id, orows * 4 * b.planes);
        up(z.x, orows * 4 * b.planes);
        H = Ho;
        cin = 4 * b.planes;
    }
    B2C_CHECK_ARG(cin == c.width * 32 && H == c.image_size / 32, "resnet: the stages end at %d channels, %d x %d (expected %d, %d)", cin, H, H,
                  c.width * 32, c.image_size / 32);
    return 0;
}

int carve_rn(const b200clip_resnet_cfg& c, const b200clip_resnet_weights& w, int batch, void* base, RnBuffers& bf) {
    RnSizes z;
    int rc;
    if ((rc = plan_rn(c, w, z)) != 0) return rc;
    const int64_t es = dtype_size(c.dtype);
    const int sp = c.image_size / 32;
    const int64_t L = static_cast<int64_t>(sp) * sp + 1;
    const int64_t E = static_cast<int64_t>(c.width) * 32;
    int64_t off = 0;
    char* b = static_cast<char*>(base);
    auto take = [&](int64_t bytes) {
        char* p = b ? b + off : nullptr;
        off += align_up(bytes > 0 ? bytes : 16, 256);
        return p;
    };
    bf.col = take(batch * z.col * es);
    bf.x[0] = take(batch * z.x * es);
    bf.x[1] = take(batch * z.x * es);
    bf.t[0] = take(batch * z.t * es);
    bf.t[1] = take(batch * z.t * es);
    bf.p[0] = take(batch * z.p * es);
    bf.p[1] = take(batch * z.p * es);
    bf.id = take(batch * z.id * es);
    bf.tok = take(batch * L * E * es);
    bf.qkv = take(batch * L * 3 * E * es);
    bf.att = take(batch * L * E * es);
    bf.sk = c.dtype != B200CLIP_F32 ? take(gemm_pair_sk_workspace_bytes()) : nullptr;
    bf.total = off;
    return 0;
}

}  // namespace

int64_t resnet_workspace_bytes(const b200clip_resnet_cfg* cfg, const b200clip_resnet_weights* w, int batch) {
    if (check_rn(cfg, w) != 0 || batch <= 0) return -1;
    RnBuffers bf;
    if (carve_rn(*cfg, *w, batch, nullptr, bf) != 0) return -1;
    return bf.total;
}

int resnet_forward_stages(const b200clip_resnet_cfg* cfg, const b200clip_resnet_weights* w, const void* image, void* out, int batch,
                          int normalize, void* workspace, int64_t workspace_bytes_, int stages, cudaStream_t s) {
    int rc;
    if ((rc = check_rn(cfg, w)) != 0) return rc;
    const b200clip_resnet_cfg& c = *cfg;
    B2C_CHECK_ARG(stages > 0 && stages <= 7, "resnet_forward: bad stage mask %d", stages);
    B2C_CHECK_ARG(batch > 0 && workspace != nullptr && reinterpret_cast<uintptr_t>(workspace) % 256 == 0,
                  "resnet_forward: empty batch or missing / misaligned workspace");
    B2C_CHECK_ARG(!(stages & B200CLIP_STAGE_INPUT) || image != nullptr, "resnet_forward: null image");
    B2C_CHECK_ARG(!(stages & B200CLIP_STAGE_OUTPUT) || out != nullptr, "resnet_forward: null output");
    B2C_CHECK_ARG(w->stem_w[0] && w->stem_w[1] && w->stem_w[2] && w->stem_b[0] && w->stem_b[1] && w->stem_b[2] && w->pos && w->qkv_w && w->qkv_b &&
                      w->c_proj_w && w->c_proj_b, "resnet_forward: null stem / attention-pool weights");
    RnBuffers bf;
    if ((rc = carve_rn(c, *w, batch, workspace, bf)) != 0) return rc;
    B2C_CHECK_ARG(bf.total <= workspace_bytes_, "resnet_forward: workspace too small (%lld < %lld bytes)", (long long)workspace_bytes_,
                  (long long)bf.total);
    const int dt = c.dtype;
    const int half = c.width / 2;
    const int s1 = c.image_size / 2;
    const int64_t rows1 = static_cast<int64_t>(batch) * s1 * s1;
    B2C_CHECK_ARG(rows1 < (1ll << 31), "resnet_forward: batch %d too large for 32-bit row indices", batch);
    const int sp = c.image_size / 32;
    const int L = sp * sp + 1;
    const int E = c.width * 32;

    // conv (+ folded BatchNorm shift) + ReLU as a GEMM over NHWC rows
    auto conv = [&](const void* A, int K, const void* W, const void* b, void* C, int64_t M, int N, int epi, const void* res) {
        return gemm_any(dt, A, K, W, K, b, res, N, C, N, static_cast<int>(M), N, K, epi, nullptr, 0, 0, s, bf.sk);
    };

    if (stages & B200CLIP_STAGE_INPUT) {
        if ((rc = stem_im2col(dt, image, bf.col, batch, c.image_size, c.stem_kpad, s)) != 0) return rc;
    }
    if (stages & B200CLIP_STAGE_BODY) {
        if (bf.sk != nullptr && (rc = gemm_pair_sk_workspace_reset(bf.sk, s)) != 0) return rc;
        // stem (modified_resnet.py:163-168): three conv-bn-relu, then AvgPool2d(2)
        if ((rc = conv(bf.col, c.stem_kpad, w->stem_w[0], w->stem_b[0], bf.t[0], rows1, half, B200CLIP_EPI_RELU, nullptr)) != 0) return rc;
        if ((rc = im2col3x3(dt, bf.t[0], bf.col, batch, s1, s1, half, s)) != 0) return rc;
        if ((rc = conv(bf.col, 9 * half, w->stem_w[1], w->stem_b[1], bf.t[1], rows1, half, B200CLIP_EPI_RELU, nullptr)) != 0) return rc;
        if ((rc = im2col3x3(dt, bf.t[1], bf.col, batch, s1, s1, half, s)) != 0) return rc;
        if ((rc = conv(bf.col, 9 * half, w->stem_w[2], w->stem_b[2], bf.t[0], rows1, c.width, B200CLIP_EPI_RELU, nullptr)) != 0) return rc;
        if ((rc = avgpool2(dt, bf.t[0], bf.x[0], batch, s1, s1, c.width, s)) != 0) return rc;
        int H = s1 / 2;
        int cin = c.width;
        int cur = 0;
        for (int i = 0; i < c.n_blocks; ++i) {
            const b200clip_resnet_block& b = w->blocks_host[i];
            const int64_t rows = static_cast<int64_t>(batch) * H * H;
            const int Ho = H / b.stride;
            const int64_t orows = static_cast<int64_t>(batch) * Ho * Ho;
            const int cout = 4 * b.planes;
            const char* x = bf.x[cur];
            char* y = bf.x[cur ^ 1];
            // conv1 1x1 -> conv2 3x3 -> [avgpool] -> conv3 1x1 (+ identity, ReLU)     (modified_resnet.py:42-56)
            if ((rc = conv(x, cin, b.conv1_w, b.conv1_b, bf.t[0], rows, b.planes, B200CLIP_EPI_RELU, nullptr)) != 0) return rc;
            // conv2: implicit GEMM where the geometry allows it (16-bit, planes % 64 == 0, power-of-two pixel blocks): the A tiles are
            // read from the NHWC tensor itself, tap by tap; otherwise im2col + GEMM
            rc = 1;
            if (dt != B200CLIP_F32 && !conv_via_im2col())
                rc = gemm_pair_conv3x3(dt == B200CLIP_BF16, bf.t[0], b.conv2_w, b.conv2_b, bf.t[1], batch, H, H, b.planes, b.planes, s);
            if (rc == 1) {
                if ((rc = im2col3x3(dt, bf.t[0], bf.col, batch, H, H, b.planes, s)) != 0) return rc;
                rc = conv(bf.col, 9 * b.planes, b.conv2_w, b.conv2_b, bf.t[1], rows, b.planes, B200CLIP_EPI_RELU, nullptr);
            }
            if (rc != 0) return rc;
            const char* main_in = bf.t[1];
            if (b.stride > 1) {
                if ((rc = avgpool2(dt, bf.t[1], bf.p[0], batch, H, H, b.planes, s)) != 0) return rc;
                main_in = bf.p[0];
            }
            const char* identity = x;
            if (b.down_w != nullptr) {
                const char* down_in = x;
                if (b.stride > 1) {
                    if ((rc = avgpool2(dt, x, bf.p[1], batch, H, H, cin, s)) != 0) return rc;
                    down_in = bf.p[1];
                }
                if ((rc = conv(down_in, cin, b.down_w, b.down_b, bf.id, orows, cout, B200CLIP_EPI_BIAS, nullptr)) != 0) return rc;
                identity = bf.id;
            }
            if ((rc = conv(main_in, b.planes, b.conv3_w, b.conv3_b, y, orows, cout, B200CLIP_EPI_RESIDUAL_RELU, identity)) != 0) return rc;
            cur ^= 1;
            H = Ho;
            cin = cout;
        }
        // attention pool (modified_resnet.py:69-92): tokens, K/V projection of every token, Q projection of the mean token only
        // (row pitch L*E picks token 0 of every image), attention, and the mean token's row is what the output stage projects
        if ((rc = attnpool_tokens(dt, bf.x[cur], w->pos, bf.tok, batch, sp * sp, E, s)) != 0) return rc;
        const int64_t es = dtype_size(dt);
        const char* qkv_w = static_cast<const char*>(w->qkv_w);
        const char* qkv_b = static_cast<const char*>(w->qkv_b);
        if ((rc = gemm_any(dt, bf.tok, E, qkv_w + static_cast<int64_t>(E) * E * es, E, qkv_b + E * es, nullptr, 0, bf.qkv + E * es, 3 * E,
                           batch * L, 2 * E, E, B200CLIP_EPI_BIAS, nullptr, 0, 0, s, bf.sk)) != 0)
            return rc;
        if ((rc = gemm_any(dt, bf.tok, static_cast<int64_t>(L) * E, qkv_w, E, qkv_b, nullptr, 0, bf.qkv, static_cast<int64_t>(L) * 3 * E, batch, E, E,
                           B200CLIP_EPI_BIAS, nullptr, 0, 0, s, bf.sk)) != 0)
            return rc;
        if ((rc = attention(dt, bf.qkv, bf.att, batch, L, c.heads, 0, s)) != 0) return rc;
    }
    if (stages & B200CLIP_STAGE_OUTPUT) {
        if ((rc = gemm_any(dt, bf.att, static_cast<int64_t>(L) * E, w->c_proj_w, E, w->c_proj_b, nullptr, 0, out, c.embed_dim, batch, c.embed_dim, E,
                           B200CLIP_EPI_BIAS, nullptr, 0, 0, s)) != 0)
            return rc;
        if (normalize && (rc = normalize_rows(dt, out, c.embed_dim, out, c.embed_dim, batch, c.embed_dim, 1e-12f, s)) != 0) return rc;
    }
    return 0;
}

}  // namespace b200clip
