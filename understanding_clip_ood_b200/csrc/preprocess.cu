// Image preprocessing in front of the uint8 tower entry: bicubic shortest-side resize + centre crop of a decoded
// uint8 HWC image into the [3, S, S] uint8 pixel tensor that b200clip_vit_forward_u8 consumes (ToTensor + Normalize run there).
//
//   reference: torchvision Resize(S, BICUBIC) + CenterCrop(S) on PIL images (deps/open_clip/src/open_clip/transform.py:372-392),
//   i.e. Pillow's ImagingResample: separable two-pass convolution (horizontal, then vertical) with an anti-aliasing bicubic
//   window (support 2 x scale), coefficients normalised in double precision and quantised to 22-bit fixed point, 8-bit
//   intermediate, result = clip8((2^21 + sum_k pixel_k * coeff_k) >> 22).  The coefficient / bounds tables are computed on the
//   host exactly as Pillow does (open_clip/gpu_transform.py); the kernels below are the integer convolutions, so the output
//   is BIT-IDENTICAL to the PIL pipeline.  Only the crop window of the resized image is ever computed.
#include "common.cuh"
#include "internal.h"

namespace b200clip {

namespace {

constexpr int kPrecisionBits = 22;

__device__ __forceinline__ uint8_t clip8(int v) {
    v >>= kPrecisionBits;
    return static_cast<uint8_t>(v < 0 ? 0 : (v > 255 ? 255 : v));
}

// horizontal pass: tmp[y][x][c] for y in [0, rows), x in [0, out_w): src row (y0 + y), window bounds[x] = (xmin, xsize)
__global__ void __launch_bounds__(256) resize_h_kernel(const uint8_t* __restrict__ src, int64_t row_stride, int y0, int rows, const int32_t* __restrict__ bounds,
                                                       const int32_t* __restrict__ kk, int ksize, int out_w, uint8_t* __restrict__ tmp) {
    const int64_t total = static_cast<int64_t>(rows) * out_w;
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const int y = static_cast<int>(i / out_w), x = static_cast<int>(i - static_cast<int64_t>(y) * out_w);
        const int xmin = bounds[2 * x], xs = bounds[2 * x + 1];
        const int32_t* k = kk + static_cast<int64_t>(x) * ksize;
        const uint8_t* p = src + static_cast<int64_t>(y0 + y) * row_stride + static_cast<int64_t>(xmin) * 3;
        int s0 = 1 << (kPrecisionBits - 1), s1 = s0, s2 = s0;
        for (int j = 0; j < xs; ++j) {
            const int w = k[j];
            s0 += p[3 * j + 0] * w;
            s1 += p[3 * j + 1] * w;
            s2 += p[3 * j + 2] * w;
        }
        uint8_t* o = tmp + i * 3;
        o[0] = clip8(s0);
        o[1] = clip8(s1);
        o[2] = clip8(s2);
    }
}

// vertical pass + HWC -> CHW: dst[c][y][x] from tmp rows (bounds[y].min - y0 ...), y in [0, out_h)
__global__ void __launch_bounds__(256) resize_v_kernel(const uint8_t* __restrict__ tmp, int y0, const int32_t* __restrict__ bounds, const int32_t* __restrict__ kk,
                                                       int ksize, int out_h, int out_w, uint8_t* __restrict__ dst) {
    const int64_t total = static_cast<int64_t>(out_h) * out_w;
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const int y = static_cast<int>(i / out_w), x = static_cast<int>(i - static_cast<int64_t>(y) * out_w);
        const int ymin = bounds[2 * y] - y0, ys = bounds[2 * y + 1];
        const int32_t* k = kk + static_cast<int64_t>(y) * ksize;
        const uint8_t* p = tmp + (static_cast<int64_t>(ymin) * out_w + x) * 3;
        int s0 = 1 << (kPrecisionBits - 1), s1 = s0, s2 = s0;
        for (int j = 0; j < ys; ++j) {
            const int w = k[j];
            const uint8_t* q = p + static_cast<int64_t>(j) * out_w * 3;
            s0 += q[0] * w;
            s1 += q[1] * w;
            s2 += q[2] * w;
        }
        dst[i] = clip8(s0);
        dst[total + i] = clip8(s1);
        dst[2 * total + i] = clip8(s2);
    }
}

// no resampling along an axis (scale 1, offset 0) still goes through the tables: Pillow skips the pass, and so does the host
// by handing over identity tables (one tap of 2^22), which reproduce the input bytes exactly.

inline int blocks_for(int64_t n) {
    int64_t b = (n + 255) / 256;
    const int64_t cap = static_cast<int64_t>(num_sms()) * 16;
    return static_cast<int>(b < 1 ? 1 : (b > cap ? cap : b));
}

}  // namespace

int resize_crop_u8(const uint8_t* src_hwc, int H, int W, int64_t row_stride, const int32_t* h_bounds, const int32_t* h_coeffs, int h_ksize,
                   const int32_t* v_bounds, const int32_t* v_coeffs, int v_ksize, int y0, int rows, uint8_t* tmp, uint8_t* dst_chw, int out_h,
                   int out_w, cudaStream_t stream) {
    B2C_CHECK_ARG(src_hwc && h_bounds && h_coeffs && v_bounds && v_coeffs && tmp && dst_chw, "resize_crop_u8: null pointer");
    B2C_CHECK_ARG(H > 0 && W > 0 && row_stride >= static_cast<int64_t>(W) * 3 && out_h > 0 && out_w > 0 && h_ksize > 0 && v_ksize > 0,
                  "resize_crop_u8: bad shape");
    B2C_CHECK_ARG(y0 >= 0 && rows > 0 && y0 + rows <= H, "resize_crop_u8: source row window [%d, %d) outside the image (H = %d)", y0, y0 + rows, H);
    resize_h_kernel<<<blocks_for(static_cast<int64_t>(rows) * out_w), 256, 0, stream>>>(src_hwc, row_stride, y0, rows, h_bounds, h_coeffs, h_ksize, out_w, tmp);
    B2C_LAUNCH_CHECK("resize_h_kernel");
    resize_v_kernel<<<blocks_for(static_cast<int64_t>(out_h) * out_w), 256, 0, stream>>>(tmp, y0, v_bounds, v_coeffs, v_ksize, out_h, out_w, dst_chw);
    B2C_LAUNCH_CHECK("resize_v_kernel");
    return 0;
}

}  // namespace b200clip
