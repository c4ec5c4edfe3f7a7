// ClipLoss forward + backward for the --local-loss --gather-with-grad configuration
// (deps/open_clip/src/open_clip/loss.py:102-131, labels :89-100):
//   logits_per_image = s * img_loc @ all_txt^T          [n, N]
//   logits_per_text  = s * txt_loc @ all_img^T          [n, N]
//   labels_i = i + n * rank;   loss = (CE(logits_per_image) + CE(logits_per_text)) / 2
// Round-1 structure: two fp32 logit GEMMs into a caller-provided workspace, one fused
// softmax / cross-entropy / d-logits kernel (one CTA per logit row, the row is read once and
// overwritten in place with s * dL/dlogit), a deterministic single-CTA reduction for the loss and
// d(scale), and four fp32 GEMMs for the feature gradients.  world_size == 1 is the same code with
// N == n and all_* == *_loc.
#include "common.cuh"
#include "internal.h"

namespace b200clip {

int gemm_f32_nt(bool ta, bool tb, const float* A, int64_t lda, const float* Bm, int64_t ldb, float* C, int64_t ldc, int M, int N,
                int K, cudaStream_t stream);

namespace {

__device__ __forceinline__ float block_reduce(float v, float* red, bool is_max) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    v = is_max ? warp_max(v) : warp_sum(v);
    __syncthreads();
    if (lane == 0) red[warp] = v;
    __syncthreads();
    float r = red[0];
    for (int i = 1; i < nw; ++i) r = is_max ? fmaxf(r, red[i]) : r + red[i];
    return r;
}

// rows [0,n): image->text logits; rows [n,2n): text->image logits.  `x` holds raw dot products.
__global__ void __launch_bounds__(256)
ce_rows_kernel(float* __restrict__ x, const float* __restrict__ logit_scale, const float* __restrict__ grad_out, int n, int N,
               int rank, int want_grad, float* __restrict__ row_loss, float* __restrict__ row_dscale) {
    __shared__ float red[8];
    const int r = blockIdx.x;
    float* xr = x + static_cast<int64_t>(r) * N;
    const float s = *logit_scale;
    const int label = (r % n) + n * rank;

    float mx = -INFINITY;
    for (int j = threadIdx.x; j < N; j += blockDim.x) mx = fmaxf(mx, s * xr[j]);
    mx = block_reduce(mx, red, true);
    float sum = 0.f;
    for (int j = threadIdx.x; j < N; j += blockDim.x) sum += expf(s * xr[j] - mx);
    sum = block_reduce(sum, red, false);
    const float lse = mx + logf(sum);
    if (threadIdx.x == 0) row_loss[r] = lse - s * xr[label];
    if (!want_grad) return;
    const float g = (grad_out != nullptr ? *grad_out : 1.f) / (2.f * static_cast<float>(n));
    float ds = 0.f;
    __syncthreads();  // row_loss read of xr[label] happens before the in-place overwrite below
    for (int j = threadIdx.x; j < N; j += blockDim.x) {
        const float raw = xr[j];
        const float p = expf(s * raw - lse);
        const float dl = g * (p - (j == label ? 1.f : 0.f));
        ds += dl * raw;
        xr[j] = s * dl;
    }
    ds = block_reduce(ds, red, false);
    if (threadIdx.x == 0) row_dscale[r] = ds;
}

__global__ void __launch_bounds__(256)
loss_finalize_kernel(const float* __restrict__ row_loss, const float* __restrict__ row_dscale, int rows, float inv_rows,
                     float* __restrict__ loss, float* __restrict__ d_scale) {
    __shared__ float red[8];
    float a = 0.f, b = 0.f;
    for (int i = threadIdx.x; i < rows; i += blockDim.x) {
        a += row_loss[i];
        if (d_scale != nullptr) b += row_dscale[i];
    }
    a = block_reduce(a, red, false);
    b = block_reduce(b, red, false);
    if (threadIdx.x == 0) {
        *loss = a * inv_rows;
        if (d_scale != nullptr) *d_scale = b;
    }
}

}  // namespace

int cliploss(const float* img_loc, const float* txt_loc, const float* all_img, const float* all_txt, const float* logit_scale,
             int rank, int n, int N, int D, float* loss, const float* grad_out, float* d_img_loc, float* d_txt_loc,
             float* d_all_img, float* d_all_txt, float* d_scale, float* workspace, cudaStream_t stream) {
    B2C_CHECK_ARG(n > 0 && N >= n && D > 0 && N % n == 0, "cliploss: bad shape n=%d N=%d D=%d", n, N, D);
    B2C_CHECK_ARG(rank >= 0 && (rank + 1) * n <= N, "cliploss: rank %d out of range for n=%d N=%d", rank, n, N);
    B2C_CHECK_ARG(n % 4 == 0 && D % 4 == 0, "cliploss: n and D must be multiples of 4");
    B2C_CHECK_ARG(img_loc && txt_loc && all_img && all_txt && logit_scale && loss && workspace, "cliploss: null pointer");
    const bool want_grad = d_img_loc || d_txt_loc || d_all_img || d_all_txt || d_scale;
    float* Li = workspace;
    float* Lt = workspace + static_cast<int64_t>(n) * N;
    float* row_loss = Lt + static_cast<int64_t>(n) * N;
    float* row_ds = row_loss + 2 * n;

    int rc;
    if ((rc = gemm_f32_nt(false, false, img_loc, D, all_txt, D, Li, N, n, N, D, stream)) != 0) return rc;
    if ((rc = gemm_f32_nt(false, false, txt_loc, D, all_img, D, Lt, N, n, N, D, stream)) != 0) return rc;
    ce_rows_kernel<<<2 * n, 256, 0, stream>>>(workspace, logit_scale, grad_out, n, N, rank, want_grad ? 1 : 0, row_loss, row_ds);
    B2C_LAUNCH_CHECK("ce_rows_kernel");
    loss_finalize_kernel<<<1, 256, 0, stream>>>(row_loss, row_ds, 2 * n, 1.f / (2.f * static_cast<float>(n)), loss,
                                                want_grad ? d_scale : nullptr);
    B2C_LAUNCH_CHECK("loss_finalize_kernel");
    if (!want_grad) return 0;
    // workspace now holds s * dL/dlogits
    if (d_img_loc && (rc = gemm_f32_nt(false, true, Li, N, all_txt, D, d_img_loc, D, n, D, N, stream)) != 0) return rc;
    if (d_all_txt && (rc = gemm_f32_nt(true, true, Li, N, img_loc, D, d_all_txt, D, N, D, n, stream)) != 0) return rc;
    if (d_txt_loc && (rc = gemm_f32_nt(false, true, Lt, N, all_img, D, d_txt_loc, D, n, D, N, stream)) != 0) return rc;
    if (d_all_img && (rc = gemm_f32_nt(true, true, Lt, N, txt_loc, D, d_all_img, D, N, D, n, stream)) != 0) return rc;
    return 0;
}

}  // namespace b200clip
