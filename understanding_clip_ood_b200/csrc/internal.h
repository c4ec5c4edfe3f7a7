// Internal C++ entry points shared between the kernel translation units and the C ABI (api.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace b200clip {

void count_launch(int n = 1);

int gemm_tc(bool is_bf16, const void* A, int64_t lda, const void* W, int64_t ldw, const void* bias, const void* residual,
            int64_t ldr, void* C, int64_t ldc, int M, int N, int K, int epilogue, const float* pos, int g_in, int g_out,
            int force_block_n, cudaStream_t stream);

// Backward-GEMM form of the pair kernel: C[M, N] = op(A) op(W)^T, plain stores, operands optionally MN-major (a_mn: A stored
// [K, M]; w_mn: W stored [K, N]) so that dgrad (G W) and wgrad (G^T X) need no transposed copies.  Stream-K as in gemm_pair.
int gemm_pair_mn(bool is_bf16, const void* A, int64_t lda, bool a_mn, const void* W, int64_t ldw, bool w_mn, void* C, int64_t ldc, int M, int N,
                 int K, cudaStream_t stream, void* sk_workspace);

// Implicit 3x3 / stride 1 / padding 1 convolution + bias + ReLU over NHWC activations on the pair kernel (no im2col matrix):
// 0 = launched, 1 = geometry not covered (caller falls back to im2col + GEMM), < 0 / > 0 = error
int gemm_pair_conv3x3(bool is_bf16, const void* in, const void* Wt, const void* bias, void* out, int batch, int H, int W, int C, int N,
                      cudaStream_t stream);

// Implicit patch embedding of the ViT for 16-bit NCHW batches (token-layout x [batch, G*G + 1, width], pos_cls = fp32 [G*G + 1, width]
// table whose row 0 holds class_embedding + positional_embedding[0]): 0 = launched, 1 = patch size not covered, otherwise an error
int gemm_pair_patch_embed(bool is_bf16, const void* image, const void* conv1_w, const float* pos_cls, void* x, int batch, int image_size,
                          int patch, int width, cudaStream_t stream);

// CTA-pair (cta_group::2) variant with the TMA-store epilogue (gemm_pair.cu); epilogues 0..3
int gemm_pair(bool is_bf16, const void* A, int64_t lda, const void* W, int64_t ldw, const void* bias, const void* residual,
              int64_t ldr, void* C, int64_t ldc, int M, int N, int K, int epilogue, int force_block_n, int pairs, cudaStream_t stream,
              const float* ln_colsum = nullptr, const float* ln_rowstats = nullptr, const float* pos_table = nullptr, int pos_period = 0,
              float* stats_out = nullptr, const float* stats_part = nullptr, int stats_slots = 0, float ln_eps = 1e-5f,
              void* sk_workspace = nullptr);
// Stream-K (gemm_pair.cu): `sk_workspace` of gemm_pair_sk_workspace_bytes() bytes lets a GEMM whose tile count does not fill
// whole rounds of the persistent grid cut its ragged part into equal runs of K-blocks (fp32 partials + flags live there).
// The flag words must be zero before the first use (gemm_pair_sk_workspace_reset); the kernels leave them at zero.
int64_t gemm_pair_sk_workspace_bytes();
int gemm_pair_sk_workspace_reset(void* ws, cudaStream_t stream);
// LayerNorm statistics fused into the residual GEMMs: `stats_out` ([M][gemm_pair_stats_slots(M, N)] float2 partial (sum x,
// sum x^2) of the output rows) is written by a residual-epilogue GEMM and read back through `stats_part` / `stats_slots` by
// the LN-fold GEMM that consumes those rows, instead of (mean, rstd) from row_stats.
int gemm_pair_stats_slots(int M, int N);

// per-row LayerNorm statistics (mean, rstd) as float2 (rowwise.cu)
int row_stats(int dtype, const void* x, int64_t ldx, float* stats, int rows, int width, float eps, cudaStream_t stream);

int gemm_f32(const float* A, int64_t lda, const float* W, int64_t ldw, const float* bias, const float* residual, int64_t ldr,
             float* C, int64_t ldc, int M, int N, int K, int epilogue, const float* pos, int g_in, int g_out,
             cudaStream_t stream);

// dtype-dispatching GEMM used by the tower drivers
int gemm_any(int dtype, const void* A, int64_t lda, const void* W, int64_t ldw, const void* bias, const void* residual,
             int64_t ldr, void* C, int64_t ldc, int M, int N, int K, int epilogue, const float* pos, int g_in, int g_out,
             cudaStream_t stream, void* sk_workspace = nullptr);

int layernorm(int dtype, const void* x, int64_t ldx, const float* gamma, const float* beta, void* y, int64_t ldy, int rows,
              int width, float eps, int row_stride_rows, const int32_t* row_idx, cudaStream_t stream);

void attention_tc_set_debug(long long* buf);
int attention(int dtype, const void* qkv, void* out, int batch, int seq_len, int heads, int causal, cudaStream_t stream);

int patchify(int dtype, const void* image, void* patches, int batch, int image_size, int patch, int kpad,
             const float* class_emb, const float* pos, void* x, int width, cudaStream_t stream, int cls_slot = 0);

// same im2col from uint8 pixels with ToTensor + Normalize(mean, std) applied on the fly (mean / std: 3 host floats each)
int patchify_u8(int dtype, const uint8_t* image, const float* mean, const float* std, void* patches, int batch, int image_size, int patch,
                int kpad, const float* class_emb, const float* pos, void* x, int width, cudaStream_t stream, int cls_slot = 0);

int text_embed(int dtype, const int64_t* text, int ctx, const float* tok_emb, const float* pos_emb, void* x, int32_t* eot,
               int T, int L, int width, cudaStream_t stream);

int eot_argmax(const int64_t* text, int ctx, int32_t* eot, int T, cudaStream_t stream);

int normalize_rows(int dtype, const void* x, int64_t ldx, void* y, int64_t ldy, int rows, int dim, float eps,
                   cudaStream_t stream);

int zeroshot(int dtype, const void* img_feat, const void* prompt_feat, float* logits, int64_t* topk_idx, float* topk_val,
             int B, int C, int D, int k, int normalize_img, float logit_scale, cudaStream_t stream);

int topk_rows(int dtype, const void* x, int64_t ldx, int B, int C, int k, int64_t* topk_idx, float* topk_val, float* logits_out,
              cudaStream_t stream);

int class_mean(int dtype, const void* txt_feat, void* prompt_feat, int classes, int templates, int D, cudaStream_t stream);

int cliploss(const float* img_loc, const float* txt_loc, const float* all_img, const float* all_txt, const float* logit_scale,
             int rank, int n, int N, int D, float* loss, const float* grad_out, float* d_img_loc, float* d_txt_loc,
             float* d_all_img, float* d_all_txt, float* d_scale, float* workspace, cudaStream_t stream);

// split form of `cliploss`: forward leaves the raw logits + per-row log-sum-exp in the workspace for the backward
int cliploss_forward(const float* img_loc, const float* txt_loc, const float* all_img, const float* all_txt, const float* logit_scale,
                     int rank, int n, int N, int D, float* loss, float* workspace, cudaStream_t stream);
int cliploss_backward(const float* img_loc, const float* txt_loc, const float* all_img, const float* all_txt, const float* logit_scale,
                      int rank, int n, int N, int D, const float* grad_out, float* d_img_loc, float* d_txt_loc, float* d_all_img,
                      float* d_all_txt, float* d_scale, float* workspace, cudaStream_t stream);

// world_size == 1: total gradients of the two feature matrices in one two-segment GEMM launch (after cliploss_forward with
// all_* == *_loc, rank 0, N == n)
int cliploss_single_backward(const float* img, const float* txt, const float* logit_scale, int n, int D, const float* grad_out,
                             float* d_img, float* d_txt, float* d_scale, float* workspace, cudaStream_t stream);

int cliploss_packed_forward(const float* gathered, const float* logit_scale, int rank, int n, int N, int D, float* loss,
                            float* workspace, cudaStream_t stream);
int cliploss_packed_backward(const float* gathered, const float* logit_scale, int rank, int n, int N, int D, const float* grad_out,
                             float* d_gathered, float* d_scale, float* workspace, cudaStream_t stream);

int cliploss_packed_backward_p2p(const float* gathered, const float* logit_scale, int rank, int n, int N, int D, const float* grad_out,
                                 float* const* d_slots, float* d_scale, float* workspace, cudaStream_t stream);

// peer-memory exchange of the distributed ClipLoss (p2p.cu)
int p2p_configure(double timeout_seconds, uint32_t* error_word);
int p2p_allgather(int dtype, const void* img, const void* txt, int n, int D, float* const* peer_dst, uint32_t* const* peer_flag,
                  const uint32_t* my_flags, uint32_t* counters, int world, uint32_t epoch, uint32_t* const* peer_busy, uint32_t* my_busy,
                  int hold, cudaStream_t stream);
int p2p_reduce_finish(const float* recv, float* out, int64_t elems, uint32_t* const* peer_flag, const uint32_t* my_flags, int world,
                      int slots, uint32_t epoch, uint32_t* my_busy, int split_cols, cudaStream_t stream);

// bicubic resize + centre crop of a decoded uint8 HWC image (preprocess.cu); tables from the host (Pillow's fixed-point coefficients)
int resize_crop_u8(const uint8_t* src_hwc, int H, int W, int64_t row_stride, const int32_t* h_bounds, const int32_t* h_coeffs, int h_ksize,
                   const int32_t* v_bounds, const int32_t* v_coeffs, int v_ksize, int y0, int rows, uint8_t* tmp, uint8_t* dst_chw, int out_h,
                   int out_w, cudaStream_t stream);

// ---- training path (backward.cu, towers_bwd.cu, optim.cu) ----
int gemm_f32_general(const float* A, int64_t lda, bool ta, const float* B, int64_t ldb, bool tb, float* C, int64_t ldc, int M, int N, int K,
                     bool accumulate, cudaStream_t stream);
int transpose16(int dtype, const void* in, int64_t ldi, void* out, int64_t ldo, int R, int C, int Rpad, cudaStream_t stream, int act = 0);
int64_t col_sum_scratch_floats(int rows, int cols);
// counters: optional kColSumCounters (= 256) zero-initialised words (left at zero): single-launch form (the last row chunk of each
// 64-column block finishes the sum); nullptr: partial sums + a second launch
int col_sum(int dtype, const void* g, int64_t ld, int rows, int cols, void* out, int out_f32, int accumulate, float* scratch, cudaStream_t stream,
            unsigned int* counters = nullptr);
int64_t ln_backward_scratch_floats(int rows, int width);
int ln_backward(int dtype, const void* g, int64_t ldg, const void* x, int64_t ldx, const float* gamma, const void* dres, int64_t ldr, void* dx,
                int64_t ldd, float* d_gamma, float* d_beta, int rows, int width, float eps, int row_stride, const int32_t* row_idx, int accumulate,
                float* scratch, cudaStream_t stream);
int act_backward(int dtype, const void* da, const void* z, void* dz, int64_t n, int quick, cudaStream_t stream);
int act_forward(int dtype, const void* z, void* a, int64_t n, int quick, cudaStream_t stream);
int attention_backward(int dtype, const void* qkv, const void* d_out, void* d_qkv, int batch, int seq_len, int heads, int causal, cudaStream_t stream);
int normalize_backward(int dtype, const void* x, const void* g, void* dx, int rows, int dim, float eps, cudaStream_t stream);
int period_sum(int dtype, const void* dx, int64_t ld, int groups, int L, int width, float* out, int64_t ldo, int accumulate, cudaStream_t stream);
int token_scatter(int dtype, const void* dx, int64_t ld, const int64_t* text, int ctx, int T_, int L, int width, int vocab, float* d_tok, cudaStream_t stream);
int convert(int dtype_in, const void* in, int dtype_out, void* out, int64_t n, cudaStream_t stream);

inline int dtype_size(int dtype) { return dtype == 0 ? 4 : 2; }

}  // namespace b200clip
struct b200clip_adamw_tensor;
struct b200clip_cast_tensor;
namespace b200clip {
int multi_cast(const b200clip_cast_tensor* items, const int32_t* chunk_item, const int64_t* chunk_off, int chunks, cudaStream_t stream);
int adamw_step(const b200clip_adamw_tensor* items, const int32_t* chunk_item, const int64_t* chunk_off, int chunks, float lr, float beta1,
               float beta2, float eps, float weight_decay, int step, float grad_scale, cudaStream_t stream);

}  // namespace b200clip
