/*
 * b200clip — C ABI of the B200-native (sm_100a) CLIP hot path.
 *
 * The reference (lmb-freiburg/understanding-clip-ood @ /root/reference) is pure Python: it has no
 * FFI/plugin interface of its own.  Its drop-in boundary is a set of Python call signatures
 * (SURVEY.md §8b); every entry point below states which reference call site(s) it replaces.  The
 * Python host side (understanding_clip_ood_b200/) binds these with ctypes, passing
 * `tensor.data_ptr()` and `torch.cuda.current_stream().cuda_stream`.
 *
 * Conventions
 *  - plain pointers and sizes only; every pointer is DEVICE memory owned by the caller unless the
 *    name ends in `_host`; the library never allocates memory that outlives a call;
 *  - `stream` is a `cudaStream_t` passed as `void*`; all work is enqueued on it, nothing
 *    synchronises, so every call is CUDA-graph capturable;
 *  - return value: 0 = ok, <0 = invalid argument, >0 = CUDA error code (`cudaError_t`);
 *    `b200clip_last_error()` returns a thread-local message for the last non-zero return;
 *  - `dtype`: activation/weight storage type of the call (B200CLIP_F32 / BF16 / F16).  In the 16-bit
 *    modes LayerNorm affine parameters, embedding tables and positional tables stay fp32, exactly
 *    like `convert_weights_to_lp` leaves them (deps/open_clip/src/open_clip/model.py:396-423).
 *  - matrices are row-major with explicit leading dimensions in elements.
 */
#ifndef B200CLIP_H_
#define B200CLIP_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200CLIP_F32 0
#define B200CLIP_BF16 1
#define B200CLIP_F16 2

/* GEMM epilogues (fused into the accumulator read-out) */
#define B200CLIP_EPI_BIAS 0      /* C = A W^T + b                     (MHA in-proj, functional.py _in_projection_packed) */
#define B200CLIP_EPI_GELU 1      /* C = gelu_erf(A W^T + b)           (mlp.c_fc + nn.GELU, transformer.py:231-235)       */
#define B200CLIP_EPI_QUICKGELU 2 /* C = x*sigmoid(1.702x)             (QuickGELU, transformer.py:33-36)                  */
#define B200CLIP_EPI_RESIDUAL 3  /* C = R + (A W^T + b)               (out_proj / c_proj + residual, transformer.py:262-263) */
#define B200CLIP_EPI_PATCH 4     /* patch-embedding: row remap + positional add (transformer.py:602-609)                 */
#define B200CLIP_EPI_RELU 5      /* C = relu(A W^T + b)               (conv + folded BatchNorm + ReLU, modified_resnet.py:42-45) */
#define B200CLIP_EPI_RESIDUAL_RELU 6 /* C = relu(R + (A W^T + b))     (bn3(conv3) + identity + ReLU, modified_resnet.py:46-55)   */

int b200clip_version(void);
const char* b200clip_last_error(void);
/* number of kernels this library has launched in the calling process (bench.py's gpu_launches) */
uint64_t b200clip_launch_count(void);

/* C[M,N] = epilogue(A[M,K] @ W[N,K]^T + bias[N]).  W is an nn.Linear weight ([out,in], K contiguous).
 * 16-bit dtypes run on tcgen05 tensor cores (TMA-fed, TMEM accumulators); F32 runs an FFMA kernel with
 * fp32 accumulation (the 1e-4 parity mode).  `bias`, `residual` may be NULL when the epilogue does not
 * use them.  EPI_PATCH: row m of the product is written to row (m / g_in) * g_out + m % g_in + 1 of C
 * after adding pos[(m % g_in) + 1, :] (fp32 table, ld = N); g_in = patches per image, g_out = g_in + 1.
 * Replaces: F.linear / nn.Linear / `@` call sites of transformer.py:224-235,249-263,638 and model.py:282. */
int b200clip_gemm(int dtype, const void* A, int64_t lda, const void* W, int64_t ldw, const void* bias,
                  const void* residual, int64_t ldr, void* C, int64_t ldc, int M, int N, int K, int epilogue,
                  const float* pos, int g_in, int g_out, void* stream);

/* The same GEMM with a caller-provided scratch buffer (b200clip_gemm_workspace_bytes() bytes, 256-byte aligned, device memory):
 * when the output tiles do not fill whole rounds of the persistent grid (e.g. 75 tiles on 74 SM pairs: the out-proj / c_proj of
 * a 128-image shard), the ragged part is cut stream-K fashion into equal runs of K-blocks; partial accumulators (fp32) and
 * flags live in the workspace, the sums are taken in a fixed order (deterministic).  Epilogues BIAS / GELU / QUICKGELU /
 * RESIDUAL; a RESIDUAL that does not alias C keeps whole tiles.  The tower drivers pass part of their own workspace. */
int64_t b200clip_gemm_workspace_bytes(void);
int b200clip_gemm_ws(int dtype, const void* A, int64_t lda, const void* W, int64_t ldw, const void* bias,
                     const void* residual, int64_t ldr, void* C, int64_t ldc, int M, int N, int K, int epilogue,
                     void* workspace, int64_t workspace_bytes, void* stream);

/* The GEMMs of the tower backward (16-bit dtypes), whose contraction index is NOT the contiguous one of the stored operands:
 *   C[M,N] = A' W    W stored [K, N] (row pitch ldw; N contiguous): the weight [out, in] of an nn.Linear as it lies in memory;
 *   a_mn == 0: A' = A stored [M, K]              -> dgrad  dX = G W            (autograd of F.linear, transformer.py:253-264)
 *   a_mn != 0: A' = A^T with A stored [K, M]     -> wgrad  dW = G^T X  (K = all token rows)
 * The tcgen05 instruction reads such operands as MN-major shared-memory tiles, so no transposed copy is made.  Plain stores
 * (no bias / activation).  workspace: optional stream-K scratch as for b200clip_gemm_ws (NULL = whole tiles only). */
int b200clip_gemm_mn(int dtype, const void* A, int64_t lda, int a_mn, const void* W, int64_t ldw, void* C, int64_t ldc,
                     int M, int N, int K, void* workspace, int64_t workspace_bytes, void* stream);

/* Patch embedding of the ViT (conv1 with kernel = stride = patch, no bias; class token; positional embedding,
 * transformer.py:602-609) as an IMPLICIT GEMM over a 16-bit NCHW batch: the patches are read from `image` [batch, 3, S, S]
 * through a 5-D tensor map, no im2col matrix is written:
 *   x[b, 1 + py*G + px, :] = round(patch(b, py, px) . conv1_w^T) + round(pos_cls[1 + py*G + px, :]),  x[b, 0, :] = round(pos_cls[0, :])
 * conv1_w [width, 3*P*P] (K order c, dy, dx), pos_cls fp32 [G*G + 1, width] with row 0 = class_embedding + positional_embedding[0],
 * x [batch, G*G + 1, width].  Patch 16 or 32.  Same values as b200clip_patchify + the token-layout GEMM; measured slower than
 * that pair on ViT-B/32 (64-byte gathers, A re-gathered per N tile), so the towers use it only with B200CLIP_PATCH_IMPLICIT=1. */
int b200clip_patch_embed_implicit(int dtype, const void* image, const void* conv1_w, const float* pos_cls, void* x, int batch,
                                  int image_size, int patch, int width, void* stream);

/* LayerNorm folded into the following GEMM (16-bit dtypes): C = act(LN(x) W^T + b) computed WITHOUT materialising LN(x):
 *   C[m,n] = act( rstd_m * (x W'^T)[m,n] - rstd_m * mean_m * colsum[n] + bias_f32[n] )
 * with W' = W diag(gamma) in `dtype`, colsum[n] = sum_k W'[n,k] and bias_f32 = b + W beta (both fp32, prepared once by
 * the caller) and rowstats[m] = (mean_m, rstd_m) from b200clip_row_stats.  `epilogue` is EPI_BIAS / EPI_GELU / EPI_QUICKGELU.
 * Replaces ln_1 + in_proj and ln_2 + c_fc (+ activation) of ResidualAttentionBlock (transformer.py:253-264). */
int b200clip_gemm_ln(int dtype, const void* x, int64_t ldx, const void* Wf, int64_t ldw, const float* colsum,
                     const float* bias_f32, const float* rowstats, void* C, int64_t ldc, int M, int N, int K, int epilogue,
                     void* stream);

/* b200clip_gemm_ln with the stream-K workspace of b200clip_gemm_ws */
int b200clip_gemm_ln_ws(int dtype, const void* x, int64_t ldx, const void* Wf, int64_t ldw, const float* colsum,
                        const float* bias_f32, const float* rowstats, void* C, int64_t ldc, int M, int N, int K,
                        int epilogue, void* workspace, int64_t workspace_bytes, void* stream);

/* The row statistics can also come out of the GEMM that WROTE the rows (the residual GEMMs of the block, whose output is the
 * input of the next LayerNorm), which removes the separate statistics pass over the residual stream:
 *   b200clip_gemm_residual_stats: C = round(A W^T + bias) + residual (EPI_RESIDUAL semantics; residual may alias C) and
 *     partials[m][s] = (sum, sum of squares) of the stored values of row m over the columns epilogue slot s covered,
 *     s < b200clip_gemm_stats_slots(M, N) (fp32 pairs, fixed slots: deterministic).
 *   b200clip_gemm_ln_partials: b200clip_gemm_ln with (mean_m, rstd_m) derived from those `slots` pairs per row:
 *     mean = sum / K, rstd = 1/sqrt(max(sumsq / K - mean^2, 0) + eps). */
int b200clip_gemm_stats_slots(int M, int N);
int b200clip_gemm_residual_stats(int dtype, const void* A, int64_t lda, const void* W, int64_t ldw, const void* bias,
                                 const void* residual, int64_t ldr, void* C, int64_t ldc, int M, int N, int K,
                                 float* partials, void* stream);
int b200clip_gemm_ln_partials(int dtype, const void* x, int64_t ldx, const void* Wf, int64_t ldw, const float* colsum,
                              const float* bias_f32, const float* partials, int slots, float eps, void* C, int64_t ldc,
                              int M, int N, int K, int epilogue, void* stream);

/* stats[r] = (mean, rstd = 1/sqrt(var + eps)) of row r of x, fp32 pairs (biased variance, like F.layer_norm). */
int b200clip_row_stats(int dtype, const void* x, int64_t ldx, float* stats, int rows, int width, float eps, void* stream);

/* y[r,:] = LayerNorm(x[row_index(r),:]) * gamma + beta, fp32 statistics, eps as given (1e-5).
 * `row_stride_rows` > 0 selects every row_stride_rows-th row starting at row_offset (CLS pooling:
 * stride L, offset 0); `row_idx` (int32, may be NULL) adds a per-output-row offset (EOT pooling).
 * Replaces LayerNorm / LayerNormFp32 (transformer.py:15-30) at ln_pre, ln_1, ln_2, ln_post, ln_final. */
int b200clip_layernorm(int dtype, const void* x, int64_t ldx, const float* gamma, const float* beta, void* y,
                       int64_t ldy, int rows, int width, float eps, int row_stride_rows, const int32_t* row_idx,
                       void* stream);

/* Multi-head softmax attention over a packed qkv buffer: qkv[M, 3W] with q|k|v column blocks, row
 * b*L + l, head h = columns [64h, 64h+64) of each block; out[M, W].  scale = 1/sqrt(64) is applied to
 * the scores; `causal` != 0 adds the strict upper-triangular -inf mask of the text tower.
 * Replaces nn.MultiheadAttention's SDPA core (transformer.py:224,249-251; torch functional.py
 * multi_head_attention_forward) and build_causal_mask (transformer.py:751-757). */
int b200clip_attention(int dtype, const void* qkv, void* out, int batch, int seq_len, int heads, int causal,
                       void* stream);

/* Patch im2col: image [B,3,H,H] NCHW -> patches [B*g*g, kpad] with K order (channel, ky, kx) zero-padded to
 * kpad, and the class-token rows x[b*(g*g+1), :] = class_emb + pos[0] written into `x` (ld = width).
 * Replaces the unfold part of conv1 + class-token cat (transformer.py:602-609). */
int b200clip_patchify(int dtype, const void* image, void* patches, int batch, int image_size, int patch,
                      int kpad, const float* class_emb, const float* pos, void* x, int width, void* stream);

/* x[t*L + l, :] = token_embedding[text[t, l], :] + positional_embedding[l, :]  (model.py:272-274);
 * also writes eot[t] = argmax_l text[t, l] (int32) for text_global_pool (transformer.py:654).
 * `text` is int64 [T, ctx]; only the first L <= ctx positions are embedded (causal truncation). */
int b200clip_text_embed(int dtype, const int64_t* text, int ctx, const float* tok_emb, const float* pos_emb,
                        void* x, int32_t* eot, int T, int L, int width, void* stream);

/* eot[t] = argmax_l text[t, l] over the full context (first maximum), int32: the EOT position used by
 * text_global_pool (transformer.py:654); lets the host pick the exact causal truncation length max(eot)+1. */
int b200clip_eot_argmax(const int64_t* text, int ctx, int32_t* eot, int T, void* stream);

/* y = x / max(||x||_2, eps) per row (F.normalize, model.py:267,284; xclip/zero_shot.py:34,50). */
int b200clip_normalize(int dtype, const void* x, int64_t ldx, void* y, int64_t ldy, int rows, int dim, float eps,
                       void* stream);

/* Zero-shot similarity stage: img_feat[B,D] (un-normalised when `normalize_img` != 0) against unit-norm
 * prompt_feat[C,D]; writes logits[B,C] (fp32, may be NULL), topk_idx[B,k] (int64, descending, ties to the
 * lower class index like torch.argmax/topk) and topk_val[B,k] (fp32, may be NULL).  k <= 8.
 * Replaces xclip/zero_shot.py:42-60,103-109 (_compute_img_feat's normalize, _compute_logits, argmax) and
 * training/zero_shot.py:11-14 (topk). */
int b200clip_zeroshot(int dtype, const void* img_feat, const void* prompt_feat, float* logits, int64_t* topk_idx,
                      float* topk_val, int B, int C, int D, int k, int normalize_img, float logit_scale,
                      void* stream);

/* Per-row top-k of a materialised logit matrix x[B,C] (row pitch ldx elements, C <= 1024): topk_idx[B,k] int64 descending,
 * ties to the lower class index; topk_val[B,k] fp32 (may be NULL); logits_out[B,C] fp32 copy of the rows (may be NULL; k may
 * then be 0).  The 16-bit zero-shot path computes x with b200clip_gemm(img_feat, prompt_feat) on the tensor cores and ends
 * here.  Replaces argmax / topk of xclip/zero_shot.py:103-109 and training/zero_shot.py:11-14. */
int b200clip_topk(int dtype, const void* x, int64_t ldx, int B, int C, int k, int64_t* topk_idx, float* topk_val,
                  float* logits_out, void* stream);

/* prompt_feat[c,:] = normalize(mean_t normalize(txt_feat[c*T + t, :])) (xclip/zero_shot.py:231-234). */
int b200clip_class_mean(int dtype, const void* txt_feat, void* prompt_feat, int classes, int templates, int D,
                        void* stream);

/* ClipLoss local-loss forward+backward on gathered features (loss.py:102-131).  fp32 features.
 *   loss      = (CE(s*img_loc@all_txt^T, lab) + CE(s*txt_loc@all_img^T, lab)) / 2, lab_i = i + n*rank
 *   d_img_loc, d_txt_loc [n,D]: gradient through the local operands,
 *   d_all_img, d_all_txt [N,D]: gradient through the gathered operands (reduce-scattered by the caller,
 *   as torch.distributed.nn.all_gather's backward does), d_scale: gradient of the logit scale.
 * Gradients are for `loss` with upstream gradient `grad_out` (device scalar, may be NULL = 1).
 * Any gradient pointer may be NULL (forward only when all are NULL).  workspace >= 2*n*N + 8*n + 8 floats. */
int b200clip_cliploss(const float* img_loc, const float* txt_loc, const float* all_img, const float* all_txt,
                      const float* logit_scale, int rank, int n, int N, int D, float* loss, const float* grad_out,
                      float* d_img_loc, float* d_txt_loc, float* d_all_img, float* d_all_txt, float* d_scale,
                      float* workspace, void* stream);

/* The same step split at the autograd boundary: `forward` writes the loss and leaves the raw logits and the per-row
 * log-sum-exp in `workspace`; `backward` (same operands, same workspace, upstream gradient `grad_out` = device scalar or
 * NULL for 1) turns them into the five gradients.  Two launches each; workspace as above. */
int b200clip_cliploss_forward(const float* img_loc, const float* txt_loc, const float* all_img, const float* all_txt,
                              const float* logit_scale, int rank, int n, int N, int D, float* loss, float* workspace,
                              void* stream);
int b200clip_cliploss_backward(const float* img_loc, const float* txt_loc, const float* all_img, const float* all_txt,
                               const float* logit_scale, int rank, int n, int N, int D, const float* grad_out,
                               float* d_img_loc, float* d_txt_loc, float* d_all_img, float* d_all_txt, float* d_scale,
                               float* workspace, void* stream);

/* Single-device form (world_size == 1, loss.py:102-131 with all_* == *_loc): the backward of
 * b200clip_cliploss_forward(img, txt, img, txt, logit_scale, 0, n, n, D, loss, workspace) that writes the TOTAL gradient of
 * each feature matrix (its row-operand and its column-operand term summed inside one GEMM launch), so the caller's autograd
 * has nothing to add up.  d_img, d_txt [n,D], d_scale may be NULL; grad_out = device scalar or NULL for 1. */
int b200clip_cliploss_single_backward(const float* img, const float* txt, const float* logit_scale, int n, int D,
                                      const float* grad_out, float* d_img, float* d_txt, float* d_scale, float* workspace,
                                      void* stream);

/* Distributed form (--local-loss --gather-with-grad, world_size > 1): `gathered` [N, 2D] fp32 is the payload of the ONE
 * all-gather the host issues (row r*n + i = img_i | txt_i of rank r; loss.py:49-50 gathers the two halves separately); the
 * local operands are its rows [rank*n, rank*n + n).  `backward` writes d_gathered [N, 2D], the gradient w.r.t. every
 * gathered row with the local-row terms already added in, i.e. the input of the reduce-scatter(SUM) that completes
 * torch.distributed.nn.all_gather's backward.  workspace as for b200clip_cliploss. */
int b200clip_cliploss_packed_forward(const float* gathered, const float* logit_scale, int rank, int n, int N, int D,
                                     float* loss, float* workspace, void* stream);
int b200clip_cliploss_packed_backward(const float* gathered, const float* logit_scale, int rank, int n, int N, int D,
                                      const float* grad_out, float* d_gathered, float* d_scale, float* workspace,
                                      void* stream);

/* ----- peer-memory form of the distributed step (NVLink / NVSwitch stores instead of the NCCL all-gather and reduce-scatter
 * of loss.py:49-50 and its autograd backward).  The host maps one symmetric allocation per rank (torch symmetric memory) and
 * passes DEVICE arrays of `world` pre-offset pointers; flags are uint32 words holding a monotonically increasing epoch.
 *
 * b200clip_p2p_allgather: rows img | txt (dtype B200CLIP_*, [n, D] each) -> fp32 [n, 2D] written into peer_dst[p] for every
 *   rank p (= slot `rank` of p's gather buffer), then *peer_flag[p] = epoch (release, system scope), then waits until
 *   my_flags[q] >= epoch for all q: when the call's kernel has finished, the local gather buffer holds every rank's rows.
 *   `counters`: `world` zero-initialised uint32 in local memory (left at zero).
 * b200clip_cliploss_packed_backward_p2p: as b200clip_cliploss_packed_backward, but the gradient block of rank j's rows
 *   ([n, 2D]) is stored to d_slots[j], j < world (= slot `rank` of j's receive buffer): the GEMM epilogue is the scatter.
 *   d_slots[world], d_slots[world + 1]: two LOCAL [n, 2D] slots for the two K halves of the local-row terms.
 * b200clip_p2p_reduce_finish: *peer_flag[p] = epoch for every p, wait my_flags[q] >= epoch for all q < world, then
 *   out[elems] = sum over s < slots of recv[s * elems ...] (recv = the local receive buffer; slots = world + 2 here).
 *   split_cols = 2D > 0: the blocks are rows of 2D floats (img | txt gradient) and `out` receives them as two contiguous
 *   [rows, D] halves (d_img, then d_txt) instead of [rows, 2D] — dense tensors for autograd; 0: same layout as the blocks.
 *
 * Ring-slot protection: `my_busy` (may be NULL) is this rank's busy word of the ring slot in use — the all-gather sets it to
 * `epoch` when `hold` != 0 (a backward will read the gathered rows) or to 0, reduce_finish clears it; `peer_busy` (DEVICE array
 * of `world` pointers, may be NULL) are the peers' busy words of the same slot: a rank stores into peer p's slot only once
 * *peer_busy[p] is 0 or `epoch`.
 * Waits are bounded in wall time.  b200clip_p2p_configure sets the bound (seconds; default 600, or the environment variable
 * B200CLIP_P2P_TIMEOUT_S) and registers `error_word`, a DEVICE-ACCESSIBLE uint32 (pinned host memory; may be NULL): an expired
 * wait stores a non-zero code there (1 = a peer's flag never arrived, 2 = a peer's ring slot was never released) and lets the
 * kernel retire with undefined results — the host checks the word before trusting them.  Without an error word an expired
 * wait traps (sticky CUDA error). */
int b200clip_p2p_configure(double timeout_seconds, uint32_t* error_word);
int b200clip_p2p_allgather(int dtype, const void* img, const void* txt, int n, int D, float* const* peer_dst,
                           uint32_t* const* peer_flag, const uint32_t* my_flags, uint32_t* counters, int world,
                           uint32_t epoch, uint32_t* const* peer_busy, uint32_t* my_busy, int hold, void* stream);
int b200clip_cliploss_packed_backward_p2p(const float* gathered, const float* logit_scale, int rank, int n, int N, int D,
                                          const float* grad_out, float* const* d_slots, float* d_scale, float* workspace,
                                          void* stream);
int b200clip_p2p_reduce_finish(const float* recv, float* out, int64_t elems, uint32_t* const* peer_flag,
                               const uint32_t* my_flags, int world, int slots, uint32_t epoch, uint32_t* my_busy,
                               int split_cols, void* stream);

/* ----- whole-tower drivers: one call per encode_image / encode_text ------------------------------ */

typedef struct b200clip_tower_cfg {
    int32_t dtype;       /* B200CLIP_* */
    int32_t width;       /* W */
    int32_t layers;
    int32_t heads;       /* W / 64 */
    int32_t mlp_width;   /* 4W */
    int32_t embed_dim;   /* D */
    int32_t seq_len;     /* L: image tokens incl. CLS, or text context length */
    int32_t quick_gelu;  /* 0 = erf GELU, 1 = QuickGELU */
    int32_t image_size;  /* vision only */
    int32_t patch_size;  /* vision only */
    int32_t patch_kpad;  /* vision only: 3*P*P rounded up to a multiple of 64 */
    int32_t vocab_size;  /* text only */
    int32_t fold_ln;     /* 16-bit dtypes: run ln_1 / ln_2 folded into the QKV / c_fc GEMMs (needs the *_f weights below) */
} b200clip_tower_cfg;

/* per-layer weights, all device pointers.  16-bit modes: matrices/biases in `dtype`, LN params fp32. */
typedef struct b200clip_block_weights {
    const float *ln1_g, *ln1_b, *ln2_g, *ln2_b;
    const void *in_proj_w, *in_proj_b;   /* [3W,W], [3W] */
    const void *out_proj_w, *out_proj_b; /* [W,W],  [W]  */
    const void *fc_w, *fc_b;             /* [4W,W], [4W] */
    const void *proj_w, *proj_b;         /* [W,4W], [W]  */
    /* LN-fold operands (cfg.fold_ln; may be NULL otherwise): W diag(gamma) in dtype, its row sums and b + W beta in fp32 */
    const void* in_proj_wf;  const float *in_proj_c, *in_proj_bf;   /* [3W,W], [3W], [3W] */
    const void* fc_wf;       const float *fc_c, *fc_bf;             /* [4W,W], [4W], [4W] */
} b200clip_block_weights;

typedef struct b200clip_vit_weights {
    const void* conv1_w;      /* [W, patch_kpad] (K zero-padded), dtype */
    const float* class_emb;   /* [W] fp32 */
    const float* pos_emb;     /* [L, W] fp32 */
    const float *ln_pre_g, *ln_pre_b, *ln_post_g, *ln_post_b;
    const void* proj_t;       /* [D, W]: visual.proj transposed, dtype */
    const b200clip_block_weights* blocks_host; /* HOST array of `layers` entries */
    /* 16-bit modes, optional: [L, W] fp32 = pos_emb with row 0 replaced by dtype(class_emb) + dtype(pos_emb[0]); when set the
     * patch embedding runs as one token-layout GEMM on the CTA-pair kernel */
    const float* pos_cls;
} b200clip_vit_weights;

typedef struct b200clip_text_weights {
    const float* tok_emb;     /* [vocab, W] fp32 */
    const float* pos_emb;     /* [ctx, W] fp32 */
    const float *ln_final_g, *ln_final_b;
    const void* proj_t;       /* [D, W]: text_projection transposed, dtype */
    const b200clip_block_weights* blocks_host;
} b200clip_text_weights;

/* bytes of scratch a forward of `batch` items needs (seq_len from cfg; text may pass a truncated L) */
int64_t b200clip_workspace_bytes(const b200clip_tower_cfg* cfg, int batch, int seq_len);

/* The same forward from uint8 pixels [B,3,S,S] (0..255, already resized / centre-cropped): the preprocessing tail of the
 * reference — ToTensor (x / 255) and Normalize((x - mean) / std), deps/open_clip/src/open_clip/transform.py:274-392 with
 * constants.py:1-2 — runs inside the im2col, in fp32 with the reference's operation order, followed by the one rounding to
 * the tower dtype that `.half()` / `.to(bfloat16)` performs in the evaluation scripts.  `mean`, `std`: 3 HOST floats each.
 * b200clip_patchify_u8 is the im2col alone (patches [B*g*g, kpad]). */
int b200clip_vit_forward_u8(const b200clip_tower_cfg* cfg, const b200clip_vit_weights* w, const uint8_t* image,
                            const float* mean, const float* std, void* out, int batch, int normalize, void* workspace,
                            int64_t workspace_bytes, void* stream);
int b200clip_patchify_u8(int dtype, const uint8_t* image, const float* mean, const float* std, void* patches, int batch,
                         int image_size, int patch, int kpad, void* stream);

/* Bicubic resize + centre crop in front of the uint8 entry: a decoded image `src_hwc` [H, W, 3] uint8 (row pitch `row_stride`
 * bytes) -> `dst_chw` [3, out_h, out_w] uint8 = the crop window of torchvision's Resize(S, BICUBIC) + CenterCrop(S) on a PIL
 * image (deps/open_clip/src/open_clip/transform.py:372-392), BIT-IDENTICAL to Pillow's ImagingResample: two integer passes
 * (horizontal, then vertical) with Pillow's anti-aliasing bicubic coefficients in 22-bit fixed point and an 8-bit intermediate.
 * The host computes the tables the way Pillow does (open_clip/gpu_transform.py) for the crop window only:
 *   h_bounds [out_w][2] = (first source column, taps), h_coeffs [out_w][h_ksize] int32; v_bounds / v_coeffs likewise for the
 *   out_h output rows (source ROW indices); [y0, y0 + rows) = the source rows the vertical windows touch;
 *   tmp: rows * out_w * 3 bytes of scratch (the horizontally resampled rows). */
int b200clip_resize_crop_u8(const uint8_t* src_hwc, int H, int W, int64_t row_stride, const int32_t* h_bounds,
                            const int32_t* h_coeffs, int h_ksize, const int32_t* v_bounds, const int32_t* v_coeffs,
                            int v_ksize, int y0, int rows, uint8_t* tmp, uint8_t* dst_chw, int out_h, int out_w,
                            void* stream);

/* VisionTransformer.forward (transformer.py:601-643): image [B,3,S,S] dtype -> out [B,D] dtype
 * (L2-normalised when `normalize` != 0, CLIP.encode_image model.py:265-267). */
int b200clip_vit_forward(const b200clip_tower_cfg* cfg, const b200clip_vit_weights* w, const void* image, void* out,
                         int batch, int normalize, void* workspace, int64_t workspace_bytes, void* stream);

/* CLIP.encode_text (model.py:269-284): text int64 [T, ctx] -> out [T,D].  `seq_len` <= ctx is the number of
 * leading positions actually run (ctx = reference behaviour; max(eot)+1 is exact by causality). */
int b200clip_text_forward(const b200clip_tower_cfg* cfg, const b200clip_text_weights* w, const int64_t* text,
                          void* out, int batch, int seq_len, int normalize, void* workspace,
                          int64_t workspace_bytes, void* stream);

/* The same two forwards split at the points where caller-owned buffers are touched, so that a host can replay the long
 * middle part as ONE captured CUDA graph that depends on nothing but the workspace, whatever buffers the inputs arrive in and
 * the outputs go to (a DataLoader loop hands over a fresh tensor per batch, evaluate_domainnet_lso_openai.py:18-36):
 *   B200CLIP_STAGE_INPUT   the only kernels that read `image` / `image_u8` / `text`: the patch embedding up to the token rows
 *                          (16-bit NCHW batches, patch 16 / 32: an implicit GEMM that reads the patches from the image itself
 *                          through a 5-D tensor map; otherwise im2col + GEMM) resp. the embedding gather,
 *   B200CLIP_STAGE_BODY    ln_pre, all blocks, ln_post / ln_final on the pooled rows (workspace -> workspace),
 *   B200CLIP_STAGE_OUTPUT  projection (+ L2 normalisation): the only kernels that write `out`.
 * `stages` is a bit mask; pointers a selected stage does not use may be NULL.  All three stages in one call, or in three
 * calls on the same stream and workspace, give bit-identical results to b200clip_vit_forward(_u8) / b200clip_text_forward.
 * Exactly one of `image` (tower dtype) and `image_u8` (+ mean, std) is used by the vision input stage. */
#define B200CLIP_STAGE_INPUT 1
#define B200CLIP_STAGE_BODY 2
#define B200CLIP_STAGE_OUTPUT 4
int b200clip_vit_forward_stages(const b200clip_tower_cfg* cfg, const b200clip_vit_weights* w, const void* image,
                                const uint8_t* image_u8, const float* mean, const float* std, void* out, int batch,
                                int normalize, void* workspace, int64_t workspace_bytes, int stages, void* stream);
int b200clip_text_forward_stages(const b200clip_tower_cfg* cfg, const b200clip_text_weights* w, const int64_t* text,
                                 void* out, int batch, int seq_len, int normalize, void* workspace,
                                 int64_t workspace_bytes, int stages, void* stream);

/* ----- ModifiedResNet image tower (RN50 family) ----------------------------------------------------------------
 * ModifiedResNet.forward in eval mode (deps/open_clip/src/open_clip/modified_resnet.py:95-181): activations are NHWC rows
 * [B*H*W, C]; every convolution is a GEMM over those rows (3x3: over an im2col of them, K order (ky, kx, cin)) with
 * BatchNorm's inference statistics folded into the weight rows and a per-channel shift by the caller:
 *   w' = w * gamma / sqrt(running_var + eps),  b' = beta - running_mean * gamma / sqrt(running_var + eps).
 * Weights are [cout, K] in the tower dtype (K contiguous), shifts [cout] in the tower dtype. */
typedef struct b200clip_resnet_cfg {
    int32_t dtype;       /* B200CLIP_F32 / BF16 / F16 */
    int32_t image_size;  /* square input, multiple of 32 */
    int32_t width;       /* stem output channels (64 for RN50); the stage widths are width * (1, 2, 4, 8) */
    int32_t embed_dim;   /* attention-pool output dimension (CLIP embed_dim) */
    int32_t heads;       /* attention-pool heads = width * 32 / 64 */
    int32_t n_blocks;    /* bottlenecks over all four stages */
    int32_t stem_kpad;   /* K of the stem conv1 GEMM: 27 taps zero-padded to whole 16-byte vectors */
} b200clip_resnet_cfg;

/* One Bottleneck (modified_resnet.py:10-56): conv1 1x1 [planes, cin], conv2 3x3 [planes, 9*planes], conv3 1x1
 * [4*planes, planes]; `stride` 2 = AvgPool2d(2) after conv2 and in front of the downsample convolution;
 * down_w [4*planes, cin] (NULL when the block has no downsample branch: stride 1 and cin == 4*planes). */
typedef struct b200clip_resnet_block {
    const void *conv1_w, *conv1_b, *conv2_w, *conv2_b, *conv3_w, *conv3_b, *down_w, *down_b;
    int32_t cin, planes, stride, reserved;
} b200clip_resnet_block;

typedef struct b200clip_resnet_weights {
    const void* stem_w[3];   /* [width/2, stem_kpad], [width/2, 9*width/2], [width, 9*width/2] */
    const void* stem_b[3];
    const b200clip_resnet_block* blocks_host;  /* HOST array of n_blocks entries (pointers inside are device pointers) */
    const float* pos;        /* attnpool.positional_embedding [HW + 1, 32*width] fp32 */
    const void* qkv_w;       /* [3*E, E] = q_proj | k_proj | v_proj weights stacked (E = 32*width), tower dtype */
    const void* qkv_b;       /* [3*E] */
    const void* c_proj_w;    /* [embed_dim, E] */
    const void* c_proj_b;    /* [embed_dim] */
} b200clip_resnet_weights;

int64_t b200clip_resnet_workspace_bytes(const b200clip_resnet_cfg* cfg, const b200clip_resnet_weights* w, int batch);
/* image [B,3,S,S] tower dtype -> out [B, embed_dim] (L2-normalised when `normalize` != 0); `stages` as for the ViT:
 * INPUT = the stem im2col (the only kernel that reads `image`), BODY = stem GEMMs ... attention pool (workspace only),
 * OUTPUT = attnpool.c_proj of the pooled token (+ normalise), the only kernels that write `out`. */
int b200clip_resnet_forward_stages(const b200clip_resnet_cfg* cfg, const b200clip_resnet_weights* w, const void* image,
                                   void* out, int batch, int normalize, void* workspace, int64_t workspace_bytes,
                                   int stages, void* stream);

/* The tower's data-movement operators on NHWC rows, also callable on their own:
 *   stem_im2col      image [B,3,S,S] -> [B*(S/2)^2, kpad], conv1's 3x3 / stride 2 / padding 1 windows, K = (ky*3+kx)*3 + c
 *                    (modified_resnet.py:107), columns 27.. zero;
 *   im2col3x3        [B*H*W, C] -> [B*H*W, 9*C], 3x3 / stride 1 / padding 1 windows (modified_resnet.py:21,110,113);
 *   avgpool2         nn.AvgPool2d(2) on [B,H,W,C] -> [B,H/2,W/2,C] (modified_resnet.py:25,36,116), fp32 sum, one rounding;
 *   attnpool_tokens  [B*HW, C] -> [B*(HW+1), C]: the mean token in front, positional embedding added
 *                    (AttentionPool2d.forward, modified_resnet.py:70-72). */
int b200clip_stem_im2col(int dtype, const void* image, void* out, int batch, int image_size, int kpad, void* stream);
int b200clip_im2col3x3(int dtype, const void* in, void* out, int batch, int H, int W, int C, void* stream);
int b200clip_avgpool2(int dtype, const void* in, void* out, int batch, int H, int W, int C, void* stream);
int b200clip_attnpool_tokens(int dtype, const void* x, const float* pos, void* tok, int batch, int HW, int C, void* stream);

/* ----- training path: tower backward + fused optimizer -------------------------------------------------------
 * What autograd does for the reference's training step (deps/open_clip/src/training/train.py:115-183) through
 * CLIP.encode_image / encode_text with --grad-checkpointing (transformer.py:353-355): the training forward keeps only the
 * residual stream at every block boundary (`saved`, b200clip_train_saved_bytes bytes), the backward recomputes each block
 * and produces the gradient of every parameter.  Gradient buffers are OVERWRITTEN (the caller accumulates); layouts and
 * dtypes mirror the weight structs: matrices / biases in `cfg.dtype`, LayerNorm parameters and embedding tables fp32. */
typedef struct b200clip_block_grads {
    float *ln1_g, *ln1_b, *ln2_g, *ln2_b;
    void *in_proj_w, *in_proj_b;   /* [3W,W], [3W] */
    void *out_proj_w, *out_proj_b; /* [W,W],  [W]  */
    void *fc_w, *fc_b;             /* [4W,W], [4W] */
    void *proj_w, *proj_b;         /* [W,4W], [W]  */
} b200clip_block_grads;

typedef struct b200clip_vit_grads {
    void* conv1_w;       /* [W, patch_kpad] (the padding columns are don't-care) */
    float* class_emb;    /* [W] */
    float* pos_emb;      /* [L, W] */
    float *ln_pre_g, *ln_pre_b, *ln_post_g, *ln_post_b;
    void* proj;          /* [W, D]: gradient of visual.proj in ITS layout (not transposed) */
    const b200clip_block_grads* blocks_host; /* HOST array of `layers` entries */
} b200clip_vit_grads;

typedef struct b200clip_text_grads {
    float* tok_emb;      /* [vocab, W] (zeroed, then scatter-added) */
    float* pos_emb;      /* [ctx, W] */
    float *ln_final_g, *ln_final_b;
    void* proj;          /* [W, D]: gradient of text_projection */
    const b200clip_block_grads* blocks_host;
} b200clip_text_grads;

int64_t b200clip_train_saved_bytes(const b200clip_tower_cfg* cfg, int batch, int seq_len);
int64_t b200clip_backward_workspace_bytes(const b200clip_tower_cfg* cfg, int batch, int seq_len);
/* forward of b200clip_vit_forward / b200clip_text_forward that also fills `saved` */
int b200clip_vit_forward_train(const b200clip_tower_cfg* cfg, const b200clip_vit_weights* w, const void* image, void* out,
                               int batch, int normalize, void* saved, int64_t saved_bytes, void* workspace,
                               int64_t workspace_bytes, void* stream);
int b200clip_text_forward_train(const b200clip_tower_cfg* cfg, const b200clip_text_weights* w, const int64_t* text,
                                void* out, int batch, int seq_len, int normalize, void* saved, int64_t saved_bytes,
                                void* workspace, int64_t workspace_bytes, void* stream);
/* d_out [batch, D] = gradient of the features the training forward returned (normalised when `normalize` != 0; same flag as
 * in the forward).  `workspace`: b200clip_backward_workspace_bytes bytes, 256-byte aligned.  The weights must not have
 * changed since the forward.  cfg.fold_ln is ignored (the recompute runs the LayerNorms as kernels). */
int b200clip_vit_backward(const b200clip_tower_cfg* cfg, const b200clip_vit_weights* w, const void* image, const void* d_out,
                          int batch, int normalize, const void* saved, const b200clip_vit_grads* grads, void* workspace,
                          int64_t workspace_bytes, void* stream);
int b200clip_text_backward(const b200clip_tower_cfg* cfg, const b200clip_text_weights* w, const int64_t* text,
                           const void* d_out, int batch, int seq_len, int normalize, const void* saved,
                           const b200clip_text_grads* grads, void* workspace, int64_t workspace_bytes, void* stream);

/* Fused multi-tensor AdamW (torch.optim.AdamW semantics; training/main.py:299-326): `items` = DEVICE table of `n` tensors,
 * `chunk_item` / `chunk_off` = DEVICE lists of `chunks` work items (tensor index, element offset; 4096 elements each,
 * b200clip_adamw_chunk()).  grad is read as grad * grad_scale (loss-scaling / accumulation averaging); `step` >= 1. */
typedef struct b200clip_adamw_tensor {
    void* param;
    const void* grad;
    float* exp_avg;
    float* exp_avg_sq;
    int64_t count;
    int32_t param_dtype; /* B200CLIP_* */
    int32_t grad_dtype;
} b200clip_adamw_tensor;
int b200clip_adamw_chunk(void);
int b200clip_adamw_step(const b200clip_adamw_tensor* items, const int32_t* chunk_item, const int64_t* chunk_off, int chunks,
                        float lr, float beta1, float beta2, float eps, float weight_decay, int step, float grad_scale,
                        void* stream);


/* Multi-tensor conversion: dst[i] = (dst dtype) src[i] for every tensor of a device table, ONE launch (chunks of
 * b200clip_adamw_chunk() elements, as for b200clip_adamw_step).  The towers' engines refresh their 16-bit operand copies from the
 * fp32 master parameters with it after an optimizer step (precision amp / amp_bf16: convert_weights_to_lp semantics,
 * open_clip/model.py:396-423, applied to the whole parameter list at once). */
typedef struct b200clip_cast_tensor {
    const void* src;
    void* dst;
    int64_t count;
    int32_t src_dtype; /* B200CLIP_* */
    int32_t dst_dtype;
} b200clip_cast_tensor;
int b200clip_multi_cast(const b200clip_cast_tensor* items, const int32_t* chunk_item, const int64_t* chunk_off, int chunks,
                        void* stream);

#ifdef __cplusplus
}
#endif
#endif /* B200CLIP_H_ */
