// extern "C" surface of libb200clip.so (declared in include/b200clip.h) + error / bookkeeping helpers.
#include "../../include/b200clip.h"

#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <mutex>

#include "common.cuh"
#include "internal.h"

namespace b200clip {

static thread_local char g_err[512] = "";
static std::atomic<uint64_t> g_launches{0};

void set_last_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what) {
    set_last_error("CUDA error %d (%s) at %s", static_cast<int>(e), cudaGetErrorString(e), what);
    return static_cast<int>(e) > 0 ? static_cast<int>(e) : 1;
}

void count_launch(int n) { g_launches.fetch_add(static_cast<uint64_t>(n), std::memory_order_relaxed); }

bool pdl_enabled() {
    static const bool v = [] {
        const char* e = getenv("B200CLIP_NO_PDL");
        return !(e != nullptr && e[0] == '1');
    }();
    return v;
}

int num_sms() {
    // per-device cache: one process drives one GPU, but stay correct if the current device changes
    static int cached_dev = -1;
    static int cached = 0;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (dev != cached_dev) {
        int v = 0;
        if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) v = 148;
        cached = v;
        cached_dev = dev;
    }
    return cached;
}

// B200CLIP_GEMM_SINGLE_CTA=1 routes every 16-bit GEMM to the single-CTA kernel (A/B measurements only)
static bool force_single_cta_gemm() {
    static const bool v = [] {
        const char* e = getenv("B200CLIP_GEMM_SINGLE_CTA");
        return e != nullptr && e[0] == '1';
    }();
    return v;
}

int gemm_any(int dtype, const void* A, int64_t lda, const void* W, int64_t ldw, const void* bias, const void* residual,
             int64_t ldr, void* C, int64_t ldc, int M, int N, int K, int epilogue, const float* pos, int g_in, int g_out,
             cudaStream_t stream, void* sk_workspace) {
    if (dtype == B200CLIP_F32)
        return gemm_f32(static_cast<const float*>(A), lda, static_cast<const float*>(W), ldw, static_cast<const float*>(bias),
                        static_cast<const float*>(residual), ldr, static_cast<float*>(C), ldc, M, N, K, epilogue, pos, g_in, g_out,
                        stream);
    if (dtype == B200CLIP_BF16 || dtype == B200CLIP_F16) {
        // the patch-embedding epilogue remaps rows (no TMA-store box), it stays on the single-CTA kernel
        if (epilogue != B200CLIP_EPI_PATCH && (!force_single_cta_gemm() || epilogue >= B200CLIP_EPI_RELU))
            return gemm_pair(dtype == B200CLIP_BF16, A, lda, W, ldw, bias, residual, ldr, C, ldc, M, N, K, epilogue, 0, 0, stream, nullptr,
                             nullptr, nullptr, 0, nullptr, nullptr, 0, 1e-5f, sk_workspace);
        return gemm_tc(dtype == B200CLIP_BF16, A, lda, W, ldw, bias, residual, ldr, C, ldc, M, N, K, epilogue, pos, g_in, g_out, 0,
                       stream);
    }
    set_last_error("gemm: unknown dtype %d", dtype);
    return -1;
}

int64_t workspace_bytes(const b200clip_tower_cfg* cfg, int batch, int seq_len);
int vit_forward(const b200clip_tower_cfg* cfg, const b200clip_vit_weights* w, const void* image, void* out, int batch,
                int normalize, void* workspace, int64_t workspace_bytes_, cudaStream_t s);
int vit_forward_u8(const b200clip_tower_cfg* cfg, const b200clip_vit_weights* w, const uint8_t* image, const float* mean,
                   const float* std, void* out, int batch, int normalize, void* workspace, int64_t workspace_bytes_, cudaStream_t s);
int text_forward(const b200clip_tower_cfg* cfg, const b200clip_text_weights* w, const int64_t* text, void* out, int batch,
                 int seq_len, int normalize, void* workspace, int64_t workspace_bytes_, cudaStream_t s);
int64_t train_saved_bytes(const b200clip_tower_cfg* cfg, int batch, int seq_len);
int64_t backward_workspace_bytes(const b200clip_tower_cfg* cfg, int batch, int seq_len);
int vit_forward_train(const b200clip_tower_cfg* cfg, const b200clip_vit_weights* w, const void* image, void* out, int batch, int normalize,
                      void* saved, int64_t saved_bytes, void* workspace, int64_t workspace_bytes_, cudaStream_t s);
int text_forward_train(const b200clip_tower_cfg* cfg, const b200clip_text_weights* w, const int64_t* text, void* out, int batch, int seq_len,
                       int normalize, void* saved, int64_t saved_bytes, void* workspace, int64_t workspace_bytes_, cudaStream_t s);
int vit_backward(const b200clip_tower_cfg* cfg, const b200clip_vit_weights* w, const void* image, const void* d_out, int batch, int normalize,
                 const void* saved, const b200clip_vit_grads* g, void* workspace, int64_t workspace_bytes_, cudaStream_t s);
int text_backward(const b200clip_tower_cfg* cfg, const b200clip_text_weights* w, const int64_t* text, const void* d_out, int batch, int seq_len,
                  int normalize, const void* saved, const b200clip_text_grads* g, void* workspace, int64_t workspace_bytes_, cudaStream_t s);
int vit_forward_stages(const b200clip_tower_cfg* cfg, const b200clip_vit_weights* w, const void* image, const uint8_t* image_u8,
                       const float* mean, const float* std, void* out, int batch, int normalize, void* workspace,
                       int64_t workspace_bytes_, int stages, cudaStream_t s);
int text_forward_stages(const b200clip_tower_cfg* cfg, const b200clip_text_weights* w, const int64_t* text, void* out, int batch,
                        int seq_len, int normalize, void* workspace, int64_t workspace_bytes_, int stages, cudaStream_t s);

int64_t resnet_workspace_bytes(const b200clip_resnet_cfg* cfg, const b200clip_resnet_weights* w, int batch);
int resnet_forward_stages(const b200clip_resnet_cfg* cfg, const b200clip_resnet_weights* w, const void* image, void* out, int batch,
                          int normalize, void* workspace, int64_t workspace_bytes_, int stages, cudaStream_t s);
int stem_im2col(int dtype, const void* image, void* out, int batch, int image_size, int kpad, cudaStream_t s);
int im2col3x3(int dtype, const void* in, void* out, int batch, int H, int W, int C, cudaStream_t s);
int avgpool2(int dtype, const void* in, void* out, int batch, int H, int W, int C, cudaStream_t s);
int attnpool_tokens(int dtype, const void* x, const float* pos, void* tok, int batch, int HW, int C, cudaStream_t s);

}  // namespace b200clip

using namespace b200clip;

static inline cudaStream_t S(void* s) { return static_cast<cudaStream_t>(s); }

extern "C" {

int b200clip_version(void) { return 100; }  // 0.1.0

const char* b200clip_last_error(void) { return g_err; }

uint64_t b200clip_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

int b200clip_gemm(int dtype, const void* A, int64_t lda, const void* W, int64_t ldw, const void* bias, const void* residual,
                  int64_t ldr, void* C, int64_t ldc, int M, int N, int K, int epilogue, const float* pos, int g_in, int g_out,
                  void* stream) {
    B2C_CHECK_ARG(A != nullptr && W != nullptr && C != nullptr, "gemm: null pointer");
    B2C_CHECK_ARG(epilogue >= 0 && epilogue <= 6, "gemm: unknown epilogue %d", epilogue);
    return gemm_any(dtype, A, lda, W, ldw, bias, residual, ldr, C, ldc, M, N, K, epilogue, pos, g_in, g_out, S(stream));
}

int64_t b200clip_gemm_workspace_bytes(void) { return gemm_pair_sk_workspace_bytes(); }

int b200clip_gemm_ws(int dtype, const void* A, int64_t lda, const void* W, int64_t ldw, const void* bias, const void* residual,
                     int64_t ldr, void* C, int64_t ldc, int M, int N, int K, int epilogue, void* workspace, int64_t workspace_bytes,
                     void* stream) {
    B2C_CHECK_ARG(A != nullptr && W != nullptr && C != nullptr, "gemm_ws: null pointer");
    B2C_CHECK_ARG((epilogue >= 0 && epilogue <= 3) || epilogue == B200CLIP_EPI_RELU || epilogue == B200CLIP_EPI_RESIDUAL_RELU,
                  "gemm_ws: epilogue must be BIAS, GELU, QUICKGELU, RESIDUAL, RELU or RESIDUAL_RELU");
    if (dtype == B200CLIP_F32)  // the FFMA parity kernel tiles finely enough: no workspace needed
        return gemm_any(dtype, A, lda, W, ldw, bias, residual, ldr, C, ldc, M, N, K, epilogue, nullptr, 0, 0, S(stream));
    B2C_CHECK_ARG(workspace != nullptr && workspace_bytes >= gemm_pair_sk_workspace_bytes(), "gemm_ws: workspace too small (%lld < %lld bytes)",
                  (long long)workspace_bytes, (long long)gemm_pair_sk_workspace_bytes());
    int rc;
    if ((rc = gemm_pair_sk_workspace_reset(workspace, S(stream))) != 0) return rc;
    return gemm_any(dtype, A, lda, W, ldw, bias, residual, ldr, C, ldc, M, N, K, epilogue, nullptr, 0, 0, S(stream), workspace);
}

int b200clip_gemm_mn(int dtype, const void* A, int64_t lda, int a_mn, const void* W, int64_t ldw, void* C, int64_t ldc, int M, int N, int K,
                     void* workspace, int64_t workspace_bytes, void* stream) {
    B2C_CHECK_ARG(A != nullptr && W != nullptr && C != nullptr, "gemm_mn: null pointer");
    B2C_CHECK_ARG(dtype == B200CLIP_BF16 || dtype == B200CLIP_F16, "gemm_mn: 16-bit dtypes only (fp32 runs b200clip's FFMA GEMM with transpose flags)");
    if (workspace != nullptr) {
        B2C_CHECK_ARG(workspace_bytes >= gemm_pair_sk_workspace_bytes(), "gemm_mn: workspace too small (%lld < %lld bytes)",
                      (long long)workspace_bytes, (long long)gemm_pair_sk_workspace_bytes());
        int rc;
        if ((rc = gemm_pair_sk_workspace_reset(workspace, S(stream))) != 0) return rc;
    }
    return gemm_pair_mn(dtype == B200CLIP_BF16, A, lda, a_mn != 0, W, ldw, true, C, ldc, M, N, K, S(stream), workspace);
}

int b200clip_patch_embed_implicit(int dtype, const void* image, const void* conv1_w, const float* pos_cls, void* x, int batch,
                                  int image_size, int patch, int width, void* stream) {
    B2C_CHECK_ARG(dtype == B200CLIP_BF16 || dtype == B200CLIP_F16, "patch_embed_implicit: 16-bit dtypes only");
    const int rc = gemm_pair_patch_embed(dtype == B200CLIP_BF16, image, conv1_w, pos_cls, x, batch, image_size, patch, width, S(stream));
    B2C_CHECK_ARG(rc != 1, "patch_embed_implicit: patch size %d is not covered (16-byte pixel rows that divide 64 elements: 16, 32)", patch);
    return rc;
}

int b200clip_gemm_ln_ws(int dtype, const void* x, int64_t ldx, const void* Wf, int64_t ldw, const float* colsum, const float* bias_f32,
                        const float* rowstats, void* C, int64_t ldc, int M, int N, int K, int epilogue, void* workspace,
                        int64_t workspace_bytes, void* stream) {
    B2C_CHECK_ARG(x && Wf && colsum && bias_f32 && rowstats && C, "gemm_ln_ws: null pointer");
    B2C_CHECK_ARG(dtype == B200CLIP_BF16 || dtype == B200CLIP_F16, "gemm_ln_ws: 16-bit dtypes only");
    B2C_CHECK_ARG(epilogue >= 0 && epilogue <= 2, "gemm_ln_ws: epilogue must be BIAS, GELU or QUICKGELU");
    B2C_CHECK_ARG(workspace != nullptr && workspace_bytes >= gemm_pair_sk_workspace_bytes(), "gemm_ln_ws: workspace too small (%lld < %lld bytes)",
                  (long long)workspace_bytes, (long long)gemm_pair_sk_workspace_bytes());
    int rc;
    if ((rc = gemm_pair_sk_workspace_reset(workspace, S(stream))) != 0) return rc;
    return gemm_pair(dtype == B200CLIP_BF16, x, ldx, Wf, ldw, bias_f32, nullptr, 0, C, ldc, M, N, K, epilogue, 0, 0, S(stream), colsum,
                     rowstats, nullptr, 0, nullptr, nullptr, 0, 1e-5f, workspace);
}

/* test hook: same as b200clip_gemm for 16-bit dtypes but with a forced N tile (128 or 256) */
int b200clip_gemm_tile(int dtype, const void* A, int64_t lda, const void* W, int64_t ldw, const void* bias, const void* residual,
                       int64_t ldr, void* C, int64_t ldc, int M, int N, int K, int epilogue, const float* pos, int g_in,
                       int g_out, int block_n, void* stream) {
    B2C_CHECK_ARG(A != nullptr && W != nullptr && C != nullptr, "gemm: null pointer");
    B2C_CHECK_ARG(dtype == B200CLIP_BF16 || dtype == B200CLIP_F16, "gemm_tile: 16-bit dtypes only");
    B2C_CHECK_ARG(epilogue >= 0 && epilogue <= 4, "gemm: unknown epilogue %d", epilogue);
    if (block_n >= 1000)  // 1000 + BLOCK_N: CTA-pair kernel (clusters of 2); 2000 + BLOCK_N: clusters of 4 with W multicast
        return gemm_pair(dtype == B200CLIP_BF16, A, lda, W, ldw, bias, residual, ldr, C, ldc, M, N, K, epilogue, block_n % 1000,
                         block_n / 1000, S(stream));
    return gemm_tc(dtype == B200CLIP_BF16, A, lda, W, ldw, bias, residual, ldr, C, ldc, M, N, K, epilogue, pos, g_in, g_out, block_n,
                   S(stream));
}

int b200clip_gemm_ln(int dtype, const void* x, int64_t ldx, const void* Wf, int64_t ldw, const float* colsum, const float* bias_f32,
                     const float* rowstats, void* C, int64_t ldc, int M, int N, int K, int epilogue, void* stream) {
    B2C_CHECK_ARG(x && Wf && colsum && bias_f32 && rowstats && C, "gemm_ln: null pointer");
    B2C_CHECK_ARG(dtype == B200CLIP_BF16 || dtype == B200CLIP_F16, "gemm_ln: 16-bit dtypes only (the fp32 mode keeps the LayerNorm kernel)");
    B2C_CHECK_ARG(epilogue >= 0 && epilogue <= 2, "gemm_ln: epilogue must be BIAS, GELU or QUICKGELU");
    return gemm_pair(dtype == B200CLIP_BF16, x, ldx, Wf, ldw, bias_f32, nullptr, 0, C, ldc, M, N, K, epilogue, 0, 0, S(stream), colsum,
                     rowstats);
}

int b200clip_gemm_stats_slots(int M, int N) {
    B2C_CHECK_ARG(M > 0 && N > 0, "gemm_stats_slots: empty problem");
    return gemm_pair_stats_slots(M, N);
}

int b200clip_gemm_residual_stats(int dtype, const void* A, int64_t lda, const void* W, int64_t ldw, const void* bias,
                                 const void* residual, int64_t ldr, void* C, int64_t ldc, int M, int N, int K, float* partials,
                                 void* stream) {
    B2C_CHECK_ARG(A && W && C && residual && partials, "gemm_residual_stats: null pointer");
    B2C_CHECK_ARG(dtype == B200CLIP_BF16 || dtype == B200CLIP_F16, "gemm_residual_stats: 16-bit dtypes only");
    return gemm_pair(dtype == B200CLIP_BF16, A, lda, W, ldw, bias, residual, ldr, C, ldc, M, N, K, B200CLIP_EPI_RESIDUAL, 0, 0, S(stream),
                     nullptr, nullptr, nullptr, 0, partials);
}

int b200clip_gemm_ln_partials(int dtype, const void* x, int64_t ldx, const void* Wf, int64_t ldw, const float* colsum,
                              const float* bias_f32, const float* partials, int slots, float eps, void* C, int64_t ldc, int M, int N,
                              int K, int epilogue, void* stream) {
    B2C_CHECK_ARG(x && Wf && colsum && bias_f32 && partials && C, "gemm_ln_partials: null pointer");
    B2C_CHECK_ARG(dtype == B200CLIP_BF16 || dtype == B200CLIP_F16, "gemm_ln_partials: 16-bit dtypes only");
    B2C_CHECK_ARG(epilogue >= 0 && epilogue <= 2, "gemm_ln_partials: epilogue must be BIAS, GELU or QUICKGELU");
    return gemm_pair(dtype == B200CLIP_BF16, x, ldx, Wf, ldw, bias_f32, nullptr, 0, C, ldc, M, N, K, epilogue, 0, 0, S(stream), colsum,
                     nullptr, nullptr, 0, nullptr, partials, slots, eps);
}

int b200clip_row_stats(int dtype, const void* x, int64_t ldx, float* stats, int rows, int width, float eps, void* stream) {
    B2C_CHECK_ARG(x && stats, "row_stats: null pointer");
    return row_stats(dtype, x, ldx, stats, rows, width, eps, S(stream));
}

int b200clip_layernorm(int dtype, const void* x, int64_t ldx, const float* gamma, const float* beta, void* y, int64_t ldy,
                       int rows, int width, float eps, int row_stride_rows, const int32_t* row_idx, void* stream) {
    B2C_CHECK_ARG(x && gamma && beta && y, "layernorm: null pointer");
    return layernorm(dtype, x, ldx, gamma, beta, y, ldy, rows, width, eps, row_stride_rows, row_idx, S(stream));
}

int b200clip_attention(int dtype, const void* qkv, void* out, int batch, int seq_len, int heads, int causal, void* stream) {
    B2C_CHECK_ARG(qkv && out, "attention: null pointer");
    return attention(dtype, qkv, out, batch, seq_len, heads, causal, S(stream));
}

int b200clip_patchify(int dtype, const void* image, void* patches, int batch, int image_size, int patch, int kpad,
                      const float* class_emb, const float* pos, void* x, int width, void* stream) {
    B2C_CHECK_ARG(image && patches, "patchify: null pointer");
    return patchify(dtype, image, patches, batch, image_size, patch, kpad, class_emb, pos, x, width, S(stream));
}

int b200clip_text_embed(int dtype, const int64_t* text, int ctx, const float* tok_emb, const float* pos_emb, void* x, int32_t* eot,
                        int T, int L, int width, void* stream) {
    B2C_CHECK_ARG(text && tok_emb && pos_emb && x, "text_embed: null pointer");
    return text_embed(dtype, text, ctx, tok_emb, pos_emb, x, eot, T, L, width, S(stream));
}

int b200clip_eot_argmax(const int64_t* text, int ctx, int32_t* eot, int T, void* stream) {
    B2C_CHECK_ARG(text && eot, "eot_argmax: null pointer");
    return eot_argmax(text, ctx, eot, T, S(stream));
}

int b200clip_normalize(int dtype, const void* x, int64_t ldx, void* y, int64_t ldy, int rows, int dim, float eps, void* stream) {
    B2C_CHECK_ARG(x && y, "normalize: null pointer");
    return normalize_rows(dtype, x, ldx, y, ldy, rows, dim, eps, S(stream));
}

int b200clip_zeroshot(int dtype, const void* img_feat, const void* prompt_feat, float* logits, int64_t* topk_idx, float* topk_val,
                      int B, int C, int D, int k, int normalize_img, float logit_scale, void* stream) {
    B2C_CHECK_ARG(img_feat && prompt_feat, "zeroshot: null pointer");
    return zeroshot(dtype, img_feat, prompt_feat, logits, topk_idx, topk_val, B, C, D, k, normalize_img, logit_scale, S(stream));
}

int b200clip_topk(int dtype, const void* x, int64_t ldx, int B, int C, int k, int64_t* topk_idx, float* topk_val, float* logits_out,
                  void* stream) {
    return topk_rows(dtype, x, ldx, B, C, k, topk_idx, topk_val, logits_out, S(stream));
}

int b200clip_class_mean(int dtype, const void* txt_feat, void* prompt_feat, int classes, int templates, int D, void* stream) {
    B2C_CHECK_ARG(txt_feat && prompt_feat, "class_mean: null pointer");
    return class_mean(dtype, txt_feat, prompt_feat, classes, templates, D, S(stream));
}

int b200clip_cliploss(const float* img_loc, const float* txt_loc, const float* all_img, const float* all_txt,
                      const float* logit_scale, int rank, int n, int N, int D, float* loss, const float* grad_out,
                      float* d_img_loc, float* d_txt_loc, float* d_all_img, float* d_all_txt, float* d_scale, float* workspace,
                      void* stream) {
    return cliploss(img_loc, txt_loc, all_img, all_txt, logit_scale, rank, n, N, D, loss, grad_out, d_img_loc, d_txt_loc, d_all_img,
                    d_all_txt, d_scale, workspace, S(stream));
}

int b200clip_cliploss_forward(const float* img_loc, const float* txt_loc, const float* all_img, const float* all_txt,
                              const float* logit_scale, int rank, int n, int N, int D, float* loss, float* workspace, void* stream) {
    return cliploss_forward(img_loc, txt_loc, all_img, all_txt, logit_scale, rank, n, N, D, loss, workspace, S(stream));
}

int b200clip_cliploss_backward(const float* img_loc, const float* txt_loc, const float* all_img, const float* all_txt,
                               const float* logit_scale, int rank, int n, int N, int D, const float* grad_out, float* d_img_loc,
                               float* d_txt_loc, float* d_all_img, float* d_all_txt, float* d_scale, float* workspace, void* stream) {
    return cliploss_backward(img_loc, txt_loc, all_img, all_txt, logit_scale, rank, n, N, D, grad_out, d_img_loc, d_txt_loc, d_all_img,
                             d_all_txt, d_scale, workspace, S(stream));
}

int b200clip_cliploss_single_backward(const float* img, const float* txt, const float* logit_scale, int n, int D,
                                      const float* grad_out, float* d_img, float* d_txt, float* d_scale, float* workspace,
                                      void* stream) {
    return cliploss_single_backward(img, txt, logit_scale, n, D, grad_out, d_img, d_txt, d_scale, workspace, S(stream));
}

int b200clip_cliploss_packed_forward(const float* gathered, const float* logit_scale, int rank, int n, int N, int D, float* loss,
                                     float* workspace, void* stream) {
    return cliploss_packed_forward(gathered, logit_scale, rank, n, N, D, loss, workspace, S(stream));
}

int b200clip_cliploss_packed_backward(const float* gathered, const float* logit_scale, int rank, int n, int N, int D,
                                      const float* grad_out, float* d_gathered, float* d_scale, float* workspace, void* stream) {
    return cliploss_packed_backward(gathered, logit_scale, rank, n, N, D, grad_out, d_gathered, d_scale, workspace, S(stream));
}

int b200clip_p2p_configure(double timeout_seconds, uint32_t* error_word) { return p2p_configure(timeout_seconds, error_word); }

int b200clip_p2p_allgather(int dtype, const void* img, const void* txt, int n, int D, float* const* peer_dst,
                           uint32_t* const* peer_flag, const uint32_t* my_flags, uint32_t* counters, int world, uint32_t epoch,
                           uint32_t* const* peer_busy, uint32_t* my_busy, int hold, void* stream) {
    return p2p_allgather(dtype, img, txt, n, D, peer_dst, peer_flag, my_flags, counters, world, epoch, peer_busy, my_busy, hold,
                         S(stream));
}

int b200clip_cliploss_packed_backward_p2p(const float* gathered, const float* logit_scale, int rank, int n, int N, int D,
                                          const float* grad_out, float* const* d_slots, float* d_scale, float* workspace,
                                          void* stream) {
    return cliploss_packed_backward_p2p(gathered, logit_scale, rank, n, N, D, grad_out, d_slots, d_scale, workspace, S(stream));
}

int b200clip_p2p_reduce_finish(const float* recv, float* out, int64_t elems, uint32_t* const* peer_flag, const uint32_t* my_flags,
                               int world, int slots, uint32_t epoch, uint32_t* my_busy, int split_cols, void* stream) {
    return p2p_reduce_finish(recv, out, elems, peer_flag, my_flags, world, slots, epoch, my_busy, split_cols, S(stream));
}

// not part of the public header: timeline capture of the tcgen05 attention kernel (tools/attn_timeline.py)
void b200clip_attention_debug(long long* buf) { attention_tc_set_debug(buf); }

int64_t b200clip_workspace_bytes(const b200clip_tower_cfg* cfg, int batch, int seq_len) {
    return workspace_bytes(cfg, batch, seq_len);
}

int b200clip_vit_forward(const b200clip_tower_cfg* cfg, const b200clip_vit_weights* w, const void* image, void* out, int batch,
                         int normalize, void* workspace, int64_t workspace_bytes, void* stream) {
    return vit_forward(cfg, w, image, out, batch, normalize, workspace, workspace_bytes, S(stream));
}

int b200clip_vit_forward_u8(const b200clip_tower_cfg* cfg, const b200clip_vit_weights* w, const uint8_t* image, const float* mean,
                            const float* std, void* out, int batch, int normalize, void* workspace, int64_t workspace_bytes,
                            void* stream) {
    return vit_forward_u8(cfg, w, image, mean, std, out, batch, normalize, workspace, workspace_bytes, S(stream));
}

int b200clip_vit_forward_stages(const b200clip_tower_cfg* cfg, const b200clip_vit_weights* w, const void* image,
                                const uint8_t* image_u8, const float* mean, const float* std, void* out, int batch, int normalize,
                                void* workspace, int64_t workspace_bytes, int stages, void* stream) {
    return vit_forward_stages(cfg, w, image, image_u8, mean, std, out, batch, normalize, workspace, workspace_bytes, stages, S(stream));
}

int b200clip_text_forward_stages(const b200clip_tower_cfg* cfg, const b200clip_text_weights* w, const int64_t* text, void* out,
                                 int batch, int seq_len, int normalize, void* workspace, int64_t workspace_bytes, int stages,
                                 void* stream) {
    return text_forward_stages(cfg, w, text, out, batch, seq_len, normalize, workspace, workspace_bytes, stages, S(stream));
}

int64_t b200clip_resnet_workspace_bytes(const b200clip_resnet_cfg* cfg, const b200clip_resnet_weights* w, int batch) {
    return resnet_workspace_bytes(cfg, w, batch);
}

int b200clip_resnet_forward_stages(const b200clip_resnet_cfg* cfg, const b200clip_resnet_weights* w, const void* image, void* out,
                                   int batch, int normalize, void* workspace, int64_t workspace_bytes, int stages, void* stream) {
    return resnet_forward_stages(cfg, w, image, out, batch, normalize, workspace, workspace_bytes, stages, S(stream));
}

int b200clip_stem_im2col(int dtype, const void* image, void* out, int batch, int image_size, int kpad, void* stream) {
    return stem_im2col(dtype, image, out, batch, image_size, kpad, S(stream));
}

int b200clip_im2col3x3(int dtype, const void* in, void* out, int batch, int H, int W, int C, void* stream) {
    return im2col3x3(dtype, in, out, batch, H, W, C, S(stream));
}

int b200clip_avgpool2(int dtype, const void* in, void* out, int batch, int H, int W, int C, void* stream) {
    return avgpool2(dtype, in, out, batch, H, W, C, S(stream));
}

int b200clip_attnpool_tokens(int dtype, const void* x, const float* pos, void* tok, int batch, int HW, int C, void* stream) {
    return attnpool_tokens(dtype, x, pos, tok, batch, HW, C, S(stream));
}

int b200clip_resize_crop_u8(const uint8_t* src_hwc, int H, int W, int64_t row_stride, const int32_t* h_bounds, const int32_t* h_coeffs,
                            int h_ksize, const int32_t* v_bounds, const int32_t* v_coeffs, int v_ksize, int y0, int rows, uint8_t* tmp,
                            uint8_t* dst_chw, int out_h, int out_w, void* stream) {
    return resize_crop_u8(src_hwc, H, W, row_stride, h_bounds, h_coeffs, h_ksize, v_bounds, v_coeffs, v_ksize, y0, rows, tmp, dst_chw, out_h, out_w,
                          S(stream));
}

int64_t b200clip_train_saved_bytes(const b200clip_tower_cfg* cfg, int batch, int seq_len) { return train_saved_bytes(cfg, batch, seq_len); }
int64_t b200clip_backward_workspace_bytes(const b200clip_tower_cfg* cfg, int batch, int seq_len) {
    return backward_workspace_bytes(cfg, batch, seq_len);
}
int b200clip_vit_forward_train(const b200clip_tower_cfg* cfg, const b200clip_vit_weights* w, const void* image, void* out, int batch,
                               int normalize, void* saved, int64_t saved_bytes, void* workspace, int64_t workspace_bytes, void* stream) {
    return vit_forward_train(cfg, w, image, out, batch, normalize, saved, saved_bytes, workspace, workspace_bytes, S(stream));
}
int b200clip_text_forward_train(const b200clip_tower_cfg* cfg, const b200clip_text_weights* w, const int64_t* text, void* out, int batch,
                                int seq_len, int normalize, void* saved, int64_t saved_bytes, void* workspace, int64_t workspace_bytes,
                                void* stream) {
    return text_forward_train(cfg, w, text, out, batch, seq_len, normalize, saved, saved_bytes, workspace, workspace_bytes, S(stream));
}
int b200clip_vit_backward(const b200clip_tower_cfg* cfg, const b200clip_vit_weights* w, const void* image, const void* d_out, int batch,
                          int normalize, const void* saved, const b200clip_vit_grads* grads, void* workspace, int64_t workspace_bytes,
                          void* stream) {
    return vit_backward(cfg, w, image, d_out, batch, normalize, saved, grads, workspace, workspace_bytes, S(stream));
}
int b200clip_text_backward(const b200clip_tower_cfg* cfg, const b200clip_text_weights* w, const int64_t* text, const void* d_out, int batch,
                           int seq_len, int normalize, const void* saved, const b200clip_text_grads* grads, void* workspace,
                           int64_t workspace_bytes, void* stream) {
    return text_backward(cfg, w, text, d_out, batch, seq_len, normalize, saved, grads, workspace, workspace_bytes, S(stream));
}
int b200clip_adamw_chunk(void) { return 4096; }
int b200clip_multi_cast(const b200clip_cast_tensor* items, const int32_t* chunk_item, const int64_t* chunk_off, int chunks, void* stream) {
    return multi_cast(items, chunk_item, chunk_off, chunks, S(stream));
}
int b200clip_adamw_step(const b200clip_adamw_tensor* items, const int32_t* chunk_item, const int64_t* chunk_off, int chunks, float lr,
                        float beta1, float beta2, float eps, float weight_decay, int step, float grad_scale, void* stream) {
    return adamw_step(items, chunk_item, chunk_off, chunks, lr, beta1, beta2, eps, weight_decay, step, grad_scale, S(stream));
}

int b200clip_patchify_u8(int dtype, const uint8_t* image, const float* mean, const float* std, void* patches, int batch, int image_size,
                         int patch, int kpad, void* stream) {
    return patchify_u8(dtype, image, mean, std, patches, batch, image_size, patch, kpad, nullptr, nullptr, nullptr, 0, S(stream));
}

int b200clip_text_forward(const b200clip_tower_cfg* cfg, const b200clip_text_weights* w, const int64_t* text, void* out, int batch,
                          int seq_len, int normalize, void* workspace, int64_t workspace_bytes, void* stream) {
    return text_forward(cfg, w, text, out, batch, seq_len, normalize, workspace, workspace_bytes, S(stream));
}

}  // extern "C"
