// Tower backward (training path): gradients of every parameter of the ViT image tower / the text transformer
// from the gradient of the (optionally L2-normalised) output features.
//
//   reference: autograd through CLIP.encode_image / encode_text (deps/open_clip/src/open_clip/model.py:265-284,
//   transformer.py:253-264,601-658) as run by training/train.py:115-183 with --grad-checkpointing (transformer.py:353-355):
//   only the input of every ResidualAttentionBlock survives the forward, the block is recomputed in the backward.
//
// Per block, from the saved input x_in and the incoming gradient dY (both [M, W], M = batch * L):
//   recompute  h1 = LN1(x_in), qkv = h1 Wqkv^T + b, att = attention(qkv), x_mid = x_in + att Wo^T + bo,
//              h2 = LN2(x_mid), z = h2 Wfc^T + bfc, a = act(z)
//   backward   d_a = dY Wproj            dWproj = dY^T a          dbproj = colsum(dY)
//              d_z = d_a * act'(z)       d_h2 = d_z Wfc           dWfc = d_z^T h2     dbfc = colsum(d_z)
//              d_xmid = dY + LN2'(d_h2)  d_att = d_xmid Wo        dWo = d_xmid^T att  dbo = colsum(d_xmid)
//              d_qkv = attention'(d_att) d_h1 = d_qkv Wqkv        dWqkv = d_qkv^T h1  dbqkv = colsum(d_qkv)
//              d_xin = d_xmid + LN1'(d_h1)
// GEMMs: 16-bit modes run every product on gemm_pair_kernel (tcgen05, C = A B^T with K-major operands): dgrad takes a
// transposed copy of the weight, wgrad transposed copies of both activations (contraction over the M token rows: few output
// tiles with a very long K -> the stream-K schedule spreads them over all SM pairs).  fp32 (parity mode) uses the general FFMA
// GEMM with transpose flags.  Everything is enqueued on the caller's stream into a caller-provided workspace.
#include "../../include/b200clip.h"
#include "common.cuh"
#include "internal.h"

#include <cstdlib>

namespace b200clip {

int eot_argmax(const int64_t* text, int ctx, int32_t* eot, int T, cudaStream_t stream);

namespace {

inline int64_t align_up(int64_t v, int64_t a) { return (v + a - 1) / a * a; }

struct BwdWs {
    char *h1, *qkv, *att, *xmid, *h2, *z, *a;     // recomputed forward of one block
    char *gA, *gB, *gC, *big;                     // gradients: three [M, W], one [M, max(4W, 3W, kpad)]
    char *tA, *tB, *wT;                           // transposed operands (16-bit modes)
    char *pooled, *feat, *dfeat, *dpooled;        // [batch, W] / [batch, D]
    int32_t* eot;
    float* scratch;                               // column-sum / LayerNorm partials
    unsigned int* counters;                       // 256 words, zero between kernels: last-chunk-done counters of the column sums
    void* sk;                                     // stream-K workspace of the CTA-pair GEMM
    int64_t ldt;                                  // leading dimension of the transposed activations (M rounded up to 8)
    int64_t total;
};

BwdWs carve_bwd(const b200clip_tower_cfg& c, int batch, int L, void* base) {
    const int64_t es = dtype_size(c.dtype);
    const int64_t M = static_cast<int64_t>(batch) * L;
    const int64_t W = c.width, H = c.mlp_width;
    int64_t wide = H > 3 * W ? H : 3 * W;
    if (c.patch_kpad > wide) wide = c.patch_kpad;
    BwdWs w;
    int64_t off = 0;
    char* b = static_cast<char*>(base);
    auto take = [&](int64_t bytes) {
        char* p = b ? b + off : nullptr;
        off += align_up(bytes, 256);
        return p;
    };
    w.h1 = take(M * W * es);
    w.qkv = take(M * 3 * W * es);
    w.att = take(M * W * es);
    w.xmid = take(M * W * es);
    w.h2 = take(M * W * es);
    w.z = take(M * wide * es);
    w.a = take(M * wide * es);
    w.gA = take(M * W * es);
    w.gB = take(M * W * es);
    w.gC = take(M * W * es);
    w.big = take(M * wide * es);
    w.ldt = align_up(M, 8);
    const bool lp = c.dtype != B200CLIP_F32;
    w.tA = lp ? take(wide * w.ldt * es) : nullptr;
    w.tB = lp ? take(wide * w.ldt * es) : nullptr;
    w.wT = lp ? take(wide * (W > c.embed_dim ? W : c.embed_dim) * es) : nullptr;
    w.pooled = take(static_cast<int64_t>(batch) * W * es);
    w.feat = take(static_cast<int64_t>(batch) * c.embed_dim * es);
    w.dfeat = take(static_cast<int64_t>(batch) * c.embed_dim * es);
    w.dpooled = take(static_cast<int64_t>(batch) * W * es);
    w.eot = reinterpret_cast<int32_t*>(take(static_cast<int64_t>(batch) * 4));
    int64_t sf = col_sum_scratch_floats(static_cast<int>(M), static_cast<int>(wide));
    const int64_t lf = ln_backward_scratch_floats(static_cast<int>(M), static_cast<int>(W));
    if (lf > sf) sf = lf;
    w.scratch = reinterpret_cast<float*>(take(sf * 4));
    w.counters = reinterpret_cast<unsigned int*>(take(256 * 4));
    w.sk = lp ? take(gemm_pair_sk_workspace_bytes()) : nullptr;
    w.total = off;
    return w;
}

struct Ctx {
    int dt;
    cudaStream_t s;
    const BwdWs* ws;
};

// B200CLIP_BWD_TRANSPOSE=1: the round-2a path (transposed operand copies + the all-K-major GEMM), kept for A/B measurements
bool bwd_transposed_operands() {
    static const bool v = [] {
        const char* e = getenv("B200CLIP_BWD_TRANSPOSE");
        return e != nullptr && e[0] == '1';
    }();
    return v;
}

// dX[M, K] = G[M, N] W[N, K]
int dgrad(const Ctx& c, const void* G, int64_t ldg, const void* Wt, int64_t ldw, void* dX, int64_t ldx, int M, int N, int K) {
    if (c.dt == B200CLIP_F32)
        return gemm_f32_general(static_cast<const float*>(G), ldg, false, static_cast<const float*>(Wt), ldw, true, static_cast<float*>(dX), ldx, M, K, N,
                                false, c.s);
    if (!bwd_transposed_operands())   // W [N, K] as it lies in memory is the MN-major operand of dX = G W: no transposed copy
        return gemm_pair_mn(c.dt == B200CLIP_BF16, G, ldg, false, Wt, ldw, true, dX, ldx, M, K, N, c.s, c.ws->sk);
    int rc;
    if ((rc = transpose16(c.dt, Wt, ldw, c.ws->wT, N, N, K, N, c.s)) != 0) return rc;          // [N, K] -> [K, N]
    return gemm_pair(c.dt == B200CLIP_BF16, G, ldg, c.ws->wT, N, nullptr, nullptr, 0, dX, ldx, M, K, N, B200CLIP_EPI_BIAS, 0, 0, c.s, nullptr, nullptr,
                     nullptr, 0, nullptr, nullptr, 0, 1e-5f, c.ws->sk);
}

// dW[N, K] = G[M, N]^T X[M, K]  (+ optional bias gradient db[N] = column sums of G).  x_act != 0 (16-bit modes only): X is given
// as the pre-activation and the activation (1 = GELU, 2 = QuickGELU) is applied inside the operand transpose.
int wgrad(const Ctx& c, const void* G, int64_t ldg, const void* X, int64_t ldx, void* dW, int64_t ldw, void* db, int M, int N, int K, int x_act = 0) {
    int rc;
    if (db != nullptr && (rc = col_sum(c.dt, G, ldg, M, N, db, 0, 0, c.ws->scratch, c.s, c.ws->counters)) != 0) return rc;
    if (c.dt == B200CLIP_F32)
        return gemm_f32_general(static_cast<const float*>(G), ldg, true, static_cast<const float*>(X), ldx, true, static_cast<float*>(dW), ldw, N, K, M,
                                false, c.s);
    if (!bwd_transposed_operands()) {
        // G [tokens, N] and X [tokens, K] as they lie in memory are the MN-major operands of dW = G^T X (contraction over the token
        // rows): no transposed copies.  An activation that used to ride on X's transpose becomes one element-wise pass.
        const void* Xa = X;
        if (x_act != 0) {
            B2C_CHECK_ARG(ldx == K, "wgrad: the activated operand must be contiguous");
            if ((rc = act_forward(c.dt, X, c.ws->tB, static_cast<int64_t>(M) * K, x_act == 2 ? 1 : 0, c.s)) != 0) return rc;
            Xa = c.ws->tB;
        }
        return gemm_pair_mn(c.dt == B200CLIP_BF16, G, ldg, true, Xa, ldx, true, dW, ldw, N, K, M, c.s, c.ws->sk);
    }
    const int64_t ldt = c.ws->ldt;
    if ((rc = transpose16(c.dt, G, ldg, c.ws->tA, ldt, M, N, static_cast<int>(ldt), c.s)) != 0) return rc;     // [M, N] -> [N, Mpad]
    if ((rc = transpose16(c.dt, X, ldx, c.ws->tB, ldt, M, K, static_cast<int>(ldt), c.s, x_act)) != 0) return rc;   // [M, K] -> [K, Mpad]
    return gemm_pair(c.dt == B200CLIP_BF16, c.ws->tA, ldt, c.ws->tB, ldt, nullptr, nullptr, 0, dW, ldw, N, K, static_cast<int>(ldt), B200CLIP_EPI_BIAS, 0,
                     0, c.s, nullptr, nullptr, nullptr, 0, nullptr, nullptr, 0, 1e-5f, c.ws->sk);
}

inline char* slot_ptr(const void* saved, int slot, const b200clip_tower_cfg& c, int batch, int L) {
    const int64_t bytes = align_up(static_cast<int64_t>(batch) * L * c.width * dtype_size(c.dtype), 256);
    return static_cast<char*>(const_cast<void*>(saved)) + slot * bytes;
}

// one block: ws.gA holds dY on entry and d(x_in) on exit
int block_backward(const b200clip_tower_cfg& c, const b200clip_block_weights& bw, const b200clip_block_grads& bg, const BwdWs& ws, const void* x_in,
                   int batch, int L, int causal, cudaStream_t s) {
    const int dt = c.dtype, W = c.width, H = c.mlp_width, M = batch * L;
    const Ctx cx{dt, s, &ws};
    int rc;
    // ---- recompute the block forward (un-fused: the LayerNorm outputs and the pre-activation are operands of the backward) ----
    if ((rc = layernorm(dt, x_in, W, bw.ln1_g, bw.ln1_b, ws.h1, W, M, W, 1e-5f, 1, nullptr, s)) != 0) return rc;
    if ((rc = gemm_any(dt, ws.h1, W, bw.in_proj_w, W, bw.in_proj_b, nullptr, 0, ws.qkv, 3 * W, M, 3 * W, W, B200CLIP_EPI_BIAS, nullptr, 0, 0, s, ws.sk)) != 0) return rc;
    if ((rc = attention(dt, ws.qkv, ws.att, batch, L, c.heads, causal, s)) != 0) return rc;
    if ((rc = gemm_any(dt, ws.att, W, bw.out_proj_w, W, bw.out_proj_b, x_in, W, ws.xmid, W, M, W, W, B200CLIP_EPI_RESIDUAL, nullptr, 0, 0, s, ws.sk)) != 0) return rc;
    if ((rc = layernorm(dt, ws.xmid, W, bw.ln2_g, bw.ln2_b, ws.h2, W, M, W, 1e-5f, 1, nullptr, s)) != 0) return rc;
    if ((rc = gemm_any(dt, ws.h2, W, bw.fc_w, W, bw.fc_b, nullptr, 0, ws.z, H, M, H, W, B200CLIP_EPI_BIAS, nullptr, 0, 0, s, ws.sk)) != 0) return rc;
    // a = act(z) is an operand of dWproj only: the 16-bit path applies the activation inside that operand's transpose
    const bool lp = dt != B200CLIP_F32;
    if (!lp && (rc = act_forward(dt, ws.z, ws.a, static_cast<int64_t>(M) * H, c.quick_gelu, s)) != 0) return rc;
    // ---- MLP branch ----
    if ((rc = wgrad(cx, ws.gA, W, lp ? ws.z : ws.a, H, bg.proj_w, H, bg.proj_b, M, W, H, lp ? (c.quick_gelu ? 2 : 1) : 0)) != 0) return rc;   // dWproj, dbproj
    if ((rc = dgrad(cx, ws.gA, W, bw.proj_w, H, ws.big, H, M, W, H)) != 0) return rc;                                // d_a [M, H]
    if ((rc = act_backward(dt, ws.big, ws.z, ws.big, static_cast<int64_t>(M) * H, c.quick_gelu, s)) != 0) return rc;  // d_z in place
    if ((rc = wgrad(cx, ws.big, H, ws.h2, W, bg.fc_w, W, bg.fc_b, M, H, W)) != 0) return rc;                          // dWfc [H, W], dbfc
    if ((rc = dgrad(cx, ws.big, H, bw.fc_w, W, ws.gB, W, M, H, W)) != 0) return rc;                                   // d_h2 [M, W]
    if ((rc = ln_backward(dt, ws.gB, W, ws.xmid, W, bw.ln2_g, ws.gA, W, ws.gC, W, bg.ln2_g, bg.ln2_b, M, W, 1e-5f, 1, nullptr, 0, ws.scratch, s)) != 0)
        return rc;                                                                                                     // d_xmid = dY + LN2'
    // ---- attention branch ----
    if ((rc = wgrad(cx, ws.gC, W, ws.att, W, bg.out_proj_w, W, bg.out_proj_b, M, W, W)) != 0) return rc;              // dWo, dbo
    if ((rc = dgrad(cx, ws.gC, W, bw.out_proj_w, W, ws.gB, W, M, W, W)) != 0) return rc;                              // d_att
    if ((rc = attention_backward(dt, ws.qkv, ws.gB, ws.big, batch, L, c.heads, causal, s)) != 0) return rc;           // d_qkv [M, 3W]
    if ((rc = wgrad(cx, ws.big, 3 * W, ws.h1, W, bg.in_proj_w, W, bg.in_proj_b, M, 3 * W, W)) != 0) return rc;        // dWqkv, dbqkv
    if ((rc = dgrad(cx, ws.big, 3 * W, bw.in_proj_w, W, ws.gB, W, M, 3 * W, W)) != 0) return rc;                      // d_h1
    return ln_backward(dt, ws.gB, W, x_in, W, bw.ln1_g, ws.gC, W, ws.gA, W, bg.ln1_g, bg.ln1_b, M, W, 1e-5f, 1, nullptr, 0, ws.scratch, s);   // d_xin
}

int check_common(const b200clip_tower_cfg* cfg, const void* saved, const void* d_out, void* workspace, int batch, const char* what) {
    B2C_CHECK_ARG(cfg && saved && d_out && workspace, "%s: null pointer", what);
    B2C_CHECK_ARG(cfg->dtype >= 0 && cfg->dtype <= 2 && cfg->width % 64 == 0 && cfg->heads * 64 == cfg->width && cfg->layers > 0, "%s: bad cfg", what);
    B2C_CHECK_ARG(batch > 0, "%s: empty batch", what);
    B2C_CHECK_ARG(reinterpret_cast<uintptr_t>(workspace) % 256 == 0 && reinterpret_cast<uintptr_t>(saved) % 256 == 0, "%s: buffers must be 256-byte aligned", what);
    return 0;
}

// head of both towers: features = [normalize](LN(x_final[pooled rows]) proj)  ->  ws.gA = d(x_final) (zero outside the pooled rows)
int head_backward(const b200clip_tower_cfg& c, const BwdWs& ws, const void* x_final, const float* ln_g, const float* ln_b, const void* proj_t,
                  const void* d_out, int batch, int L, int normalize, const int32_t* row_idx, float* d_ln_g, float* d_ln_b, void* d_proj,
                  cudaStream_t s) {
    const int dt = c.dtype, W = c.width, D = c.embed_dim;
    const Ctx cx{dt, s, &ws};
    int rc;
    if ((rc = layernorm(dt, x_final, W, ln_g, ln_b, ws.pooled, W, batch, W, 1e-5f, L, row_idx, s)) != 0) return rc;
    const void* dfeat = d_out;
    if (normalize) {
        if ((rc = gemm_any(dt, ws.pooled, W, proj_t, W, nullptr, nullptr, 0, ws.feat, D, batch, D, W, B200CLIP_EPI_BIAS, nullptr, 0, 0, s)) != 0) return rc;
        if ((rc = normalize_backward(dt, ws.feat, d_out, ws.dfeat, batch, D, 1e-12f, s)) != 0) return rc;
        dfeat = ws.dfeat;
    }
    // features = pooled proj with proj [W, D]:  d_proj[W, D] = pooled^T dfeat,  d_pooled[batch, W] = dfeat proj_t
    if ((rc = wgrad(cx, ws.pooled, W, dfeat, D, d_proj, D, nullptr, batch, W, D)) != 0) return rc;
    if ((rc = dgrad(cx, dfeat, D, proj_t, W, ws.dpooled, W, batch, D, W)) != 0) return rc;
    B2C_CUDA(cudaMemsetAsync(ws.gA, 0, static_cast<size_t>(batch) * L * W * dtype_size(dt), s));
    return ln_backward(dt, ws.dpooled, W, x_final, W, ln_g, nullptr, 0, ws.gA, W, d_ln_g, d_ln_b, batch, W, 1e-5f, L, row_idx, 0, ws.scratch, s);
}

}  // namespace

int64_t backward_workspace_bytes(const b200clip_tower_cfg* cfg, int batch, int seq_len) {
    if (cfg == nullptr || batch <= 0 || seq_len <= 0) return -1;
    return carve_bwd(*cfg, batch, seq_len, nullptr).total;
}

int vit_backward(const b200clip_tower_cfg* cfg, const b200clip_vit_weights* w, const void* image, const void* d_out, int batch, int normalize,
                 const void* saved, const b200clip_vit_grads* g, void* workspace, int64_t workspace_bytes_, cudaStream_t s) {
    int rc;
    if ((rc = check_common(cfg, saved, d_out, workspace, batch, "vit_backward")) != 0) return rc;
    B2C_CHECK_ARG(w && image && g && w->blocks_host && g->blocks_host, "vit_backward: null pointer");
    const b200clip_tower_cfg& c = *cfg;
    const int L = c.seq_len, W = c.width, dt = c.dtype, M = batch * L;
    const BwdWs ws = carve_bwd(c, batch, L, workspace);
    B2C_CHECK_ARG(ws.total <= workspace_bytes_, "vit_backward: workspace too small (%lld < %lld bytes)", (long long)workspace_bytes_, (long long)ws.total);
    if (ws.sk != nullptr && (rc = gemm_pair_sk_workspace_reset(ws.sk, s)) != 0) return rc;
    B2C_CUDA(cudaMemsetAsync(ws.counters, 0, 256 * 4, s));
    const Ctx cx{dt, s, &ws};
    if ((rc = head_backward(c, ws, slot_ptr(saved, c.layers + 1, c, batch, L), w->ln_post_g, w->ln_post_b, w->proj_t, d_out, batch, L, normalize, nullptr,
                            g->ln_post_g, g->ln_post_b, g->proj, s)) != 0)
        return rc;
    for (int l = c.layers - 1; l >= 0; --l)
        if ((rc = block_backward(c, w->blocks_host[l], g->blocks_host[l], ws, slot_ptr(saved, 1 + l, c, batch, L), batch, L, 0, s)) != 0) return rc;
    // ln_pre: gA = d(ln_pre output) -> gB = d(x0), x0 = token matrix [class + pos | patch embedding + pos]
    if ((rc = ln_backward(dt, ws.gA, W, slot_ptr(saved, 0, c, batch, L), W, w->ln_pre_g, nullptr, 0, ws.gB, W, g->ln_pre_g, g->ln_pre_b, M, W, 1e-5f, 1,
                          nullptr, 0, ws.scratch, s)) != 0)
        return rc;
    // positional embedding: sum over the batch per token position; the class embedding enters row 0 only
    if ((rc = period_sum(dt, ws.gB, W, batch, L, W, g->pos_emb, W, 0, s)) != 0) return rc;
    B2C_CUDA(cudaMemcpyAsync(g->class_emb, g->pos_emb, static_cast<size_t>(W) * 4, cudaMemcpyDeviceToDevice, s));
    // conv1 (patch embedding as a GEMM over the im2col): token-layout patches with an all-zero row in every class-token slot
    if ((rc = patchify(dt, image, ws.z, batch, c.image_size, c.patch_size, c.patch_kpad, nullptr, nullptr, nullptr, W, s, 1)) != 0) return rc;
    return wgrad(cx, ws.gB, W, ws.z, c.patch_kpad, g->conv1_w, c.patch_kpad, nullptr, M, W, c.patch_kpad);
}

int text_backward(const b200clip_tower_cfg* cfg, const b200clip_text_weights* w, const int64_t* text, const void* d_out, int batch, int seq_len,
                  int normalize, const void* saved, const b200clip_text_grads* g, void* workspace, int64_t workspace_bytes_, cudaStream_t s) {
    int rc;
    if ((rc = check_common(cfg, saved, d_out, workspace, batch, "text_backward")) != 0) return rc;
    B2C_CHECK_ARG(w && text && g && w->blocks_host && g->blocks_host, "text_backward: null pointer");
    const b200clip_tower_cfg& c = *cfg;
    B2C_CHECK_ARG(seq_len > 0 && seq_len <= c.seq_len, "text_backward: seq_len %d outside (0, %d]", seq_len, c.seq_len);
    const int L = seq_len, W = c.width, dt = c.dtype;
    const BwdWs ws = carve_bwd(c, batch, L, workspace);
    B2C_CHECK_ARG(ws.total <= workspace_bytes_, "text_backward: workspace too small (%lld < %lld bytes)", (long long)workspace_bytes_, (long long)ws.total);
    if (ws.sk != nullptr && (rc = gemm_pair_sk_workspace_reset(ws.sk, s)) != 0) return rc;
    B2C_CUDA(cudaMemsetAsync(ws.counters, 0, 256 * 4, s));
    if ((rc = eot_argmax(text, c.seq_len, ws.eot, batch, s)) != 0) return rc;
    if ((rc = head_backward(c, ws, slot_ptr(saved, c.layers + 1, c, batch, L), w->ln_final_g, w->ln_final_b, w->proj_t, d_out, batch, L, normalize, ws.eot,
                            g->ln_final_g, g->ln_final_b, g->proj, s)) != 0)
        return rc;
    for (int l = c.layers - 1; l >= 0; --l)
        if ((rc = block_backward(c, w->blocks_host[l], g->blocks_host[l], ws, slot_ptr(saved, 1 + l, c, batch, L), batch, L, 1, s)) != 0) return rc;
    // x0[t, l] = token_embedding[text[t, l]] + positional_embedding[l]: positions >= L took no part in the forward
    B2C_CUDA(cudaMemsetAsync(g->pos_emb, 0, static_cast<size_t>(c.seq_len) * W * 4, s));
    if ((rc = period_sum(dt, ws.gA, W, batch, L, W, g->pos_emb, W, 0, s)) != 0) return rc;
    const int vocab = c.vocab_size > 0 ? c.vocab_size : 0x7fffffff;
    B2C_CHECK_ARG(c.vocab_size > 0, "text_backward: cfg.vocab_size is needed for the embedding gradient");
    B2C_CUDA(cudaMemsetAsync(g->tok_emb, 0, static_cast<size_t>(c.vocab_size) * W * 4, s));
    return token_scatter(dt, ws.gA, W, text, c.seq_len, batch, L, W, vocab, g->tok_emb, s);
}

}  // namespace b200clip
