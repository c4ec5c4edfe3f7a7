// Whole-tower drivers: one C-ABI call per CLIP.encode_image / CLIP.encode_text.
//
//   vit_forward  = VisionTransformer.forward (deps/open_clip/src/open_clip/transformer.py:601-643)
//   text_forward = CLIP.encode_text          (deps/open_clip/src/open_clip/model.py:269-284)
//   both run the ResidualAttentionBlock stack (transformer.py:253-264, Transformer.forward :350-359)
//
// The driver only enqueues kernels on the caller's stream into a caller-provided workspace: no allocation,
// no synchronisation, so a whole forward is CUDA-graph capturable.  Per layer: LN1 -> QKV GEMM(+bias) ->
// attention -> out-proj GEMM(+bias+residual, in place on x) -> LN2 -> c_fc GEMM(+bias+GELU) -> c_proj
// GEMM(+bias+residual, in place) = 7 launches, residual adds / activations / bias all fused in GEMM epilogues.
// 16-bit modes with folded LayerNorms: both LayerNorms live in the QKV / c_fc GEMM epilogues; their row statistics come from
// row_stats_kernel (default) or from the epilogue of the residual GEMM that wrote those rows (B200CLIP_FUSED_STATS=1).
// Token layout is batch-major [B*L, W] (the reference's LND transpose, transformer.py:351, is an
// implementation detail of nn.MultiheadAttention, not a contract).
#include "../../include/b200clip.h"
#include "common.cuh"
#include "internal.h"

#include <cstdlib>

namespace b200clip {

// B200CLIP_PATCH_IMPLICIT=1: the patch embedding of 16-bit NCHW batches as an implicit GEMM (patches read from the image through a
// 5-D tensor map, b200clip_patch_embed_implicit).  Correct (bit-identical token rows) but measured SLOWER than im2col + GEMM on
// ViT-B/32 at batch 1024: 351 us against 109 + 191 us -- the gather moves 64-byte pixel rows, and every A tile is gathered once per
// N tile (three times at width 768).  Off by default.
static bool patch_embed_via_im2col() {
    static const bool v = [] {
        const char* e = getenv("B200CLIP_PATCH_IMPLICIT");
        return !(e != nullptr && e[0] == '1');
    }();
    return v;
}

int text_embed_v(int dtype, const int64_t* text, int ctx, const float* tok_emb, const float* pos_emb, void* x, int32_t* eot,
                 int T, int L, int width, int vocab, cudaStream_t stream);

namespace {

inline int64_t align_up(int64_t v, int64_t a) { return (v + a - 1) / a * a; }

struct Workspace {
    char* x;
    char* h;
    char* qkv;
    char* mlp;
    char* pooled;
    int32_t* eot;
    float* stats;   // [rows, 2] (mean, rstd) for the LN-fold GEMM of the first layer
    float* part[2]; // [rows, slots, 2] partial (sum x, sum x^2) written by the residual GEMMs (after attention / after the MLP)
    void* sk;       // stream-K partial accumulators + flags of the CTA-pair GEMM (16-bit modes)
    int64_t total;
};

Workspace carve(const b200clip_tower_cfg& c, int batch, int seq_len, void* base) {
    const int64_t es = dtype_size(c.dtype);
    const int64_t rows = static_cast<int64_t>(batch) * seq_len;
    int64_t mlp_cols = c.mlp_width;
    if (c.patch_kpad > mlp_cols) mlp_cols = c.patch_kpad;
    Workspace w;
    int64_t off = 0;
    char* b = static_cast<char*>(base);
    auto take = [&](int64_t bytes) {
        char* p = b ? b + off : nullptr;
        off += align_up(bytes, 256);
        return p;
    };
    w.x = take(rows * c.width * es);
    w.h = take(rows * c.width * es);
    w.qkv = take(rows * 3 * c.width * es);
    w.mlp = take(rows * mlp_cols * es);
    w.pooled = take(static_cast<int64_t>(batch) * c.width * es);
    w.eot = reinterpret_cast<int32_t*>(take(static_cast<int64_t>(batch) * 4));
    w.stats = reinterpret_cast<float*>(take(rows * 2 * 4));
    const int64_t max_slots = c.width / 64 + 2;   // >= 2 * ceil(width / 128), the narrowest N tile
    w.part[0] = reinterpret_cast<float*>(take(rows * max_slots * 8));
    w.part[1] = reinterpret_cast<float*>(take(rows * max_slots * 8));
    w.sk = c.dtype != B200CLIP_F32 ? take(gemm_pair_sk_workspace_bytes()) : nullptr;
    w.total = off;
    return w;
}

int check_cfg(const b200clip_tower_cfg* c) {
    B2C_CHECK_ARG(c != nullptr, "tower: null cfg");
    B2C_CHECK_ARG(c->dtype >= 0 && c->dtype <= 2, "tower: unknown dtype %d", c->dtype);
    B2C_CHECK_ARG(c->width > 0 && c->width % 64 == 0 && c->heads * 64 == c->width, "tower: width %d must be heads*64 (heads=%d)",
                  c->width, c->heads);
    B2C_CHECK_ARG(c->layers > 0 && c->mlp_width > 0 && c->mlp_width % 8 == 0 && c->embed_dim > 0 && c->embed_dim % 8 == 0,
                  "tower: bad layers/mlp_width/embed_dim");
    B2C_CHECK_ARG(c->seq_len > 0, "tower: bad seq_len");
    return 0;
}

// Training forward: the input of every block (slot 1 + l) and the final residual stream (slot layers + 1) are copied into
// `saved` (slot 0 = the vision tower's token matrix before ln_pre); the backward recomputes everything else per block, the
// way --grad-checkpointing does in the reference (transformer.py:353-355).
inline int64_t saved_slot_bytes(const b200clip_tower_cfg& c, int batch, int L) {
    return align_up(static_cast<int64_t>(batch) * L * c.width * dtype_size(c.dtype), 256);
}
inline int save_slot(void* saved, int slot, const void* x, const b200clip_tower_cfg& c, int batch, int L, cudaStream_t s) {
    if (saved == nullptr) return 0;
    const int64_t bytes = static_cast<int64_t>(batch) * L * c.width * dtype_size(c.dtype);
    B2C_CUDA(cudaMemcpyAsync(static_cast<char*>(saved) + slot * saved_slot_bytes(c, batch, L), x, bytes, cudaMemcpyDeviceToDevice, s));
    return 0;
}

int run_blocks(const b200clip_tower_cfg& c, const b200clip_block_weights* blocks, const Workspace& ws, int batch, int L,
               int causal, cudaStream_t s, void* saved = nullptr) {
    const int dt = c.dtype;
    const int M = batch * L;
    const int W = c.width;
    const int act = c.quick_gelu ? B200CLIP_EPI_QUICKGELU : B200CLIP_EPI_GELU;
    int rc;
    const bool fold = c.fold_ln != 0 && dt != B200CLIP_F32;
    const bool bf = dt == B200CLIP_BF16;
    // B200CLIP_FUSED_STATS=1: take the LayerNorm statistics out of the residual GEMMs' epilogues instead of running the
    // row-statistics kernel in front of every LN-fold GEMM (5 launches per layer instead of 7).  Off by default: the
    // epilogue then has to LOAD the residual (the default in-place form lets the L2 do the add through a TMA reduce-add
    // store), and that extra L2 -> SM traffic costs the out-proj / c_proj GEMMs what the two statistics passes cost
    // (ViT-B/32 batch 1024: 9.65 ms either way; ViT-L/14: 43.2 vs 43.8 ms; ViT-B/16 batch 256: 9.79 vs 9.61 ms).
    static const bool fused_stats = [] {
        const char* e = getenv("B200CLIP_FUSED_STATS");
        return e != nullptr && e[0] == '1';
    }();
    const int slots = fold ? gemm_pair_stats_slots(M, W) : 0;
    B2C_CHECK_ARG(slots <= W / 64 + 2, "tower: statistics slot count %d exceeds the workspace carve-up", slots);
    for (int l = 0; l < c.layers; ++l) {
        const b200clip_block_weights& bw = blocks[l];
        if ((rc = save_slot(saved, 1 + l, ws.x, c, batch, L, s)) != 0) return rc;
        if (fold) {
            // LN-fold: the GEMM reads x itself; per-row statistics come from row_stats (first layer) or from the epilogue of the
            // c_proj GEMM of the previous layer
            B2C_CHECK_ARG(bw.in_proj_wf && bw.in_proj_c && bw.in_proj_bf && bw.fc_wf && bw.fc_c && bw.fc_bf,
                          "tower: cfg.fold_ln is set but layer %d has no folded weights", l);
            const bool have = fused_stats && l > 0;
            if (!have && (rc = row_stats(dt, ws.x, W, ws.stats, M, W, 1e-5f, s)) != 0) return rc;
            if ((rc = gemm_pair(bf, ws.x, W, bw.in_proj_wf, W, bw.in_proj_bf, nullptr, 0, ws.qkv, 3 * W, M, 3 * W, W, B200CLIP_EPI_BIAS, 0,
                                0, s, bw.in_proj_c, have ? nullptr : ws.stats, nullptr, 0, nullptr, have ? ws.part[1] : nullptr, slots,
                                1e-5f, ws.sk)) != 0)
                return rc;
        } else {
            if ((rc = layernorm(dt, ws.x, W, bw.ln1_g, bw.ln1_b, ws.h, W, M, W, 1e-5f, 1, nullptr, s)) != 0) return rc;
            if ((rc = gemm_any(dt, ws.h, W, bw.in_proj_w, W, bw.in_proj_b, nullptr, 0, ws.qkv, 3 * W, M, 3 * W, W, B200CLIP_EPI_BIAS,
                               nullptr, 0, 0, s, ws.sk)) != 0) return rc;
        }
        if ((rc = attention(dt, ws.qkv, ws.h, batch, L, c.heads, causal, s)) != 0) return rc;
        if (fold && fused_stats) {
            if ((rc = gemm_pair(bf, ws.h, W, bw.out_proj_w, W, bw.out_proj_b, ws.x, W, ws.x, W, M, W, W, B200CLIP_EPI_RESIDUAL, 0, 0, s,
                                nullptr, nullptr, nullptr, 0, ws.part[0])) != 0)
                return rc;
        } else if ((rc = gemm_any(dt, ws.h, W, bw.out_proj_w, W, bw.out_proj_b, ws.x, W, ws.x, W, M, W, W, B200CLIP_EPI_RESIDUAL, nullptr,
                                  0, 0, s, ws.sk)) != 0) {
            return rc;
        }
        if (fold) {
            if (!fused_stats && (rc = row_stats(dt, ws.x, W, ws.stats, M, W, 1e-5f, s)) != 0) return rc;
            if ((rc = gemm_pair(bf, ws.x, W, bw.fc_wf, W, bw.fc_bf, nullptr, 0, ws.mlp, c.mlp_width, M, c.mlp_width, W, act, 0, 0, s, bw.fc_c,
                                fused_stats ? nullptr : ws.stats, nullptr, 0, nullptr, fused_stats ? ws.part[0] : nullptr, slots,
                                1e-5f, ws.sk)) != 0)
                return rc;
        } else {
            if ((rc = layernorm(dt, ws.x, W, bw.ln2_g, bw.ln2_b, ws.h, W, M, W, 1e-5f, 1, nullptr, s)) != 0) return rc;
            if ((rc = gemm_any(dt, ws.h, W, bw.fc_w, W, bw.fc_b, nullptr, 0, ws.mlp, c.mlp_width, M, c.mlp_width, W, act, nullptr, 0, 0,
                               s, ws.sk)) != 0) return rc;
        }
        if (fold && fused_stats && l + 1 < c.layers) {
            if ((rc = gemm_pair(bf, ws.mlp, c.mlp_width, bw.proj_w, c.mlp_width, bw.proj_b, ws.x, W, ws.x, W, M, W, c.mlp_width,
                                B200CLIP_EPI_RESIDUAL, 0, 0, s, nullptr, nullptr, nullptr, 0, ws.part[1])) != 0)
                return rc;
        } else if ((rc = gemm_any(dt, ws.mlp, c.mlp_width, bw.proj_w, c.mlp_width, bw.proj_b, ws.x, W, ws.x, W, M, W, c.mlp_width,
                                  B200CLIP_EPI_RESIDUAL, nullptr, 0, 0, s, ws.sk)) != 0) {
            return rc;
        }
    }
    return save_slot(saved, 1 + c.layers, ws.x, c, batch, L, s);
}

}  // namespace

int64_t workspace_bytes(const b200clip_tower_cfg* cfg, int batch, int seq_len) {
    if (cfg == nullptr || batch <= 0 || seq_len <= 0) return -1;
    return carve(*cfg, batch, seq_len, nullptr).total;
}

// image_u8 != nullptr: uint8 pixels, ToTensor + Normalize(mean, std) fused into the im2col (`image` is ignored)
// `stages`: B200CLIP_STAGE_INPUT (the patch embedding: the only kernels that read `image`) | _BODY (patch GEMM ... ln_post on the pooled rows,
// workspace -> workspace) | _OUTPUT (projection + optional normalise: the only kernels that write `out`).
static int vit_forward_impl(const b200clip_tower_cfg* cfg, const b200clip_vit_weights* w, const void* image, const uint8_t* image_u8,
                            const float* mean, const float* std, void* out, int batch, int normalize, void* workspace,
                            int64_t workspace_bytes_, int stages, cudaStream_t s, void* saved = nullptr) {
    int rc;
    if ((rc = check_cfg(cfg)) != 0) return rc;
    const b200clip_tower_cfg& c = *cfg;
    B2C_CHECK_ARG(stages > 0 && stages <= 7, "vit_forward: bad stage mask %d", stages);
    B2C_CHECK_ARG(w != nullptr && workspace != nullptr && w->blocks_host != nullptr, "vit_forward: null pointer");
    B2C_CHECK_ARG(!(stages & B200CLIP_STAGE_INPUT) || image != nullptr || image_u8 != nullptr, "vit_forward: null image");
    B2C_CHECK_ARG(!(stages & B200CLIP_STAGE_OUTPUT) || out != nullptr, "vit_forward: null output");
    B2C_CHECK_ARG(batch > 0, "vit_forward: empty batch");
    B2C_CHECK_ARG(c.patch_size > 0 && c.image_size % c.patch_size == 0, "vit_forward: image %d not divisible by patch %d",
                  c.image_size, c.patch_size);
    const int g = c.image_size / c.patch_size;
    const int L = g * g + 1;
    B2C_CHECK_ARG(L == c.seq_len, "vit_forward: seq_len %d != (image/patch)^2+1 = %d", c.seq_len, L);
    B2C_CHECK_ARG(c.patch_kpad >= 3 * c.patch_size * c.patch_size && c.patch_kpad % 8 == 0, "vit_forward: bad patch_kpad %d",
                  c.patch_kpad);
    B2C_CHECK_ARG(reinterpret_cast<uintptr_t>(workspace) % 256 == 0, "vit_forward: workspace must be 256-byte aligned");
    const Workspace ws = carve(c, batch, L, workspace);
    B2C_CHECK_ARG(ws.total <= workspace_bytes_, "vit_forward: workspace too small (%lld < %lld bytes)", (long long)workspace_bytes_,
                  (long long)ws.total);
    const int dt = c.dtype;
    const int W = c.width;
    const int M = batch * L;

    const bool token_layout = dt != B200CLIP_F32 && w->pos_cls != nullptr;
    if (stages & B200CLIP_STAGE_INPUT) {
        // Everything that depends on the caller's batch: the patch embedding up to the token rows x [B*L, W] (+ class / positional rows).
        // 16-bit NCHW batches with patch 16 / 32: implicit GEMM, the patches are read from the image itself through a 5-D tensor map
        // (no im2col matrix); otherwise im2col (from uint8 pixels: ToTensor + Normalize fused) + GEMM.
        rc = 1;
        if (token_layout && image_u8 == nullptr && c.patch_kpad == 3 * c.patch_size * c.patch_size && !patch_embed_via_im2col())
            rc = gemm_pair_patch_embed(dt == B200CLIP_BF16, image, w->conv1_w, w->pos_cls, ws.x, batch, c.image_size, c.patch_size, W, s);
        if (rc == 1) {
            // token layout: im2col with an all-zero row in every class-token slot; otherwise the im2col also writes the class-token
            // rows of x (class_emb + pos[0])
            if (token_layout)
                rc = image_u8 != nullptr
                         ? patchify_u8(dt, image_u8, mean, std, ws.mlp, batch, c.image_size, c.patch_size, c.patch_kpad, nullptr, nullptr, nullptr, W, s, 1)
                         : patchify(dt, image, ws.mlp, batch, c.image_size, c.patch_size, c.patch_kpad, nullptr, nullptr, nullptr, W, s, 1);
            else
                rc = image_u8 != nullptr
                         ? patchify_u8(dt, image_u8, mean, std, ws.mlp, batch, c.image_size, c.patch_size, c.patch_kpad, w->class_emb, w->pos_emb, ws.x, W, s)
                         : patchify(dt, image, ws.mlp, batch, c.image_size, c.patch_size, c.patch_kpad, w->class_emb, w->pos_emb, ws.x, W, s);
            if (rc != 0) return rc;
            if (token_layout) {
                // ONE CTA-pair GEMM over all B*L rows whose epilogue adds the (class + positional) table row of each token
                if (ws.sk != nullptr && (rc = gemm_pair_sk_workspace_reset(ws.sk, s)) != 0) return rc;
                rc = gemm_pair(dt == B200CLIP_BF16, ws.mlp, c.patch_kpad, w->conv1_w, c.patch_kpad, nullptr, nullptr, 0, ws.x, W, M, W,
                               c.patch_kpad, B200CLIP_EPI_BIAS, 0, 0, s, nullptr, nullptr, w->pos_cls, L, nullptr, nullptr, 0, 1e-5f, ws.sk);
            } else {
                // GEMM whose epilogue scatters to token rows 1..L-1 and adds the positional embedding
                rc = gemm_any(dt, ws.mlp, c.patch_kpad, w->conv1_w, c.patch_kpad, nullptr, nullptr, 0, ws.x, W, batch * g * g, W, c.patch_kpad,
                              B200CLIP_EPI_PATCH, w->pos_emb, g * g, L, s);
            }
        }
        if (rc != 0) return rc;
    }
    if (stages & B200CLIP_STAGE_BODY) {
        // stream-K flag words start at zero (the kernels leave them at zero; this also heals a workspace a failed launch left dirty)
        if (ws.sk != nullptr && (rc = gemm_pair_sk_workspace_reset(ws.sk, s)) != 0) return rc;
        if ((rc = save_slot(saved, 0, ws.x, c, batch, L, s)) != 0) return rc;
        if ((rc = layernorm(dt, ws.x, W, w->ln_pre_g, w->ln_pre_b, ws.x, W, M, W, 1e-5f, 1, nullptr, s)) != 0) return rc;
        if ((rc = run_blocks(c, w->blocks_host, ws, batch, L, 0, s, saved)) != 0) return rc;
        // pool_type 'tok': ln_post on the class token only (LN is per-row, so pooling first is exact), then @ proj
        if ((rc = layernorm(dt, ws.x, W, w->ln_post_g, w->ln_post_b, ws.pooled, W, batch, W, 1e-5f, L, nullptr, s)) != 0) return rc;
    }
    if (stages & B200CLIP_STAGE_OUTPUT) {
        if ((rc = gemm_any(dt, ws.pooled, W, w->proj_t, W, nullptr, nullptr, 0, out, c.embed_dim, batch, c.embed_dim, W,
                           B200CLIP_EPI_BIAS, nullptr, 0, 0, s)) != 0)
            return rc;
        if (normalize && (rc = normalize_rows(dt, out, c.embed_dim, out, c.embed_dim, batch, c.embed_dim, 1e-12f, s)) != 0) return rc;
    }
    return 0;
}

int vit_forward(const b200clip_tower_cfg* cfg, const b200clip_vit_weights* w, const void* image, void* out, int batch,
                int normalize, void* workspace, int64_t workspace_bytes_, cudaStream_t s) {
    B2C_CHECK_ARG(image != nullptr && out != nullptr, "vit_forward: null image / output");
    return vit_forward_impl(cfg, w, image, nullptr, nullptr, nullptr, out, batch, normalize, workspace, workspace_bytes_, 7, s);
}

int vit_forward_stages(const b200clip_tower_cfg* cfg, const b200clip_vit_weights* w, const void* image, const uint8_t* image_u8,
                       const float* mean, const float* std, void* out, int batch, int normalize, void* workspace,
                       int64_t workspace_bytes_, int stages, cudaStream_t s) {
    B2C_CHECK_ARG(image_u8 == nullptr || (mean != nullptr && std != nullptr), "vit_forward_stages: uint8 input needs mean / std");
    return vit_forward_impl(cfg, w, image_u8 != nullptr ? nullptr : image, image_u8, mean, std, out, batch, normalize, workspace,
                            workspace_bytes_, stages, s);
}

int vit_forward_u8(const b200clip_tower_cfg* cfg, const b200clip_vit_weights* w, const uint8_t* image, const float* mean,
                   const float* std, void* out, int batch, int normalize, void* workspace, int64_t workspace_bytes_, cudaStream_t s) {
    B2C_CHECK_ARG(image != nullptr && mean != nullptr && std != nullptr && out != nullptr, "vit_forward_u8: null pointer");
    return vit_forward_impl(cfg, w, nullptr, image, mean, std, out, batch, normalize, workspace, workspace_bytes_, 7, s);
}

static int text_forward_impl(const b200clip_tower_cfg* cfg, const b200clip_text_weights* w, const int64_t* text, void* out, int batch,
                             int seq_len, int normalize, void* workspace, int64_t workspace_bytes_, int stages, cudaStream_t s, void* saved) {
    int rc;
    if ((rc = check_cfg(cfg)) != 0) return rc;
    const b200clip_tower_cfg& c = *cfg;
    B2C_CHECK_ARG(stages > 0 && stages <= 7, "text_forward: bad stage mask %d", stages);
    B2C_CHECK_ARG(w != nullptr && workspace != nullptr && w->blocks_host != nullptr, "text_forward: null pointer");
    B2C_CHECK_ARG(!(stages & B200CLIP_STAGE_INPUT) || text != nullptr, "text_forward: null token ids");
    B2C_CHECK_ARG(!(stages & B200CLIP_STAGE_OUTPUT) || out != nullptr, "text_forward: null output");
    B2C_CHECK_ARG(batch > 0, "text_forward: empty batch");
    B2C_CHECK_ARG(seq_len > 0 && seq_len <= c.seq_len, "text_forward: seq_len %d outside (0, %d]", seq_len, c.seq_len);
    B2C_CHECK_ARG(reinterpret_cast<uintptr_t>(workspace) % 256 == 0, "text_forward: workspace must be 256-byte aligned");
    const int L = seq_len;
    const Workspace ws = carve(c, batch, L, workspace);
    B2C_CHECK_ARG(ws.total <= workspace_bytes_, "text_forward: workspace too small (%lld < %lld bytes)",
                  (long long)workspace_bytes_, (long long)ws.total);
    const int dt = c.dtype;
    const int W = c.width;
    if ((stages & B200CLIP_STAGE_INPUT) &&
        (rc = text_embed_v(dt, text, c.seq_len, w->tok_emb, w->pos_emb, ws.x, ws.eot, batch, L, W,
                           c.vocab_size > 0 ? c.vocab_size : 0x7fffffff, s)) != 0)
        return rc;
    if (stages & B200CLIP_STAGE_BODY) {
        if (ws.sk != nullptr && (rc = gemm_pair_sk_workspace_reset(ws.sk, s)) != 0) return rc;
        if ((rc = run_blocks(c, w->blocks_host, ws, batch, L, 1, s, saved)) != 0) return rc;
        // ln_final only on the pooled (EOT) rows: LN is per-row, so this equals pooling after ln_final
        if ((rc = layernorm(dt, ws.x, W, w->ln_final_g, w->ln_final_b, ws.pooled, W, batch, W, 1e-5f, L, ws.eot, s)) != 0) return rc;
    }
    if (stages & B200CLIP_STAGE_OUTPUT) {
        if ((rc = gemm_any(dt, ws.pooled, W, w->proj_t, W, nullptr, nullptr, 0, out, c.embed_dim, batch, c.embed_dim, W,
                           B200CLIP_EPI_BIAS, nullptr, 0, 0, s)) != 0)
            return rc;
        if (normalize && (rc = normalize_rows(dt, out, c.embed_dim, out, c.embed_dim, batch, c.embed_dim, 1e-12f, s)) != 0) return rc;
    }
    return 0;
}

int text_forward_stages(const b200clip_tower_cfg* cfg, const b200clip_text_weights* w, const int64_t* text, void* out, int batch,
                        int seq_len, int normalize, void* workspace, int64_t workspace_bytes_, int stages, cudaStream_t s) {
    return text_forward_impl(cfg, w, text, out, batch, seq_len, normalize, workspace, workspace_bytes_, stages, s, nullptr);
}

// ---- training forwards: the same kernels, plus the copies of the residual stream the backward starts from -----------------
int64_t train_saved_bytes(const b200clip_tower_cfg* cfg, int batch, int seq_len) {
    if (cfg == nullptr || batch <= 0 || seq_len <= 0) return -1;
    return (cfg->layers + 2) * saved_slot_bytes(*cfg, batch, seq_len);
}

int vit_forward_train(const b200clip_tower_cfg* cfg, const b200clip_vit_weights* w, const void* image, void* out, int batch, int normalize,
                      void* saved, int64_t saved_bytes, void* workspace, int64_t workspace_bytes_, cudaStream_t s) {
    B2C_CHECK_ARG(cfg && image && out && saved, "vit_forward_train: null pointer");
    B2C_CHECK_ARG(saved_bytes >= train_saved_bytes(cfg, batch, cfg->seq_len) && reinterpret_cast<uintptr_t>(saved) % 256 == 0,
                  "vit_forward_train: activation buffer too small or misaligned");
    return vit_forward_impl(cfg, w, image, nullptr, nullptr, nullptr, out, batch, normalize, workspace, workspace_bytes_, 7, s, saved);
}

int text_forward_train(const b200clip_tower_cfg* cfg, const b200clip_text_weights* w, const int64_t* text, void* out, int batch, int seq_len,
                       int normalize, void* saved, int64_t saved_bytes, void* workspace, int64_t workspace_bytes_, cudaStream_t s) {
    B2C_CHECK_ARG(cfg && text && out && saved, "text_forward_train: null pointer");
    B2C_CHECK_ARG(saved_bytes >= train_saved_bytes(cfg, batch, seq_len) && reinterpret_cast<uintptr_t>(saved) % 256 == 0,
                  "text_forward_train: activation buffer too small or misaligned");
    return text_forward_impl(cfg, w, text, out, batch, seq_len, normalize, workspace, workspace_bytes_, 7, s, saved);
}

int text_forward(const b200clip_tower_cfg* cfg, const b200clip_text_weights* w, const int64_t* text, void* out, int batch,
                 int seq_len, int normalize, void* workspace, int64_t workspace_bytes_, cudaStream_t s) {
    B2C_CHECK_ARG(text != nullptr && out != nullptr, "text_forward: null pointer");
    return text_forward_stages(cfg, w, text, out, batch, seq_len, normalize, workspace, workspace_bytes_, 7, s);
}

}  // namespace b200clip
