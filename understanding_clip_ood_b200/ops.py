"""Tensor-level wrappers over the C ABI (one function per `b200clip_*` kernel entry point).

PyTorch is only used for device memory and streams here; all arithmetic happens in libb200clip.so.
"""
from __future__ import annotations

import torch

from . import _lib as L


def _c(t: torch.Tensor) -> torch.Tensor:
    return t if t.is_contiguous() else t.contiguous()


def gemm(a: torch.Tensor, w: torch.Tensor, bias: torch.Tensor | None = None, *, epilogue: int = L.EPI_BIAS,
         residual: torch.Tensor | None = None, out: torch.Tensor | None = None, pos: torch.Tensor | None = None,
         g_in: int = 0, g_out: int = 0, block_n: int = 0) -> torch.Tensor:
    """out[M,N] = epilogue(a[M,K] @ w[N,K]^T + bias).  See include/b200clip.h:b200clip_gemm."""
    L.require_cuda(a, w, bias, residual, out, pos)
    a, w = _c(a), _c(w)
    M, K = a.shape
    N, K2 = w.shape
    if K != K2:
        raise L.B200ClipError(f"gemm: inner dimensions differ ({K} vs {K2})")
    if w.dtype != a.dtype:
        raise L.B200ClipError("gemm: a and w must share a dtype")
    if out is None:
        rows = M if epilogue != L.EPI_PATCH else (M // g_in) * g_out
        out = torch.empty((rows, N), dtype=a.dtype, device=a.device)
    lib = L.load()
    args = [L.dtype_code(a.dtype), a.data_ptr(), a.stride(0), w.data_ptr(), w.stride(0), L.ptr(bias), L.ptr(residual),
            residual.stride(0) if residual is not None else 0, out.data_ptr(), out.stride(0), M, N, K, epilogue,
            L.ptr(pos), g_in, g_out]
    if block_n:
        rc = lib.b200clip_gemm_tile(*args, block_n, L.stream_ptr())
    else:
        rc = lib.b200clip_gemm(*args, L.stream_ptr())
    L.check(rc, "b200clip_gemm")
    return out


_gemm_ws: dict = {}


def gemm_workspace(device) -> torch.Tensor:
    """Per-device scratch buffer of the stream-K GEMMs (`b200clip_gemm_workspace_bytes`), allocated once."""
    dev = torch.device(device)
    ws = _gemm_ws.get(dev)
    if ws is None:
        ws = torch.zeros(int(L.load().b200clip_gemm_workspace_bytes()), dtype=torch.uint8, device=dev)
        _gemm_ws[dev] = ws
    return ws


def gemm_ws(a: torch.Tensor, w: torch.Tensor, bias: torch.Tensor | None = None, *, epilogue: int = L.EPI_BIAS,
            residual: torch.Tensor | None = None, out: torch.Tensor | None = None, workspace: torch.Tensor | None = None) -> torch.Tensor:
    """`gemm` with the stream-K workspace (ragged tile grids are cut into equal runs of K-blocks); see b200clip_gemm_ws."""
    L.require_cuda(a, w, bias, residual, out)
    a, w = _c(a), _c(w)
    M, K = a.shape
    N = w.shape[0]
    if out is None:
        out = torch.empty((M, N), dtype=a.dtype, device=a.device)
    ws = workspace if workspace is not None else gemm_workspace(a.device)
    rc = L.load().b200clip_gemm_ws(L.dtype_code(a.dtype), a.data_ptr(), a.stride(0), w.data_ptr(), w.stride(0), L.ptr(bias), L.ptr(residual),
                                   residual.stride(0) if residual is not None else 0, out.data_ptr(), out.stride(0), M, N, K, epilogue,
                                   ws.data_ptr(), ws.numel(), L.stream_ptr())
    L.check(rc, "b200clip_gemm_ws")
    return out


def gemm_mn(a: torch.Tensor, w: torch.Tensor, *, a_transposed: bool = False, out: torch.Tensor | None = None,
            workspace: torch.Tensor | None = None, stream_k: bool = True) -> torch.Tensor:
    """Backward GEMMs without transposed copies: `a @ w` (dgrad: a [M, K], w [K, N]) or `a.t() @ w` (wgrad, a_transposed: a [K, M],
    w [K, N]); see b200clip_gemm_mn."""
    L.require_cuda(a, w, out)
    a, w = _c(a), _c(w)
    K, N = w.shape
    M = a.shape[1] if a_transposed else a.shape[0]
    if (a.shape[0] if a_transposed else a.shape[1]) != K:
        raise L.B200ClipError(f"gemm_mn: contraction extents differ ({tuple(a.shape)} vs {tuple(w.shape)})")
    if out is None:
        out = torch.empty((M, N), dtype=a.dtype, device=a.device)
    ws = (workspace if workspace is not None else gemm_workspace(a.device)) if stream_k else None
    rc = L.load().b200clip_gemm_mn(L.dtype_code(a.dtype), a.data_ptr(), a.stride(0), 1 if a_transposed else 0, w.data_ptr(), w.stride(0),
                                   out.data_ptr(), out.stride(0), M, N, K, L.ptr(ws), ws.numel() if ws is not None else 0, L.stream_ptr())
    L.check(rc, "b200clip_gemm_mn")
    return out


def patch_embed_implicit(image: torch.Tensor, conv1_w: torch.Tensor, pos_cls: torch.Tensor, patch: int) -> torch.Tensor:
    """ViT patch embedding as an implicit GEMM over a 16-bit NCHW batch (no im2col matrix); see b200clip_patch_embed_implicit.
    conv1_w [width, 3*P*P], pos_cls fp32 [G*G + 1, width] (row 0 = class + positional[0]) -> x [B, G*G + 1, width]."""
    L.require_cuda(image, conv1_w, pos_cls)
    image, conv1_w, pos_cls = _c(image), _c(conv1_w), _c(pos_cls)
    B, _, S, _ = image.shape
    width = conv1_w.shape[0]
    G = S // patch
    x = torch.empty((B, G * G + 1, width), dtype=image.dtype, device=image.device)
    rc = L.load().b200clip_patch_embed_implicit(L.dtype_code(image.dtype), image.data_ptr(), conv1_w.data_ptr(), pos_cls.data_ptr(), x.data_ptr(),
                                               B, S, patch, width, L.stream_ptr())
    L.check(rc, "b200clip_patch_embed_implicit")
    return x


def gemm_ln_ws(x: torch.Tensor, wf: torch.Tensor, colsum: torch.Tensor, bias_f32: torch.Tensor, stats: torch.Tensor, *,
               epilogue: int = L.EPI_BIAS, out: torch.Tensor | None = None, workspace: torch.Tensor | None = None) -> torch.Tensor:
    """`gemm_ln` with the stream-K workspace; see b200clip_gemm_ln_ws."""
    L.require_cuda(x, wf, colsum, bias_f32, stats, out)
    x, wf = _c(x), _c(wf)
    M, K = x.shape
    N = wf.shape[0]
    if out is None:
        out = torch.empty((M, N), dtype=x.dtype, device=x.device)
    ws = workspace if workspace is not None else gemm_workspace(x.device)
    rc = L.load().b200clip_gemm_ln_ws(L.dtype_code(x.dtype), x.data_ptr(), x.stride(0), wf.data_ptr(), wf.stride(0), colsum.data_ptr(),
                                      bias_f32.data_ptr(), stats.data_ptr(), out.data_ptr(), out.stride(0), M, N, K, epilogue,
                                      ws.data_ptr(), ws.numel(), L.stream_ptr())
    L.check(rc, "b200clip_gemm_ln_ws")
    return out


def row_stats(x: torch.Tensor, eps: float = 1e-5) -> torch.Tensor:
    """-> [rows, 2] fp32 (mean, rstd) per row; see b200clip_row_stats."""
    L.require_cuda(x)
    x = _c(x)
    x2 = x.view(-1, x.shape[-1])
    stats = torch.empty((x2.shape[0], 2), dtype=torch.float32, device=x.device)
    rc = L.load().b200clip_row_stats(L.dtype_code(x.dtype), x2.data_ptr(), x2.stride(0), stats.data_ptr(), x2.shape[0], x2.shape[1],
                                     eps, L.stream_ptr())
    L.check(rc, "b200clip_row_stats")
    return stats


def fold_layernorm(weight: torch.Tensor, bias: torch.Tensor | None, gamma: torch.Tensor, beta: torch.Tensor, dtype: torch.dtype):
    """Host-side preparation of the LN-fold operands of `gemm_ln`: -> (W diag(gamma) in `dtype`, row sums fp32, b + W beta fp32)."""
    w32 = weight.detach().float()
    wf = (w32 * gamma.detach().float()[None, :]).to(dtype).contiguous()
    colsum = wf.float().sum(dim=1).contiguous()
    bf = w32 @ beta.detach().float()
    if bias is not None:
        bf = bf + bias.detach().float()
    return wf, colsum, bf.contiguous()


def gemm_ln(x: torch.Tensor, wf: torch.Tensor, colsum: torch.Tensor, bias_f32: torch.Tensor, stats: torch.Tensor, *,
            epilogue: int = L.EPI_BIAS, out: torch.Tensor | None = None) -> torch.Tensor:
    """act(LayerNorm(x) W^T + b) with the LayerNorm folded into the GEMM epilogue; see b200clip_gemm_ln."""
    L.require_cuda(x, wf, colsum, bias_f32, stats, out)
    x, wf = _c(x), _c(wf)
    M, K = x.shape
    N = wf.shape[0]
    if out is None:
        out = torch.empty((M, N), dtype=x.dtype, device=x.device)
    rc = L.load().b200clip_gemm_ln(L.dtype_code(x.dtype), x.data_ptr(), x.stride(0), wf.data_ptr(), wf.stride(0), colsum.data_ptr(),
                                   bias_f32.data_ptr(), stats.data_ptr(), out.data_ptr(), out.stride(0), M, N, K, epilogue,
                                   L.stream_ptr())
    L.check(rc, "b200clip_gemm_ln")
    return out


def gemm_residual_stats(a: torch.Tensor, w: torch.Tensor, bias: torch.Tensor | None, residual: torch.Tensor, *,
                        out: torch.Tensor | None = None):
    """out = round(a W^T + bias) + residual (residual may be `out`) and the LayerNorm partial sums of the stored rows.
    -> (out [M, N], partials [M, slots, 2] fp32); see b200clip_gemm_residual_stats."""
    L.require_cuda(a, w, bias, residual, out)
    a, w = _c(a), _c(w)
    M, K = a.shape
    N = w.shape[0]
    if out is None:
        out = torch.empty((M, N), dtype=a.dtype, device=a.device)
    lib = L.load()
    slots = lib.b200clip_gemm_stats_slots(M, N)
    partials = torch.empty((M, slots, 2), dtype=torch.float32, device=a.device)
    rc = lib.b200clip_gemm_residual_stats(L.dtype_code(a.dtype), a.data_ptr(), a.stride(0), w.data_ptr(), w.stride(0), L.ptr(bias),
                                          residual.data_ptr(), residual.stride(0), out.data_ptr(), out.stride(0), M, N, K,
                                          partials.data_ptr(), L.stream_ptr())
    L.check(rc, "b200clip_gemm_residual_stats")
    return out, partials


def gemm_ln_partials(x: torch.Tensor, wf: torch.Tensor, colsum: torch.Tensor, bias_f32: torch.Tensor, partials: torch.Tensor, *,
                     eps: float = 1e-5, epilogue: int = L.EPI_BIAS, out: torch.Tensor | None = None) -> torch.Tensor:
    """gemm_ln with the row statistics derived from the partial sums a residual GEMM left behind; see b200clip_gemm_ln_partials."""
    L.require_cuda(x, wf, colsum, bias_f32, partials, out)
    x, wf = _c(x), _c(wf)
    M, K = x.shape
    N = wf.shape[0]
    if out is None:
        out = torch.empty((M, N), dtype=x.dtype, device=x.device)
    rc = L.load().b200clip_gemm_ln_partials(L.dtype_code(x.dtype), x.data_ptr(), x.stride(0), wf.data_ptr(), wf.stride(0),
                                            colsum.data_ptr(), bias_f32.data_ptr(), partials.data_ptr(), partials.shape[1], eps,
                                            out.data_ptr(), out.stride(0), M, N, K, epilogue, L.stream_ptr())
    L.check(rc, "b200clip_gemm_ln_partials")
    return out


def layernorm(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, eps: float = 1e-5, *, rows: int | None = None,
              row_stride_rows: int = 1, row_idx: torch.Tensor | None = None, out: torch.Tensor | None = None) -> torch.Tensor:
    L.require_cuda(x, gamma, beta, row_idx, out)
    x = _c(x)
    width = x.shape[-1]
    x2 = x.view(-1, width)
    if rows is None:
        rows = x2.shape[0] // row_stride_rows
    if out is None:
        out = torch.empty((rows, width), dtype=x.dtype, device=x.device)
    rc = L.load().b200clip_layernorm(L.dtype_code(x.dtype), x2.data_ptr(), x2.stride(0), gamma.data_ptr(), beta.data_ptr(),
                                     out.data_ptr(), out.stride(0), rows, width, eps, row_stride_rows, L.ptr(row_idx),
                                     L.stream_ptr())
    L.check(rc, "b200clip_layernorm")
    return out


def attention(qkv: torch.Tensor, batch: int, seq_len: int, heads: int, causal: bool = False,
              out: torch.Tensor | None = None) -> torch.Tensor:
    """qkv [batch*seq_len, 3*heads*64] -> out [batch*seq_len, heads*64]."""
    L.require_cuda(qkv, out)
    qkv = _c(qkv)
    W = heads * 64
    if qkv.shape != (batch * seq_len, 3 * W):
        raise L.B200ClipError(f"attention: qkv shape {tuple(qkv.shape)} != {(batch * seq_len, 3 * W)}")
    if out is None:
        out = torch.empty((batch * seq_len, W), dtype=qkv.dtype, device=qkv.device)
    rc = L.load().b200clip_attention(L.dtype_code(qkv.dtype), qkv.data_ptr(), out.data_ptr(), batch, seq_len, heads,
                                     int(causal), L.stream_ptr())
    L.check(rc, "b200clip_attention")
    return out


def patchify(image: torch.Tensor, patch: int, kpad: int, class_emb: torch.Tensor | None = None,
             pos: torch.Tensor | None = None, x: torch.Tensor | None = None) -> torch.Tensor:
    L.require_cuda(image, class_emb, pos, x)
    image = _c(image)
    B, Cc, S, S2 = image.shape
    if Cc != 3 or S != S2:
        raise L.B200ClipError("patchify: image must be [B,3,S,S]")
    g = S // patch
    patches = torch.empty((B * g * g, kpad), dtype=image.dtype, device=image.device)
    width = x.shape[-1] if x is not None else 0
    rc = L.load().b200clip_patchify(L.dtype_code(image.dtype), image.data_ptr(), patches.data_ptr(), B, S, patch, kpad,
                                    L.ptr(class_emb), L.ptr(pos), L.ptr(x), width, L.stream_ptr())
    L.check(rc, "b200clip_patchify")
    return patches


def patchify_u8(image: torch.Tensor, patch: int, kpad: int, dtype: torch.dtype, mean, std) -> torch.Tensor:
    """uint8 pixels [B,3,S,S] -> normalised patch matrix [B*g*g, kpad] in `dtype`; see b200clip_patchify_u8."""
    L.require_cuda(image)
    if image.dtype != torch.uint8:
        raise L.B200ClipError("patchify_u8: uint8 image required")
    image = _c(image)
    B, Cc, S, S2 = image.shape
    if Cc != 3 or S != S2:
        raise L.B200ClipError("patchify: image must be [B,3,S,S]")
    g = S // patch
    patches = torch.empty((B * g * g, kpad), dtype=dtype, device=image.device)
    import ctypes as C
    m = (C.c_float * 3)(*[float(v) for v in mean])
    sd = (C.c_float * 3)(*[float(v) for v in std])
    rc = L.load().b200clip_patchify_u8(L.dtype_code(dtype), image.data_ptr(), m, sd, patches.data_ptr(), B, S, patch, kpad, L.stream_ptr())
    L.check(rc, "b200clip_patchify_u8")
    return patches


def text_embed(text: torch.Tensor, tok_emb: torch.Tensor, pos_emb: torch.Tensor, dtype: torch.dtype, seq_len: int | None = None):
    L.require_cuda(text, tok_emb, pos_emb)
    text = _c(text)
    T, ctx = text.shape
    Lq = ctx if seq_len is None else seq_len
    width = tok_emb.shape[1]
    x = torch.empty((T * Lq, width), dtype=dtype, device=text.device)
    eot = torch.empty((T,), dtype=torch.int32, device=text.device)
    rc = L.load().b200clip_text_embed(L.dtype_code(dtype), text.data_ptr(), ctx, tok_emb.data_ptr(), pos_emb.data_ptr(),
                                      x.data_ptr(), eot.data_ptr(), T, Lq, width, L.stream_ptr())
    L.check(rc, "b200clip_text_embed")
    return x, eot


def normalize(x: torch.Tensor, eps: float = 1e-12, out: torch.Tensor | None = None) -> torch.Tensor:
    L.require_cuda(x, out)
    x = _c(x)
    x2 = x.view(-1, x.shape[-1])
    if out is None:
        out = torch.empty_like(x2)
    rc = L.load().b200clip_normalize(L.dtype_code(x.dtype), x2.data_ptr(), x2.stride(0), out.data_ptr(), out.stride(0),
                                     x2.shape[0], x2.shape[1], eps, L.stream_ptr())
    L.check(rc, "b200clip_normalize")
    return out.view(x.shape)


def zeroshot(img_feat: torch.Tensor, prompt_feat: torch.Tensor, k: int = 1, *, normalize_img: bool = True,
             want_logits: bool = True, logit_scale: float = 1.0):
    """-> (logits fp32 [B,C] | None, topk_idx int64 [B,k] | None, topk_val fp32 [B,k] | None)."""
    L.require_cuda(img_feat, prompt_feat)
    img_feat, prompt_feat = _c(img_feat), _c(prompt_feat)
    B, D = img_feat.shape
    Cn, D2 = prompt_feat.shape
    if D != D2 or img_feat.dtype != prompt_feat.dtype:
        raise L.B200ClipError("zeroshot: feature dims / dtypes differ")
    dev = img_feat.device
    logits = torch.empty((B, Cn), dtype=torch.float32, device=dev) if want_logits else None
    idx = torch.empty((B, k), dtype=torch.int64, device=dev) if k > 0 else None
    val = torch.empty((B, k), dtype=torch.float32, device=dev) if k > 0 else None
    if img_feat.dtype != torch.float32 and logit_scale == 1.0 and Cn <= 1024 and D % 8 == 0 and (k > 0 or want_logits):
        # 16-bit path: logits on the tensor cores (bf16/fp16 operands, fp32 accumulation, rounded to the storage type — the
        # reference's own bf16 `tensordot` semantics), then the per-row warp top-k.  fp32 keeps the fused FFMA kernel below
        # (index-exact parity gate).
        feat = normalize(img_feat) if normalize_img else img_feat
        ld = (Cn + 7) // 8 * 8
        lg = torch.empty((B, ld), dtype=img_feat.dtype, device=dev)
        lib = L.load()
        rc = lib.b200clip_gemm(L.dtype_code(feat.dtype), feat.data_ptr(), feat.stride(0), prompt_feat.data_ptr(), prompt_feat.stride(0),
                               None, None, 0, lg.data_ptr(), ld, B, Cn, D, L.EPI_BIAS, None, 0, 0, L.stream_ptr())
        L.check(rc, "b200clip_gemm (zero-shot logits)")
        rc = lib.b200clip_topk(L.dtype_code(lg.dtype), lg.data_ptr(), ld, B, Cn, k, L.ptr(idx), L.ptr(val), L.ptr(logits), L.stream_ptr())
        L.check(rc, "b200clip_topk")
        return logits, idx, val
    rc = L.load().b200clip_zeroshot(L.dtype_code(img_feat.dtype), img_feat.data_ptr(), prompt_feat.data_ptr(), L.ptr(logits),
                                    L.ptr(idx), L.ptr(val), B, Cn, D, k, int(normalize_img), logit_scale, L.stream_ptr())
    L.check(rc, "b200clip_zeroshot")
    return logits, idx, val


def class_mean(txt_feat: torch.Tensor, classes: int, templates: int) -> torch.Tensor:
    L.require_cuda(txt_feat)
    txt_feat = _c(txt_feat)
    D = txt_feat.shape[-1]
    out = torch.empty((classes, D), dtype=txt_feat.dtype, device=txt_feat.device)
    rc = L.load().b200clip_class_mean(L.dtype_code(txt_feat.dtype), txt_feat.data_ptr(), out.data_ptr(), classes, templates, D,
                                      L.stream_ptr())
    L.check(rc, "b200clip_class_mean")
    return out


def cliploss_workspace_floats(n: int, N: int) -> int:
    return 2 * n * N + 8 * n + 8


def cliploss_forward(img_loc, txt_loc, all_img, all_txt, logit_scale, rank: int):
    """fp32, contiguous CUDA operands (checked by the caller).  -> (loss 0-dim, workspace) ; see b200clip_cliploss_forward."""
    n, D = img_loc.shape
    N = all_img.shape[0]
    loss = torch.empty((), dtype=torch.float32, device=img_loc.device)
    ws = torch.empty((cliploss_workspace_floats(n, N),), dtype=torch.float32, device=img_loc.device)
    rc = L.load().b200clip_cliploss_forward(img_loc.data_ptr(), txt_loc.data_ptr(), all_img.data_ptr(), all_txt.data_ptr(),
                                            logit_scale.data_ptr(), rank, n, N, D, loss.data_ptr(), ws.data_ptr(), L.stream_ptr())
    L.check(rc, "b200clip_cliploss_forward")
    return loss, ws


def cliploss_backward(img_loc, txt_loc, all_img, all_txt, logit_scale, rank: int, ws, grad_out, needs):
    """-> [d_img_loc, d_txt_loc, d_all_img, d_all_txt, d_scale] (None where `needs[i]` is False)."""
    n, D = img_loc.shape
    N = all_img.shape[0]
    dev = img_loc.device
    shapes = ((n, D), (n, D), (N, D), (N, D), ())
    grads = [torch.empty(sh, dtype=torch.float32, device=dev) if need else None for sh, need in zip(shapes, needs)]
    rc = L.load().b200clip_cliploss_backward(img_loc.data_ptr(), txt_loc.data_ptr(), all_img.data_ptr(), all_txt.data_ptr(),
                                             logit_scale.data_ptr(), rank, n, N, D, L.ptr(grad_out), L.ptr(grads[0]), L.ptr(grads[1]),
                                             L.ptr(grads[2]), L.ptr(grads[3]), L.ptr(grads[4]), ws.data_ptr(), L.stream_ptr())
    L.check(rc, "b200clip_cliploss_backward")
    return grads


def cliploss_single_backward(img, txt, logit_scale, ws, grad_out, needs):
    """world_size == 1 (after cliploss_forward(img, txt, img, txt, scale, 0)): -> [d_img, d_txt, d_scale], the TOTAL gradients
    (None where `needs[i]` is False); see b200clip_cliploss_single_backward."""
    n, D = img.shape
    dev = img.device
    d_img = torch.empty((n, D), dtype=torch.float32, device=dev) if needs[0] else None
    d_txt = torch.empty((n, D), dtype=torch.float32, device=dev) if needs[1] else None
    d_s = torch.empty((), dtype=torch.float32, device=dev) if needs[2] else None
    rc = L.load().b200clip_cliploss_single_backward(img.data_ptr(), txt.data_ptr(), logit_scale.data_ptr(), n, D, L.ptr(grad_out),
                                                    L.ptr(d_img), L.ptr(d_txt), L.ptr(d_s), ws.data_ptr(), L.stream_ptr())
    L.check(rc, "b200clip_cliploss_single_backward")
    return [d_img, d_txt, d_s]


def cliploss_packed_forward(gathered: torch.Tensor, logit_scale: torch.Tensor, rank: int, n: int, ws: torch.Tensor | None = None):
    """gathered [N, 2D] fp32 (img | txt of every rank).  -> (loss 0-dim, workspace); see b200clip_cliploss_packed_forward."""
    N, D2 = gathered.shape
    loss = torch.empty((), dtype=torch.float32, device=gathered.device)
    if ws is None:
        ws = torch.empty((cliploss_workspace_floats(n, N),), dtype=torch.float32, device=gathered.device)
    rc = L.load().b200clip_cliploss_packed_forward(gathered.data_ptr(), logit_scale.data_ptr(), rank, n, N, D2 // 2, loss.data_ptr(),
                                                   ws.data_ptr(), L.stream_ptr())
    L.check(rc, "b200clip_cliploss_packed_forward")
    return loss, ws


def cliploss_packed_backward(gathered: torch.Tensor, logit_scale: torch.Tensor, rank: int, n: int, ws: torch.Tensor,
                             grad_out: torch.Tensor | None, want_scale: bool = True):
    """-> (d_gathered [N, 2D] incl. the local-row terms, d_scale 0-dim | None); see b200clip_cliploss_packed_backward."""
    N, D2 = gathered.shape
    d_g = torch.empty_like(gathered)
    d_s = torch.empty((), dtype=torch.float32, device=gathered.device) if want_scale else None
    rc = L.load().b200clip_cliploss_packed_backward(gathered.data_ptr(), logit_scale.data_ptr(), rank, n, N, D2 // 2, L.ptr(grad_out),
                                                    d_g.data_ptr(), L.ptr(d_s), ws.data_ptr(), L.stream_ptr())
    L.check(rc, "b200clip_cliploss_packed_backward")
    return d_g, d_s


def cliploss_packed_backward_p2p(gathered: torch.Tensor, logit_scale: torch.Tensor, rank: int, n: int, ws: torch.Tensor,
                                 grad_out: torch.Tensor | None, d_slots: torch.Tensor, want_scale: bool = True):
    """Slot-addressed form: the gradient block of rank j's rows is stored to the address in d_slots[j] (int64 device table of
    peer-memory pointers, see open_clip/peer.py).  -> d_scale 0-dim | None; see b200clip_cliploss_packed_backward_p2p."""
    N, D2 = gathered.shape
    d_s = torch.empty((), dtype=torch.float32, device=gathered.device) if want_scale else None
    rc = L.load().b200clip_cliploss_packed_backward_p2p(gathered.data_ptr(), logit_scale.data_ptr(), rank, n, N, D2 // 2, L.ptr(grad_out),
                                                        d_slots.data_ptr(), L.ptr(d_s), ws.data_ptr(), L.stream_ptr())
    L.check(rc, "b200clip_cliploss_packed_backward_p2p")
    return d_s


def cliploss_fwd_bwd(img_loc: torch.Tensor, txt_loc: torch.Tensor, all_img: torch.Tensor, all_txt: torch.Tensor,
                     logit_scale: torch.Tensor, rank: int, *, want_grad: bool = True, grad_out: torch.Tensor | None = None):
    """fp32 features.  -> loss (0-dim), (d_img_loc, d_txt_loc, d_all_img, d_all_txt, d_scale) or None."""
    L.require_cuda(img_loc, txt_loc, all_img, all_txt, logit_scale, grad_out)
    for t in (img_loc, txt_loc, all_img, all_txt, logit_scale):
        if t.dtype != torch.float32:
            raise L.B200ClipError("cliploss: fp32 tensors required")
    img_loc, txt_loc, all_img, all_txt = _c(img_loc), _c(txt_loc), _c(all_img), _c(all_txt)
    n, D = img_loc.shape
    N = all_img.shape[0]
    dev = img_loc.device
    loss = torch.empty((), dtype=torch.float32, device=dev)
    ws = torch.empty((cliploss_workspace_floats(n, N),), dtype=torch.float32, device=dev)
    grads = None
    if want_grad:
        grads = (torch.empty_like(img_loc), torch.empty_like(txt_loc), torch.empty_like(all_img), torch.empty_like(all_txt),
                 torch.empty((), dtype=torch.float32, device=dev))
    g = grads if grads is not None else (None,) * 5
    rc = L.load().b200clip_cliploss(img_loc.data_ptr(), txt_loc.data_ptr(), all_img.data_ptr(), all_txt.data_ptr(),
                                    logit_scale.data_ptr(), rank, n, N, D, loss.data_ptr(), L.ptr(grad_out), L.ptr(g[0]),
                                    L.ptr(g[1]), L.ptr(g[2]), L.ptr(g[3]), L.ptr(g[4]), ws.data_ptr(), L.stream_ptr())
    L.check(rc, "b200clip_cliploss")
    return loss, grads
