"""Per-kernel numerics on the B200: every C-ABI kernel entry point against a plain PyTorch fp32 restatement
of the same op (these are floating-point kernels; tolerances are written next to each assert).
Tower-level parity against the oracle / the reference's golden vectors lives in test_parity_gpu.py."""
import math
import time

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

from understanding_clip_ood_b200 import _lib as L  # noqa: E402
from understanding_clip_ood_b200 import ops  # noqa: E402

DEV = "cuda"


def _gen(seed=0):
    return torch.Generator(device=DEV).manual_seed(seed)


def _rel(a, b):
    return ((a.float() - b.float()).norm() / b.float().norm().clamp_min(1e-30)).item()


# ------------------------------------------------------------------ GEMM (tcgen05) ------------------
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("M,N,K,epi,bn", [
    (128, 128, 64, L.EPI_BIAS, 128),
    (128, 256, 64, L.EPI_BIAS, 256),
    (6400, 2304, 768, L.EPI_BIAS, 0),        # ViT-B/32 QKV, 128 images
    (6400, 3072, 768, L.EPI_GELU, 0),        # c_fc + GELU
    (6400, 3072, 768, L.EPI_QUICKGELU, 0),   # c_fc + QuickGELU
    (6400, 768, 3072, L.EPI_RESIDUAL, 0),    # c_proj + residual
    (3200, 768, 768, L.EPI_RESIDUAL, 128),   # out_proj + residual
    (6622, 1536, 512, L.EPI_BIAS, 0),        # text tower QKV, 86 prompts (ragged M)
    (77, 512, 512, L.EPI_BIAS, 0),           # single ragged tile
    (1000, 264, 72, L.EPI_BIAS, 128),        # ragged M, N and K tails
    (64, 512, 768, L.EPI_BIAS, 0),           # pooled projection
])
def test_gemm_tc(dtype, M, N, K, epi, bn):
    g = _gen(1)
    a = (torch.randn(M, K, device=DEV, generator=g) * 0.5).to(dtype)
    w = (torch.randn(N, K, device=DEV, generator=g) * 0.05).to(dtype)
    bias = (torch.randn(N, device=DEV, generator=g) * 0.1).to(dtype)
    res = torch.randn(M, N, device=DEV, generator=g).to(dtype) if epi == L.EPI_RESIDUAL else None
    out = ops.gemm(a, w, bias, epilogue=epi, residual=res, block_n=bn)
    lin32 = a.float() @ w.float().t() + bias.float()
    lin = lin32.to(dtype).float()   # the reference rounds the linear output before the residual add ...
    if epi == L.EPI_GELU:           # ... the activations are applied to the fp32 value here (one rounding, gemm_pair.cu B2C_ACT_ROUND)
        ref = F.gelu(lin32)
    elif epi == L.EPI_QUICKGELU:
        ref = lin32 * torch.sigmoid(1.702 * lin32)
    elif epi == L.EPI_RESIDUAL:
        ref = lin + res.float()
    else:
        ref = lin
    ref = ref.to(dtype).float()
    # one ulp of the 16-bit output type on top of fp32-accumulation-order noise
    ulp = 2 ** -8 if dtype == torch.bfloat16 else 2 ** -11
    tol = 2 * ulp * ref.abs().max().item() + 1e-3
    assert (out.float() - ref).abs().max().item() <= tol
    assert _rel(out, ref) < 3e-3


def test_gemm_tc_inplace_residual_and_nobias():
    g = _gen(2)
    M, N, K = 640, 768, 768
    a = (torch.randn(M, K, device=DEV, generator=g) * 0.5).bfloat16()
    w = (torch.randn(N, K, device=DEV, generator=g) * 0.05).bfloat16()
    x = torch.randn(M, N, device=DEV, generator=g).bfloat16()
    ref = ((a.float() @ w.float().t()).bfloat16().float() + x.float()).bfloat16()
    out = ops.gemm(a, w, None, epilogue=L.EPI_RESIDUAL, residual=x, out=x)
    assert out.data_ptr() == x.data_ptr()
    assert _rel(out, ref) < 3e-3


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("M,N,K,epi,inplace", [
    (6400, 768, 3072, L.EPI_RESIDUAL, True),     # c_proj of a 128-image shard: 75 pair tiles on 74 clusters -> stream-K
    (6400, 768, 768, L.EPI_RESIDUAL, True),      # out-proj of the same shard
    (6400, 2304, 768, L.EPI_BIAS, False),        # QKV: 3 whole rounds + 3 tiles
    (6400, 3072, 768, L.EPI_GELU, False),        # c_fc
    (12800, 768, 3072, L.EPI_RESIDUAL, True),    # 256-image shard: 150 tiles
    (6300, 768, 3072, L.EPI_RESIDUAL, True),     # ragged last M tile inside the stream-K region
    (51200, 768, 768, L.EPI_RESIDUAL, True),     # full batch: 600 tiles = 8 rounds + 8
    (2000, 512, 2048, L.EPI_BIAS, False),        # fewer tiles (16) than clusters: every tile is shared by several clusters
    (6400, 768, 3072, L.EPI_RESIDUAL, False),    # separate residual: keeps whole tiles (TMA-loaded residual path)
])
def test_gemm_stream_k(dtype, M, N, K, epi, inplace):
    """b200clip_gemm_ws: ragged tile grids are cut into equal runs of K-blocks (fp32 partials through the workspace, fixed
    summation order).  Same tolerance as the whole-tile kernel, bit-identical when repeated (deterministic), and the flag words
    of the workspace are zero again afterwards."""
    g = _gen(3)
    a = (torch.randn(M, K, device=DEV, generator=g) * 0.5).to(dtype)
    w = (torch.randn(N, K, device=DEV, generator=g) * 0.05).to(dtype)
    bias = (torch.randn(N, device=DEV, generator=g) * 0.1).to(dtype)
    res = torch.randn(M, N, device=DEV, generator=g).to(dtype) if epi == L.EPI_RESIDUAL else None
    lin32 = a.float() @ w.float().t() + bias.float()
    lin = lin32.to(dtype).float()
    ref = {L.EPI_GELU: lambda: F.gelu(lin32), L.EPI_RESIDUAL: lambda: lin + res.float(), L.EPI_BIAS: lambda: lin}[epi]().to(dtype).float()
    ws = torch.zeros(int(L.load().b200clip_gemm_workspace_bytes()), dtype=torch.uint8, device=DEV)
    outs = []
    for _ in range(2):
        x = res.clone() if inplace else None
        outs.append(ops.gemm_ws(a, w, bias, epilogue=epi, residual=x if inplace else res, out=x, workspace=ws))
    torch.cuda.synchronize()
    ulp = 2 ** -8 if dtype == torch.bfloat16 else 2 ** -11
    assert (outs[0].float() - ref).abs().max().item() <= 2 * ulp * ref.abs().max().item() + 1e-3
    assert _rel(outs[0], ref) < 3e-3
    assert torch.equal(outs[0], outs[1])
    # workspace = one fp32 accumulator slot (2 CTAs x 64 float4 columns x 128 rows x 16 B) per SM pair, then the flag words
    off = (torch.cuda.get_device_properties(0).multi_processor_count // 2) * 2 * 64 * 128 * 16
    assert int(ws[off:].view(torch.int32).abs().sum()) == 0


def test_gemm_stream_k_with_folded_layernorm():
    """LN-fold + GELU epilogue on top of a stream-K fix-up (c_fc of a 256-image shard: 600 tiles = 8 rounds + 8)."""
    g = _gen(4)
    M, N, K = 12800, 3072, 768
    x = (torch.randn(M, K, device=DEV, generator=g) * 0.7 + 0.1).bfloat16()
    w = (torch.randn(N, K, device=DEV, generator=g) * 0.04).bfloat16()
    b = (torch.randn(N, device=DEV, generator=g) * 0.1).bfloat16()
    gamma = 1 + 0.1 * torch.randn(K, device=DEV, generator=g)
    beta = 0.1 * torch.randn(K, device=DEV, generator=g)
    wf, colsum, bf = ops.fold_layernorm(w, b, gamma, beta, torch.bfloat16)
    stats = ops.row_stats(x)
    got = ops.gemm_ln_ws(x, wf, colsum, bf, stats, epilogue=L.EPI_GELU)
    want = ops.gemm_ln(x, wf, colsum, bf, stats, epilogue=L.EPI_GELU)            # whole-tile schedule of the same kernel
    ln = F.layer_norm(x.float(), (K,), gamma, beta, 1e-5)
    ref = F.gelu((ln @ w.float().t() + b.float()).bfloat16().float())
    assert _rel(got, ref) < 8e-3 and _rel(got, want) < 2e-3
    assert torch.equal(got, ops.gemm_ln_ws(x, wf, colsum, bf, stats, epilogue=L.EPI_GELU))


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("kind,M,N,K,stream_k", [
    ("dgrad", 6400, 768, 3072, True),      # d(act) = dY Wproj: G [M, 4W] x W [4W... ] -> contraction over the weight's OUT index
    ("dgrad", 6400, 2304, 768, True),
    ("dgrad", 6400, 768, 768, False),
    ("dgrad", 1000, 264, 72, True),        # ragged everything
    ("wgrad", 768, 3072, 6400, True),      # dWproj [W, 4W] = dY^T a: contraction over the 6400 token rows (100 K-blocks, 36 tiles: stream-K)
    ("wgrad", 3072, 768, 6400, True),
    ("wgrad", 2304, 768, 6400, False),     # whole tiles only
    ("wgrad", 1536, 512, 9856, True),      # text tower, 128 prompts x 77
    ("wgrad", 768, 512, 128, True),        # projection: contraction over the batch
    ("wgrad", 264, 1000, 6300, True),      # ragged token count / tile edges
])
def test_gemm_mn_major_operands(dtype, kind, M, N, K, stream_k):
    """b200clip_gemm_mn: the backward GEMMs read their operands as they lie in memory (MN-major tcgen05 operand tiles), no
    transposed copies: dgrad = a @ w with w [K, N]; wgrad = a.t() @ w with a [K, M], w [K, N].  fp32 torch matmul of the same
    16-bit operands is the reference; stream-K runs are bit-identical when repeated."""
    g = _gen(51)
    w = (torch.randn(K, N, device=DEV, generator=g) * 0.05).to(dtype)
    if kind == "dgrad":
        a = (torch.randn(M, K, device=DEV, generator=g) * 0.5).to(dtype)
        ref = a.float() @ w.float()
    else:
        a = (torch.randn(K, M, device=DEV, generator=g) * 0.5).to(dtype)
        ref = a.float().t() @ w.float()
    out = ops.gemm_mn(a, w, a_transposed=kind == "wgrad", stream_k=stream_k)
    ref16 = ref.to(dtype).float()
    ulp = 2 ** -8 if dtype == torch.bfloat16 else 2 ** -11
    assert torch.isfinite(out.float()).all()
    assert (out.float() - ref16).abs().max().item() <= 2 * ulp * ref16.abs().max().item() + 1e-3
    assert _rel(out, ref16) < 3e-3
    assert torch.equal(out, ops.gemm_mn(a, w, a_transposed=kind == "wgrad", stream_k=stream_k))


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("B,S,P,width", [(5, 224, 32, 768), (3, 224, 16, 768), (130, 64, 32, 128), (2, 224, 16, 1024)])
def test_patch_embedding_as_implicit_gemm(dtype, B, S, P, width):
    """b200clip_patch_embed_implicit: the ViT's conv1 (kernel = stride = patch) + class token + positional embedding with the
    patches read from the NCHW batch through a 5-D tensor map (narrow-swizzle K-major sub-tiles), against F.conv2d in fp32 on
    the same 16-bit operands with the reference's rounding points (transformer.py:602-609)."""
    g = _gen(61)
    G = S // P
    image = torch.randn(B, 3, S, S, device=DEV, generator=g).to(dtype)
    w = (torch.randn(width, 3, P, P, device=DEV, generator=g) * 0.02).to(dtype)
    pos_cls = torch.randn(G * G + 1, width, device=DEV, generator=g) * 0.1
    x = ops.patch_embed_implicit(image, w.reshape(width, -1), pos_cls, P)
    conv = F.conv2d(image.float(), w.float(), stride=P).to(dtype).float()               # [B, width, G, G], rounded like the 16-bit conv
    want = torch.cat([torch.zeros(B, 1, width, device=DEV), conv.flatten(2).transpose(1, 2)], dim=1) + pos_cls.to(dtype).float()[None]
    want = want.to(dtype)
    assert x.shape == want.shape and torch.isfinite(x.float()).all()
    assert torch.equal(x[:, 0], want[:, 0])
    ulp = 2 ** -8 if dtype == torch.bfloat16 else 2 ** -11
    assert (x.float() - want.float()).abs().max().item() <= 2 * ulp * want.float().abs().max().item() + 1e-3
    assert _rel(x, want) < 3e-3


def test_gemm_bad_args_raise():
    a = torch.zeros(8, 12, device=DEV, dtype=torch.bfloat16)   # K % 8 != 0
    w = torch.zeros(16, 12, device=DEV, dtype=torch.bfloat16)
    with pytest.raises(L.B200ClipError):
        ops.gemm(a, w)
    with pytest.raises(L.B200ClipError):
        ops.gemm(torch.zeros(8, 16), torch.zeros(16, 16))      # CPU tensors: no fallback


# ------------------------------------------------------------------ GEMM (fp32 parity path) ----------
@pytest.mark.parametrize("M,N,K,epi", [
    (3200, 2304, 768, L.EPI_BIAS), (3200, 3072, 768, L.EPI_GELU), (3200, 768, 3072, L.EPI_RESIDUAL),
    (130, 516, 100, L.EPI_QUICKGELU), (64, 512, 768, L.EPI_BIAS),
])
def test_gemm_f32(M, N, K, epi):
    g = _gen(3)
    a = torch.randn(M, K, device=DEV, generator=g) * 0.5
    w = torch.randn(N, K, device=DEV, generator=g) * 0.05
    bias = torch.randn(N, device=DEV, generator=g) * 0.1
    res = torch.randn(M, N, device=DEV, generator=g) if epi == L.EPI_RESIDUAL else None
    out = ops.gemm(a, w, bias, epilogue=epi, residual=res)
    lin = (a.double() @ w.double().t() + bias.double())
    if epi == L.EPI_GELU:
        ref = F.gelu(lin)
    elif epi == L.EPI_QUICKGELU:
        ref = lin * torch.sigmoid(1.702 * lin)
    elif epi == L.EPI_RESIDUAL:
        ref = lin + res.double()
    else:
        ref = lin
    assert _rel(out, ref) < 2e-6   # fp32 FMA accumulation vs fp64


# ------------------------------------------------------------------ patch embedding ------------------
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
@pytest.mark.parametrize("S,P,W", [(224, 32, 768), (224, 16, 768), (224, 14, 1024)])
def test_patch_embed(dtype, S, P, W):
    g = _gen(4)
    B = 5
    grid = S // P
    Lq = grid * grid + 1
    kreal = 3 * P * P
    kpad = (kreal + 63) // 64 * 64
    img = torch.randn(B, 3, S, S, device=DEV, generator=g).to(dtype)
    conv_w = (torch.randn(W, 3, P, P, device=DEV, generator=g) * 0.02).to(dtype)
    cls = torch.randn(W, device=DEV, generator=g) * 0.05
    pos = torch.randn(Lq, W, device=DEV, generator=g) * 0.05
    wpad = torch.zeros(W, kpad, device=DEV, dtype=dtype)
    wpad[:, :kreal] = conv_w.reshape(W, kreal)
    x = torch.empty(B * Lq, W, device=DEV, dtype=dtype)
    patches = ops.patchify(img, P, kpad, cls, pos, x)
    ref_p = img.view(B, 3, grid, P, grid, P).permute(0, 2, 4, 1, 3, 5).reshape(B * grid * grid, kreal)
    assert torch.equal(patches[:, :kreal], ref_p)
    assert patches[:, kreal:].abs().sum().item() == 0
    ops.gemm(patches, wpad, None, epilogue=L.EPI_PATCH, out=x, pos=pos, g_in=grid * grid, g_out=Lq)
    conv = F.conv2d(img.double(), conv_w.double(), stride=P).to(dtype).float()   # [B,W,g,g]; fp64: cuDNN fp32 conv may use TF32
    tok = conv.reshape(B, W, -1).permute(0, 2, 1)
    full = torch.cat([cls.to(dtype).float().expand(B, 1, W), tok], dim=1) + pos.to(dtype).float()
    ref = full.to(dtype).reshape(B * Lq, W)
    tol = 3e-3 if dtype == torch.bfloat16 else 2e-6
    assert _rel(x, ref) < tol


# ------------------------------------------------------------------ LayerNorm / normalise ------------
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16, torch.float32])
@pytest.mark.parametrize("rows,width", [(6400, 768), (77 * 3, 512), (1000, 1024), (5, 64)])
def test_layernorm(dtype, rows, width):
    g = _gen(5)
    x = (torch.randn(rows, width, device=DEV, generator=g) * 2 + 0.3).to(dtype)
    gamma = torch.randn(width, device=DEV, generator=g)
    beta = torch.randn(width, device=DEV, generator=g)
    out = ops.layernorm(x, gamma, beta)
    ref = F.layer_norm(x.float(), (width,), gamma, beta, 1e-5).to(dtype)
    tol = {torch.bfloat16: 2 ** -8, torch.float16: 2 ** -11, torch.float32: 1e-6}[dtype]
    assert (out.float() - ref.float()).abs().max().item() <= tol * ref.float().abs().max().item() + 1e-6
    # in place
    y = x.clone()
    ops.layernorm(y, gamma, beta, out=y)
    assert torch.equal(y, out)


def test_layernorm_pooling_gather():
    g = _gen(6)
    B, Lq, W = 7, 50, 768
    x = torch.randn(B * Lq, W, device=DEV, generator=g).bfloat16()
    gamma = torch.randn(W, device=DEV, generator=g)
    beta = torch.randn(W, device=DEV, generator=g)
    out = ops.layernorm(x, gamma, beta, rows=B, row_stride_rows=Lq)
    ref = F.layer_norm(x.view(B, Lq, W)[:, 0].float(), (W,), gamma, beta, 1e-5).bfloat16()
    assert _rel(out, ref) < 3e-3
    idx = torch.randint(0, Lq, (B,), device=DEV, generator=g, dtype=torch.int32)
    out = ops.layernorm(x, gamma, beta, rows=B, row_stride_rows=Lq, row_idx=idx)
    ref = F.layer_norm(x.view(B, Lq, W)[torch.arange(B, device=DEV), idx.long()].float(), (W,), gamma, beta, 1e-5).bfloat16()
    assert _rel(out, ref) < 3e-3


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
def test_normalize(dtype):
    g = _gen(7)
    x = (torch.randn(333, 512, device=DEV, generator=g) * 3).to(dtype)
    x[5] = 0  # eps clamp path
    out = ops.normalize(x)
    ref = F.normalize(x.float(), dim=-1).to(dtype)
    assert (out.float() - ref.float()).abs().max().item() <= (2 ** -7 if dtype == torch.bfloat16 else 1e-6)
    assert out[5].abs().sum().item() == 0


# ------------------------------------------------------------------ LN folded into the GEMM -----------
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("M,N,K,epi", [(6400, 2304, 768, L.EPI_BIAS), (6400, 3072, 768, L.EPI_GELU), (1000, 1536, 512, L.EPI_QUICKGELU),
                                       (77, 264, 72, L.EPI_BIAS), (51200, 768, 768, L.EPI_GELU)])
def test_gemm_with_folded_layernorm(dtype, M, N, K, epi):
    """act(LN(x) W^T + b) through row_stats + gemm_ln vs fp32 torch on the same (16-bit) operands."""
    g = _gen(21)
    x = (torch.randn(M, K, device=DEV, generator=g) * 1.7 + 0.3).to(dtype)
    w = (torch.randn(N, K, device=DEV, generator=g) * 0.05).to(dtype)
    b = (torch.randn(N, device=DEV, generator=g) * 0.1).to(dtype)
    gamma = 1.0 + 0.2 * torch.randn(K, device=DEV, generator=g)
    beta = 0.1 * torch.randn(K, device=DEV, generator=g)
    stats = ops.row_stats(x)
    xf = x.float()
    mean = xf.mean(dim=1)
    rstd = torch.rsqrt(xf.var(dim=1, unbiased=False) + 1e-5)
    assert torch.allclose(stats[:, 0], mean, atol=1e-5, rtol=1e-5)
    assert torch.allclose(stats[:, 1], rstd, atol=1e-6, rtol=1e-5)
    wf, colsum, bf = ops.fold_layernorm(w, b, gamma, beta, dtype)
    out = ops.gemm_ln(x, wf, colsum, bf, stats, epilogue=epi)
    ref = torch.nn.functional.layer_norm(xf, (K,), gamma, beta, 1e-5) @ w.float().t() + b.float()
    if epi == L.EPI_GELU:
        ref = torch.nn.functional.gelu(ref)
    elif epi == L.EPI_QUICKGELU:
        ref = ref * torch.sigmoid(1.702 * ref)
    assert torch.isfinite(out.float()).all()
    assert _rel(out, ref) < (1e-2 if dtype == torch.bfloat16 else 3e-3)


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("M,N,K,N2,epi", [(6400, 768, 768, 3072, L.EPI_GELU), (6400, 768, 3072, 2304, L.EPI_BIAS), (1000, 512, 512, 1536, L.EPI_QUICKGELU),
                                          (300, 1024, 256, 512, L.EPI_BIAS), (51200, 768, 768, 768, L.EPI_GELU)])
def test_row_statistics_fused_into_the_residual_gemm(dtype, M, N, K, N2, epi):
    """x += a W^T + b with the LayerNorm partial sums coming out of the same epilogue (b200clip_gemm_residual_stats), then
    act(LN(x) W2^T + b2) from those sums (b200clip_gemm_ln_partials): the stored rows are bit-identical to the plain residual
    GEMM, the statistics match a two-pass fp32 computation on them, and the LN-fold result matches torch."""
    g = _gen(31)
    a = (torch.randn(M, K, device=DEV, generator=g) * 0.8).to(dtype)
    w = (torch.randn(N, K, device=DEV, generator=g) * 0.05).to(dtype)
    b = (torch.randn(N, device=DEV, generator=g) * 0.1).to(dtype)
    x0 = (torch.randn(M, N, device=DEV, generator=g) * 1.5 + 0.4).to(dtype)
    x0[:, 5] += 30.0                                      # an outlier channel, as real CLIP residual streams have
    x_plain = x0.clone()
    ops.gemm(a, w, b, epilogue=L.EPI_RESIDUAL, residual=x_plain, out=x_plain)       # TMA reduce-add route
    x = x0.clone()
    _, partials = ops.gemm_residual_stats(a, w, b, x, out=x)
    assert torch.equal(x, x_plain)
    xf = x.float()
    sums = partials.double().sum(dim=1)
    assert torch.allclose(sums[:, 0], xf.double().sum(dim=1), rtol=1e-5, atol=1e-3)
    assert torch.allclose(sums[:, 1], (xf.double() ** 2).sum(dim=1), rtol=1e-5, atol=1e-3)
    # consumer
    w2 = (torch.randn(N2, N, device=DEV, generator=g) * 0.05).to(dtype)
    b2 = (torch.randn(N2, device=DEV, generator=g) * 0.1).to(dtype)
    gamma = 1.0 + 0.2 * torch.randn(N, device=DEV, generator=g)
    beta = 0.1 * torch.randn(N, device=DEV, generator=g)
    wf, colsum, bf = ops.fold_layernorm(w2, b2, gamma, beta, dtype)
    out = ops.gemm_ln_partials(x, wf, colsum, bf, partials, epilogue=epi)
    out_two_pass = ops.gemm_ln(x, wf, colsum, bf, ops.row_stats(x), epilogue=epi)
    ref = torch.nn.functional.layer_norm(xf, (N,), gamma, beta, 1e-5) @ w2.float().t() + b2.float()
    if epi == L.EPI_GELU:
        ref = torch.nn.functional.gelu(ref)
    elif epi == L.EPI_QUICKGELU:
        ref = ref * torch.sigmoid(1.702 * ref)
    assert torch.isfinite(out.float()).all()
    assert _rel(out, ref) < (1e-2 if dtype == torch.bfloat16 else 3e-3)
    assert _rel(out, out_two_pass.float()) < (4e-3 if dtype == torch.bfloat16 else 1e-3)


# ------------------------------------------------------------------ attention ------------------------
def _attn_ref(qkv, B, Lq, H, causal):
    W = H * 64
    q, k, v = qkv.float().view(B, Lq, 3, H, 64).permute(2, 0, 3, 1, 4)
    s = q @ k.transpose(-1, -2) / 8.0
    if causal:
        s = s + torch.full((Lq, Lq), float("-inf"), device=qkv.device).triu_(1)
    o = torch.softmax(s, dim=-1) @ v
    return o.permute(0, 2, 1, 3).reshape(B * Lq, W)


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16, torch.float32])
@pytest.mark.parametrize("B,Lq,H,causal", [(3, 50, 12, False), (2, 77, 8, True), (2, 197, 12, False), (1, 257, 16, False),
                                           (4, 16, 8, True), (2, 64, 2, True), (2, 130, 2, True),
                                           (700, 50, 12, False), (33, 7, 8, True), (5, 1, 8, False), (300, 33, 8, True),
                                           # tcgen05 kernel (64 < L <= 257): tile-exact, M=128 / M=64 tail tiles, the CUDA-core key of
                                           # L = 257, persistent multi-item CTAs; L > 257 falls back to the generic kernel
                                           (3, 128, 4, False), (2, 256, 2, True), (2, 288, 2, False), (3, 65, 3, True), (2, 288, 1, True),
                                           (40, 197, 12, False), (64, 77, 8, True), (1, 300, 2, False), (20, 257, 16, False),
                                           (3, 257, 2, True), (2, 192, 3, False), (2, 129, 2, True), (2, 193, 1, False), (150, 80, 3, True),
                                           (149, 257, 1, False)])
def test_attention(dtype, B, Lq, H, causal):
    g = _gen(8)
    qkv = (torch.randn(B * Lq, 3 * H * 64, device=DEV, generator=g) * 1.5).to(dtype)
    out = ops.attention(qkv, B, Lq, H, causal)
    ref = _attn_ref(qkv, B, Lq, H, causal)
    tol = {torch.bfloat16: 1e-2, torch.float16: 2e-3, torch.float32: 2e-6}[dtype]
    assert torch.isfinite(out.float()).all()
    assert _rel(out, ref) < tol


# ------------------------------------------------------------------ text embedding -------------------
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
def test_text_embed(dtype):
    g = _gen(9)
    T, ctx, W, V = 9, 77, 512, 49408
    text = torch.zeros(T, ctx, dtype=torch.int64, device=DEV)
    eot_ref = []
    for t in range(T):
        n = 3 + t
        text[t, 0] = 49406
        text[t, 1:n] = torch.randint(1, 40000, (n - 1,), device=DEV, generator=g)
        text[t, n] = 49407
        eot_ref.append(n)
    tok = torch.randn(V, W, device=DEV, generator=g) * 0.02
    pos = torch.randn(ctx, W, device=DEV, generator=g) * 0.01
    for Lq in (ctx, 16):
        x, eot = ops.text_embed(text, tok, pos, dtype, seq_len=Lq)
        ref = (tok[text[:, :Lq]].to(dtype) + pos[:Lq].to(dtype)).reshape(T * Lq, W)
        assert torch.equal(x, ref)
        assert eot.tolist() == eot_ref


# ------------------------------------------------------------------ zero-shot stage ------------------
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("B,C,D,k", [(64, 345, 512, 5), (1024, 345, 512, 1), (13, 1000, 768, 5), (3, 7, 64, 7)])
def test_zeroshot(dtype, B, C, D, k):
    g = _gen(10)
    img = (torch.randn(B, D, device=DEV, generator=g) * 4).to(dtype)
    prm = F.normalize(torch.randn(C, D, device=DEV, generator=g), dim=-1).to(dtype)
    logits, idx, val = ops.zeroshot(img, prm, k)
    n = F.normalize(img.float(), dim=-1).to(dtype).float()
    ref = (n.double() @ prm.double().t()).float()
    if dtype == torch.float32:
        assert (logits - ref).abs().max().item() < 2e-6
    else:
        assert (logits - ref).abs().max().item() < 2 ** -8
    # the indices must be exactly the top-k of the logits this kernel wrote (ties -> lower index)
    order = torch.sort(logits, dim=1, descending=True, stable=True).indices[:, :k]
    assert torch.equal(idx, order)
    assert torch.equal(val, torch.gather(logits, 1, idx))
    assert torch.equal(idx[:, 0], logits.argmax(dim=1))


def test_zeroshot_tie_break_and_prenormalized():
    D, C = 64, 40
    prm = torch.zeros(C, D, device=DEV)
    prm[:, 0] = 1.0                      # every class has the same logit
    img = torch.zeros(4, D, device=DEV)
    img[:, 0] = 2.0
    logits, idx, _ = ops.zeroshot(img, prm, 5)
    assert idx.tolist() == [[0, 1, 2, 3, 4]] * 4
    assert torch.allclose(logits, torch.ones_like(logits))
    logits2, _, _ = ops.zeroshot(img, prm, 1, normalize_img=False)
    assert torch.allclose(logits2, 2 * torch.ones_like(logits2))


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
def test_zeroshot_16bit_route_ties_and_logit_copy(dtype):
    """16-bit route = tcgen05 logits GEMM + per-row warp top-k (b200clip_topk): ties -> lower index, fp32 logit copy exact."""
    D, C = 64, 345
    prm = torch.zeros(C, D, device=DEV, dtype=dtype)
    prm[:, 0] = 1.0
    prm[100:103, 1] = 0.5                # three classes share the (strictly larger) best logit
    img = torch.zeros(9, D, device=DEV, dtype=dtype)
    img[:, 0] = 1.0
    img[:, 1] = 1.0
    logits, idx, val = ops.zeroshot(img, prm, 5, normalize_img=False)
    assert idx.tolist() == [[100, 101, 102, 0, 1]] * 9
    assert torch.equal(val, torch.gather(logits, 1, idx))
    ref = (img.float() @ prm.float().t()).to(dtype).float()
    assert torch.equal(logits, ref)
    only_logits, none_idx, _ = ops.zeroshot(img, prm, 0, normalize_img=False)
    assert none_idx is None and torch.equal(only_logits, ref)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_class_mean(dtype):
    g = _gen(11)
    Cn, T, D = 11, 86, 512
    txt = (torch.randn(Cn * T, D, device=DEV, generator=g) * 3).to(dtype)
    out = ops.class_mean(txt, Cn, T)
    e = F.normalize(txt.float().view(Cn, T, D), dim=-1).to(dtype).float()
    m = e.mean(dim=1).to(dtype).float()
    ref = F.normalize(m, dim=-1).to(dtype)
    assert _rel(out, ref) < (1e-2 if dtype == torch.bfloat16 else 1e-6)


# ------------------------------------------------------------------ ClipLoss -------------------------
@pytest.mark.parametrize("n,world,rank,D", [(256, 1, 0, 512), (128, 4, 2, 512), (256, 8, 7, 512), (12, 2, 1, 64), (6, 1, 0, 64), (7, 3, 1, 30)])
def test_cliploss_fwd_bwd(n, world, rank, D):
    g = _gen(12)
    N = n * world
    all_img = F.normalize(torch.randn(N, D, device=DEV, generator=g), dim=-1)
    all_txt = F.normalize(torch.randn(N, D, device=DEV, generator=g) + 0.5 * all_img, dim=-1)
    img_loc = all_img[rank * n:(rank + 1) * n].clone()
    txt_loc = all_txt[rank * n:(rank + 1) * n].clone()
    scale = torch.tensor(1 / 0.07, device=DEV)

    # torch restatement of loss.py:102-131 with separate leaves for local and gathered operands
    il, tl, ai, at, s = [t.clone().double().requires_grad_(True) for t in (img_loc, txt_loc, all_img, all_txt, scale)]
    labels = torch.arange(n, device=DEV) + n * rank
    ref_loss = (F.cross_entropy(s * il @ at.t(), labels) + F.cross_entropy(s * tl @ ai.t(), labels)) / 2
    ref_loss.backward()

    loss, grads = ops.cliploss_fwd_bwd(img_loc, txt_loc, all_img, all_txt, scale, rank)
    assert abs(loss.item() - ref_loss.item()) / abs(ref_loss.item()) < 1e-5   # north_star: loss within 1e-3 relative
    for got, want in zip(grads, (il.grad, tl.grad, ai.grad, at.grad, s.grad)):
        assert _rel(got, want) < 1e-4
    loss2, none = ops.cliploss_fwd_bwd(img_loc, txt_loc, all_img, all_txt, scale, rank, want_grad=False)
    assert none is None and abs(loss2.item() - loss.item()) < 1e-6
    g_out = torch.tensor(0.25, device=DEV)
    _, grads3 = ops.cliploss_fwd_bwd(img_loc, txt_loc, all_img, all_txt, scale, rank, grad_out=g_out)
    assert _rel(grads3[0], 0.25 * il.grad) < 1e-4


@pytest.mark.parametrize("n,D,dtype", [(256, 512, torch.float32), (6, 64, torch.float32), (7, 30, torch.float32), (130, 512, torch.bfloat16)])
def test_cliploss_single_device_node(n, D, dtype):
    """world_size == 1 through the public ClipLoss: one autograd node whose backward writes the TOTAL feature gradients with one
    two-segment GEMM launch (b200clip_cliploss_single_backward) — against float64 autograd of loss.py:102-131, with an upstream
    gradient other than 1 and with only some inputs requiring a gradient."""
    from understanding_clip_ood_b200 import open_clip
    g = _gen(13)
    img0 = F.normalize(torch.randn(n, D, device=DEV, generator=g), dim=-1).to(dtype)
    txt0 = F.normalize(torch.randn(n, D, device=DEV, generator=g) + 0.5 * img0.float(), dim=-1).to(dtype)
    i64, t64 = img0.double().requires_grad_(True), txt0.double().requires_grad_(True)
    s64 = torch.tensor(1 / 0.07, device=DEV, dtype=torch.float64, requires_grad=True)
    labels = torch.arange(n, device=DEV)
    ref = (F.cross_entropy(s64 * i64 @ t64.t(), labels) + F.cross_entropy(s64 * t64 @ i64.t(), labels)) / 2
    (0.25 * ref).backward()
    img, txt = img0.clone().requires_grad_(True), txt0.clone().requires_grad_(True)
    scale = torch.tensor(1 / 0.07, device=DEV, requires_grad=True)
    loss = open_clip.ClipLoss()(img, txt, scale)
    assert type(loss.grad_fn).__name__ == "_SingleClipLossBackward"
    (0.25 * loss).backward()
    assert abs(loss.item() - ref.item()) / abs(ref.item()) < 1e-5
    tol = 1e-4 if dtype == torch.float32 else 1e-2          # bf16 features: the gradients are returned in bf16
    assert img.grad.dtype == dtype and _rel(img.grad, i64.grad) < tol and _rel(txt.grad, t64.grad) < tol
    assert float(scale.grad) == pytest.approx(float(s64.grad), rel=1e-4)
    # only the logit scale requires a gradient (frozen towers): the feature gradients are not computed
    scale2 = torch.tensor(1 / 0.07, device=DEV, requires_grad=True)
    open_clip.ClipLoss()(img0, txt0, scale2).backward()
    assert float(scale2.grad) == pytest.approx(4 * float(s64.grad), rel=1e-4)


# ------------------------------------------------------------------ uint8 preprocessing fused into the im2col --------
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16, torch.float32])
@pytest.mark.parametrize("S,P,kpad", [(224, 32, 3072), (224, 16, 768), (224, 14, 640), (64, 8, 192)])
def test_patchify_u8_is_totensor_normalize_patchify(dtype, S, P, kpad):
    from oracle import clip_oracle as O
    from understanding_clip_ood_b200.open_clip import OPENAI_DATASET_MEAN as MEAN, OPENAI_DATASET_STD as STD
    g = _gen(41)
    img = torch.randint(0, 256, (5, 3, S, S), device=DEV, generator=g, dtype=torch.uint8)
    got = ops.patchify_u8(img, P, kpad, dtype, MEAN, STD)
    ref_img = O.preprocess_u8(img.cpu(), MEAN, STD).to(dtype).to(DEV)          # fp32 ToTensor + Normalize, one rounding
    want = ops.patchify(ref_img, P, kpad)
    assert torch.equal(got, want)


# ------------------------------------------------------------------ peer-memory exchange kernels (csrc/p2p.cu) on ONE device -
# The multi-process form runs in tests/test_peer_gpu.py (>= 2 GPUs).  Here the "peers" are local buffers and the flags of
# the other ranks are set by hand, which exercises the same kernels, addressing and flag protocol.
def _ptr_table(ptrs):
    return torch.tensor(ptrs, dtype=torch.int64, device=DEV)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16, torch.float16])
@pytest.mark.parametrize("n,D,world,rank", [(256, 512, 2, 1), (36, 64, 3, 0), (128, 512, 8, 5), (4, 4, 1, 0)])
def test_p2p_allgather_kernel(dtype, n, D, world, rank):
    g = _gen(21)
    img = torch.randn(n, D, device=DEV, generator=g).to(dtype)
    txt = torch.randn(n, D, device=DEV, generator=g).to(dtype)
    S = n * 2 * D
    bufs = [torch.full((world * S,), float("nan"), device=DEV) for _ in range(world)]       # every rank's gather buffer
    pads = [torch.zeros(64, dtype=torch.int32, device=DEV) for _ in range(world)]
    epoch = 7
    pads[rank][:world] = epoch            # the other ranks "have already published" into this rank's pad
    pads[rank][rank] = 0                  # ... except this rank itself: set by the kernel
    dst = _ptr_table([bufs[p].data_ptr() + 4 * rank * S for p in range(world)])
    flag = _ptr_table([pads[p].data_ptr() + 4 * rank for p in range(world)])
    # ring-slot busy words (word 48 of every pad): free on the peers, one of them "already runs this epoch"
    if world > 1:
        pads[(rank + 1) % world][48] = epoch
    busy = _ptr_table([pads[p].data_ptr() + 4 * 48 for p in range(world)])
    rc = L.load().b200clip_p2p_allgather(L.dtype_code(dtype), img.data_ptr(), txt.data_ptr(), n, D, dst.data_ptr(), flag.data_ptr(),
                                         pads[rank].data_ptr(), pads[rank].data_ptr() + 128, world, epoch, busy.data_ptr(),
                                         pads[rank].data_ptr() + 4 * 48, 1, L.stream_ptr())
    L.check(rc, "b200clip_p2p_allgather")
    torch.cuda.synchronize()
    want = torch.cat([img.float(), txt.float()], dim=1)
    for p in range(world):
        got = bufs[p].view(world, n, 2 * D)
        assert torch.equal(got[rank], want)                              # bit-exact fp32 conversion, every peer
        assert torch.isnan(got[[q for q in range(world) if q != rank]]).all()   # nothing else touched
        assert int(pads[p][rank]) == epoch                               # flag published on every peer
    assert int(pads[rank][32:48].abs().sum()) == 0                       # block counters left at zero
    assert int(pads[rank][48]) == epoch                                  # hold=1: this rank's slot is marked busy for its backward


def test_p2p_wait_is_bounded_and_reported():
    """A peer that never releases its ring slot / never publishes its flag must not hang or trap the GPU: the wait expires after
    the configured wall time, the registered (pinned) error word is raised and the kernel retires."""
    n, D, world, rank = 8, 16, 2, 0
    lib = L.load()
    err = torch.zeros(1, dtype=torch.int32).pin_memory()
    L.check(lib.b200clip_p2p_configure(0.05, err.data_ptr()), "b200clip_p2p_configure")
    try:
        img = torch.randn(n, D, device=DEV)
        txt = torch.randn(n, D, device=DEV)
        S = n * 2 * D
        bufs = [torch.zeros(world * S, device=DEV) for _ in range(world)]
        pads = [torch.zeros(64, dtype=torch.int32, device=DEV) for _ in range(world)]
        epoch = 9
        dst = _ptr_table([bufs[p].data_ptr() + 4 * rank * S for p in range(world)])
        flag = _ptr_table([pads[p].data_ptr() + 4 * rank for p in range(world)])
        busy = _ptr_table([pads[p].data_ptr() + 4 * 48 for p in range(world)])
        # (a) peer 1 still holds the slot for epoch 5 and never releases it -> code 2
        pads[1][48] = 5
        pads[rank][1] = epoch                  # its flag is there, only the slot is blocked
        t0 = time.perf_counter()
        L.check(lib.b200clip_p2p_allgather(L.F32, img.data_ptr(), txt.data_ptr(), n, D, dst.data_ptr(), flag.data_ptr(), pads[rank].data_ptr(),
                                           pads[rank].data_ptr() + 128, world, epoch, busy.data_ptr(), pads[rank].data_ptr() + 4 * 48, 0,
                                           L.stream_ptr()), "b200clip_p2p_allgather")
        torch.cuda.synchronize()               # no sticky error: the context survives
        assert time.perf_counter() - t0 < 5.0
        assert int(err[0]) == 2
        # (b) slot free, but peer 1's flag never arrives -> code 1
        err.zero_()
        pads[1][48] = 0
        pads[rank][1] = 0
        L.check(lib.b200clip_p2p_allgather(L.F32, img.data_ptr(), txt.data_ptr(), n, D, dst.data_ptr(), flag.data_ptr(), pads[rank].data_ptr(),
                                           pads[rank].data_ptr() + 128, world, epoch + 1, busy.data_ptr(), pads[rank].data_ptr() + 4 * 48, 0,
                                           L.stream_ptr()), "b200clip_p2p_allgather")
        torch.cuda.synchronize()
        assert int(err[0]) == 1
        assert torch.isfinite(torch.ones(1, device=DEV) + 1).all()      # the device still works
    finally:
        L.check(lib.b200clip_p2p_configure(600.0, None), "b200clip_p2p_configure")


@pytest.mark.parametrize("n,D,world,rank", [(256, 512, 2, 1), (36, 64, 3, 0), (128, 512, 8, 5), (64, 512, 1, 0)])
def test_slot_addressed_backward_and_reduce_finish(n, D, world, rank):
    """b200clip_cliploss_packed_backward_p2p + b200clip_p2p_reduce_finish against b200clip_cliploss_packed_backward:
    slot j must hold the gathered-row terms of rank j's block, the two local slots the local-row terms."""
    g = _gen(22)
    N = n * world
    gathered = F.normalize(torch.randn(N, 2 * D, device=DEV, generator=g), dim=-1)
    scale = torch.tensor(1 / 0.07, device=DEV)
    gout = torch.tensor(0.7, device=DEV)
    loss, ws = ops.cliploss_packed_forward(gathered, scale, rank, n)
    d_ref, ds_ref = ops.cliploss_packed_backward(gathered, scale, rank, n, ws.clone(), gout, True)
    S = n * 2 * D
    # rank j's receive buffer: world + 2 slots; this rank writes slot `rank` of each, and its own two local slots
    recv = [torch.zeros((world + 2) * S, device=DEV) for _ in range(world)]
    slots = _ptr_table([recv[j].data_ptr() + 4 * rank * S for j in range(world)] +
                       [recv[rank].data_ptr() + 4 * (world + h) * S for h in range(2)])
    ds = ops.cliploss_packed_backward_p2p(gathered, scale, rank, n, ws, gout, slots, True)
    torch.cuda.synchronize()
    assert float((ds - ds_ref).abs()) <= 1e-6 * max(1.0, float(ds_ref.abs()))
    for j in range(world):
        got = recv[j].view(world + 2, n, 2 * D)[rank]
        if j == rank:
            got = got + recv[rank].view(world + 2, n, 2 * D)[world:].sum(0)
        want = d_ref[j * n:(j + 1) * n]
        assert float((got - want).norm() / want.norm()) < 2e-6, j
    # reduce-finish on this rank's buffer (the other ranks' flags set by hand)
    pad = torch.zeros(64, dtype=torch.int32, device=DEV)
    others = [torch.zeros(64, dtype=torch.int32, device=DEV) for _ in range(world)]
    others[rank] = pad
    epoch = 3
    pad[16:16 + world] = epoch
    pad[16 + rank] = 0
    flag = _ptr_table([others[p].data_ptr() + 4 * (16 + rank) for p in range(world)])
    out = torch.empty(n, 2 * D, device=DEV)
    pad[48] = 11                         # this rank's ring slot was held for the backward that ends here
    rc = L.load().b200clip_p2p_reduce_finish(recv[rank].data_ptr(), out.data_ptr(), S, flag.data_ptr(), pad.data_ptr() + 64, world,
                                             world + 2, epoch, pad.data_ptr() + 4 * 48, 0, L.stream_ptr())
    L.check(rc, "b200clip_p2p_reduce_finish")
    torch.cuda.synchronize()
    assert int(pad[48]) == 0             # ... and is released
    assert torch.allclose(out, recv[rank].view(world + 2, n, 2 * D).sum(0), rtol=1e-6, atol=1e-7)
    assert all(int(others[p][16 + rank]) == epoch for p in range(world))
    # split form: the same sums as two dense [n, D] halves (d_img, d_txt), what _PeerLocalClipLoss hands to autograd
    out2 = torch.empty(2, n, D, device=DEV)
    rc = L.load().b200clip_p2p_reduce_finish(recv[rank].data_ptr(), out2.data_ptr(), S, flag.data_ptr(), pad.data_ptr() + 64, world,
                                             world + 2, epoch, pad.data_ptr() + 4 * 48, 2 * D, L.stream_ptr())
    L.check(rc, "b200clip_p2p_reduce_finish")
    torch.cuda.synchronize()
    assert torch.equal(out2[0], out[:, :D]) and torch.equal(out2[1], out[:, D:])


# ------------------------------------------------------------------ bicubic resize + centre crop (csrc/preprocess.cu) -------
@pytest.mark.parametrize("h,w,S", [(300, 400, 224), (500, 333, 224), (224, 224, 224), (1080, 1920, 224), (100, 150, 224), (768, 512, 336)])
def test_resize_center_crop_is_pillow_bicubic_bit_for_bit(h, w, S):
    """The GPU preprocessing front end against the PIL pipeline of the reference's eval transform (torchvision Resize(S, BICUBIC) +
    CenterCrop(S), transform.py:372-392): integer arithmetic, so byte-identical."""
    import numpy as np
    from PIL import Image
    from torchvision import transforms as T
    from torchvision.transforms import InterpolationMode
    from understanding_clip_ood_b200.open_clip import gpu_transform as G
    rng = np.random.default_rng(h + w)
    img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    ref = np.asarray(T.Compose([T.Resize(S, interpolation=InterpolationMode.BICUBIC), T.CenterCrop(S)])(Image.fromarray(img))).transpose(2, 0, 1)
    got = G.resize_center_crop(torch.from_numpy(img).to(DEV), S)
    assert got.dtype == torch.uint8 and tuple(got.shape) == (3, S, S)
    assert np.array_equal(got.cpu().numpy(), ref)
    tf = G.GpuEvalTransform(S)
    assert torch.equal(tf(Image.fromarray(img)), got)                      # PIL image / numpy array inputs take the same route
    view = torch.from_numpy(np.concatenate([img, img], axis=1)).to(DEV)[:, :w]   # a strided view (row pitch 2 W)
    assert torch.equal(G.resize_center_crop(view, S), got)


# ------------------------------------------------------------------ ModifiedResNet operators (csrc/resnet.cu) ---------
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16, torch.float32])
@pytest.mark.parametrize("M,N,K,epi", [
    (3 * 56 * 56, 64, 256, L.EPI_RELU),              # layer1 conv1 (1x1): narrow N
    (2 * 112 * 112, 32, 288, L.EPI_RELU),            # stem conv2 as im2col GEMM: N = 32, K = 4.5 K-blocks
    (2 * 112 * 112, 32, 32, L.EPI_RELU),             # stem conv1: K = 32 (27 taps padded)
    (3 * 56 * 56, 256, 64, L.EPI_RESIDUAL_RELU),     # layer1 conv3 + identity + ReLU
    (5 * 49, 2048, 512, L.EPI_RESIDUAL_RELU),        # layer4 conv3, ragged M
    (1000, 264, 72, L.EPI_RESIDUAL_RELU),            # ragged M, N and K tails
    (6400, 1024, 2304, L.EPI_RELU),                  # long K (stream-K when the workspace is given)
])
def test_gemm_relu_epilogues(dtype, M, N, K, epi):
    """conv + folded BatchNorm (+ identity) + ReLU as GEMM epilogues (modified_resnet.py:42-56), vs fp32 torch."""
    g = _gen(3)
    a = (torch.randn(M, K, device=DEV, generator=g) * 0.5).to(dtype)
    w = (torch.randn(N, K, device=DEV, generator=g) * 0.05).to(dtype)
    bias = (torch.randn(N, device=DEV, generator=g) * 0.1).to(dtype)
    res = torch.randn(M, N, device=DEV, generator=g).to(dtype) if epi == L.EPI_RESIDUAL_RELU else None
    lin = a.float() @ w.float().t() + bias.float()
    if dtype != torch.float32:
        lin = lin.to(dtype).float()
    ref = (lin + res.float() if res is not None else lin).clamp_min(0).to(dtype).float()
    ulp = {torch.bfloat16: 2 ** -8, torch.float16: 2 ** -11, torch.float32: 2 ** -22}[dtype]
    tol = 2 * ulp * ref.abs().max().item() + (1e-3 if dtype != torch.float32 else 1e-5)
    outs = [ops.gemm(a, w, bias, epilogue=epi, residual=res)]
    if dtype != torch.float32:
        outs.append(ops.gemm_ws(a, w, bias, epilogue=epi, residual=res))
    for out in outs:
        assert (out.float() - ref).abs().max().item() <= tol
        assert (out >= 0).all()


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16, torch.float32])
@pytest.mark.parametrize("B,H,W,C", [(2, 56, 56, 64), (3, 7, 7, 512), (1, 14, 10, 8), (2, 112, 112, 32)])
def test_im2col3x3_is_unfold(dtype, B, H, W, C):
    x = torch.randn(B, H, W, C, device=DEV, generator=_gen(4)).to(dtype)
    out = torch.empty(B * H * W, 9 * C, dtype=dtype, device=DEV)
    L.check(L.load().b200clip_im2col3x3(L.dtype_code(dtype), x.data_ptr(), out.data_ptr(), B, H, W, C, L.stream_ptr()), "im2col3x3")
    cols = F.unfold(x.permute(0, 3, 1, 2).float(), 3, padding=1)                    # [B, C*9, H*W], K order (c, ky, kx)
    ref = cols.reshape(B, C, 9, H * W).permute(0, 3, 2, 1).reshape(B * H * W, 9 * C)  # -> (tap, c)
    assert torch.equal(out.float(), ref)


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
@pytest.mark.parametrize("B,S,kpad", [(2, 224, 32), (1, 64, 32), (3, 32, 64)])
def test_stem_im2col_is_strided_unfold(dtype, B, S, kpad):
    img = torch.randn(B, 3, S, S, device=DEV, generator=_gen(5)).to(dtype)
    Ho = S // 2
    out = torch.full((B * Ho * Ho, kpad), float("nan"), dtype=dtype, device=DEV)
    L.check(L.load().b200clip_stem_im2col(L.dtype_code(dtype), img.data_ptr(), out.data_ptr(), B, S, kpad, L.stream_ptr()), "stem_im2col")
    cols = F.unfold(img.float(), 3, padding=1, stride=2)                            # [B, 3*9, Ho*Ho], K order (c, tap)
    ref = cols.reshape(B, 3, 9, Ho * Ho).permute(0, 3, 2, 1).reshape(B * Ho * Ho, 27)
    assert torch.equal(out[:, :27].float(), ref)
    assert (out[:, 27:] == 0).all()


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16, torch.float32])
@pytest.mark.parametrize("B,H,W,C", [(2, 56, 56, 256), (3, 14, 14, 8), (1, 112, 112, 64)])
def test_avgpool2_nhwc(dtype, B, H, W, C):
    x = torch.randn(B, H, W, C, device=DEV, generator=_gen(6)).to(dtype)
    out = torch.empty(B, H // 2, W // 2, C, dtype=dtype, device=DEV)
    L.check(L.load().b200clip_avgpool2(L.dtype_code(dtype), x.data_ptr(), out.data_ptr(), B, H, W, C, L.stream_ptr()), "avgpool2")
    ref = F.avg_pool2d(x.permute(0, 3, 1, 2).float(), 2).permute(0, 2, 3, 1).to(dtype)
    assert (out.float() - ref.float()).abs().max().item() <= (1e-6 if dtype == torch.float32 else 2 ** -8 * ref.float().abs().max().item())


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16, torch.float32])
@pytest.mark.parametrize("B,HW,C", [(4, 49, 2048), (2, 4, 1024), (3, 81, 2560)])
def test_attnpool_tokens(dtype, B, HW, C):
    g = _gen(7)
    x = torch.randn(B, HW, C, device=DEV, generator=g).to(dtype)
    pos = torch.randn(HW + 1, C, device=DEV, generator=g) * 0.05
    tok = torch.empty(B, HW + 1, C, dtype=dtype, device=DEV)
    L.check(L.load().b200clip_attnpool_tokens(L.dtype_code(dtype), x.data_ptr(), pos.data_ptr(), tok.data_ptr(), B, HW, C, L.stream_ptr()),
            "attnpool_tokens")
    # the reference's ops on `dtype` tensors (modified_resnet.py:70-72): mean, cat, + pos.to(dtype), one rounding each
    ref = torch.cat([x.float().mean(dim=1, keepdim=True).to(dtype), x], dim=1) + pos.to(dtype)
    ulp = {torch.bfloat16: 2 ** -8, torch.float16: 2 ** -11, torch.float32: 2 ** -22}[dtype]
    assert (tok.float() - ref.float()).abs().max().item() <= ulp * ref.float().abs().max().item()
