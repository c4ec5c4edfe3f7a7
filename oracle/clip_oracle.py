"""CPU ORACLE for the CLIP hot path — TEST INFRASTRUCTURE ONLY.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` legs may import
this module, and only as the checker or the reported CPU baseline — never as the shipped path (the product
path is `understanding_clip_ood_b200/` -> libb200clip.so and has no CPU fallback).

It is a plain-op fp32 restatement (matmul / softmax / mean / rsqrt on torch CPU tensors; no nn.Module, no
F.layer_norm / nn.MultiheadAttention / F.cross_entropy) of the reference algorithm.  Where the arithmetic
lives in the third-party dependency PyTorch (pinned torch==2.4.1 in /root/reference/pyproject.toml:21;
2.11.0 in this image) the published semantics of those ops are restated here:

  vit_forward        VisionTransformer.forward       deps/open_clip/src/open_clip/transformer.py:601-643
  text_forward       CLIP.encode_text                deps/open_clip/src/open_clip/model.py:269-284
  _block             ResidualAttentionBlock.forward  transformer.py:253-264 (+ .attention :238-251, torch
                     F.multi_head_attention_forward: packed q|k|v in-projection, heads = contiguous 64-wide
                     column slices, scale 1/sqrt(head_dim), additive mask, out-projection)
  _layer_norm        LayerNorm / LayerNormFp32       transformer.py:15-30 (eps 1e-5, biased variance)
  _act               nn.GELU (erf) / QuickGELU       model.py:116,192 ; transformer.py:33-36
  normalize          F.normalize                     model.py:267,284 ; xclip/zero_shot.py:34,50
  zero_shot_logits   _compute_logits                 xclip/zero_shot.py:54-60
  predict / topk     predict_from_features / accuracy  xclip/zero_shot.py:103-109 ; training/zero_shot.py:11-14
  class_prompt_feat  OpenAIZeroShotClassifier.__init__ xclip/zero_shot.py:223-240
  clip_loss_local    ClipLoss(local_loss, gather_with_grad)  deps/open_clip/src/open_clip/loss.py:89-131
  resnet_forward     ModifiedResNet.forward (eval)   deps/open_clip/src/open_clip/modified_resnet.py:95-181 (Bottleneck :42-56,
                     AttentionPool2d :69-92; torch nn.Conv2d = unfold + matmul, nn.BatchNorm2d on running statistics,
                     nn.AvgPool2d(2) = mean of 2x2 windows, F.multi_head_attention_forward with separate q/k/v weights)

PINNING: the reference ships no golden tensors for this path (tests/data is absent), so the
oracle is pinned against outputs of the reference itself, generated in the build container by
`oracle/make_golden.py` (imports the unmodified reference from /root/reference) and committed under
`tests/golden/`; `tests/test_oracle_cpu.py` checks the oracle against them on every CPU run.
"""
from __future__ import annotations

import math

import torch

__all__ = [
    "vit_forward", "text_forward", "normalize", "zero_shot_logits", "predict", "topk", "class_prompt_feat",
    "clip_loss_local", "clip_loss_local_grads", "cfg_from_state_dict", "preprocess_u8", "train_step_grads", "resnet_forward", "randomize_batchnorm_", "test_images",
]


def _layer_norm(x: torch.Tensor, g: torch.Tensor, b: torch.Tensor, eps: float = 1e-5) -> torch.Tensor:
    mu = x.mean(dim=-1, keepdim=True)
    xc = x - mu
    var = (xc * xc).mean(dim=-1, keepdim=True)
    return xc * torch.rsqrt(var + eps) * g + b


def _act(x: torch.Tensor, quick_gelu: bool) -> torch.Tensor:
    if quick_gelu:
        return x * torch.sigmoid(1.702 * x)
    return 0.5 * x * (1.0 + torch.erf(x * (1.0 / math.sqrt(2.0))))


def _block(x: torch.Tensor, sd: dict, prefix: str, heads: int, mask: torch.Tensor | None, quick_gelu: bool) -> torch.Tensor:
    """x: [B, L, W] (batch-major; the reference's LND layout is an nn.MultiheadAttention detail)."""
    B, L, W = x.shape
    hd = W // heads
    h = _layer_norm(x, sd[prefix + "ln_1.weight"], sd[prefix + "ln_1.bias"])
    qkv = h @ sd[prefix + "attn.in_proj_weight"].t() + sd[prefix + "attn.in_proj_bias"]
    q, k, v = qkv.split(W, dim=-1)
    q = q.reshape(B, L, heads, hd).transpose(1, 2)
    k = k.reshape(B, L, heads, hd).transpose(1, 2)
    v = v.reshape(B, L, heads, hd).transpose(1, 2)
    s = (q @ k.transpose(-1, -2)) * (1.0 / math.sqrt(hd))
    if mask is not None:
        s = s + mask
    s = s - s.amax(dim=-1, keepdim=True)
    p = torch.exp(s)
    p = p / p.sum(dim=-1, keepdim=True)
    o = (p @ v).transpose(1, 2).reshape(B, L, W)
    x = x + o @ sd[prefix + "attn.out_proj.weight"].t() + sd[prefix + "attn.out_proj.bias"]
    h = _layer_norm(x, sd[prefix + "ln_2.weight"], sd[prefix + "ln_2.bias"])
    h = _act(h @ sd[prefix + "mlp.c_fc.weight"].t() + sd[prefix + "mlp.c_fc.bias"], quick_gelu)
    x = x + h @ sd[prefix + "mlp.c_proj.weight"].t() + sd[prefix + "mlp.c_proj.bias"]
    return x


def _num_layers(sd: dict, prefix: str) -> int:
    n = 0
    while f"{prefix}{n}.ln_1.weight" in sd:
        n += 1
    return n


def cfg_from_state_dict(sd: dict) -> dict:
    """Recover the architecture from an open_clip CLIP state_dict (keys listed in SURVEY.md §8b)."""
    conv = sd["visual.conv1.weight"]
    W, _, P, _ = conv.shape
    Lv = sd["visual.positional_embedding"].shape[0]
    g = int(round(math.sqrt(Lv - 1)))
    Wt = sd["token_embedding.weight"].shape[1]
    return {
        "vision_width": W, "patch": P, "image_size": g * P, "vision_layers": _num_layers(sd, "visual.transformer.resblocks."),
        "vision_heads": W // 64, "text_width": Wt, "text_layers": _num_layers(sd, "transformer.resblocks."),
        "text_heads": Wt // 64, "context_length": sd["positional_embedding"].shape[0],
        "vocab_size": sd["token_embedding.weight"].shape[0], "embed_dim": sd["text_projection"].shape[1],
    }


def _f32(sd: dict) -> dict:
    return {k: v.detach().to(torch.float32).cpu() for k, v in sd.items()}


def vit_forward(sd: dict, image: torch.Tensor, *, heads: int | None = None, quick_gelu: bool = False) -> torch.Tensor:
    """image [B,3,S,S] -> [B,D]   (transformer.py:601-643, pool_type 'tok', final_ln_after_pool False)."""
    return _vit(_f32(sd), image.detach().to(torch.float32).cpu(), heads, quick_gelu)


def _vit(sd: dict, x: torch.Tensor, heads: int | None, quick_gelu: bool) -> torch.Tensor:
    conv = sd["visual.conv1.weight"]
    W, _, P, _ = conv.shape
    B, _, S, _ = x.shape
    g = S // P
    heads = heads or W // 64
    # conv with kernel == stride == P and no bias is a GEMM over non-overlapping patches, K order (c, ky, kx)
    patches = x.reshape(B, 3, g, P, g, P).permute(0, 2, 4, 1, 3, 5).reshape(B, g * g, 3 * P * P)
    tok = patches @ conv.reshape(W, 3 * P * P).t()
    cls = sd["visual.class_embedding"].reshape(1, 1, W).expand(B, 1, W)
    t = torch.cat([cls, tok], dim=1) + sd["visual.positional_embedding"]
    t = _layer_norm(t, sd["visual.ln_pre.weight"], sd["visual.ln_pre.bias"])
    for i in range(_num_layers(sd, "visual.transformer.resblocks.")):
        t = _block(t, sd, f"visual.transformer.resblocks.{i}.", heads, None, quick_gelu)
    t = _layer_norm(t, sd["visual.ln_post.weight"], sd["visual.ln_post.bias"])
    return t[:, 0] @ sd["visual.proj"]


def text_forward(sd: dict, text: torch.Tensor, *, heads: int | None = None, quick_gelu: bool = False,
                 seq_len: int | None = None) -> torch.Tensor:
    """text int64 [T, ctx] -> [T, D]   (model.py:269-284; causal mask transformer.py:751-757; EOT pool :654).
    `seq_len` < ctx runs only the leading positions (exact when it exceeds every EOT index, by causality)."""
    return _text(_f32(sd), text.detach().cpu().long(), heads, quick_gelu, seq_len)


def _text(sd: dict, text: torch.Tensor, heads: int | None, quick_gelu: bool, seq_len: int | None) -> torch.Tensor:
    Wt = sd["token_embedding.weight"].shape[1]
    heads = heads or Wt // 64
    L = text.shape[1] if seq_len is None else seq_len
    eot = text.argmax(dim=-1)
    assert int(eot.max()) < L, "seq_len must exceed every EOT index"
    x = sd["token_embedding.weight"][text[:, :L]] + sd["positional_embedding"][:L]
    mask = torch.full((L, L), float("-inf")).triu_(1)
    for i in range(_num_layers(sd, "transformer.resblocks.")):
        x = _block(x, sd, f"transformer.resblocks.{i}.", heads, mask, quick_gelu)
    x = _layer_norm(x, sd["ln_final.weight"], sd["ln_final.bias"])
    pooled = x[torch.arange(x.shape[0]), eot]
    return pooled @ sd["text_projection"]


def randomize_batchnorm_(visual: torch.nn.Module, seed: int = 5, branch_gain: float = 0.25) -> None:
    """Test helper: give every BatchNorm2d of a ModifiedResNet tower (the reference's or this repo's: same module order)
    seeded non-trivial affine parameters.  At initialisation bn3.weight is zero (modified_resnet.py:143-146), which would
    make every bottleneck's main branch vanish and the parity check blind."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for name, m in visual.named_modules():
            if isinstance(m, torch.nn.BatchNorm2d):
                n = m.weight.shape
                m.weight.copy_(torch.rand(n, generator=g) * 0.8 + 0.6)
                m.bias.copy_(torch.randn(n, generator=g) * 0.1)
                if name.startswith("layer") and name.endswith(".bn3"):
                    m.weight.mul_(branch_gain)      # residual branches weaker than the identity path, as in a trained tower


def test_images(n: int, size: int, seed: int) -> torch.Tensor:
    """Seeded images with low-frequency structure (a random 7 x 7 field upsampled, plus pixel noise): unlike white noise they
    survive the tower's average pools, so the embeddings of different images differ."""
    g = torch.Generator().manual_seed(seed)
    low = torch.nn.functional.interpolate(torch.randn(n, 3, 7, 7, generator=g), size=size, mode="bilinear", align_corners=False)
    return 2.0 * low + 0.5 * torch.randn(n, 3, size, size, generator=g)


def _conv_bn(x: torch.Tensor, sd: dict, conv: str, bn: str, *, stride: int = 1, relu: bool = True) -> torch.Tensor:
    """nn.Conv2d (no bias, padding = k // 2) as unfold + matmul, then eval-mode nn.BatchNorm2d (eps 1e-5) and ReLU."""
    w = sd[conv + ".weight"]
    cout, cin, k, _ = w.shape
    B, _, H, W = x.shape
    pad = k // 2
    Ho, Wo = (H + 2 * pad - k) // stride + 1, (W + 2 * pad - k) // stride + 1
    cols = torch.nn.functional.unfold(x, k, padding=pad, stride=stride)            # [B, cin*k*k, Ho*Wo], K order (c, ky, kx)
    y = (w.reshape(cout, cin * k * k) @ cols).reshape(B, cout, Ho, Wo)
    g, b, m, v = (sd[f"{bn}.{n}"].reshape(1, cout, 1, 1) for n in ("weight", "bias", "running_mean", "running_var"))
    y = (y - m) / torch.sqrt(v + 1e-5) * g + b
    return y.clamp_min(0) if relu else y


def _avgpool2(x: torch.Tensor) -> torch.Tensor:
    B, C, H, W = x.shape
    return x.reshape(B, C, H // 2, 2, W // 2, 2).mean(dim=(3, 5))


def resnet_forward(sd: dict, image: torch.Tensor, *, heads: int | None = None) -> torch.Tensor:
    """image [B,3,S,S] -> [B,D]: ModifiedResNet.forward in eval mode (modified_resnet.py:163-181)."""
    sd = _f32(sd)
    x = image.detach().to(torch.float32).cpu()
    v = "visual."
    x = _conv_bn(x, sd, v + "conv1", v + "bn1", stride=2)
    x = _conv_bn(x, sd, v + "conv2", v + "bn2")
    x = _conv_bn(x, sd, v + "conv3", v + "bn3")
    x = _avgpool2(x)
    for li in range(1, 5):
        bi = 0
        while f"{v}layer{li}.{bi}.conv1.weight" in sd:
            p = f"{v}layer{li}.{bi}."
            stride = 2 if (li > 1 and bi == 0) else 1
            out = _conv_bn(x, sd, p + "conv1", p + "bn1")
            out = _conv_bn(out, sd, p + "conv2", p + "bn2")
            if stride > 1:
                out = _avgpool2(out)
            out = _conv_bn(out, sd, p + "conv3", p + "bn3", relu=False)
            identity = x
            if p + "downsample.0.weight" in sd:
                identity = _conv_bn(_avgpool2(x) if stride > 1 else x, sd, p + "downsample.0", p + "downsample.1", relu=False)
            x = (out + identity).clamp_min(0)
            bi += 1
    # AttentionPool2d (modified_resnet.py:69-92): the mean token queries all HW + 1 tokens
    a = v + "attnpool."
    B, E, H, W = x.shape
    heads = heads or E // 64
    t = x.reshape(B, E, H * W).permute(0, 2, 1)                                    # [B, HW, E]
    t = torch.cat([t.mean(dim=1, keepdim=True), t], dim=1) + sd[a + "positional_embedding"]
    q = t[:, :1] @ sd[a + "q_proj.weight"].t() + sd[a + "q_proj.bias"]
    k = t @ sd[a + "k_proj.weight"].t() + sd[a + "k_proj.bias"]
    val = t @ sd[a + "v_proj.weight"].t() + sd[a + "v_proj.bias"]
    hd = E // heads
    q = q.reshape(B, 1, heads, hd).transpose(1, 2) * hd ** -0.5
    k = k.reshape(B, -1, heads, hd).transpose(1, 2)
    val = val.reshape(B, -1, heads, hd).transpose(1, 2)
    att = torch.softmax(q @ k.transpose(-1, -2), dim=-1) @ val                     # [B, heads, 1, hd]
    att = att.transpose(1, 2).reshape(B, E)
    return att @ sd[a + "c_proj.weight"].t() + sd[a + "c_proj.bias"]


def normalize(x: torch.Tensor, eps: float = 1e-12) -> torch.Tensor:
    x = x.detach().to(torch.float32).cpu()
    return x / x.pow(2).sum(dim=-1, keepdim=True).sqrt().clamp_min(eps)


def zero_shot_logits(img_feat: torch.Tensor, prompt_feat: torch.Tensor, normalize_img: bool = True) -> torch.Tensor:
    """[B,D] x [C,D] -> [B,C]; no logit scale (xclip/zero_shot.py:54-60)."""
    f = normalize(img_feat) if normalize_img else img_feat.detach().float().cpu()
    return f @ prompt_feat.detach().float().cpu().t()


def predict(logits: torch.Tensor) -> torch.Tensor:
    """argmax over classes, first (lowest) index on ties (xclip/zero_shot.py:107)."""
    B, C = logits.shape
    best = logits.amax(dim=1, keepdim=True)
    idx = torch.arange(C).expand(B, C)
    return torch.where(logits == best, idx, torch.full_like(idx, C)).amin(dim=1)


def topk(logits: torch.Tensor, k: int) -> torch.Tensor:
    """indices of the k largest logits per row, descending, lower index first on ties (training/zero_shot.py:11-14)."""
    return torch.sort(logits, dim=1, descending=True, stable=True).indices[:, :k]


def class_prompt_feat(txt_feat: torch.Tensor, classes: int, templates: int) -> torch.Tensor:
    """normalize -> mean over templates -> normalize (xclip/zero_shot.py:231-234); txt_feat [classes*templates, D]."""
    e = normalize(txt_feat).reshape(classes, templates, -1)
    return normalize(e.mean(dim=1))


def _log_softmax_rows(z: torch.Tensor) -> torch.Tensor:
    m = z.amax(dim=1, keepdim=True)
    return z - m - torch.log(torch.exp(z - m).sum(dim=1, keepdim=True))


def clip_loss_local(img_loc, txt_loc, all_img, all_txt, logit_scale, rank: int) -> torch.Tensor:
    """loss.py:102-131 with local_loss=True: rows of this rank against all gathered columns, labels i + n*rank."""
    img_loc, txt_loc, all_img, all_txt = [t.detach().double().cpu() for t in (img_loc, txt_loc, all_img, all_txt)]
    s = float(logit_scale)
    n = img_loc.shape[0]
    labels = torch.arange(n) + n * rank
    li = _log_softmax_rows(s * img_loc @ all_txt.t())
    lt = _log_softmax_rows(s * txt_loc @ all_img.t())
    rows = torch.arange(n)
    return (-(li[rows, labels]).mean() - (lt[rows, labels]).mean()) / 2


def clip_loss_local_grads(img_loc, txt_loc, all_img, all_txt, logit_scale, rank: int):
    """Closed-form gradients of clip_loss_local w.r.t. (img_loc, txt_loc, all_img, all_txt, logit_scale), treating local
    and gathered operands as separate leaves (the reduce-scatter of d_all_* is the caller's collective)."""
    img_loc, txt_loc, all_img, all_txt = [t.detach().double().cpu() for t in (img_loc, txt_loc, all_img, all_txt)]
    s = float(logit_scale)
    n = img_loc.shape[0]
    labels = torch.arange(n) + n * rank
    rows = torch.arange(n)
    out = []
    d_scale = 0.0
    for a_loc, b_all in ((img_loc, all_txt), (txt_loc, all_img)):
        raw = a_loc @ b_all.t()
        p = torch.exp(_log_softmax_rows(s * raw))
        p[rows, labels] -= 1.0
        dl = p / (2 * n)
        out.append((s * dl @ b_all, s * dl.t() @ a_loc))
        d_scale += float((dl * raw).sum())
    (d_img_loc, d_all_txt), (d_txt_loc, d_all_img) = out
    return d_img_loc, d_txt_loc, d_all_img, d_all_txt, d_scale


def preprocess_u8(image_u8: torch.Tensor, mean, std) -> torch.Tensor:
    """ToTensor + Normalize on already resized / cropped uint8 pixels [B,3,H,W] (deps/open_clip/src/open_clip/transform.py:
    274-392: `ToTensor()` = x.float().div(255), `Normalize(mean, std)` = (x - mean) / std), fp32."""
    x = image_u8.float().div(255.0)
    m = torch.tensor(mean, dtype=torch.float32).view(1, 3, 1, 1)
    s = torch.tensor(std, dtype=torch.float32).view(1, 3, 1, 1)
    return x.sub(m).div(s)


def train_step_grads(sd: dict, image: torch.Tensor, text: torch.Tensor, *, quick_gelu: bool = False, dtype=torch.float32):
    """Loss and parameter gradients of ONE contrastive training step on one rank without accumulation — `model(image, text)`
    (CLIP.forward, model.py:295-315: both towers, F.normalize, logit_scale.exp()), ClipLoss.forward with world_size 1
    (loss.py:120-131), `backward()` (training/train.py:115-183) — by autograd through the plain-op restatement above.
    -> (loss float, {state_dict key: gradient}).  The checker for the tower backward of SURVEY §8(f)-1."""
    leaf = {k: v.detach().to(dtype).cpu().clone().requires_grad_(v.is_floating_point()) for k, v in sd.items()}
    img = _vit(leaf, image.detach().to(dtype).cpu(), None, quick_gelu)
    txt = _text(leaf, text.detach().cpu().long(), None, quick_gelu, None)
    img = img / img.norm(dim=-1, keepdim=True).clamp_min(1e-12)
    txt = txt / txt.norm(dim=-1, keepdim=True).clamp_min(1e-12)
    scale = leaf["logit_scale"].exp()
    li = _log_softmax_rows(scale * img @ txt.t())
    lt = _log_softmax_rows(scale * txt @ img.t())
    n = img.shape[0]
    rows = torch.arange(n)
    loss = (-(li[rows, rows]).mean() - (lt[rows, rows]).mean()) / 2
    loss.backward()
    return float(loss), {k: v.grad.detach().clone() for k, v in leaf.items() if v.requires_grad and v.grad is not None}
