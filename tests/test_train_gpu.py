"""Training path on the B200: tower backward + ClipLoss + fused AdamW through the public drop-in API
(`model(image, text)` -> `ClipLoss` -> `.backward()` -> `optimizer.step()`), against
  (1) the committed gradients of the UNMODIFIED reference's training step (tests/golden/tiny_train_grads.pt),
  (2) the oracle's autograd through its plain-op restatement on seeded inputs (16-bit modes, larger shapes),
  (3) torch.optim.AdamW for the optimizer.
Tolerances: fp32 gradients 1e-4 relative per tensor (north_star's fp32 gate), loss 1e-3 relative; 16-bit gradients are compared
per tensor with a cosine / relative-L2 bound that reflects bf16 round-off through the depth of the model."""
from pathlib import Path

import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import clip_oracle as O  # noqa: E402  (the checker)
from understanding_clip_ood_b200 import open_clip  # noqa: E402

GOLD = Path(__file__).resolve().parent / "golden"
DEV = "cuda"


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-300))


@pytest.fixture(scope="module")
def tiny():
    return torch.load(GOLD / "tiny_clip.pt", weights_only=False)


def _train_step(model, image, text):
    model.train()
    model.zero_grad(set_to_none=True)
    fi, ft, scale = model(image, text)
    loss = open_clip.ClipLoss()(fi, ft, scale)
    loss.backward()
    return loss, {k: p.grad for k, p in model.named_parameters() if p.grad is not None}


def test_tiny_fp32_training_step_matches_reference_gradients(tiny):
    """All 62 parameter gradients of the reference's own training step (model(image, text) -> ClipLoss -> backward, fp32)."""
    gold = torch.load(GOLD / "tiny_train_grads.pt", weights_only=False)
    m = open_clip.create_model("ViT-B-32", precision="fp32", device=DEV, **tiny["cfg"])
    m.load_state_dict(tiny["state_dict"])
    loss, grads = _train_step(m, tiny["image"].to(DEV), tiny["text"][:6].to(DEV))
    assert abs(float(loss) - float(gold["loss"])) / float(gold["loss"]) < 1e-5
    assert set(grads) == set(gold["grads"])
    worst = max((rel(grads[k], gold["grads"][k]), k) for k in gold["grads"])
    assert worst[0] < 1e-4, worst
    for k, g in grads.items():
        assert g.dtype == torch.float32 and g.shape == gold["grads"][k].shape


def test_training_forward_equals_inference_forward(tiny):
    """The training forward runs the inference kernels (plus the activation copies): same features."""
    m = open_clip.create_model("ViT-B-32", precision="fp32", device=DEV, **tiny["cfg"])
    m.load_state_dict(tiny["state_dict"])
    image, text = tiny["image"].to(DEV), tiny["text"][:6].to(DEV)
    m.eval()
    with torch.no_grad():
        fi0, ft0, _ = m(image, text)
    m.train()
    fi1, ft1, _ = m(image, text)
    assert fi1.requires_grad and ft1.requires_grad and fi1.grad_fn is not None
    assert torch.equal(fi0, fi1.detach()) and torch.equal(ft0, ft1.detach())
    m.lock_image_tower()
    fi2, _, _ = m(image, text)
    assert not fi2.requires_grad                      # a locked tower takes the inference path


@pytest.mark.parametrize("precision,dtype,quick", [("bf16", torch.bfloat16, False), ("fp16", torch.float16, True), ("amp_bf16", torch.bfloat16, False)])
def test_low_precision_training_step_matches_oracle(precision, dtype, quick):
    """A ViT-B/32-shaped model (width 768 / 512, 3 layers, 50 / 77 tokens) in the 16-bit modes: tcgen05 dgrad / wgrad GEMMs
    (stream-K over the long token contraction), against the oracle's fp32 autograd on the same weights and inputs."""
    cfg = dict(embed_dim=512, vision_cfg={"image_size": 224, "layers": 3, "width": 768, "patch_size": 32},
               text_cfg={"context_length": 77, "vocab_size": 1000, "width": 512, "heads": 8, "layers": 3})
    torch.manual_seed(11)
    m = open_clip.create_model("ViT-B-32", precision=precision, device="cpu", force_quick_gelu=quick, **cfg)
    sd = {k: v.detach().float().clone() for k, v in m.state_dict().items()}
    m = m.to(DEV)
    g = torch.Generator().manual_seed(12)
    B = 24
    image = torch.randn(B, 3, 224, 224, generator=g)
    text = torch.zeros(B, 77, dtype=torch.long)
    for i in range(B):
        n = 3 + int(torch.randint(0, 12, (1,), generator=g))
        text[i, 0], text[i, n + 1] = 998, 999
        text[i, 1:n + 1] = torch.randint(1, 998, (n,), generator=g)
    img_in = image.to(DEV) if precision == "amp_bf16" else image.to(dtype).to(DEV)
    loss, grads = _train_step(m, img_in, text.to(DEV))
    want_loss, want = O.train_step_grads(sd, image.to(dtype).float(), text, quick_gelu=quick)
    assert abs(float(loss) - want_loss) / want_loss < 2e-2
    assert set(grads) == set(want)
    bad = []
    for k, gw in want.items():
        gg = grads[k].float().cpu()
        want_dtype = torch.float32 if precision == "amp_bf16" else m.state_dict()[k].dtype
        assert grads[k].dtype == want_dtype, (k, grads[k].dtype)
        if float(gw.norm()) < 1e-7:
            continue
        cos = float((gg.double().flatten() @ gw.double().flatten()) / (gg.double().norm() * gw.double().norm()).clamp_min(1e-300))
        if cos < 0.98 or rel(gg, gw) > 0.2:
            bad.append((k, cos, rel(gg, gw)))
    assert not bad, bad[:8]


def test_fused_adamw_matches_torch_adamw():
    g = torch.Generator(device=DEV).manual_seed(5)
    shapes = [(768, 768), (3072,), (5, 7, 3), (1,), (4097,)]
    for dtype, tol in ((torch.float32, 2e-6), (torch.bfloat16, 1e-2)):
        ours = [torch.randn(s, device=DEV, generator=g).to(dtype).requires_grad_(True) for s in shapes]
        ref = [p.detach().float().clone().requires_grad_(True) for p in ours]
        o1 = open_clip.AdamW([{"params": ours[:2], "weight_decay": 0.0}, {"params": ours[2:], "weight_decay": 0.2}], lr=1e-2, betas=(0.9, 0.98), eps=1e-6)
        o2 = torch.optim.AdamW([{"params": ref[:2], "weight_decay": 0.0}, {"params": ref[2:], "weight_decay": 0.2}], lr=1e-2, betas=(0.9, 0.98), eps=1e-6)
        hold = []
        for step in range(4):
            for p, r in zip(ours, ref):
                gr = torch.randn(p.shape, device=DEV, generator=g)
                if step % 2 == 0:
                    hold.append(p.grad)      # keeps the old buffer alive: the next gradient lands at a NEW address (the optimizer
                p.grad = gr.to(dtype)        # re-points its device table), on odd steps the allocator may hand the old one back
                r.grad = gr.to(dtype).float()
            v0 = ours[0]._version
            o1.step()
            o2.step()
            assert ours[0]._version > v0
        for p, r in zip(ours, ref):
            assert rel(p.float(), r) < tol, (dtype, p.shape, rel(p.float(), r))


def test_multi_tensor_cast_refreshes_the_operand_copies():
    """_Keep.cast_pairs (b200clip_multi_cast): every (copy, source) pair converted by ONE launch — what the engines run after an
    optimizer step in the amp modes — equals torch's per-tensor copy_, for aligned and unaligned / odd-sized tensors, fp32 -> 16-bit
    and 16-bit -> fp32, and again after the sources changed (cached device table)."""
    from understanding_clip_ood_b200.open_clip.model import _Keep
    g = torch.Generator(device=DEV).manual_seed(6)
    keep = _Keep()
    flat = torch.randn(10000, device=DEV, generator=g)
    srcs = [torch.randn(768, 768, device=DEV, generator=g), torch.randn(4097, device=DEV, generator=g), torch.randn(3, device=DEV, generator=g),
            flat[1:1 + 2049], torch.randn(512, 77, device=DEV, generator=g).bfloat16(), torch.randn(9, device=DEV, generator=g).half()]
    dsts = [torch.empty(768, 768, device=DEV, dtype=torch.bfloat16), torch.empty(4097, device=DEV, dtype=torch.float16),
            torch.empty(3, device=DEV, dtype=torch.bfloat16), torch.empty(2049, device=DEV, dtype=torch.bfloat16),
            torch.empty(512, 77, device=DEV), torch.empty(9, device=DEV)]
    keep.pairs = list(zip(dsts, srcs))
    for _ in range(2):
        keep.cast_pairs()
        for d, s_ in keep.pairs:
            assert torch.equal(d, s_.to(d.dtype))
        for s_ in srcs:
            s_.mul_(1.5)


def test_training_loop_reduces_the_loss(tiny):
    """Drop-in train loop on the tiny model: forward, ClipLoss, backward, fused AdamW, logit_scale clamp (train.py:190-191)."""
    m = open_clip.create_model("ViT-B-32", precision="fp32", device=DEV, **tiny["cfg"])
    m.load_state_dict(tiny["state_dict"])
    m.train()
    opt = open_clip.AdamW(m.parameters(), lr=3e-4, betas=(0.9, 0.98), eps=1e-6, weight_decay=0.1)
    image, text = tiny["image"].to(DEV), tiny["text"][:6].to(DEV)
    loss_fn = open_clip.ClipLoss()
    losses = []
    for _ in range(12):
        opt.zero_grad(set_to_none=True)
        fi, ft, scale = m(image, text)
        loss = loss_fn(fi, ft, scale)
        loss.backward()
        opt.step()
        with torch.no_grad():
            m.logit_scale.clamp_(0, 4.6052)
        losses.append(float(loss))
    assert losses[-1] < losses[0] - 0.2, losses
