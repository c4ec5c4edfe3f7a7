"""Generate the committed golden fixtures under tests/golden/ by RUNNING THE UNMODIFIED REFERENCE
(/root/reference, CPU) in the build container — TEST INFRASTRUCTURE ONLY, never imported by the product.

    python -m oracle.make_golden [--skip-full] [--skip-bf16]

Fixtures (all outputs of reference code on seeded inputs):
  tiny_clip.pt        small ViT/text CLIP built by the reference's own `create_model` with cfg overrides
                      (factory.py:260): state_dict + inputs + `encode_image` / `encode_text` / `forward` outputs
                      (GELU and QuickGELU), `xclip.zero_shot.ZeroShotClassifier` / `OpenAIZeroShotClassifier`
                      prompt features, logits and predictions, and `training.zero_shot.accuracy`-style top-5.
  cliploss.pt         reference `ClipLoss`: world_size 1 (loss + grads) and a 2-rank gloo run with
                      local_loss=True, gather_with_grad=True (per-rank loss, feature / logit_scale grads).
  tiny_train_grads.pt reference training step on the tiny model (model(image, text) -> ClipLoss -> backward): loss and the
                      gradient of every parameter; pins oracle.train_step_grads (checker of the tower backward).
  domainnet_prompts.npz   reference tokenizer output for the 345 DomainNet classes x 86 OpenAI templates
                      (class names = keys of data/in_to_dn_mapping.json, label order), truncated to the
                      first 24 context positions (max EOT index is 15), uint16.
  vitb32_seed0.pt     BASELINE config 1: reference ViT-B-32, `torch.manual_seed(0)` random init, fp32, CPU:
                      images = randn(64,3,224,224; seed 1) -> normalised image features, OpenAIZeroShotClassifier
                      prompt_feat for the 345 classes (86 templates), logits, top-1 / top-5.  The weights are NOT
                      stored: `create_model` of this repo reproduces the reference's seed-0 init bit-exactly
                      (checked in tests/test_reference_pins_cpu.py), so the GPU box rebuilds them from the seed.
  vitb32_seed0_bf16.pt  same images through the reference instantiated with precision='bf16' (CPU): the
                      "ours-bf16 vs reference-bf16" anchor and the reference's own bf16<->fp32 agreement floor.
Also writes understanding_clip_ood_b200/data/{openai_templates,domainnet_classes}.json (prompt DATA).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from oracle import ref_loader  # noqa: E402

GOLD = ROOT / "tests" / "golden"
DATA = ROOT / "understanding_clip_ood_b200" / "data"

TINY = dict(embed_dim=64,
            vision_cfg={"image_size": 64, "layers": 2, "width": 128, "patch_size": 16},
            text_cfg={"context_length": 77, "vocab_size": 300, "width": 64, "heads": 1, "layers": 2})


def fake_tokens(n: int, vocab: int, ctx: int, seed: int, min_len: int = 2, max_len: int = 12) -> torch.Tensor:
    """[n, ctx] int64 in the tokenizer's layout: sot, words..., eot (= vocab-1, the arg-max), zero padding."""
    g = torch.Generator().manual_seed(seed)
    out = torch.zeros(n, ctx, dtype=torch.long)
    for i in range(n):
        ln = int(torch.randint(min_len, max_len + 1, (1,), generator=g))
        out[i, 0] = vocab - 2
        out[i, 1:1 + ln] = torch.randint(1, vocab - 2, (ln,), generator=g)
        out[i, 1 + ln] = vocab - 1
    return out


class FakeTokenizer:
    """Deterministic stand-in for a BPE tokenizer over a tiny vocabulary: ids derive from a hash of the text."""

    def __init__(self, vocab: int, ctx: int = 77):
        self.vocab, self.ctx = vocab, ctx

    def __call__(self, texts):
        if isinstance(texts, str):
            texts = [texts]
        out = torch.zeros(len(texts), self.ctx, dtype=torch.long)
        for i, t in enumerate(texts):
            words = t.lower().split()[: self.ctx - 2]
            ids = [1 + (sum(ord(c) * (j + 7) for j, c in enumerate(w)) % (self.vocab - 3)) for w in words]
            out[i, 0] = self.vocab - 2
            out[i, 1:1 + len(ids)] = torch.tensor(ids, dtype=torch.long)
            out[i, 1 + len(ids)] = self.vocab - 1
        return out


def make_tiny(open_clip, zs, xo):
    torch.manual_seed(0)
    ref = open_clip.create_model("ViT-B-32", precision="fp32", **TINY).eval()
    sd = {k: v.clone() for k, v in ref.state_dict().items()}
    image = torch.randn(6, 3, 64, 64, generator=torch.Generator().manual_seed(1))
    text = fake_tokens(10, 300, 77, seed=2)
    out = {"cfg": TINY, "state_dict": sd, "image": image, "text": text}
    with torch.no_grad():
        out["image_features"] = ref.encode_image(image)
        out["text_features"] = ref.encode_text(text)
        fi, ft, scale = ref(image, text[:6])
        out["forward"] = (fi, ft, scale)
    refq = open_clip.create_model("ViT-B-32", precision="fp32", force_quick_gelu=True, **TINY).eval()
    refq.load_state_dict(sd)
    with torch.no_grad():
        out["image_features_quickgelu"] = refq.encode_image(image)
        out["text_features_quickgelu"] = refq.encode_text(text)

    # the reference's zero-shot classifiers on top of the reference model
    clip = xo.OpenCLIP(ref)
    names = ["aircraft carrier", "alarm clock", "ant", "apple", "axe", "banana", "The Eiffel Tower"]
    tok = FakeTokenizer(300)
    z = zs.ZeroShotClassifier(clip, tok, names, prompt_fn=lambda c: f"a photo of a {c}.")
    out["zs_names"] = names
    out["zs_prompt_feat"] = z.prompt_feat.clone()
    out["zs_logits"] = z.predict(image, return_scores=True)["pred"].clone()
    out["zs_pred"] = z.predict(image)["pred"].clone()
    feat = z._compute_img_feat(image)
    out["zs_img_feat"] = feat.clone()
    out["zs_pred_from_features"] = z.predict_from_features(feat)["pred"].clone()
    out["zs_top5"] = (100.0 * out["zs_logits"]).topk(5, 1, True, True)[1].clone()   # training/zero_shot.py:11-14
    zo = zs.OpenAIZeroShotClassifier(clip, tok, names)
    out["openai_prompt_feat"] = zo.prompt_feat.clone()
    out["openai_pred"] = zo.predict(image)["pred"].clone()
    out["openai_logits"] = zo.predict(image, return_scores=True)["pred"].clone()
    zd = zs.OpenAIZeroShotClassifier(clip, tok, names, domain_invariant=True)
    out["openai_di_templates"] = len(zd.templates)
    out["openai_di_prompt_feat"] = zd.prompt_feat.clone()
    torch.save(out, GOLD / "tiny_clip.pt")
    print("tiny_clip.pt", {k: tuple(v.shape) for k, v in out.items() if torch.is_tensor(v)})


def _cliploss_worker(rank, world, port, feats, scale0, ret):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    open_clip, _, _ = ref_loader.load()
    n = feats[0].shape[0] // world
    img = feats[0][rank * n:(rank + 1) * n].clone().requires_grad_(True)
    txt = feats[1][rank * n:(rank + 1) * n].clone().requires_grad_(True)
    scale = torch.tensor(scale0, requires_grad=True)
    loss_fn = open_clip.loss.ClipLoss(local_loss=True, gather_with_grad=True, cache_labels=True, rank=rank, world_size=world)
    loss = loss_fn(img, txt, scale)
    loss.backward()
    ret[rank] = {"loss": loss.detach(), "d_img": img.grad.clone(), "d_txt": txt.grad.clone(), "d_scale": scale.grad.clone()}
    dist.destroy_process_group()


def make_cliploss(open_clip):
    import torch.multiprocessing as mp
    g = torch.Generator().manual_seed(3)
    N, D = 32, 64
    img = torch.nn.functional.normalize(torch.randn(N, D, generator=g), dim=-1)
    txt = torch.nn.functional.normalize(torch.randn(N, D, generator=g) + 0.7 * img, dim=-1)
    scale0 = float(1 / 0.07)
    out = {"img": img, "txt": txt, "scale": scale0}
    # world_size == 1
    a, b = img.clone().requires_grad_(True), txt.clone().requires_grad_(True)
    s = torch.tensor(scale0, requires_grad=True)
    loss = open_clip.loss.ClipLoss()(a, b, s)
    loss.backward()
    out["w1"] = {"loss": loss.detach(), "d_img": a.grad.clone(), "d_txt": b.grad.clone(), "d_scale": s.grad.clone()}
    # 2 ranks, gloo, local loss + gather with grad
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_cliploss_worker, args=(2, 29533, (img, txt), scale0, ret), nprocs=2, join=True)
    out["w2"] = {r: dict(ret[r]) for r in range(2)}
    torch.save(out, GOLD / "cliploss.pt")
    print("cliploss.pt  w1 loss", float(out["w1"]["loss"]), " w2 losses", [float(out["w2"][r]["loss"]) for r in range(2)])


def domainnet_names():
    return list(json.load(open(ref_loader.REFERENCE / "data" / "in_to_dn_mapping.json")).keys())


def make_prompts(open_clip, zs):
    names = domainnet_names()
    templates = list(zs.OpenAIZeroShotClassifier.templates)
    DATA.mkdir(exist_ok=True)
    (DATA / "openai_templates.json").write_text(json.dumps(templates, indent=0))
    (DATA / "domainnet_classes.json").write_text(json.dumps(names, indent=0))
    tok = open_clip.get_tokenizer("ViT-B-32")
    texts = [t.format(c) for c in names for t in templates]
    tokens = tok(texts)
    eot = tokens.argmax(-1)
    keep = 24
    assert int(eot.max()) < keep and int(tokens[:, keep:].abs().sum()) == 0
    np.savez_compressed(GOLD / "domainnet_prompts.npz", tokens=tokens[:, :keep].numpy().astype(np.uint16),
                        classes=len(names), templates=len(templates), context_length=77, eot_max=int(eot.max()))
    print("domainnet_prompts.npz", tokens.shape, "eot max", int(eot.max()), "mean", float(eot.float().mean()))
    return names, templates, tokens


def make_full(open_clip, zs, xo, skip_bf16: bool):
    names = domainnet_names()
    tok = open_clip.get_tokenizer("ViT-B-32")
    torch.set_num_threads(os.cpu_count())
    image = torch.randn(64, 3, 224, 224, generator=torch.Generator().manual_seed(1))
    torch.manual_seed(0)
    ref = open_clip.create_model("ViT-B-32", precision="fp32").eval()
    clip = xo.OpenCLIP(ref)
    t0 = time.time()
    z = zs.OpenAIZeroShotClassifier(clip, tok, names)
    t_cls = time.time() - t0
    t0 = time.time()
    feat = z._compute_img_feat(image)
    t_img = time.time() - t0
    logits = z.predict_from_features(feat, return_scores=True)["pred"]
    out = {"seed_weights": 0, "seed_images": 1, "image_features_normalized": feat.clone(), "prompt_feat": z.prompt_feat.clone(),
           "logits": logits.clone(), "pred": z.predict_from_features(feat)["pred"].clone(),
           "top5": logits.topk(5, 1, True, True)[1].clone(), "cpu_seconds_classifier_build": t_cls, "cpu_seconds_encode_image_64": t_img,
           "cpu_threads": torch.get_num_threads()}
    with torch.no_grad():
        out["text_features_first_class"] = ref.encode_text(tok([t.format(names[0]) for t in z.templates])).clone()
    torch.save(out, GOLD / "vitb32_seed0.pt")
    print(f"vitb32_seed0.pt  classifier build {t_cls:.0f}s  encode_image(64) {t_img:.1f}s")
    if skip_bf16:
        return
    torch.manual_seed(0)
    refb = open_clip.create_model("ViT-B-32", precision="bf16").eval()
    with torch.no_grad():
        fb = torch.nn.functional.normalize(refb.encode_image(image.bfloat16()), dim=-1)
        prompt_b = z.prompt_feat.bfloat16()
        lb = torch.tensordot(fb, prompt_b.movedim(-1, 0), dims=1)
    torch.save({"image_features_normalized": fb.clone(), "logits_vs_fp32_prompts": lb.clone(), "pred": lb.argmax(dim=1).clone(),
                "top5": lb.float().topk(5, 1, True, True)[1].clone()}, GOLD / "vitb32_seed0_bf16.pt")
    agree = (lb.argmax(1) == out["pred"]).float().mean().item()
    print(f"vitb32_seed0_bf16.pt  reference bf16 vs fp32 top-1 agreement {agree:.3f}")


def make_train_grads(open_clip):
    """One training step of the reference on the tiny model: model(image, text) -> ClipLoss -> backward.  Stores the loss and
    the gradient of every parameter (the golden vectors of the tower backward); inputs / weights are the ones
    of tiny_clip.pt."""
    tiny = torch.load(GOLD / "tiny_clip.pt", weights_only=False)
    ref = open_clip.create_model("ViT-B-32", precision="fp32", **TINY)
    ref.load_state_dict(tiny["state_dict"])
    ref.train()
    image, text = tiny["image"], tiny["text"][:6]
    fi, ft, scale = ref(image, text)
    loss = open_clip.ClipLoss()(fi, ft, scale)
    loss.backward()
    grads = {k: p.grad.detach().clone() for k, p in ref.named_parameters() if p.grad is not None}
    torch.save({"loss": loss.detach().clone(), "grads": grads}, GOLD / "tiny_train_grads.pt")
    print("tiny_train_grads.pt", float(loss), len(grads), "gradients,", sum(g.numel() for g in grads.values()), "values")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only-train-grads", action="store_true")
    ap.add_argument("--skip-full", action="store_true")
    ap.add_argument("--skip-bf16", action="store_true")
    ap.add_argument("--only-full", action="store_true")
    args = ap.parse_args()
    GOLD.mkdir(parents=True, exist_ok=True)
    open_clip, zs, xo = ref_loader.load()
    if args.only_train_grads:
        make_train_grads(open_clip)
        return
    if not args.only_full:
        make_prompts(open_clip, zs)
        make_tiny(open_clip, zs, xo)
        make_cliploss(open_clip)
        make_train_grads(open_clip)
    if not args.skip_full:
        make_full(open_clip, zs, xo, args.skip_bf16)


if __name__ == "__main__":
    main()
