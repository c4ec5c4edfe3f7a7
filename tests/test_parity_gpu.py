"""Parity tests proper (B200): the CUDA path, called through the drop-in Python boundary -> C ABI, against
  (1) the committed golden vectors = outputs of the unmodified reference (tests/golden/, oracle/make_golden.py),
  (2) the CPU oracle on the same seeded inputs at sizes it finishes in seconds,
  (3) size-independent properties at BASELINE.json's full sizes.
Tolerances are north_star's: embeddings 1e-4 relative (fp32) / 2e-2 (bf16), loss 1e-3 relative, top-1/top-5 indices.
"""
import json
from pathlib import Path

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import clip_oracle as O  # noqa: E402  (the checker)
from oracle.make_golden import FakeTokenizer  # noqa: E402
from tests.parity_metrics import prediction_parity_b1024  # noqa: E402
from understanding_clip_ood_b200 import open_clip, ops  # noqa: E402
from understanding_clip_ood_b200.xclip import zero_shot as zs  # noqa: E402
from understanding_clip_ood_b200.xclip.open_clip import OpenCLIP  # noqa: E402

GOLD = Path(__file__).resolve().parent / "golden"
DEV = "cuda"


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-300))


def row_rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float(((a - b).norm(dim=-1) / b.norm(dim=-1).clamp_min(1e-300)).max())


@pytest.fixture(scope="module")
def tiny():
    return torch.load(GOLD / "tiny_clip.pt", weights_only=False)


def tiny_model(tiny, precision="fp32", quick=False):
    m = open_clip.create_model("ViT-B-32", precision=precision, device=DEV, force_quick_gelu=quick, **tiny["cfg"])
    m.load_state_dict(tiny["state_dict"])
    return m.eval()


# ------------------------------------------------------------------ uint8 pixel batches (preprocessing tail on the GPU)
@pytest.mark.parametrize("precision,dtype", [("fp32", torch.float32), ("bf16", torch.bfloat16), ("fp16", torch.float16)])
def test_uint8_images_equal_reference_preprocessing_then_forward(tiny, precision, dtype):
    """encode_image(uint8 pixels) == encode_image(ToTensor + Normalize of the same pixels, cast like the scripts do): the
    normalisation is fused into the im2col with the reference's operation order, so the two routes agree bit for bit; and
    the fp32 route matches the CPU oracle run on the oracle-preprocessed image."""
    m = tiny_model(tiny, precision)
    S = m.visual.image_size[0]
    g = torch.Generator().manual_seed(5)
    u8 = torch.randint(0, 256, (6, 3, S, S), generator=g, dtype=torch.uint8)
    pp = m.visual.preprocess_cfg
    ref_img = O.preprocess_u8(u8, pp["mean"], pp["std"])
    a = m.encode_image(u8.to(DEV))
    b = m.encode_image(ref_img.to(dtype).to(DEV))
    assert a.dtype == dtype and torch.equal(a, b)
    a2 = m.encode_image(u8.to(DEV))                  # second sighting of the same buffer: CUDA-graph replay
    assert torch.equal(a2, a)
    if precision == "fp32":
        want = O.vit_forward(tiny["state_dict"], ref_img)
        assert row_rel(a, want) < 1e-4
    clf = zs.ZeroShotClassifier(OpenCLIP(m), FakeTokenizer(300), [f"class {i}" for i in range(7)])
    assert torch.equal(clf.predict(u8.to(DEV))["pred"], clf.predict(ref_img.to(dtype).to(DEV))["pred"])


# ------------------------------------------------------------------ (1) golden vectors of the reference
def test_tiny_fp32_embeddings_match_reference(tiny):
    m = tiny_model(tiny)
    img = m.encode_image(tiny["image"].to(DEV))
    txt = m.encode_text(tiny["text"].to(DEV))
    assert img.dtype == torch.float32 and img.shape == (6, 64) and txt.shape == (10, 64)
    assert row_rel(img, tiny["image_features"]) < 1e-4
    assert row_rel(txt, tiny["text_features"]) < 1e-4
    fi, ft, scale = m(tiny["image"].to(DEV), tiny["text"][:6].to(DEV))
    assert row_rel(fi, tiny["forward"][0]) < 1e-4 and row_rel(ft, tiny["forward"][1]) < 1e-4
    assert float(scale) == pytest.approx(float(tiny["forward"][2]), rel=1e-6)
    m.output_dict = True
    d = m(tiny["image"].to(DEV), tiny["text"][:6].to(DEV))
    assert set(d) == {"image_features", "text_features", "logit_scale"} and torch.equal(d["image_features"], fi)


def test_tiny_quickgelu_matches_reference(tiny):
    m = tiny_model(tiny, quick=True)
    assert row_rel(m.encode_image(tiny["image"].to(DEV)), tiny["image_features_quickgelu"]) < 1e-4
    assert row_rel(m.encode_text(tiny["text"].to(DEV)), tiny["text_features_quickgelu"]) < 1e-4


@pytest.mark.parametrize("precision,dtype", [("bf16", torch.bfloat16), ("fp16", torch.float16)])
def test_tiny_low_precision_within_tolerance(tiny, precision, dtype):
    m = tiny_model(tiny, precision)
    img = m.encode_image(tiny["image"].to(DEV, dtype))
    txt = m.encode_text(tiny["text"].to(DEV))
    assert img.dtype == dtype and txt.dtype == dtype
    tol = 2e-2 if dtype == torch.bfloat16 else 5e-3
    assert row_rel(img, tiny["image_features"]) < tol
    assert row_rel(txt, tiny["text_features"]) < tol
    with pytest.raises(RuntimeError):
        m.encode_image(tiny["image"].to(DEV))          # fp32 input into a 16-bit model: same contract as the reference


def test_text_truncation_is_exact(tiny):
    m = tiny_model(tiny)
    full = m.encode_text(tiny["text"].to(DEV))
    m.truncate_text_at_eot = True
    trunc = m.encode_text(tiny["text"].to(DEV))
    assert row_rel(trunc, full) < 2e-6


def test_zero_shot_classifiers_match_reference(tiny):
    clip = OpenCLIP(tiny_model(tiny))
    tok = FakeTokenizer(300)
    z = zs.ZeroShotClassifier(clip, tok, tiny["zs_names"], prompt_fn=lambda c: f"a photo of a {c}.")
    assert row_rel(z.prompt_feat, tiny["zs_prompt_feat"]) < 1e-4
    image = tiny["image"].to(DEV)
    scores = z.predict(image, return_scores=True)["pred"]
    assert scores.shape == (6, 7) and rel(scores, tiny["zs_logits"]) < 1e-4
    assert torch.equal(z.predict(image)["pred"].cpu(), tiny["zs_pred"])
    feat = z._compute_img_feat(image)
    assert row_rel(feat, tiny["zs_img_feat"]) < 1e-4
    assert torch.equal(z.predict_from_features(feat)["pred"].cpu(), tiny["zs_pred_from_features"])
    assert torch.equal(z.predict_topk_from_features(feat, 5)["pred"].cpu(), tiny["zs_top5"])
    assert z.predict(image[0])["pred"].shape == (1,)                       # [3,H,W] input is accepted (:45-46)
    assert float(z.variance_from_features(feat)["variance"]) == pytest.approx(float(tiny["zs_logits"].var()), rel=1e-3)
    zo = zs.OpenAIZeroShotClassifier(clip, tok, tiny["zs_names"])
    assert len(zo.templates) == 86
    assert row_rel(zo.prompt_feat, tiny["openai_prompt_feat"]) < 1e-4
    assert rel(zo.predict(image, return_scores=True)["pred"], tiny["openai_logits"]) < 1e-4
    assert torch.equal(zo.predict(image)["pred"].cpu(), tiny["openai_pred"])
    zd = zs.OpenAIZeroShotClassifier(clip, tok, tiny["zs_names"], domain_invariant=True)
    assert len(zd.templates) == tiny["openai_di_templates"]
    assert row_rel(zd.prompt_feat, tiny["openai_di_prompt_feat"]) < 1e-4


def test_multi_checkpoint_eval_driver(tiny, tmp_path):
    """xclip/evaluate.py: several checkpoints x datasets through ONE model instance; accuracies equal those of a
    fresh model + classifier per checkpoint (the reference loop, scripts/evaluate_domainnet_lso_openai.py:214-228)."""
    from torch.utils.data import TensorDataset
    from understanding_clip_ood_b200.xclip.evaluate import evaluate_checkpoints
    tok = FakeTokenizer(300)
    names = tiny["zs_names"]
    g = torch.Generator().manual_seed(31)
    ds = {"a": TensorDataset(torch.randn(37, 3, 64, 64, generator=g), torch.randint(0, len(names), (37,), generator=g)),
          "b": TensorDataset(torch.randint(0, 256, (20, 3, 64, 64), generator=g, dtype=torch.uint8), torch.randint(0, len(names), (20,), generator=g))}
    ckpts = []
    for i in range(3):
        torch.manual_seed(40 + i)
        m = open_clip.create_model("ViT-B-32", precision="fp32", device="cpu", **tiny["cfg"])
        path = tmp_path / f"epoch_{i}.pt"
        torch.save({"state_dict": m.state_dict(), "epoch": i}, path)
        ckpts.append(str(path))
    res = evaluate_checkpoints("ViT-B-32", ckpts, ds, {"a": names, "b": names}, tokenizer=tok, precision="fp32", batch_size=8, **tiny["cfg"])
    for i, path in enumerate(ckpts):
        m = open_clip.create_model("ViT-B-32", precision="fp32", device=DEV, pretrained=path, **tiny["cfg"]).eval()
        z = zs.OpenAIZeroShotClassifier(OpenCLIP(m), tok, names)
        for name, d in ds.items():
            img, lab = d.tensors
            pred = z.predict(img.to(DEV))["pred"].cpu()
            assert res[name][path]["num-samples"] == len(lab)
            assert res[name][path]["top1"] == pytest.approx(int((pred == lab).sum()) / len(lab), abs=1e-12)
            assert res[name][path]["top5"] >= res[name][path]["top1"]


def test_compute_scores_matches_reference_formula(tiny):
    """_compute_scores (xclip/zero_shot.py:62-67): softmax(clip.logit_scale * logits) over the class axis; the golden logits
    are the reference's, logit_scale is the wrapper's exp().clamp(0, 100) (xclip/open_clip/model.py:25-27)."""
    clip = OpenCLIP(tiny_model(tiny))
    z = zs.ZeroShotClassifier(clip, FakeTokenizer(300), tiny["zs_names"], prompt_fn=lambda c: f"a photo of a {c}.")
    feat = z._compute_img_feat(tiny["image"].to(DEV))
    scores = z._compute_scores(feat)
    scale = float(tiny["state_dict"]["logit_scale"].exp().clamp(0, 100))
    want = torch.softmax(scale * tiny["zs_logits"].double().flatten(1), dim=1).reshape_as(tiny["zs_logits"])
    assert scores.shape == tiny["zs_logits"].shape
    assert float((scores.double().cpu() - want).abs().max()) < 1e-5
    assert float((scores.sum(dim=1) - 1).abs().max()) < 1e-5
    assert torch.equal(scores.argmax(dim=1).cpu(), tiny["zs_pred"])


def test_cliploss_second_backward_is_refused():
    """The fused backward consumes its saved logits in place: a second backward through the same node must raise, not
    return gradients computed from gradients."""
    g = torch.load(GOLD / "cliploss.pt", weights_only=False)
    img = g["img"].to(DEV).requires_grad_(True)
    txt = g["txt"].to(DEV).requires_grad_(True)
    scale = torch.tensor(g["scale"], device=DEV, requires_grad=True)
    loss = open_clip.ClipLoss()(img, txt, scale)
    loss.backward(retain_graph=True)
    first = img.grad.clone()
    assert rel(first, g["w1"]["d_img"]) < 1e-4
    with pytest.raises(RuntimeError, match="already run"):
        loss.backward()
    assert torch.equal(img.grad, first)                  # nothing was accumulated by the refused call


def test_cliploss_matches_reference_world1():
    g = torch.load(GOLD / "cliploss.pt", weights_only=False)
    img = g["img"].to(DEV).requires_grad_(True)
    txt = g["txt"].to(DEV).requires_grad_(True)
    scale = torch.tensor(g["scale"], device=DEV, requires_grad=True)
    loss = open_clip.ClipLoss()(img, txt, scale)
    loss.backward()
    assert abs(float(loss) - float(g["w1"]["loss"])) / float(g["w1"]["loss"]) < 1e-3      # north_star: 1e-3 relative
    assert abs(float(loss) - float(g["w1"]["loss"])) / float(g["w1"]["loss"]) < 1e-5      # what fp32 actually achieves
    assert rel(img.grad, g["w1"]["d_img"]) < 1e-4 and rel(txt.grad, g["w1"]["d_txt"]) < 1e-4
    assert float(scale.grad) == pytest.approx(float(g["w1"]["d_scale"]), rel=1e-4)
    out = open_clip.ClipLoss()(img.detach(), txt.detach(), scale.detach(), output_dict=True)
    assert set(out) == {"contrastive_loss"} and float(out["contrastive_loss"]) == pytest.approx(float(loss), rel=1e-6)


def test_cliploss_matches_reference_two_rank_local_loss():
    """The reference ran 2 gloo ranks with local_loss + gather_with_grad.  Both ranks' kernels are run here one
    after the other on one GPU (no cross-rank waiting), the reduce-scatter of the gathered-side gradients is the sum."""
    g = torch.load(GOLD / "cliploss.pt", weights_only=False)
    img, txt = g["img"].to(DEV), g["txt"].to(DEV)
    scale = torch.tensor(g["scale"], device=DEV)
    n = img.shape[0] // 2
    res = []
    for r in range(2):
        loss, grads = ops.cliploss_fwd_bwd(img[r * n:(r + 1) * n], txt[r * n:(r + 1) * n], img, txt, scale, r)
        assert abs(float(loss) - float(g["w2"][r]["loss"])) / float(g["w2"][r]["loss"]) < 1e-5
        res.append(grads)
    all_img = res[0][2] + res[1][2]
    all_txt = res[0][3] + res[1][3]
    for r in range(2):
        assert rel(res[r][0] + all_img[r * n:(r + 1) * n], g["w2"][r]["d_img"]) < 1e-4
        assert rel(res[r][1] + all_txt[r * n:(r + 1) * n], g["w2"][r]["d_txt"]) < 1e-4
        assert float(res[r][4]) == pytest.approx(float(g["w2"][r]["d_scale"]), rel=1e-4)


def test_cliploss_packed_distributed_form_matches_reference_two_rank():
    """Packed (img | txt) form used by the NCCL path: same golden two-rank run; the reduce-scatter is emulated as the sum of
    both ranks' d_gathered, whose rows [r*n, r*n+n) must equal rank r's feature gradients (local-row terms folded in)."""
    g = torch.load(GOLD / "cliploss.pt", weights_only=False)
    img, txt = g["img"].to(DEV), g["txt"].to(DEV)
    scale = torch.tensor(g["scale"], device=DEV)
    n, D = img.shape[0] // 2, img.shape[1]
    gathered = torch.cat([img, txt], dim=1).contiguous()
    d_sum = torch.zeros_like(gathered)
    d_scales = []
    for r in range(2):
        loss, ws = ops.cliploss_packed_forward(gathered, scale, r, n)
        assert abs(float(loss) - float(g["w2"][r]["loss"])) / float(g["w2"][r]["loss"]) < 1e-5
        d_g, d_s = ops.cliploss_packed_backward(gathered, scale, r, n, ws, None)
        d_sum += d_g
        d_scales.append(float(d_s))
    for r in range(2):
        assert rel(d_sum[r * n:(r + 1) * n, :D], g["w2"][r]["d_img"]) < 1e-4
        assert rel(d_sum[r * n:(r + 1) * n, D:], g["w2"][r]["d_txt"]) < 1e-4
        assert d_scales[r] == pytest.approx(float(g["w2"][r]["d_scale"]), rel=1e-4)
    # upstream gradient is read from device memory
    loss, ws = ops.cliploss_packed_forward(gathered, scale, 0, n)
    d_g2, d_s2 = ops.cliploss_packed_backward(gathered, scale, 0, n, ws, torch.tensor(0.5, device=DEV))
    loss, ws = ops.cliploss_packed_forward(gathered, scale, 0, n)
    d_g1, d_s1 = ops.cliploss_packed_backward(gathered, scale, 0, n, ws, None)
    assert rel(d_g2, 0.5 * d_g1) < 1e-6 and float(d_s2) == pytest.approx(0.5 * float(d_s1), rel=1e-6)


# ------------------------------------------------------------------ BASELINE config 1 at full size ---
def _domainnet_tokens():
    z = np.load(GOLD / "domainnet_prompts.npz")
    tok = torch.zeros((z["tokens"].shape[0], int(z["context_length"])), dtype=torch.long)
    tok[:, : z["tokens"].shape[1]] = torch.from_numpy(z["tokens"].astype(np.int64))
    return tok, int(z["classes"]), int(z["templates"])


@pytest.fixture(scope="module")
def vitb32_fp32():
    torch.manual_seed(0)                      # reproduces the reference's seed-0 init bit-exactly (tests/test_host_cpu.py)
    return open_clip.create_model("ViT-B-32", precision="fp32", device="cpu").to(DEV).eval()


def test_vit_b_32_fp32_zero_shot_matches_reference_config1(vitb32_fp32):
    path = GOLD / "vitb32_seed0.pt"
    if not path.exists():
        pytest.skip("vitb32_seed0.pt not generated")
    g = torch.load(path, weights_only=False)
    m = vitb32_fp32
    m.truncate_text_at_eot = True
    image = torch.randn(64, 3, 224, 224, generator=torch.Generator().manual_seed(1)).to(DEV)
    tokens, C, T = _domainnet_tokens()
    z = zs.OpenAIZeroShotClassifier.from_tokens(OpenCLIP(m), tokens, C, T)
    assert row_rel(z.prompt_feat, g["prompt_feat"]) < 1e-4
    first = m.encode_text(tokens[:T].to(DEV))
    assert row_rel(first, g["text_features_first_class"]) < 1e-4
    feat = z._compute_img_feat(image)
    assert row_rel(feat, g["image_features_normalized"]) < 1e-4
    logits = z.predict_from_features(feat, return_scores=True)["pred"]
    assert float((logits.cpu() - g["logits"]).abs().max()) < 2e-6
    pred = z.predict_from_features(feat)["pred"].cpu()
    top5 = z.predict_topk_from_features(feat, 5)["pred"].cpu()
    ref_sorted = torch.sort(g["logits"], dim=1, descending=True).values
    # index agreement; a disagreement is only tolerated where the reference's own margin is below the fp32 noise floor
    for i in range(64):
        if pred[i] != g["pred"][i]:
            assert float(ref_sorted[i, 0] - ref_sorted[i, 1]) < 1e-6, f"top-1 differs on sample {i} with a real margin"
        if set(top5[i].tolist()) != set(g["top5"][i].tolist()):
            assert float(ref_sorted[i, 4] - ref_sorted[i, 5]) < 1e-6, f"top-5 differs on sample {i} with a real margin"
    assert float((pred == g["pred"]).float().mean()) >= 0.98


def test_vit_b_32_bf16_against_reference_bf16_and_fp32(vitb32_fp32):
    p32, p16 = GOLD / "vitb32_seed0.pt", GOLD / "vitb32_seed0_bf16.pt"
    if not (p32.exists() and p16.exists()):
        pytest.skip("full-size golden fixtures not generated")
    g32, g16 = torch.load(p32, weights_only=False), torch.load(p16, weights_only=False)
    torch.manual_seed(0)
    m = open_clip.create_model("ViT-B-32", precision="bf16", device="cpu").to(DEV).eval()
    image = torch.randn(64, 3, 224, 224, generator=torch.Generator().manual_seed(1)).bfloat16().to(DEV)
    feat = m.encode_image(image, normalize=True)
    assert row_rel(feat, g32["image_features_normalized"]) < 2e-2         # north_star: bf16 embeddings within 2e-2
    assert row_rel(feat, g16["image_features_normalized"]) < 2e-2         # vs the reference's own bf16 path
    # prediction agreement is reported against the reference's own bf16<->fp32 floor (SURVEY.md §7 hard part 1)
    logits, idx, _ = ops.zeroshot(feat, g32["prompt_feat"].bfloat16().to(DEV), 5, normalize_img=False)
    ours_vs_fp32 = float((idx[:, 0].cpu() == g32["pred"]).float().mean())
    ref16_vs_fp32 = float((g16["pred"] == g32["pred"]).float().mean())
    ours_vs_ref16 = float((idx[:, 0].cpu() == g16["pred"]).float().mean())
    print(f"top-1 agreement: ours-bf16 vs ref-fp32 {ours_vs_fp32:.3f}; ref-bf16 vs ref-fp32 {ref16_vs_fp32:.3f}; "
          f"ours-bf16 vs ref-bf16 {ours_vs_ref16:.3f}")
    assert ours_vs_fp32 >= ref16_vs_fp32 - 0.1


def test_vit_b_32_bf16_prediction_parity_b1024(record_property):
    """north_star's prediction gate at BASELINE config 2's batch, in numbers: ours-bf16 must agree with the fp32 reference at
    least as often as the reference's own bf16 run does (minus 1 point of sampling slack), and every disagreement must lie
    inside the reference's own bf16 noise band (margin-aware agreement >= 99.9 %)."""
    if not (GOLD / "vitb32_seed0_b1024.pt").exists():
        pytest.skip("vitb32_seed0_b1024.pt not generated")
    torch.manual_seed(0)
    m = open_clip.create_model("ViT-B-32", precision="bf16", device="cpu").to(DEV).eval()
    r = prediction_parity_b1024(m)
    for k, v in r.items():
        record_property(k, v)
    print("bf16 prediction parity at B=1024: " + json.dumps(r))
    (Path(__file__).resolve().parent.parent / "gpurun_out").mkdir(exist_ok=True)
    (Path(__file__).resolve().parent.parent / "gpurun_out" / "parity_b1024.json").write_text(json.dumps(r, indent=1))
    assert r["embedding_rel_l2_vs_ref_fp32"] < 2e-2 and r["embedding_rel_l2_vs_ref_bf16"] < 2e-2
    assert r["ours_max_logit_err_vs_ref_fp32"] <= 2.0 * r["ref_bf16_logit_noise_band"]
    assert r["top1_ours_vs_ref_fp32"] >= r["top1_ref_bf16_vs_ref_fp32"] - 0.01
    assert r["top5_ours_vs_ref_fp32"] >= r["top5_ref_bf16_vs_ref_fp32"] - 0.02
    assert r["top1_margin_aware_vs_ref_fp32"] >= 0.999
    assert r["top5_margin_aware_vs_ref_fp32"] >= 0.999


# ------------------------------------------------------------------ (2) oracle on seeded inputs -------
@pytest.mark.parametrize("name,B,T", [("ViT-B-32", 8, 12), ("ViT-B-16", 2, 4)])
def test_towers_match_oracle_fp32(name, B, T):
    torch.manual_seed(3)
    m = open_clip.create_model(name, precision="fp32", device="cpu")
    sd = {k: v.clone() for k, v in m.state_dict().items()}
    m = m.to(DEV).eval()
    image = torch.randn(B, 3, 224, 224, generator=torch.Generator().manual_seed(4))
    tokens, _, _ = _domainnet_tokens()
    text = tokens[torch.randperm(tokens.shape[0], generator=torch.Generator().manual_seed(5))[:T]]
    assert row_rel(m.encode_image(image.to(DEV)), O.vit_forward(sd, image)) < 1e-4
    assert row_rel(m.encode_text(text.to(DEV)), O.text_forward(sd, text)) < 1e-4


def test_vit_l_14_bf16_matches_oracle():
    torch.manual_seed(6)
    m = open_clip.create_model("ViT-L-14", precision="bf16", device="cpu")
    sd = {k: v.float() for k, v in m.state_dict().items()}
    m = m.to(DEV).eval()
    image = torch.randn(2, 3, 224, 224, generator=torch.Generator().manual_seed(7)).bfloat16()
    assert row_rel(m.encode_image(image.to(DEV)), O.vit_forward(sd, image.float())) < 2e-2      # K=588 padded patch GEMM
    tokens, _, _ = _domainnet_tokens()
    assert row_rel(m.encode_text(tokens[:3].to(DEV)), O.text_forward(sd, tokens[:3])) < 2e-2


@pytest.mark.parametrize("precision,dtype", [("bf16", torch.bfloat16), ("fp16", torch.float16)])
def test_vit_b_16_low_precision_matches_oracle(precision, dtype):
    """ViT-B/16 in the 16-bit modes: L = 197 runs the tcgen05 attention kernel (attention_tc_kernel), which the fp32 case
    above never reaches."""
    torch.manual_seed(8)
    m = open_clip.create_model("ViT-B-16", precision=precision, device="cpu")
    sd = {k: v.float() for k, v in m.state_dict().items()}
    m = m.to(DEV).eval()
    image = torch.randn(3, 3, 224, 224, generator=torch.Generator().manual_seed(9)).to(dtype)
    assert row_rel(m.encode_image(image.to(DEV)), O.vit_forward(sd, image.float())) < 2e-2
    tokens, _, _ = _domainnet_tokens()
    assert row_rel(m.encode_text(tokens[5:9].to(DEV)), O.text_forward(sd, tokens[5:9])) < 2e-2


def test_vit_b_32_fp16_full_size_matches_oracle_sample():
    """fp16 is the precision the reference's evaluation scripts default to (xclip/open_clip/model.py:35).  Full-size batch
    (1024 images) on the GPU; the oracle checks a 12-image sample of it, the rest through batch independence."""
    torch.manual_seed(0)
    m = open_clip.create_model("ViT-B-32", precision="fp16", device="cpu")
    sd = {k: v.float() for k, v in m.state_dict().items()}
    m = m.to(DEV).eval()
    g = torch.Generator().manual_seed(12)
    image = torch.randn(1024, 3, 224, 224, generator=g).half()
    feat = m.encode_image(image.to(DEV))
    assert torch.isfinite(feat.float()).all()
    pick = torch.tensor([0, 1, 100, 101, 511, 512, 513, 777, 900, 1021, 1022, 1023])
    assert row_rel(feat[pick.to(DEV)], O.vit_forward(sd, image[pick].float())) < 2e-2
    assert row_rel(m.encode_image(image[pick].to(DEV)), feat[pick.to(DEV)]) < 1e-2      # batch independence up to summation order


@pytest.mark.parametrize("name", open_clip.list_models())
def test_every_registered_config_runs_in_fp32(name):
    """create_model's default precision on every registered architecture (P = 14 gives K = 588 for the patch
    GEMM, which needs padding to 16-byte rows in fp32 as well)."""
    torch.manual_seed(1)
    m = open_clip.create_model(name, precision="fp32", device="cpu")
    sd = {k: v.clone() for k, v in m.state_dict().items()}
    m = m.to(DEV).eval()
    S = m.visual.image_size[0]
    image = torch.randn(1, 3, S, S, generator=torch.Generator().manual_seed(2))
    if name == "ViT-L-14-336":
        # 577 tokens: the fp32 parity kernel keeps K and V of a whole (image, head) in shared memory (<= 435 tokens) and says
        # so; the 16-bit modes (the ones that configuration is run in) take the streaming attention kernel
        from understanding_clip_ood_b200 import _lib as L
        with pytest.raises(L.B200ClipError, match="too long for the shared-memory path"):
            m.encode_image(image.to(DEV))
        mb = open_clip.create_model(name, precision="bf16", device="cpu")
        mb.load_state_dict(sd)
        got = mb.to(DEV).eval().encode_image(image.bfloat16().to(DEV))
        assert row_rel(got, O.vit_forward(sd, image)) < 2e-2
        return
    got = m.encode_image(image.to(DEV))
    assert got.shape == (1, m.visual.output_dim) and torch.isfinite(got).all()
    if name in ("ViT-L-14", "ViT-B-32-256", "ViT-L-14-336"):         # the geometries no other test covers in fp32
        assert row_rel(got, O.vit_forward(sd, image, quick_gelu="quickgelu" in name)) < 1e-4


def test_graph_replay_is_independent_of_input_and_output_buffers():
    """The captured graph covers only the workspace-to-workspace body (B200CLIP_STAGE_BODY): a loop that hands over a FRESH
    tensor per batch — what a DataLoader loop does, evaluate_domainnet_lso_openai.py:18-36 — replays it, results equal the
    un-graphed path bit for bit, and every call returns its own output tensor (no aliasing of a static buffer)."""
    from understanding_clip_ood_b200 import _lib as L
    torch.manual_seed(4)
    m = open_clip.create_model("ViT-B-32", precision="bf16", device=DEV,
                               vision_cfg={"image_size": 224, "layers": 3, "width": 768, "patch_size": 32}).eval()
    g = torch.Generator(device=DEV).manual_seed(5)
    batches = [torch.randn(32, 3, 224, 224, device=DEV, generator=g).bfloat16() for _ in range(4)]
    m.visual.use_cuda_graphs = False
    want = [m.encode_image(b, normalize=True).clone() for b in batches]
    m.visual.use_cuda_graphs = True
    outs = [m.encode_image(b.clone(), normalize=True) for b in batches]          # fresh input buffer every call
    assert len(m.visual._engine.graphs) == 1                                      # captured on the 2nd call, replayed after
    assert len({o.data_ptr() for o in outs}) == len(outs)
    for o, w in zip(outs, want):
        assert torch.equal(o, w)
    n0 = L.launch_count()
    m.encode_image(batches[0].clone())
    assert L.launch_count() - n0 > 20                                             # replayed kernels are counted
    # text tower: same mechanism, keyed on (batch, sequence length)
    tokens, _, _ = _domainnet_tokens()
    m.use_cuda_graphs = False
    tw = [m.encode_text(tokens[i * 16:(i + 1) * 16].to(DEV)).clone() for i in range(3)]
    m.use_cuda_graphs = True
    tg = [m.encode_text(tokens[i * 16:(i + 1) * 16].to(DEV)) for i in range(3)]
    assert len(m._text_engine.graphs) == 1
    for o, w in zip(tg, tw):
        assert torch.equal(o, w)


def test_engine_rebuilds_when_public_switches_change():
    """fold_layernorm / quick_gelu are public attributes baked into the cached weight structs and graphs."""
    torch.manual_seed(3)
    m = open_clip.create_model("ViT-B-32", precision="bf16", device=DEV,
                               vision_cfg={"image_size": 224, "layers": 2, "width": 768, "patch_size": 32}).eval()
    image = torch.randn(4, 3, 224, 224, device=DEV).bfloat16()
    a = m.encode_image(image).clone()
    a2 = m.encode_image(image).clone()            # second call: graph replay
    assert torch.equal(a, a2)
    m.visual.fold_layernorm = False
    b = m.encode_image(image).clone()
    assert not torch.equal(a, b) and row_rel(b, a) < 2e-2      # a different (unfolded) path ran, same function
    m.visual.quick_gelu = True
    c = m.encode_image(image).clone()
    assert row_rel(c, b) > 1e-3                    # a different activation ran
    m.visual.quick_gelu = False
    m.visual.fold_layernorm = True
    assert torch.equal(m.encode_image(image), a)


# ------------------------------------------------------------------ (3) properties at full size -------
def test_full_size_bf16_properties():
    """BASELINE config 2 shapes: ViT-B-32 bf16, 1024 images, 345 classes."""
    torch.manual_seed(0)
    m = open_clip.create_model("ViT-B-32", precision="bf16", device=DEV).eval()
    g = torch.Generator(device=DEV).manual_seed(1)
    image = torch.randn(1024, 3, 224, 224, device=DEV, generator=g).bfloat16()
    feat = m.encode_image(image)
    assert torch.isfinite(feat.float()).all()
    # determinism: the same batch gives the same bits (stream-K partial sums are added in a fixed order, no atomics)
    assert torch.equal(m.encode_image(image), feat)
    # batch independence: an image's embedding does not depend on what else is in the batch.  Up to fp32 summation order only:
    # the stream-K split of a tile's K range depends on the tile grid, i.e. on the batch size and the row's position
    # (well inside one bf16 ulp of the 12-layer residual stream's noise)
    part = torch.cat([m.encode_image(image[:100]), m.encode_image(image[100:612]), m.encode_image(image[612:])])
    assert row_rel(part, feat) < 1e-2
    # permutation equivariance (same caveat)
    perm = torch.randperm(1024, device=DEV, generator=g)
    assert row_rel(m.encode_image(image[perm]), feat[perm]) < 1e-2
    # normalize is idempotent up to one bf16 ulp and produces unit rows
    n1 = ops.normalize(feat)
    assert float((n1.float().norm(dim=-1) - 1).abs().max()) < 1e-2
    assert float((ops.normalize(n1).float() - n1.float()).abs().max()) <= 2 ** -8
    prompt = ops.normalize(torch.randn(345, 512, device=DEV, generator=g).bfloat16())
    logits, idx, val = ops.zeroshot(feat, prompt, 5)
    assert (val[:, :-1] >= val[:, 1:]).all()                                  # sorted descending
    assert torch.equal(val, torch.gather(logits, 1, idx))                     # indices point at the values
    assert torch.equal(idx[:, 0], logits.argmax(dim=1))                       # checksum-of-argmax against the written logits
    assert float(logits.abs().max()) <= 1.0 + 2 ** -7                         # cosine range
    # fused normalize flag == explicit normalize
    _, idx2, _ = ops.zeroshot(n1, prompt, 5, normalize_img=False)
    assert torch.equal(idx, idx2)


def test_cliploss_full_size_properties():
    """BASELINE config 4 shape: n=256 local rows, world 8 -> N=2048, D=512."""
    g = torch.Generator(device=DEV).manual_seed(9)
    n, world, D = 256, 8, 512
    all_img = ops.normalize(torch.randn(n * world, D, device=DEV, generator=g))
    all_txt = ops.normalize(torch.randn(n * world, D, device=DEV, generator=g))
    scale = torch.tensor(1 / 0.07, device=DEV)
    losses, d_scale = [], []
    sum_d_all_img = torch.zeros_like(all_img)
    for r in range(world):
        loss, grads = ops.cliploss_fwd_bwd(all_img[r * n:(r + 1) * n], all_txt[r * n:(r + 1) * n], all_img, all_txt, scale, r)
        losses.append(float(loss))
        d_scale.append(float(grads[4]))
        sum_d_all_img += grads[2]
        # softmax gradients sum to zero over every logit row => each rank's feature gradients are orthogonal to ... the
        # cheap invariant: total gradient mass through the local image rows equals minus that through the gathered text rows
        assert torch.isfinite(grads[0]).all()
    # mean of the local losses == the world_size-1 loss over all gathered rows (the survey's probe identity)
    full, _ = ops.cliploss_fwd_bwd(all_img, all_txt, all_img, all_txt, scale, 0, want_grad=False)
    assert abs(sum(losses) / world - float(full)) / float(full) < 1e-5
    ref = float(O.clip_loss_local(all_img[:n], all_txt[:n], all_img, all_txt, float(scale), 0))
    assert abs(losses[0] - ref) / ref < 1e-3


# ------------------------------------------------------------------ ModifiedResNet tower
@pytest.fixture(scope="module")
def rn_gold():
    return torch.load(GOLD / "rn_seed0.pt", weights_only=False)


def rn_model(gold, which, precision):
    """Seed-0 weights (bit-identical to the reference's, tests/test_host_cpu.py), the fixture's BatchNorm recipe + statistics."""
    torch.manual_seed(gold["seed_weights"])
    kw = gold["tiny_cfg"] if which == "tiny" else {}
    m = open_clip.create_model("RN50", precision=precision, device="cpu", **kw)
    O.randomize_batchnorm_(m.visual, gold["seed_bn"])
    missing, unexpected = m.load_state_dict({"visual." + k: v for k, v in gold[which + "_bn"].items()}, strict=False)
    assert not unexpected
    return m.to(DEV).eval()


def centered_rel(a, b):
    """relative error of the image-dependent part (features minus their batch mean)"""
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float(((a - a.mean(0)) - (b - b.mean(0))).norm() / (b - b.mean(0)).norm())


@pytest.mark.parametrize("which,size,n", [("tiny", 64, 6), ("rn50", 224, 8)])
def test_resnet_fp32_matches_reference_golden(rn_gold, which, size, n):
    m = rn_model(rn_gold, which, "fp32")
    img = O.test_images(n, size, rn_gold["seed_images"])
    want = rn_gold[which + "_fp32"]
    got = m.encode_image(img.to(DEV))
    assert row_rel(got, want) < 1e-4 and centered_rel(got, want) < 1e-4
    got2 = m.encode_image(img.to(DEV).clone())        # second sighting of the shape: CUDA-graph replay, fresh input tensor
    assert torch.equal(got2, got)
    assert row_rel(O.resnet_forward(m.state_dict(), img), want) < 1e-4      # the oracle on the same weights
    nrm = m.encode_image(img.to(DEV), normalize=True)
    assert row_rel(nrm, torch.nn.functional.normalize(want, dim=-1)) < 1e-4
    if which == "tiny":
        from oracle.make_golden_rn import tiny_text
        assert row_rel(m.encode_text(tiny_text().to(DEV)), rn_gold["tiny_text"]) < 1e-4


@pytest.mark.parametrize("precision,dtype", [("bf16", torch.bfloat16), ("fp16", torch.float16)])
def test_resnet_16bit_matches_reference(rn_gold, precision, dtype):
    """RN50 in the 16-bit modes vs the reference's fp32 golden (gate 2e-2, north_star) — the reference's own bf16 run is the
    yardstick for what 16-bit arithmetic costs on this tower (fixture rn50_bf16)."""
    m = rn_model(rn_gold, "rn50", precision)
    img = O.test_images(8, 224, rn_gold["seed_images"])
    want = rn_gold["rn50_fp32"]
    got = m.encode_image(img.to(dtype).to(DEV))
    assert got.dtype == dtype
    ref16 = row_rel(rn_gold["rn50_bf16"].float(), want)
    ours = row_rel(got.float(), want)
    print(f"RN50 {precision}: ours vs ref-fp32 {ours:.3e}, ref-bf16 vs ref-fp32 {ref16:.3e}, centered {centered_rel(got.float(), want):.3e}")
    assert ours < 2e-2
    assert centered_rel(got.float(), want) < 4e-2


def test_resnet_large_batch_is_batch_invariant_and_classifies(rn_gold):
    """Full-size property checks (no oracle at this size): a 256-image bf16 batch reproduces the rows of its 8-image prefix run
    alone up to the batch-dependent fp32 summation order of the stream-K GEMMs, and the zero-shot stage runs on the tower."""
    m = rn_model(rn_gold, "rn50", "bf16")
    img = O.test_images(256, 224, 11).bfloat16().to(DEV)
    big = m.encode_image(img, normalize=True)
    small = m.encode_image(img[:8].contiguous(), normalize=True)
    assert torch.isfinite(big.float()).all()
    assert row_rel(big[:8].float(), small.float()) < 1e-2
    clf = zs.ZeroShotClassifier(OpenCLIP(m), FakeTokenizer(300), [f"class {i}" for i in range(7)])
    out = clf.predict(img[:32].contiguous())
    assert out["pred"].shape == (32,)


def test_resnet_rejects_what_is_not_on_the_path(rn_gold):
    m = rn_model(rn_gold, "tiny", "fp32")
    img = O.test_images(2, 64, 1).to(DEV)
    m.train()
    with pytest.raises(RuntimeError, match="BatchNorm"):
        m.encode_image(img)
    m.visual.lock(freeze_bn_stats=True)               # frozen statistics + no gradients: the eval path again
    assert m.encode_image(img).shape == (2, 128)
    m.eval()
    with pytest.raises(RuntimeError, match="expected images"):
        m.encode_image(img[:, :, :32, :32])
