"""Full-size prediction-parity fixture (BASELINE config 2's batch): RUN THE UNMODIFIED REFERENCE (/root/reference, CPU) on
1024 seeded images in fp32 AND in bf16 — TEST INFRASTRUCTURE ONLY, never imported by the product.

    python -m oracle.make_golden_b1024

Writes tests/golden/vitb32_seed0_b1024.pt:
  image_features_fp32   [1024, 512] fp32   F.normalize(reference ViT-B-32 fp32 encode_image), seed-0 weights, randn images seed 1
  logits_fp32           [1024, 345] fp32   against the fp32 class-prompt features of vitb32_seed0.pt (OpenAIZeroShotClassifier)
  image_features_bf16   [1024, 512] bf16   the reference instantiated with precision='bf16' on the bf16-cast images
  logits_bf16           [1024, 345] bf16   against the bf16-cast prompt features (xclip/zero_shot.py:54-60 on bf16 operands)
These give, at B = 1024: the reference's OWN bf16 <-> fp32 top-1 / top-5 agreement floor and logit noise band, and the anchor
for "ours-bf16 vs reference-bf16" (tests/test_parity_gpu.py, bench.py's `parity` block).  The weights are not stored: this
repo's create_model reproduces the reference's seed-0 init bit-exactly (tests/test_host_cpu.py).
"""
from __future__ import annotations

import os
import sys
import time
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from oracle import ref_loader  # noqa: E402

GOLD = ROOT / "tests" / "golden"
B = 1024


def main():
    open_clip, zs, xo = ref_loader.load()
    torch.set_num_threads(os.cpu_count())
    base = torch.load(GOLD / "vitb32_seed0.pt", weights_only=False)
    prompt = base["prompt_feat"].float()
    image = torch.randn(B, 3, 224, 224, generator=torch.Generator().manual_seed(1))
    out = {"seed_weights": 0, "seed_images": 1, "batch": B}
    with torch.no_grad():
        torch.manual_seed(0)
        ref = open_clip.create_model("ViT-B-32", precision="fp32").eval()
        t0 = time.time()
        f32 = torch.cat([torch.nn.functional.normalize(ref.encode_image(image[i:i + 64]), dim=-1) for i in range(0, B, 64)])
        print(f"fp32 encode_image({B}) {time.time() - t0:.0f}s", flush=True)
        out["image_features_fp32"] = f32.clone()
        out["logits_fp32"] = torch.tensordot(f32, prompt.movedim(-1, 0), dims=1).clone()
        del ref
        torch.manual_seed(0)
        refb = open_clip.create_model("ViT-B-32", precision="bf16").eval()
        t0 = time.time()
        fb = torch.cat([torch.nn.functional.normalize(refb.encode_image(image[i:i + 64].bfloat16()), dim=-1) for i in range(0, B, 64)])
        print(f"bf16 encode_image({B}) {time.time() - t0:.0f}s", flush=True)
        out["image_features_bf16"] = fb.clone()
        out["logits_bf16"] = torch.tensordot(fb, prompt.bfloat16().movedim(-1, 0), dims=1).clone()
    l32, l16 = out["logits_fp32"], out["logits_bf16"].float()
    agree1 = (l32.argmax(1) == l16.argmax(1)).float().mean().item()
    t5a, t5b = l32.topk(5, 1)[1].sort(1)[0], l16.topk(5, 1)[1].sort(1)[0]
    agree5 = (t5a == t5b).all(1).float().mean().item()
    print(f"reference bf16 vs fp32 at B={B}: top-1 {agree1:.4f}, top-5 set {agree5:.4f}, max |dlogit| {(l32 - l16).abs().max():.2e}")
    torch.save(out, GOLD / "vitb32_seed0_b1024.pt")


if __name__ == "__main__":
    main()
