"""ModifiedResNet parity fixtures: RUN THE UNMODIFIED REFERENCE (/root/reference, CPU) — TEST INFRASTRUCTURE ONLY, never
imported by the product.

    python -m oracle.make_golden_rn

Writes tests/golden/rn_seed0.pt:
  rn50_fp32   [8, 1024] fp32   reference RN50 (precision='fp32', eval) encode_image on 8 seeded images (randn, seed 1)
  rn50_bf16   [8, 1024] bf16   reference RN50 instantiated with precision='bf16' on the bf16-cast images
  tiny_fp32   [6, 128]  fp32   reference RN50 architecture overridden to layers (1,2,1,1), width 32, 64 x 64 images, embed 128
  tiny_text   [4, 128]  fp32   its encode_text on 4 seeded prompts (text width 64, 1 head, 2 layers, vocab 100)
  rn50_bn / tiny_bn  {name: tensor}  the BatchNorm running statistics the outputs were computed with
Weights: torch.manual_seed(0) before create_model (this repo's create_model reproduces them bit-exactly,
tests/test_host_cpu.py), then oracle.clip_oracle.randomize_batchnorm_(visual, 5) — at init bn3.weight is zero and the
bottlenecks would be identities — and the running statistics CALIBRATED by one training-mode forward of the reference tower
over 8 other seeded images (momentum None = plain batch statistics): with the default statistics (0, 1) the random-init tower
saturates and its output is the same vector for every image to 2e-5, which would leave a parity check blind to everything
that depends on the input.  Images: oracle.clip_oracle.test_images.  The convolution / linear weights are not stored.
"""
from __future__ import annotations

import os
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from oracle import clip_oracle as O  # noqa: E402
from oracle import ref_loader  # noqa: E402

GOLD = ROOT / "tests" / "golden"
TINY = dict(embed_dim=128, vision_cfg={"image_size": 64, "layers": [1, 2, 1, 1], "width": 32, "patch_size": None},
            text_cfg={"context_length": 77, "vocab_size": 100, "width": 64, "heads": 1, "layers": 2})


def tiny_text():
    g = torch.Generator().manual_seed(3)
    text = torch.zeros(4, 77, dtype=torch.long)
    for i in range(4):
        n = 3 + 2 * i
        text[i, 0], text[i, n + 1] = 98, 99
        text[i, 1:n + 1] = torch.randint(1, 98, (n,), generator=g)
    return text


def calibrate(visual, images):
    """One training-mode forward with momentum None: running_mean / running_var := the batch statistics."""
    bns = [m for m in visual.modules() if isinstance(m, torch.nn.BatchNorm2d)]
    for m in bns:
        m.reset_running_stats()
        m.momentum = None
    visual.train()
    visual(images)
    visual.eval()
    for m in bns:
        m.momentum = 0.1
    return {k: v.clone() for k, v in visual.state_dict().items() if k.endswith(("running_mean", "running_var"))}


def main():
    open_clip, _, _ = ref_loader.load()
    torch.set_num_threads(os.cpu_count())
    out = {"seed_weights": 0, "seed_images": 1, "seed_bn": 5, "tiny_cfg": TINY}
    image = O.test_images(8, 224, 1)
    with torch.no_grad():
        torch.manual_seed(0)
        ref = open_clip.create_model("RN50", precision="fp32").eval()
        O.randomize_batchnorm_(ref.visual, 5)
        out["rn50_bn"] = calibrate(ref.visual, O.test_images(8, 224, 7))
        out["rn50_fp32"] = ref.encode_image(image).clone()
        mine = O.resnet_forward(ref.state_dict(), image)
        print("oracle vs reference (RN50 fp32):", float(((mine - out["rn50_fp32"]).norm(dim=-1) / out["rn50_fp32"].norm(dim=-1)).max()))
        f = out["rn50_fp32"]
        print("embedding norm", float(f.norm(dim=-1).mean()), "norm of the image-dependent part", float((f - f.mean(0)).norm(dim=-1).mean()))
        torch.manual_seed(0)
        refb = open_clip.create_model("RN50", precision="bf16").eval()
        O.randomize_batchnorm_(refb.visual, 5)
        refb.load_state_dict({"visual." + k: v for k, v in out["rn50_bn"].items()}, strict=False)
        out["rn50_bf16"] = refb.encode_image(image.bfloat16()).clone()
        d = out["rn50_bf16"].float() - out["rn50_fp32"]
        print("reference bf16 vs fp32 rel-L2:", float((d.norm(dim=-1) / out["rn50_fp32"].norm(dim=-1)).max()))
        torch.manual_seed(0)
        tiny = open_clip.create_model("RN50", precision="fp32", **TINY).eval()
        O.randomize_batchnorm_(tiny.visual, 5)
        out["tiny_bn"] = calibrate(tiny.visual, O.test_images(8, 64, 7))
        timg = O.test_images(6, 64, 1)
        out["tiny_fp32"] = tiny.encode_image(timg).clone()
        out["tiny_text"] = tiny.encode_text(tiny_text()).clone()
        mine = O.resnet_forward(tiny.state_dict(), timg)
        print("oracle vs reference (tiny fp32):", float(((mine - out["tiny_fp32"]).norm(dim=-1) / out["tiny_fp32"].norm(dim=-1)).max()))
    torch.save(out, GOLD / "rn_seed0.pt")
    print("wrote", GOLD / "rn_seed0.pt")


if __name__ == "__main__":
    main()
