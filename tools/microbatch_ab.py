"""A/B: one 1024-image ViT-B-32 forward vs the same batch as sequential micro-batches (L2 residency of the activations).
Alternating blocks of ~1 s each so that the power-capped clock state affects both arms alike."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from understanding_clip_ood_b200 import open_clip  # noqa: E402

torch.manual_seed(0)
model = open_clip.create_model("ViT-B-32", precision="bf16", device="cuda").eval()
g = torch.Generator(device="cuda").manual_seed(1)
image = torch.randn(1024, 3, 224, 224, device="cuda", generator=g).bfloat16()
splits = [int(v) for v in sys.argv[1:]] or [1024, 512, 342, 256]


def run(mb):
    if mb >= 1024:
        return model.encode_image(image)
    return [model.encode_image(image[i:i + mb]) for i in range(0, 1024, mb)]


for mb in splits:
    for _ in range(3):
        run(mb)
torch.cuda.synchronize()
res = {mb: [] for mb in splits}
for rep in range(4):
    for mb in splits:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(60):
            run(mb)
        e1.record()
        torch.cuda.synchronize()
        res[mb].append(e0.elapsed_time(e1) / 60)
for mb in splits:
    ms = res[mb]
    print(f"micro-batch {mb:5d}: " + " ".join(f"{v:6.2f}" for v in ms) + f" ms per 1024 images -> {1024 / (sum(ms) / len(ms)) * 1e3:8.0f} img/s", flush=True)
