"""Per-step device times of the benchmark step (CUDA events around every step) to see clock ramp / power-cap behaviour."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from understanding_clip_ood_b200 import open_clip, ops  # noqa: E402

batch = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 60
torch.manual_seed(0)
model = open_clip.create_model("ViT-B-32", precision="bf16", device="cuda").eval()
g = torch.Generator(device="cuda").manual_seed(1)
image = torch.randn(batch, 3, 224, 224, device="cuda", generator=g).bfloat16()
prompt = ops.normalize(torch.randn(345, 512, device="cuda", generator=g).bfloat16())
ev = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
import time
t0 = time.perf_counter()
ev[0].record()
for i in range(steps):
    feat = model.encode_image(image, normalize=True)
    ops.zeroshot(feat, prompt, 5, normalize_img=False, want_logits=False)
    ev[i + 1].record()
host = time.perf_counter() - t0
torch.cuda.synchronize()
ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(steps)]
print("host enqueue ms/step", host / steps * 1e3)
print("per-step ms:", " ".join(f"{m:.2f}" for m in ms))
print("mean of last half", sum(ms[steps // 2:]) / (steps - steps // 2))
