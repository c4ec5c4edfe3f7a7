"""Contrastive training step of BASELINE config 4 on one GPU: ViT-B-32, 128 image / text pairs per step, ClipLoss, tower backward,
fused AdamW (the reference: training/train.py:115-183 with --grad-checkpointing).  python tools/train_bench.py [precision] [batch]"""
import sys
import time
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from understanding_clip_ood_b200 import _lib as L, open_clip  # noqa: E402

precision = sys.argv[1] if len(sys.argv) > 1 else "amp_bf16"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 128
torch.manual_seed(0)
model = open_clip.create_model("ViT-B-32", precision=precision, device="cuda").train()
opt = open_clip.AdamW(model.parameters(), lr=5e-4, betas=(0.9, 0.98), eps=1e-6, weight_decay=0.2)
loss_fn = open_clip.ClipLoss()
g = torch.Generator(device="cuda").manual_seed(1)
in_dtype = torch.float32 if precision.startswith("amp") or precision == "fp32" else (torch.bfloat16 if "bf16" in precision else torch.float16)
image = torch.randn(B, 3, 224, 224, device="cuda", generator=g).to(in_dtype)
text = torch.zeros(B, 77, dtype=torch.long, device="cuda")
text[:, 0] = 49406
text[:, 1:9] = torch.randint(1000, 40000, (B, 8), device="cuda", generator=g)
text[:, 9] = 49407


def step():
    opt.zero_grad(set_to_none=True)
    fi, ft, scale = model(image, text)
    loss = loss_fn(fi, ft, scale)
    loss.backward()
    opt.step()
    return loss


for _ in range(3):
    loss = step()
torch.cuda.synchronize()
n0 = L.launch_count()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter()
e0.record()
n = 10
for _ in range(n):
    loss = step()
e1.record()
torch.cuda.synchronize()
wall = (time.perf_counter() - t0) / n * 1e3
ms = e0.elapsed_time(e1) / n
# algorithmic FLOPs of one step: forward (8.818 + 5.960 GFLOP per pair) x (1 forward + 1 recompute + 2 backward)
flops = B * (8.818e9 + 5.960e9) * 4
print(f"{precision} batch {B}: {ms:.2f} ms/step (wall {wall:.2f}), {B / ms * 1e3:.0f} pairs/s, {flops / ms / 1e9:.0f} TFLOP/s executed "
      f"(fwd + recompute + bwd), loss {float(loss):.4f}, {(L.launch_count() - n0) // n} kernel launches per step")
