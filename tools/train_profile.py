"""Kernel-time breakdown of the contrastive training step (tools/train_bench.py's step) from the CUDA activity records of
torch.profiler (CUPTI): which kernels of the backward dominate.  python tools/train_profile.py [precision] [batch]"""
import sys
from pathlib import Path

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from understanding_clip_ood_b200 import open_clip  # noqa: E402

precision = sys.argv[1] if len(sys.argv) > 1 else "bf16"
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
    loss_fn(fi, ft, scale).backward()
    opt.step()


for _ in range(3):
    step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(2):
        step()
    torch.cuda.synchronize()
rows = sorted(prof.key_averages(), key=lambda e: -e.device_time_total)
total = sum(e.device_time_total for e in rows)
print(f"{precision} batch {B}: {total / 2 / 1e3:.2f} ms of kernel time per step")
for e in rows[:28]:
    print(f"{e.device_time_total / 2 / 1e3:9.3f} ms {100 * e.device_time_total / total:5.1f}%  x{e.count // 2:4d}  {e.key[:110]}")
