"""Timing of the tcgen05 attention kernel shapes only (text L=77, ViT-B/16 L=197, ViT-L/14 L=257)."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from understanding_clip_ood_b200 import ops  # noqa: E402

for (B, L, H, causal) in [(1024, 77, 8, True), (256, 197, 12, False), (128, 257, 16, False), (256, 257, 16, False)]:
    W = H * 64
    g = torch.Generator(device="cuda").manual_seed(0)
    qkv = torch.randn(B * L, 3 * W, device="cuda", generator=g).bfloat16()
    out = torch.empty(B * L, W, device="cuda", dtype=torch.bfloat16)
    for _ in range(3):
        ops.attention(qkv, B, L, H, causal, out=out)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    iters = 30
    torch.cuda.synchronize()
    e0.record()
    for _ in range(iters):
        ops.attention(qkv, B, L, H, causal, out=out)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / iters * 1e3
    bytes_ = qkv.numel() * 2 + out.numel() * 2
    print(f"B={B} L={L} H={H} causal={causal}: {us:8.1f} us  {bytes_ / us / 1e3:8.1f} GB/s", flush=True)
