"""patchify (im2col) kernel time at the benchmark shape, graph-timed."""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from understanding_clip_ood_b200 import ops  # noqa: E402
from understanding_clip_ood_b200.open_clip import OPENAI_DATASET_MEAN as MEAN, OPENAI_DATASET_STD as STD  # noqa: E402

for (B, P, kpad) in [(1024, 32, 3072), (256, 16, 768), (256, 14, 640)]:
    img = torch.randn(B, 3, 224, 224, device="cuda").bfloat16()
    u8 = torch.randint(0, 256, (B, 3, 224, 224), device="cuda", dtype=torch.uint8)
    for name, fn in (("patchify", lambda: ops.patchify(img, P, kpad)), ("patchify_u8", lambda: ops.patchify_u8(u8, P, kpad, torch.bfloat16, MEAN, STD))):
        for _ in range(3):
            fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(20):
            fn()
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / 20 * 1e3
        g = 224 // P
        byts = B * 3 * 224 * 224 * (2 if name == "patchify" else 1) + B * g * g * kpad * 2
        print(f"{name:12s} B={B} P={P}: {us:7.1f} us  {byts / us / 1e3:7.1f} GB/s (incl. torch.empty of the output)", flush=True)
