"""encode_image / encode_text throughput of the configured models (BASELINE configs 3 and 5 shapes, reduced sweeps)."""
import sys
import time
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from understanding_clip_ood_b200 import open_clip  # noqa: E402

FLOPS = {"ViT-B-32": (8.818e9, 5.960e9), "ViT-B-16": (35.127e9, 5.960e9), "ViT-L-14": (162.026e9, 13.300e9)}
cases = [("ViT-B-32", (128, 1024)), ("ViT-B-16", (256, 1024)), ("ViT-L-14", (256,))]
if len(sys.argv) > 1:
    cases = [(sys.argv[1], tuple(int(v) for v in sys.argv[2:]) or (256,))]
for name, batches in cases:
    torch.manual_seed(0)
    model = open_clip.create_model(name, precision="bf16", device="cuda").eval()
    g = torch.Generator(device="cuda").manual_seed(1)
    for B in batches:
        image = torch.randn(B, 3, 224, 224, device="cuda", generator=g).bfloat16()
        for _ in range(3):
            model.encode_image(image)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 10
        torch.cuda.synchronize()
        e0.record()
        for _ in range(n):
            model.encode_image(image)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        print(f"{name} encode_image B={B}: {ms:8.2f} ms  {B / ms * 1e3:9.0f} img/s  {FLOPS[name][0] * B / ms / 1e9:7.1f} TFLOP/s", flush=True)
    T = 1024
    text = torch.zeros(T, 77, dtype=torch.long, device="cuda")
    text[:, 0] = 49406
    text[:, 1:9] = torch.randint(1000, 40000, (T, 8), device="cuda", generator=g)
    text[:, 9] = 49407
    for trunc in (False, True):
        model.truncate_text_at_eot = trunc
        for _ in range(3):
            model.encode_text(text)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(5):
            model.encode_text(text)
        torch.cuda.synchronize()
        ms = (time.perf_counter() - t0) / 5 * 1e3
        print(f"{name} encode_text T={T} truncate_at_eot={trunc}: {ms:8.2f} ms  {T / ms * 1e3:9.0f} prompts/s", flush=True)
    del model
    torch.cuda.empty_cache()
