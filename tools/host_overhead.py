"""Host-side cost of one encode_image call (tiny batch: GPU time negligible) eager vs CUDA-graph replay, and the
benchmark step at batch 1024 with / without graphs."""
import sys
import time
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from understanding_clip_ood_b200 import open_clip, ops  # noqa: E402

torch.manual_seed(0)
model = open_clip.create_model("ViT-B-32", precision="bf16", device="cuda").eval()
g = torch.Generator(device="cuda").manual_seed(1)
prompt = ops.normalize(torch.randn(345, 512, device="cuda", generator=g).bfloat16())
for batch in (8, 128, 1024):
    image = torch.randn(batch, 3, 224, 224, device="cuda", generator=g).bfloat16()
    for graphs in (False, True):
        model.visual.use_cuda_graphs = graphs
        for _ in range(4):
            model.encode_image(image, normalize=True)
        torch.cuda.synchronize()
        n = 30
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        for _ in range(n):
            feat = model.encode_image(image, normalize=True)
            ops.zeroshot(feat, prompt, 5, normalize_img=False, want_logits=False)
        e1.record()
        host = (time.perf_counter() - t0) / n * 1e3
        torch.cuda.synchronize()
        print(f"batch {batch:5d} graphs={graphs!s:5}: host enqueue {host:7.3f} ms/step, device {e0.elapsed_time(e1) / n:7.3f} ms/step", flush=True)
