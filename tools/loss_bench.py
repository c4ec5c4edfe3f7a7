"""ClipLoss fwd+bwd step rate on one GPU (world_size 1) for the BASELINE config-4 shapes."""
import sys
import time
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from understanding_clip_ood_b200 import open_clip, ops  # noqa: E402

dev = "cuda"
for n in (128, 256, 1024, 2048):
    g = torch.Generator(device=dev).manual_seed(100)
    fi = ops.normalize(torch.randn(n, 512, device=dev, generator=g)).requires_grad_(True)
    ft = ops.normalize(torch.randn(n, 512, device=dev, generator=g)).requires_grad_(True)
    ls = torch.tensor(1 / 0.07, device=dev, requires_grad=True)
    loss_fn = open_clip.ClipLoss(local_loss=True, gather_with_grad=True, cache_labels=True, rank=0, world_size=1)

    def step():
        fi.grad = ft.grad = ls.grad = None
        loss_fn(fi, ft, ls).backward()

    for _ in range(10):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    iters = 200
    t0 = time.perf_counter()
    e0.record()
    for _ in range(iters):
        step()
    e1.record()
    host = (time.perf_counter() - t0) / iters * 1e6
    torch.cuda.synchronize()
    dev_us = e0.elapsed_time(e1) / iters * 1e3
    print(f"n=N={n}: host {host:7.1f} us/step, device {dev_us:7.1f} us/step -> {1e6 / dev_us:8.0f} step/s", flush=True)
    # torch eager reference of the same step on the same GPU (for context)
    def ref_step():
        fi.grad = ft.grad = ls.grad = None
        li = ls * fi @ ft.t()
        lt = ls * ft @ fi.t()
        lab = torch.arange(n, device=dev)
        ((torch.nn.functional.cross_entropy(li, lab) + torch.nn.functional.cross_entropy(lt, lab)) / 2).backward()
    for _ in range(10):
        ref_step()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(iters):
        ref_step()
    e1.record()
    torch.cuda.synchronize()
    print(f"          torch eager (fp32, allow_tf32={torch.backends.cuda.matmul.allow_tf32}) {e0.elapsed_time(e1) / iters * 1e3:7.1f} us/step", flush=True)
