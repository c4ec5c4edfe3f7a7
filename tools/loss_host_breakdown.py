"""Wall-clock breakdown of one ClipLoss fwd+bwd step on the host (world_size 1), without a profiler."""
import sys
import time
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from understanding_clip_ood_b200 import open_clip, ops  # noqa: E402
from understanding_clip_ood_b200.open_clip import loss as LM  # noqa: E402

dev = "cuda"
n = 256
g = torch.Generator(device=dev).manual_seed(100)
fi = ops.normalize(torch.randn(n, 512, device=dev, generator=g)).requires_grad_(True)
ft = ops.normalize(torch.randn(n, 512, device=dev, generator=g)).requires_grad_(True)
ls = torch.tensor(1 / 0.07, device=dev, requires_grad=True)
loss_fn = open_clip.ClipLoss(local_loss=True, gather_with_grad=True, cache_labels=True, rank=0, world_size=1)
N = 2000
pc = time.perf_counter


def bench(name, fn):
    for _ in range(50):
        fn()
    torch.cuda.synchronize()
    t0 = pc()
    for _ in range(N):
        fn()
    t1 = pc()
    torch.cuda.synchronize()
    print(f"{name:55s}: {(t1 - t0) / N * 1e6:7.1f} us host per call", flush=True)


def full():
    fi.grad = ft.grad = ls.grad = None
    loss_fn(fi, ft, ls).backward()


def fwd_only():
    loss_fn(fi, ft, ls)


def fwd_nograd():
    with torch.no_grad():
        loss_fn(fi, ft, ls)


f32 = [fi.detach(), ft.detach(), fi.detach(), ft.detach()]
sc = ls.detach()


def ops_fwd():
    return ops.cliploss_forward(*f32, sc, 0)


loss0, ws0 = ops_fwd()
gout = torch.tensor(1.0, device=dev)


def ops_bwd():
    return ops.cliploss_backward(*f32, sc, 0, ws0, gout, [True] * 5)


def trivial_autograd():
    x = fi * 1.0
    x.sum().backward()
    fi.grad = None


bench("full step (zero grads, forward, backward)", full)
bench("forward only (autograd graph built)", fwd_only)
bench("forward under no_grad", fwd_nograd)
bench("ops.cliploss_forward (ctypes + 2 empty + 3 launches)", ops_fwd)
bench("ops.cliploss_backward (ctypes + 5 empty + 3 launches)", ops_bwd)
bench("reference point: (x*1).sum().backward() in torch", trivial_autograd)
bench("torch.empty x1", lambda: torch.empty((256, 512), device=dev))
bench("L.stream_ptr", lambda: __import__("understanding_clip_ood_b200")._lib.stream_ptr())
