"""Distributed ClipLoss (--local-loss --gather-with-grad) fwd+bwd step rate under torchrun: our peer-memory path,
our NCCL path (B200CLIP_P2P=0) and the reference formulation in stock PyTorch (torch.distributed.nn.all_gather x2 + eager ops)."""
import os
import sys
import time
from pathlib import Path

import torch
import torch.distributed as dist

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from understanding_clip_ood_b200 import open_clip, ops  # noqa: E402
from understanding_clip_ood_b200.open_clip import loss as loss_mod  # noqa: E402

rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); lr = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)


def timeit(step, iters=200):
    for _ in range(10):
        step()
    dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(iters):
        step()
    e1.record()
    host = (time.perf_counter() - t0) / iters * 1e6
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / iters * 1e3], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t), host


for n in (128, 256):
    g = torch.Generator(device=dev).manual_seed(100 + rank)
    fi = ops.normalize(torch.randn(n, 512, device=dev, generator=g)).requires_grad_(True)
    ft = ops.normalize(torch.randn(n, 512, device=dev, generator=g)).requires_grad_(True)
    ls = torch.tensor(1 / 0.07, device=dev, requires_grad=True)
    loss_fn = open_clip.ClipLoss(local_loss=True, gather_with_grad=True, cache_labels=True, rank=rank, world_size=world)

    def step():
        fi.grad = ft.grad = ls.grad = None
        loss_fn(fi, ft, ls).backward()

    def step_nccl():
        fi.grad = ft.grad = ls.grad = None
        loss_mod._DistLocalClipLoss.apply(fi, ft, ls, rank, world, None).backward()

    def step_ref():
        import torch.distributed.nn
        fi.grad = ft.grad = ls.grad = None
        all_i = torch.cat(torch.distributed.nn.all_gather(fi), dim=0)
        all_t = torch.cat(torch.distributed.nn.all_gather(ft), dim=0)
        li = ls * fi @ all_t.t()
        lt = ls * ft @ all_i.t()
        lab = torch.arange(n, device=dev) + n * rank
        ((torch.nn.functional.cross_entropy(li, lab) + torch.nn.functional.cross_entropy(lt, lab)) / 2).backward()

    for name, fn in (("peer memory", step), ("ours over NCCL", step_nccl), ("reference formulation, torch eager + NCCL", step_ref)):
        d, h = timeit(fn)
        if rank == 0:
            print(f"world={world} n={n} N={n * world} {name:45s}: device {d:7.1f} us/step (max over ranks), host {h:7.1f} us -> {1e6 / d:7.0f} step/s", flush=True)
dist.barrier()
dist.destroy_process_group()
