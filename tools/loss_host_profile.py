"""cProfile of the ClipLoss fwd+bwd host path (world_size 1): where the per-step Python time goes."""
import cProfile
import pstats
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from understanding_clip_ood_b200 import open_clip, ops  # noqa: E402

dev = "cuda"
n = 256
g = torch.Generator(device=dev).manual_seed(100)
fi = ops.normalize(torch.randn(n, 512, device=dev, generator=g)).requires_grad_(True)
ft = ops.normalize(torch.randn(n, 512, device=dev, generator=g)).requires_grad_(True)
ls = torch.tensor(1 / 0.07, device=dev, requires_grad=True)
loss_fn = open_clip.ClipLoss(local_loss=True, gather_with_grad=True, cache_labels=True, rank=0, world_size=1)


def step():
    fi.grad = ft.grad = ls.grad = None
    loss_fn(fi, ft, ls).backward()


for _ in range(20):
    step()
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
for _ in range(500):
    step()
torch.cuda.synchronize()
pr.disable()
st = pstats.Stats(pr)
st.sort_stats("cumulative").print_stats(35)
