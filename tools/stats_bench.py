"""row_stats / layernorm kernel time measured from a CUDA graph of 20 back-to-back launches (no host launch overhead)."""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from understanding_clip_ood_b200 import ops  # noqa: E402


def graph_time(fn, n=20, reps=5):
    fn(); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(n):
            fn()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        g.replay()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (n * reps) * 1e3


for (rows, W) in [(51200, 768), (6400, 768), (65792, 1024), (78848, 512)]:
    x = torch.randn(rows, W, device="cuda").bfloat16()
    st = torch.empty(rows, 2, device="cuda")
    gam = torch.ones(W, device="cuda"); bet = torch.zeros(W, device="cuda")
    out = torch.empty_like(x)
    us = graph_time(lambda: ops.row_stats(x))
    print(f"row_stats rows={rows} W={W}: {us:6.1f} us  {x.numel() * 2 / us / 1e3:7.1f} GB/s")
    us = graph_time(lambda: ops.layernorm(x, gam, bet, out=out))
    print(f"layernorm rows={rows} W={W}: {us:6.1f} us  {2 * x.numel() * 2 / us / 1e3:7.1f} GB/s")
