"""LayerNorm kernel timing at the benchmark shape (51200 x 768 bf16) + the CLS-pooling shape."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from understanding_clip_ood_b200 import ops  # noqa: E402

for (rows, W) in [(51200, 768), (6400, 768), (65792, 1024), (474720, 512)]:
    g = torch.Generator(device="cuda").manual_seed(0)
    x = torch.randn(rows, W, device="cuda", generator=g).bfloat16()
    gam = torch.randn(W, device="cuda", generator=g)
    bet = torch.randn(W, device="cuda", generator=g)
    out = torch.empty_like(x)
    for _ in range(3):
        ops.layernorm(x, gam, bet, out=out)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    iters = 50
    torch.cuda.synchronize()
    e0.record()
    for _ in range(iters):
        ops.layernorm(x, gam, bet, out=out)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / iters * 1e3
    print(f"rows={rows} W={W}: {us:7.1f} us  {2 * x.numel() * 2 / us / 1e3:7.1f} GB/s", flush=True)
