"""How much of the N = 768 residual GEMMs is the last partial round?  Time x += a W^T + b (in place, reduce-add store) at
M = 51200 (600 tiles = 8.1 rounds on 74 clusters) against M = 50432 (591 tiles = 7.99 rounds) and a few more row counts."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from understanding_clip_ood_b200 import _lib as L, ops  # noqa: E402

g = torch.Generator(device="cuda").manual_seed(1)
for (N, K) in [(768, 768), (768, 3072), (2304, 768), (3072, 768)]:
    for M in (47360, 50432, 51200, 56832):
        a = (torch.randn(M, K, device="cuda", generator=g) * 0.5).bfloat16()
        w = (torch.randn(N, K, device="cuda", generator=g) * 0.04).bfloat16()
        b = torch.zeros(N, device="cuda", dtype=torch.bfloat16)
        x = torch.zeros(M, N, device="cuda", dtype=torch.bfloat16)
        epi = L.EPI_RESIDUAL if N == 768 else L.EPI_BIAS
        kw = dict(residual=x) if N == 768 else {}
        for _ in range(3):
            ops.gemm(a, w, b, epilogue=epi, out=x, **kw)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(20):
            ops.gemm(a, w, b, epilogue=epi, out=x, **kw)
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / 20 * 1e3
        tiles = ((M + 255) // 256) * ((N + 255) // 256)
        print(f"N={N:5d} K={K:5d} M={M:6d}: {us:7.1f} us  {2.0 * M * N * K / us / 1e6:7.1f} TFLOP/s   {tiles} tiles = {tiles / 74:5.2f} rounds", flush=True)
