"""A few launches of the tower GEMM shapes at batch 1024 (for ncu --set full).  python tools/profile_gemm.py [block_n]"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from understanding_clip_ood_b200 import ops  # noqa: E402

bn = int(sys.argv[1]) if len(sys.argv) > 1 else 0
M = 1024 * 50
g = torch.Generator(device="cuda").manual_seed(1)
for (N, K, epi) in [(2304, 768, 0), (768, 768, 3), (3072, 768, 1), (768, 3072, 3)]:
    a = (torch.randn(M, K, device="cuda", generator=g) * 0.5).bfloat16()
    w = (torch.randn(N, K, device="cuda", generator=g) * 0.04).bfloat16()
    b = torch.zeros(N, device="cuda", dtype=torch.bfloat16)
    res = torch.zeros(M, N, device="cuda", dtype=torch.bfloat16) if epi == 3 else None
    out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    for _ in range(2):
        ops.gemm(a, w, b, epilogue=epi, residual=res, out=out, block_n=bn)
    torch.cuda.synchronize()
print("ok")
