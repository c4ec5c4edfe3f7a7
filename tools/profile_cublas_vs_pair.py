"""cuBLAS (F.linear + bias) and gemm_pair_kernel on the c_proj / out-proj / c_fc shapes at batch 1024, two launches each, for ONE
`ncu --set full` capture: which tile / cluster shape does cuBLAS pick, and how do L2 traffic and tensor-pipe activity compare?"""
import sys
from pathlib import Path

import torch
import torch.nn.functional as F

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from understanding_clip_ood_b200 import ops  # noqa: E402

M = 1024 * 50
g = torch.Generator(device="cuda").manual_seed(1)
for (N, K, epi) in [(768, 3072, 0), (768, 768, 0), (3072, 768, 0)]:
    a = (torch.randn(M, K, device="cuda", generator=g) * 0.5).bfloat16()
    w = (torch.randn(N, K, device="cuda", generator=g) * 0.04).bfloat16()
    b = torch.zeros(N, device="cuda", dtype=torch.bfloat16)
    out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    for _ in range(2):
        F.linear(a, w, b)
    for _ in range(2):
        ops.gemm(a, w, b, epilogue=epi, out=out, block_n=1256)
    torch.cuda.synchronize()
print("ok")
