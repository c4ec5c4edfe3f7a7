"""Two launches of the step's dominant kernel (LN-folded c_fc GEMM + GELU at batch 1024) for `ncu --set full`."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from understanding_clip_ood_b200 import _lib as L, ops  # noqa: E402

M, N, K = 1024 * 50, 3072, 768
g = torch.Generator(device="cuda").manual_seed(7)
x = (torch.randn(M, K, device="cuda", generator=g) * 0.5).bfloat16()
w = (torch.randn(N, K, device="cuda", generator=g) * 0.04).bfloat16()
b = torch.zeros(N, device="cuda", dtype=torch.bfloat16)
wf, colsum, bf = ops.fold_layernorm(w, b, torch.ones(K, device="cuda"), torch.zeros(K, device="cuda"), torch.bfloat16)
stats = ops.row_stats(x)
out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
for _ in range(3):
    ops.gemm_ln(x, wf, colsum, bf, stats, epilogue=L.EPI_GELU, out=out)
torch.cuda.synchronize()
print("ok")
