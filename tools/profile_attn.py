"""One launch of the attention kernel at a given shape, for `ncu --set full -k regex:attention` captures."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from understanding_clip_ood_b200 import ops  # noqa: E402

B, L, H, causal = (int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), bool(int(sys.argv[4]))) if len(sys.argv) > 4 else (256, 197, 12, False)
W = H * 64
g = torch.Generator(device="cuda").manual_seed(0)
qkv = torch.randn(B * L, 3 * W, device="cuda", generator=g).bfloat16()
out = torch.empty(B * L, W, device="cuda", dtype=torch.bfloat16)
for _ in range(3):
    ops.attention(qkv, B, L, H, causal, out=out)
torch.cuda.synchronize()
print("ok")
