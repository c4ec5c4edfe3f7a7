"""Timeline of CTA 0 of the tcgen05 attention kernel: clock64 stamps per tile (debug hook b200clip_attention_debug).
events: 0 S issue (MMA thread) | 1 s_full seen (softmax) | 2 PV issue | 3 P written | 4 o_full seen | 5 O drained | 6 store issued"""
import ctypes as C
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from understanding_clip_ood_b200 import _lib as L, ops  # noqa: E402

B, Lq, H, causal = (int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), bool(int(sys.argv[4]))) if len(sys.argv) > 4 else (256, 197, 12, False)
W = H * 64
qkv = torch.randn(B * Lq, 3 * W, device="cuda").bfloat16()
out = torch.empty(B * Lq, W, device="cuda", dtype=torch.bfloat16)
for _ in range(2):
    ops.attention(qkv, B, Lq, H, causal, out=out)
dbg = torch.zeros(64 * 8, dtype=torch.int64, device="cuda")
lib = L.load()
lib.b200clip_attention_debug.argtypes = [C.c_void_p]
lib.b200clip_attention_debug.restype = None
lib.b200clip_attention_debug(dbg.data_ptr())
ops.attention(qkv, B, Lq, H, causal, out=out)
torch.cuda.synchronize()
lib.b200clip_attention_debug(None)
d = dbg.view(64, 8).cpu()
t0 = int(d[0, 0])
names = ["S_issue", "s_full", "PV_issue", "P_done", "o_full", "O_drained", "store", "PV_issued"]
print("tile " + " ".join(f"{n:>10s}" for n in names) + "   (cycles since S_issue of tile 0)")
for t in range(4, 20):
    print(f"{t:4d} " + " ".join(f"{int(d[t, e]) - t0:10d}" for e in range(8)))
