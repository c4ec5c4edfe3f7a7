"""A few launches of the GEMM variants added at the end of round 2, for ONE `ncu --set full` capture: MN-major dgrad / wgrad of the
training path (b200clip_gemm_mn) and the implicit 3x3 convolution of the ModifiedResNet tower (through an RN50 forward)."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from understanding_clip_ood_b200 import open_clip, ops  # noqa: E402

g = torch.Generator(device="cuda").manual_seed(1)
tokens = 6400
G = (torch.randn(tokens, 768, device="cuda", generator=g) * 0.5).bfloat16()
X = (torch.randn(tokens, 3072, device="cuda", generator=g) * 0.5).bfloat16()
W = (torch.randn(768, 3072, device="cuda", generator=g) * 0.04).bfloat16()
for _ in range(2):
    ops.gemm_mn(G, W)                          # dgrad: d_a [6400, 3072] = dY [6400, 768] W [768, 3072]
    ops.gemm_mn(G, X, a_transposed=True)       # wgrad: dW [768, 3072] = dY^T a  (contraction over 6400 tokens, stream-K)
model = open_clip.create_model("RN50", precision="bf16", device="cuda").eval()
image = torch.randn(64, 3, 224, 224, device="cuda", generator=g).bfloat16()
with torch.no_grad():
    for _ in range(2):
        model.encode_image(image)
torch.cuda.synchronize()
print("ok")
