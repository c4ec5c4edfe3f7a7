"""One zero-shot step of the benchmark workload between cudaProfilerStart/Stop (for ncu --profile-from-start off)."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from understanding_clip_ood_b200 import open_clip, ops  # noqa: E402

batch = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
torch.manual_seed(0)
model = open_clip.create_model("ViT-B-32", precision="bf16", device="cuda").eval()
g = torch.Generator(device="cuda").manual_seed(1)
image = torch.randn(batch, 3, 224, 224, device="cuda", generator=g).bfloat16()
prompt = ops.normalize(torch.randn(345, 512, device="cuda", generator=g).bfloat16())


def step():
    feat = model.encode_image(image, normalize=True)
    return ops.zeroshot(feat, prompt, 5, normalize_img=False, want_logits=False)


for _ in range(3):
    step()
torch.cuda.synchronize()
torch.cuda.profiler.start()
step()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("ok")
