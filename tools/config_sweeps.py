"""BASELINE configs 3 and 5 on one B200: ViT-B-16 encode_image / encode_text sweep over batch 256..4096 (feature extraction,
scripts/save_domainnet_features.py-style) and ViT-L-14 zero-shot (tower -> normalise -> 345-class logits -> top-5) at the
per-GPU shard of batch 2048 over 8 GPUs (256) and at 1024.  CUDA events over back-to-back calls after warm-up; bf16."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from understanding_clip_ood_b200 import open_clip, ops  # noqa: E402

FLOPS = {"ViT-B-16": (35.127e9, 5.960e9), "ViT-L-14": (162.026e9, 13.300e9)}


def timed(fn, n):
    for _ in range(2):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


g = torch.Generator(device="cuda").manual_seed(1)
torch.manual_seed(0)
model = open_clip.create_model("ViT-B-16", precision="bf16", device="cuda").eval()
print("# config 3: ViT-B-16 forward sweep")
for B in (256, 512, 1024, 2048, 4096):
    image = torch.randn(B, 3, 224, 224, device="cuda", generator=g).bfloat16()
    ms = timed(lambda: model.encode_image(image), 5 if B <= 1024 else 3)
    out = model.encode_image(image)
    assert torch.isfinite(out.float()).all()
    print(f"ViT-B-16 encode_image B={B:5d}: {ms:8.2f} ms  {B / ms * 1e3:8.0f} img/s  {FLOPS['ViT-B-16'][0] * B / ms / 1e9:6.1f} TFLOP/s", flush=True)
    del image, out
for T in (256, 512, 1024, 2048, 4096):
    text = torch.zeros(T, 77, dtype=torch.long, device="cuda")
    text[:, 0] = 49406
    text[:, 1:9] = torch.randint(1000, 40000, (T, 8), device="cuda", generator=g)
    text[:, 9] = 49407
    for trunc in (False, True):
        model.truncate_text_at_eot = trunc
        ms = timed(lambda: model.encode_text(text), 5)
        print(f"ViT-B-16 encode_text  T={T:5d} truncate_at_eot={str(trunc):5s}: {ms:8.2f} ms  {T / ms * 1e3:8.0f} prompts/s", flush=True)
del model
torch.cuda.empty_cache()

print("# config 5: ViT-L-14 zero-shot + top-5 (345 classes)")
torch.manual_seed(0)
model = open_clip.create_model("ViT-L-14", precision="bf16", device="cuda").eval()
prompt = ops.normalize(torch.randn(345, 768, device="cuda", generator=g).bfloat16())
for B in (256, 1024):
    image = torch.randn(B, 3, 224, 224, device="cuda", generator=g).bfloat16()

    def step():
        feat = model.encode_image(image, normalize=True)
        return ops.zeroshot(feat, prompt, 5, normalize_img=False, want_logits=False)[1]

    ms = timed(step, 5 if B == 256 else 3)
    idx = step()
    assert idx.shape == (B, 5) and int(idx.min()) >= 0 and int(idx.max()) < 345
    print(f"ViT-L-14 zero-shot B={B:5d}: {ms:8.2f} ms  {B / ms * 1e3:8.0f} img/s  {FLOPS['ViT-L-14'][0] * B / ms / 1e9:6.1f} TFLOP/s", flush=True)
