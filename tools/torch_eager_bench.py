"""Context number, not part of the product or the oracle: the same ViT-B/32 zero-shot step written with stock PyTorch ops
(F.layer_norm, F.linear -> cuBLAS, F.scaled_dot_product_attention -> flash/cuDNN, F.gelu) on the same GPU in bf16 —
i.e. what the reference's eager OpenCLIP path executes on a B200 (SURVEY.md §8d "the real bar").  Weights come from this
repo's create_model (same seed-0 init as the reference)."""
import sys
from pathlib import Path

import torch
import torch.nn.functional as F

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from understanding_clip_ood_b200 import open_clip, ops  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
torch.manual_seed(0)
model = open_clip.create_model("ViT-B-32", precision="bf16", device="cuda").eval()
sd = {k: v.detach() for k, v in model.state_dict().items()}
W, H, Lyr, P = 768, 12, 12, 32


def ln(x, pfx):
    return F.layer_norm(x.float(), (x.shape[-1],), sd[pfx + ".weight"], sd[pfx + ".bias"], 1e-5).to(x.dtype)   # LayerNormFp32


@torch.no_grad()
def eager_forward(img):
    x = F.conv2d(img, sd["visual.conv1.weight"], stride=P)                       # [B, W, 7, 7]
    x = x.reshape(x.shape[0], W, -1).permute(0, 2, 1)
    cls = sd["visual.class_embedding"].to(x.dtype).expand(x.shape[0], 1, W)
    x = torch.cat([cls, x], dim=1) + sd["visual.positional_embedding"].to(x.dtype)
    x = ln(x, "visual.ln_pre")
    for i in range(Lyr):
        p = f"visual.transformer.resblocks.{i}."
        h = ln(x, p + "ln_1")
        qkv = F.linear(h, sd[p + "attn.in_proj_weight"], sd[p + "attn.in_proj_bias"])
        q, k, v = qkv.view(x.shape[0], -1, 3, H, 64).permute(2, 0, 3, 1, 4)
        o = F.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(x.shape[0], -1, W)
        x = x + F.linear(o, sd[p + "attn.out_proj.weight"], sd[p + "attn.out_proj.bias"])
        h = ln(x, p + "ln_2")
        h = F.gelu(F.linear(h, sd[p + "mlp.c_fc.weight"], sd[p + "mlp.c_fc.bias"]))
        x = x + F.linear(h, sd[p + "mlp.c_proj.weight"], sd[p + "mlp.c_proj.bias"])
    pooled = ln(x[:, 0], "visual.ln_post")
    return F.normalize(pooled @ sd["visual.proj"], dim=-1)


g = torch.Generator(device="cuda").manual_seed(1)
image = torch.randn(B, 3, 224, 224, device="cuda", generator=g).bfloat16()
prompt = ops.normalize(torch.randn(345, 512, device="cuda", generator=g).bfloat16())


def eager_step():
    feat = eager_forward(image)
    return (feat @ prompt.t()).topk(5, dim=1).indices


def ours_step():
    feat = model.encode_image(image, normalize=True)
    return ops.zeroshot(feat, prompt, 5, normalize_img=False, want_logits=False)[1]


a, b = eager_forward(image).float(), model.encode_image(image, normalize=True).float()
print(f"embedding rel-L2 ours vs torch-eager bf16: {float(((a - b).norm(dim=1) / a.norm(dim=1)).max()):.3e}")
print(f"top-1 agreement: {float((eager_step()[:, 0] == ours_step()[:, 0]).float().mean()):.4f}")
for name, fn in (("torch eager (stock ops)", eager_step), ("b200clip", ours_step)):
    for _ in range(5):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 20
    torch.cuda.synchronize()
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    print(f"{name:>24}: {ms:7.2f} ms/step  {B / ms * 1e3:9.0f} img/s")
