"""ModifiedResNet encode_image throughput: this path vs the unmodified reference (baseline/_ref) in eager PyTorch on the same GPU.
   python tools/rn_bench.py [model] [batch ...]"""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from understanding_clip_ood_b200 import open_clip  # noqa: E402

GFLOP = {"RN50": 12.22, "RN101": 19.54}     # docs/model_profile.csv of the reference (GFLOPs = 2 x MACs, image tower)


def timeit(fn, n=10, warm=3):
    for _ in range(warm):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "RN50"
    batches = tuple(int(v) for v in sys.argv[2:]) or (128, 256, 1024)
    torch.manual_seed(0)
    model = open_clip.create_model(name, precision="bf16", device="cuda").eval()
    ref = None
    if (ROOT / "baseline" / "_ref" / "open_clip").exists():
        sys.path.insert(0, str(ROOT / "baseline" / "_ref"))
        import types
        if "ftfy" not in sys.modules:
            stub = types.ModuleType("ftfy")
            stub.fix_text = lambda s: s
            sys.modules["ftfy"] = stub
        import open_clip as ref_oc
        torch.manual_seed(0)
        ref = ref_oc.create_model(name, precision="bf16", device="cuda").eval()
    g = torch.Generator(device="cuda").manual_seed(1)
    for B in batches:
        image = torch.randn(B, 3, model.visual.image_size[0], model.visual.image_size[0], device="cuda", generator=g).bfloat16()
        with torch.no_grad():
            ms = timeit(lambda: model.encode_image(image))
            line = f"{name} encode_image B={B}: {ms:8.2f} ms  {B / ms * 1e3:9.0f} img/s"
            if name in GFLOP:
                line += f"  {GFLOP[name] * B / ms:7.1f} TFLOP/s"
            if ref is not None:
                rms = timeit(lambda: ref.encode_image(image), n=5, warm=2)
                rcl = ref.to(memory_format=torch.channels_last)
                icl = image.contiguous(memory_format=torch.channels_last)
                rms2 = timeit(lambda: rcl.encode_image(icl), n=5, warm=2)
                line += f"   reference eager {rms:8.2f} ms (channels_last {rms2:8.2f} ms)  -> {rms / ms:5.2f}x"
                d = (model.encode_image(image).float() - ref.encode_image(image).float()).norm(dim=-1) / ref.encode_image(image).float().norm(dim=-1)
                line += f"   rel-L2 vs reference bf16 {float(d.max()):.2e}"
            print(line, flush=True)


if __name__ == "__main__":
    main()
