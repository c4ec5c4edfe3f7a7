"""Device time of the packed ClipLoss kernels alone at the 8-rank shape (n = 256 local rows, N = 2048 gathered), one GPU."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from understanding_clip_ood_b200 import ops  # noqa: E402

dev = "cuda"
for n, N in ((256, 256), (256, 512), (256, 2048), (128, 1024)):
    D = 512
    g = torch.Generator(device=dev).manual_seed(0)
    gathered = torch.nn.functional.normalize(torch.randn(N, 2 * D, device=dev, generator=g), dim=-1)
    scale = torch.tensor(1 / 0.07, device=dev)
    gout = torch.tensor(1.0, device=dev)

    def fwd():
        return ops.cliploss_packed_forward(gathered, scale, 0, n)

    loss, ws = fwd()

    def bwd():
        return ops.cliploss_packed_backward(gathered, scale, 0, n, ws, gout, True)

    world = N // n
    S = n * 2 * D
    recv = torch.zeros((world + 2) * S, device=dev)
    slots = torch.tensor([recv.data_ptr() + 4 * j * S for j in range(world + 2)], dtype=torch.int64, device=dev)

    def bwd_p2p():
        return ops.cliploss_packed_backward_p2p(gathered, scale, 0, n, ws, gout, slots, True)

    for name, fn in (("forward (2 launches)", fwd), ("backward (3 launches)", bwd), ("backward slot-addressed (2 launches)", bwd_p2p)):
        for _ in range(5):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        # capture into a graph so that the figure is device time, not Python launch time
        gr = torch.cuda.CUDAGraph()
        s = torch.cuda.Stream()
        with torch.cuda.stream(s):
            fn()
            torch.cuda.synchronize()
            with torch.cuda.graph(gr, stream=s):
                for _ in range(10):
                    fn()
        torch.cuda.synchronize()
        gr.replay()
        torch.cuda.synchronize()
        e0.record()
        for _ in range(10):
            gr.replay()
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / 100 * 1e3
        flops = (4 if name.startswith("forward") else 8) * n * N * D
        print(f"n={n} N={N} {name:38s}: {us:7.1f} us  {flops / us / 1e6:6.1f} TFLOP/s fp32", flush=True)
