"""Pinned-host -> device copy bandwidth for the e2e batch size (308 MB), by chunking / stream count."""
import torch

N = 1024 * 3 * 224 * 224 * 2
host = torch.empty(N, dtype=torch.uint8).pin_memory()
dev = torch.empty(N, dtype=torch.uint8, device="cuda")


def run(chunks, streams):
    ss = [torch.cuda.Stream() for _ in range(streams)]
    step = N // chunks
    def once():
        for i in range(chunks):
            with torch.cuda.stream(ss[i % streams]):
                dev[i * step:(i + 1) * step].copy_(host[i * step:(i + 1) * step], non_blocking=True)
    for _ in range(2):
        once()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for s in ss:
        s.wait_event(e0)
    for _ in range(5):
        once()
    for s in ss:
        torch.cuda.current_stream().wait_stream(s)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(f"chunks={chunks} streams={streams}: {ms:7.2f} ms  {N / ms / 1e6:6.1f} GB/s", flush=True)


for c, s in [(1, 1), (4, 1), (4, 2), (8, 4), (16, 4)]:
    run(c, s)
# device -> host for comparison
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize()
e0.record()
host.copy_(dev, non_blocking=True)
e1.record()
torch.cuda.synchronize()
print(f"d2h: {N / e0.elapsed_time(e1) / 1e6:6.1f} GB/s")
import subprocess
print(subprocess.run(["nvidia-smi", "--query-gpu=pcie.link.gen.current,pcie.link.gen.max,pcie.link.width.current,pcie.link.width.max", "--format=csv"], capture_output=True, text=True).stdout)
