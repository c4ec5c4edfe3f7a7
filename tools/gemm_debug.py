"""Standalone diagnostic for the tcgen05 GEMM: prints an error map per 32x32 block on mismatch."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from understanding_clip_ood_b200 import _lib as L, ops  # noqa: E402


def run(M, N, K, epi, block_n, dtype=torch.bfloat16, seed=0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    a = (torch.randn(M, K, device="cuda", generator=g) * 0.5).to(dtype)
    w = (torch.randn(N, K, device="cuda", generator=g) * 0.05).to(dtype)
    bias = (torch.randn(N, device="cuda", generator=g) * 0.1).to(dtype)
    res = (torch.randn(M, N, device="cuda", generator=g)).to(dtype) if epi == L.EPI_RESIDUAL else None
    out = ops.gemm(a, w, bias, epilogue=epi, residual=res, block_n=block_n)
    torch.cuda.synchronize()
    ref = a.float() @ w.float().t() + bias.float()
    ref = ref.to(dtype).float()
    if epi == L.EPI_GELU:
        ref = torch.nn.functional.gelu(ref)
    elif epi == L.EPI_QUICKGELU:
        ref = ref * torch.sigmoid(1.702 * ref)
    elif epi == L.EPI_RESIDUAL:
        ref = ref + res.float()
    err = (out.float() - ref).abs()
    tol = 2e-2 * ref.abs().max().item()
    ok = err.max().item() <= tol
    print(f"M={M} N={N} K={K} epi={epi} bn={block_n} {dtype}: max_err={err.max().item():.4e} "
          f"ref_max={ref.abs().max().item():.3f} -> {'OK' if ok else 'MISMATCH'}", flush=True)
    if not ok:
        Mb, Nb = (M + 31) // 32, (N + 31) // 32
        pad = torch.zeros(Mb * 32, Nb * 32, device="cuda")
        pad[:M, :N] = err
        blk = pad.view(Mb, 32, Nb, 32).amax(dim=(1, 3))
        bad = (blk > tol).cpu()
        print("bad 32x32 blocks (rows = M blocks, cols = N blocks), first 16x16:")
        for r in range(min(Mb, 16)):
            print("".join("X" if bad[r, c] else "." for c in range(min(Nb, 16))))
        print("out[0,:8] ", out[0, :8].float().tolist())
        print("ref[0,:8] ", ref[0, :8].tolist())
        print("out[1,:8] ", out[1, :8].float().tolist())
        print("ref[1,:8] ", ref[1, :8].tolist())
    return ok


if __name__ == "__main__":
    print(torch.cuda.get_device_name(0), flush=True)
    allok = True
    for (M, N, K, epi, bn) in [
        (128, 128, 64, 0, 128), (128, 256, 64, 0, 256), (128, 128, 128, 0, 128), (256, 256, 768, 0, 256),
        (6400, 2304, 768, 0, 0), (6400, 3072, 768, 1, 0), (6400, 768, 3072, 3, 0), (200, 512, 768, 0, 0),
        (3200, 768, 768, 3, 128), (1000, 264, 72, 2, 128),
    ]:
        allok &= run(M, N, K, epi, bn)
    allok &= run(1024, 512, 512, 0, 0, dtype=torch.float16)
    sys.exit(0 if allok else 1)
