"""Standalone diagnostic for the tcgen05 GEMMs: correctness sweep (error map per 32x32 block on mismatch) and a
timing table of the tower's GEMM shapes for the single-CTA kernel (block_n 128/256) and the CTA-pair kernel
(block_n 1000 + N tile).   python tools/gemm_debug.py [check] [bench]"""
import os
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from understanding_clip_ood_b200 import _lib as L, ops  # noqa: E402


def reference(a, w, bias, res, epi, dtype):
    ref = a.float() @ w.float().t()
    if bias is not None:
        ref = ref + bias.float()
    if epi == L.EPI_GELU:          # activations on the fp32 linear output (gemm_pair.cu B2C_ACT_ROUND = 0)
        ref = torch.nn.functional.gelu(ref)
    elif epi == L.EPI_QUICKGELU:
        ref = ref * torch.sigmoid(1.702 * ref)
    elif epi == L.EPI_RESIDUAL:
        ref = ref.to(dtype).float() + res.float()
    else:
        ref = ref.to(dtype).float()
    return ref


def run(M, N, K, epi, block_n, dtype=torch.bfloat16, seed=0, inplace=False, with_bias=True):
    g = torch.Generator(device="cuda").manual_seed(seed)
    a = (torch.randn(M, K, device="cuda", generator=g) * 0.5).to(dtype)
    w = (torch.randn(N, K, device="cuda", generator=g) * 0.05).to(dtype)
    bias = (torch.randn(N, device="cuda", generator=g) * 0.1).to(dtype) if with_bias else None
    res = (torch.randn(M, N, device="cuda", generator=g)).to(dtype) if epi == L.EPI_RESIDUAL else None
    ref = reference(a, w, bias, res, epi, dtype)
    out = res if inplace else None
    out = ops.gemm(a, w, bias, epilogue=epi, residual=res, block_n=block_n, out=out)
    torch.cuda.synchronize()
    err = (out.float() - ref).abs()
    tol = 2e-2 * ref.abs().max().item()
    ok = err.max().item() <= tol and bool(torch.isfinite(out.float()).all())
    print(f"M={M} N={N} K={K} epi={epi} bn={block_n} {str(dtype)[6:]}{' inplace' if inplace else ''}: max_err={err.max().item():.4e} "
          f"ref_max={ref.abs().max().item():.3f} -> {'OK' if ok else 'MISMATCH'}", flush=True)
    if not ok:
        Mb, Nb = (M + 31) // 32, (N + 31) // 32
        pad = torch.zeros(Mb * 32, Nb * 32, device="cuda")
        pad[:M, :N] = torch.nan_to_num(err, nan=1e9)
        blk = pad.view(Mb, 32, Nb, 32).amax(dim=(1, 3))
        bad = (blk > tol).cpu()
        print("bad 32x32 blocks (rows = M blocks, cols = N blocks), first 24x16:")
        for r in range(min(Mb, 24)):
            print("".join("X" if bad[r, c] else "." for c in range(min(Nb, 16))))
        print("out[0,:8] ", out[0, :8].float().tolist())
        print("ref[0,:8] ", ref[0, :8].tolist())
    return ok


def check():
    allok = True
    for (M, N, K, epi, bn) in [
        (128, 128, 64, 0, 128), (256, 256, 768, 0, 256), (6400, 3072, 768, 1, 0), (1000, 264, 72, 2, 128),
    ]:
        allok &= run(M, N, K, epi, bn)
    # CTA-pair kernel: every N tile x every epilogue, M / N / K tails, tiny and multi-round problems
    for bn in (1256, 1192, 1128, 2256, 2192, 2128):
        for (M, N, K, epi) in [
            (256, 256, 64, 0), (256, 256, 128, 0), (512, 512, 768, 0), (200, 512, 768, 0), (16, 512, 768, 0),
            (6400, 2304, 768, 0), (6400, 3072, 768, 1), (6400, 768, 3072, 3), (3200, 768, 768, 3),
            (1000, 264, 72, 2), (51200, 768, 768, 3), (777, 1000, 200, 3), (12800, 2304, 768, 0),
        ]:
            allok &= run(M, N, K, epi, bn)
    allok &= run(6400, 768, 768, 3, 1256, inplace=True)
    allok &= run(6400, 768, 3072, 3, 1192, inplace=True)
    allok &= run(6400, 768, 768, 3, 2256, inplace=True)
    allok &= run(51200, 768, 3072, 3, 2192, inplace=True)
    allok &= run(1024, 512, 512, 0, 2256, dtype=torch.float16)
    allok &= run(1024, 512, 768, 0, 1256, with_bias=False)
    allok &= run(1024, 512, 512, 0, 1256, dtype=torch.float16)
    allok &= run(6400, 3072, 768, 1, 0, dtype=torch.float16)
    return allok


# 1000 + N tile: CTA-pair kernel with whole tiles; -1: the same kernel with the stream-K workspace (b200clip_gemm_ws, tile picked
# by the library); 0: b200clip_gemm as the towers call it without a workspace
VARIANTS = tuple(int(v) for v in os.environ.get('GEMM_VARIANTS', '0,-1,1256,1192').split(','))
BATCHES = tuple(int(v) for v in os.environ.get('GEMM_BATCHES', '128,256,512,1024').split(','))


def sm_clock():
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(0)
        return pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)
    except Exception:  # noqa: BLE001
        return -1


def bench():
    print(f"{'shape':>34} {'kernel':>10} {'us':>9} {'TFLOP/s':>9}")
    for batch in BATCHES:
        M = batch * 50
        for (name, N, K, epi, inplace) in [("qkv", 2304, 768, 0, False), ("out_proj", 768, 768, 3, False), ("out_proj_inplace", 768, 768, 3, True),
                                           ("c_fc", 3072, 768, 1, False), ("c_proj", 768, 3072, 3, False), ("c_proj_inplace", 768, 3072, 3, True)]:
            if os.environ.get("GEMM_SHAPES") and name not in os.environ["GEMM_SHAPES"].split(","):
                continue
            g = torch.Generator(device="cuda").manual_seed(1)
            a = (torch.randn(M, K, device="cuda", generator=g) * 0.5).bfloat16()
            w = (torch.randn(N, K, device="cuda", generator=g) * 0.04).bfloat16()
            b = torch.zeros(N, device="cuda", dtype=torch.bfloat16)
            out = torch.zeros(M, N, device="cuda", dtype=torch.bfloat16)
            res = None
            if epi == 3:
                res = out if inplace else torch.zeros(M, N, device="cuda", dtype=torch.bfloat16)

            def timeit(fn, iters=40):
                for _ in range(3):
                    fn()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                torch.cuda.synchronize()
                e0.record()
                for _ in range(iters):
                    fn()
                e1.record()
                torch.cuda.synchronize()
                return e0.elapsed_time(e1) / iters * 1e3

            label = name + f" M={M} N={N} K={K}"
            for bn in VARIANTS:
                if bn == -1:
                    if epi == 3 and not inplace:
                        continue
                    us = timeit(lambda: ops.gemm_ws(a, w, b, epilogue=epi, residual=res, out=out))
                else:
                    us = timeit(lambda: ops.gemm(a, w, b, epilogue=epi, residual=res, out=out, block_n=bn))
                print(f"{label:>34} {bn:>10} {us:9.1f} {2.0 * M * N * K / us / 1e6:9.1f}", flush=True)
            if not inplace and os.environ.get("GEMM_CUBLAS", "1") == "1":
                us = timeit(lambda: torch.nn.functional.linear(a, w, b))
                print(f"{label:>34} {'cublas':>10} {us:9.1f} {2.0 * M * N * K / us / 1e6:9.1f}   (F.linear + bias only, no act/residual)", flush=True)


def bench_ln():
    """c_fc exactly as the tower runs it: LayerNorm folded into the epilogue (+bias +GELU), whole tiles vs stream-K."""
    for batch in BATCHES:
        M, N, K = batch * 50, 3072, 768
        g = torch.Generator(device="cuda").manual_seed(7)
        x = (torch.randn(M, K, device="cuda", generator=g) * 0.5).bfloat16()
        w = (torch.randn(N, K, device="cuda", generator=g) * 0.04).bfloat16()
        b = torch.zeros(N, device="cuda", dtype=torch.bfloat16)
        wf, colsum, bf = ops.fold_layernorm(w, b, torch.ones(K, device="cuda"), torch.zeros(K, device="cuda"), torch.bfloat16)
        stats = ops.row_stats(x)
        out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
        for name, fn in (("ln+gelu", lambda: ops.gemm_ln(x, wf, colsum, bf, stats, epilogue=L.EPI_GELU, out=out)),
                         ("ln+gelu ws", lambda: ops.gemm_ln_ws(x, wf, colsum, bf, stats, epilogue=L.EPI_GELU, out=out))):
            for _ in range(3):
                fn()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            for _ in range(40):
                fn()
            e1.record()
            torch.cuda.synchronize()
            us = e0.elapsed_time(e1) / 40 * 1e3
            print(f"{'c_fc M=%d N=%d K=%d' % (M, N, K):>34} {name:>10} {us:9.1f} {2.0 * M * N * K / us / 1e6:9.1f}", flush=True)


if __name__ == "__main__":
    print(torch.cuda.get_device_name(0), flush=True)
    what = sys.argv[1:] or ["check"]
    ok = True
    if "check" in what:
        ok = check()
    if "bench" in what:
        bench()
        bench_ln()
    sys.exit(0 if ok else 1)
