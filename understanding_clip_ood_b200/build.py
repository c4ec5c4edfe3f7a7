"""In-tree build of libb200clip.so (hand-written sm_100a CUDA behind the C ABI of include/b200clip.h).

`python -m understanding_clip_ood_b200.build` (or `__graft_entry__.build()`) cross-compiles every
`csrc/*.cu` for sm_100a with nvcc (no GPU needed) and links one shared library next to this file,
so that the built `.so` travels with the repo snapshot to the GPU box.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
CSRC = PKG_DIR / "csrc"
BUILD_DIR = PKG_DIR / "build"
LIB_PATH = PKG_DIR / "libb200clip.so"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr",
    "-Xptxas", "-v",
    # experiment switches (-DNAME=value), e.g. B200CLIP_NVCC_EXTRA="-DB2C_ATTN_NO_PIPE=1"
    *os.environ.get("B200CLIP_NVCC_EXTRA", "").split(),
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found (needed to build libb200clip.so)")


def _digest(src: Path) -> str:
    h = hashlib.sha256()
    h.update(" ".join(NVCC_FLAGS).encode())
    h.update(src.read_bytes())
    for hdr in sorted(list(CSRC.glob("*.cuh")) + list(CSRC.glob("*.h")) + [PKG_DIR.parent / "include" / "b200clip.h"]):
        h.update(hdr.read_bytes())
    return h.hexdigest()


def _compile(src: Path, verbose: bool) -> Path:
    obj = BUILD_DIR / (src.stem + ".o")
    stamp = BUILD_DIR / (src.stem + ".sha")
    dig = _digest(src)
    if obj.exists() and stamp.exists() and stamp.read_text() == dig:
        return obj
    cmd = [_nvcc(), *NVCC_FLAGS, "-c", str(src), "-o", str(obj)]
    res = subprocess.run(cmd, capture_output=True, text=True)
    (BUILD_DIR / (src.stem + ".ptxas.log")).write_text(res.stderr)
    if res.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src.name}:\n{res.stderr}\n{res.stdout}")
    if verbose:
        sys.stderr.write(f"[b200clip build] compiled {src.name}\n")
    stamp.write_text(dig)
    return obj


def build(verbose: bool = True, force: bool = False) -> Path:
    BUILD_DIR.mkdir(exist_ok=True)
    if force:
        for f in BUILD_DIR.glob("*.sha"):
            f.unlink()
    srcs = sorted(CSRC.glob("*.cu"))
    with ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        objs = list(ex.map(lambda s: _compile(s, verbose), srcs))
    newest = max(o.stat().st_mtime for o in objs)
    if not LIB_PATH.exists() or LIB_PATH.stat().st_mtime < newest:
        cmd = [_nvcc(), "-shared", "-o", str(LIB_PATH), *map(str, objs), "-gencode", "arch=compute_100a,code=sm_100a",
               "-Xcompiler", "-fPIC", "-cudart", "shared"]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError(f"link failed:\n{res.stderr}\n{res.stdout}")
        if verbose:
            sys.stderr.write(f"[b200clip build] linked {LIB_PATH}\n")
    return LIB_PATH


if __name__ == "__main__":
    build(verbose=True, force="--force" in sys.argv)
    print(LIB_PATH)
