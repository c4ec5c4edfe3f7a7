"""Build an A/B variant of libb200clip.so: recompile the named sources with extra -D switches, link them with the objects of
the regular build into tmp_variants/libb200clip_<name>.so (git-ignored; travels to the GPU box), to be selected with
B200CLIP_LIB=tmp_variants/libb200clip_<name>.so.   python tools/build_variant.py <name> "<-DX=1 ...>" [source.cu ...]"""
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from understanding_clip_ood_b200 import build as B  # noqa: E402


def main():
    name, defs = sys.argv[1], sys.argv[2].split()
    srcs = sys.argv[3:] or ["gemm_pair.cu"]
    out_dir = ROOT / "tmp_variants"
    out_dir.mkdir(exist_ok=True)
    B.build(verbose=False)           # the regular objects must exist
    objs = []
    for src in sorted(B.CSRC.glob("*.cu")):
        if src.name in srcs:
            obj = out_dir / f"{src.stem}_{name}.o"
            cmd = [B._nvcc(), *B.NVCC_FLAGS, *defs, "-c", str(src), "-o", str(obj)]
            res = subprocess.run(cmd, capture_output=True, text=True)
            (out_dir / f"{src.stem}_{name}.ptxas.log").write_text(res.stderr)
            if res.returncode != 0:
                sys.exit(res.stderr)
            objs.append(obj)
        else:
            objs.append(B.BUILD_DIR / (src.stem + ".o"))
    lib = out_dir / f"libb200clip_{name}.so"
    cmd = [B._nvcc(), "-shared", "-o", str(lib), *map(str, objs), "-gencode", "arch=compute_100a,code=sm_100a", "-Xcompiler", "-fPIC",
           "-cudart", "shared"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.exit(res.stderr)
    print(lib)


if __name__ == "__main__":
    main()
