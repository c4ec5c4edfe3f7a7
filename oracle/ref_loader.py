"""Import the UNMODIFIED reference (/root/reference) in the build container — TEST INFRASTRUCTURE ONLY.

Used by `oracle/make_golden.py` (fixture generation) and by the CPU tests that pin the oracle / the host-side
mirror against the live reference.  `/root/reference` does not exist on the GPU box: nothing that runs there
may call `load()`; callers must check `available()` first.

Two shims, neither of which touches reference code:
  * `ftfy` (hard import of open_clip/tokenizer.py:14, not installed here) -> identity `fix_text` stub;
  * `xclip/__init__.py` imports datasets/learner (textacy, lightning: not installed) -> register an empty
    namespace package so `xclip.zero_shot` / `xclip.open_clip` / `xclip.utils` import on their own.
"""
from __future__ import annotations

import sys
import types
from pathlib import Path

REFERENCE = Path("/root/reference")
_OPEN_CLIP_SRC = REFERENCE / "deps" / "open_clip" / "src"


def available() -> bool:
    return (_OPEN_CLIP_SRC / "open_clip" / "model.py").exists()


def load():
    """-> (open_clip, xclip.zero_shot, xclip.open_clip) modules of the reference."""
    if not available():
        raise RuntimeError("/root/reference is not present (the reference only exists in the build container)")
    if "ftfy" not in sys.modules:
        try:
            import ftfy  # noqa: F401
        except ImportError:
            stub = types.ModuleType("ftfy")
            stub.fix_text = lambda s: s
            sys.modules["ftfy"] = stub
    if str(_OPEN_CLIP_SRC) not in sys.path:
        sys.path.insert(0, str(_OPEN_CLIP_SRC))
    import open_clip  # the vendored 2.24.0
    if "xclip" not in sys.modules:
        pkg = types.ModuleType("xclip")
        pkg.__path__ = [str(REFERENCE / "xclip")]
        sys.modules["xclip"] = pkg
    import xclip.open_clip as xo
    import xclip.zero_shot as zs
    return open_clip, zs, xo
