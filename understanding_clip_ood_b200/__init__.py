"""B200-native CLIP hot path: the host-side mirror of the reference's `open_clip` / `xclip` API (`open_clip/`, `xclip/`) over
the C ABI of `include/b200clip.h` (`_lib.py`: ctypes binding, `ops.py`: one Python function per entry point), implemented
by the hand-written sm_100a kernels under `csrc/` and built in-tree into `libb200clip.so` by `build.py`.

Importing works without a GPU; calling an encoder or an op with CPU tensors raises `B200ClipError` — there is no CPU path.
"""

__version__ = "0.1.0"
