"""ctypes binding of libb200clip.so (include/b200clip.h).

There is NO fallback: if the shared library is missing or a call fails, this module raises.  The product
path never routes around the CUDA extension (no PyTorch-op / CPU substitute).
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import torch

import os

_PKG = Path(__file__).resolve().parent
# B200CLIP_LIB=<path>: load another build of the same ABI (A/B measurements of kernel variants only)
LIB_PATH = Path(os.environ["B200CLIP_LIB"]).resolve() if os.environ.get("B200CLIP_LIB") else _PKG / "libb200clip.so"

F32, BF16, F16 = 0, 1, 2
EPI_BIAS, EPI_GELU, EPI_QUICKGELU, EPI_RESIDUAL, EPI_PATCH, EPI_RELU, EPI_RESIDUAL_RELU = 0, 1, 2, 3, 4, 5, 6
STAGE_INPUT, STAGE_BODY, STAGE_OUTPUT = 1, 2, 4

_DTYPE_CODE = {torch.float32: F32, torch.bfloat16: BF16, torch.float16: F16}


class B200ClipError(RuntimeError):
    pass


class TowerCfg(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "dtype", "width", "layers", "heads", "mlp_width", "embed_dim", "seq_len", "quick_gelu",
        "image_size", "patch_size", "patch_kpad", "vocab_size", "fold_ln")]


class BlockWeights(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in (
        "ln1_g", "ln1_b", "ln2_g", "ln2_b", "in_proj_w", "in_proj_b", "out_proj_w", "out_proj_b",
        "fc_w", "fc_b", "proj_w", "proj_b", "in_proj_wf", "in_proj_c", "in_proj_bf", "fc_wf", "fc_c", "fc_bf")]


class VitWeights(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in (
        "conv1_w", "class_emb", "pos_emb", "ln_pre_g", "ln_pre_b", "ln_post_g", "ln_post_b", "proj_t",
        "blocks_host", "pos_cls")]


class TextWeights(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in (
        "tok_emb", "pos_emb", "ln_final_g", "ln_final_b", "proj_t", "blocks_host")]


class ResnetCfg(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("dtype", "image_size", "width", "embed_dim", "heads", "n_blocks", "stem_kpad")]


class ResnetBlock(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("conv1_w", "conv1_b", "conv2_w", "conv2_b", "conv3_w", "conv3_b", "down_w", "down_b")] + \
               [(n, C.c_int32) for n in ("cin", "planes", "stride", "reserved")]


class ResnetWeights(C.Structure):
    _fields_ = [("stem_w", C.c_void_p * 3), ("stem_b", C.c_void_p * 3), ("blocks_host", C.c_void_p), ("pos", C.c_void_p),
                ("qkv_w", C.c_void_p), ("qkv_b", C.c_void_p), ("c_proj_w", C.c_void_p), ("c_proj_b", C.c_void_p)]


class BlockGrads(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in (
        "ln1_g", "ln1_b", "ln2_g", "ln2_b", "in_proj_w", "in_proj_b", "out_proj_w", "out_proj_b", "fc_w", "fc_b", "proj_w", "proj_b")]


class VitGrads(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in (
        "conv1_w", "class_emb", "pos_emb", "ln_pre_g", "ln_pre_b", "ln_post_g", "ln_post_b", "proj", "blocks_host")]


class TextGrads(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("tok_emb", "pos_emb", "ln_final_g", "ln_final_b", "proj", "blocks_host")]


class AdamWTensor(C.Structure):
    _fields_ = [("param", C.c_void_p), ("grad", C.c_void_p), ("exp_avg", C.c_void_p), ("exp_avg_sq", C.c_void_p), ("count", C.c_int64),
                ("param_dtype", C.c_int32), ("grad_dtype", C.c_int32)]


class CastTensor(C.Structure):
    _fields_ = [("src", C.c_void_p), ("dst", C.c_void_p), ("count", C.c_int64), ("src_dtype", C.c_int32), ("dst_dtype", C.c_int32)]


_P, _I, _L, _F = C.c_void_p, C.c_int, C.c_int64, C.c_float

# name -> (restype, argtypes); must list every symbol include/b200clip.h declares (tests check this)
SIGNATURES = {
    "b200clip_version": (C.c_int, []),
    "b200clip_last_error": (C.c_char_p, []),
    "b200clip_launch_count": (C.c_uint64, []),
    "b200clip_gemm": (C.c_int, [_I, _P, _L, _P, _L, _P, _P, _L, _P, _L, _I, _I, _I, _I, _P, _I, _I, _P]),
    "b200clip_gemm_workspace_bytes": (C.c_int64, []),
    "b200clip_gemm_mn": (C.c_int, [_I, _P, _L, _I, _P, _L, _P, _L, _I, _I, _I, _P, _L, _P]),
    "b200clip_patch_embed_implicit": (C.c_int, [_I, _P, _P, _P, _P, _I, _I, _I, _I, _P]),
    "b200clip_gemm_ws": (C.c_int, [_I, _P, _L, _P, _L, _P, _P, _L, _P, _L, _I, _I, _I, _I, _P, _L, _P]),
    "b200clip_gemm_ln_ws": (C.c_int, [_I, _P, _L, _P, _L, _P, _P, _P, _P, _L, _I, _I, _I, _I, _P, _L, _P]),
    "b200clip_gemm_ln": (C.c_int, [_I, _P, _L, _P, _L, _P, _P, _P, _P, _L, _I, _I, _I, _I, _P]),
    "b200clip_gemm_stats_slots": (C.c_int, [_I, _I]),
    "b200clip_gemm_residual_stats": (C.c_int, [_I, _P, _L, _P, _L, _P, _P, _L, _P, _L, _I, _I, _I, _P, _P]),
    "b200clip_gemm_ln_partials": (C.c_int, [_I, _P, _L, _P, _L, _P, _P, _P, _I, _F, _P, _L, _I, _I, _I, _I, _P]),
    "b200clip_row_stats": (C.c_int, [_I, _P, _L, _P, _I, _I, _F, _P]),
    "b200clip_layernorm": (C.c_int, [_I, _P, _L, _P, _P, _P, _L, _I, _I, _F, _I, _P, _P]),
    "b200clip_attention": (C.c_int, [_I, _P, _P, _I, _I, _I, _I, _P]),
    "b200clip_patchify": (C.c_int, [_I, _P, _P, _I, _I, _I, _I, _P, _P, _P, _I, _P]),
    "b200clip_text_embed": (C.c_int, [_I, _P, _I, _P, _P, _P, _P, _I, _I, _I, _P]),
    "b200clip_eot_argmax": (C.c_int, [_P, _I, _P, _I, _P]),
    "b200clip_normalize": (C.c_int, [_I, _P, _L, _P, _L, _I, _I, _F, _P]),
    "b200clip_zeroshot": (C.c_int, [_I, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _F, _P]),
    "b200clip_topk": (C.c_int, [_I, _P, _L, _I, _I, _I, _P, _P, _P, _P]),
    "b200clip_class_mean": (C.c_int, [_I, _P, _P, _I, _I, _I, _P]),
    "b200clip_cliploss": (C.c_int, [_P, _P, _P, _P, _P, _I, _I, _I, _I, _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    "b200clip_cliploss_forward": (C.c_int, [_P, _P, _P, _P, _P, _I, _I, _I, _I, _P, _P, _P]),
    "b200clip_cliploss_backward": (C.c_int, [_P, _P, _P, _P, _P, _I, _I, _I, _I, _P, _P, _P, _P, _P, _P, _P, _P]),
    "b200clip_cliploss_single_backward": (C.c_int, [_P, _P, _P, _I, _I, _P, _P, _P, _P, _P, _P]),
    "b200clip_cliploss_packed_forward": (C.c_int, [_P, _P, _I, _I, _I, _I, _P, _P, _P]),
    "b200clip_cliploss_packed_backward": (C.c_int, [_P, _P, _I, _I, _I, _I, _P, _P, _P, _P, _P]),
    "b200clip_p2p_configure": (C.c_int, [C.c_double, _P]),
    "b200clip_p2p_allgather": (C.c_int, [_I, _P, _P, _I, _I, _P, _P, _P, _P, _I, C.c_uint32, _P, _P, _I, _P]),
    "b200clip_cliploss_packed_backward_p2p": (C.c_int, [_P, _P, _I, _I, _I, _I, _P, _P, _P, _P, _P]),
    "b200clip_p2p_reduce_finish": (C.c_int, [_P, _P, _L, _P, _P, _I, _I, C.c_uint32, _P, _I, _P]),
    "b200clip_workspace_bytes": (C.c_int64, [C.POINTER(TowerCfg), _I, _I]),
    "b200clip_vit_forward": (C.c_int, [C.POINTER(TowerCfg), C.POINTER(VitWeights), _P, _P, _I, _I, _P, _L, _P]),
    "b200clip_vit_forward_u8": (C.c_int, [C.POINTER(TowerCfg), C.POINTER(VitWeights), _P, C.POINTER(C.c_float), C.POINTER(C.c_float), _P, _I,
                                          _I, _P, _L, _P]),
    "b200clip_patchify_u8": (C.c_int, [_I, _P, C.POINTER(C.c_float), C.POINTER(C.c_float), _P, _I, _I, _I, _I, _P]),
    "b200clip_text_forward": (C.c_int, [C.POINTER(TowerCfg), C.POINTER(TextWeights), _P, _P, _I, _I, _I, _P, _L, _P]),
    "b200clip_vit_forward_stages": (C.c_int, [C.POINTER(TowerCfg), C.POINTER(VitWeights), _P, _P, C.POINTER(C.c_float), C.POINTER(C.c_float),
                                              _P, _I, _I, _P, _L, _I, _P]),
    "b200clip_text_forward_stages": (C.c_int, [C.POINTER(TowerCfg), C.POINTER(TextWeights), _P, _P, _I, _I, _I, _P, _L, _I, _P]),
    "b200clip_resnet_workspace_bytes": (C.c_int64, [C.POINTER(ResnetCfg), C.POINTER(ResnetWeights), _I]),
    "b200clip_resnet_forward_stages": (C.c_int, [C.POINTER(ResnetCfg), C.POINTER(ResnetWeights), _P, _P, _I, _I, _P, _L, _I, _P]),
    "b200clip_stem_im2col": (C.c_int, [_I, _P, _P, _I, _I, _I, _P]),
    "b200clip_im2col3x3": (C.c_int, [_I, _P, _P, _I, _I, _I, _I, _P]),
    "b200clip_avgpool2": (C.c_int, [_I, _P, _P, _I, _I, _I, _I, _P]),
    "b200clip_attnpool_tokens": (C.c_int, [_I, _P, _P, _P, _I, _I, _I, _P]),
    "b200clip_resize_crop_u8": (C.c_int, [_P, _I, _I, _L, _P, _P, _I, _P, _P, _I, _I, _I, _P, _P, _I, _I, _P]),
    "b200clip_train_saved_bytes": (C.c_int64, [C.POINTER(TowerCfg), _I, _I]),
    "b200clip_backward_workspace_bytes": (C.c_int64, [C.POINTER(TowerCfg), _I, _I]),
    "b200clip_vit_forward_train": (C.c_int, [C.POINTER(TowerCfg), C.POINTER(VitWeights), _P, _P, _I, _I, _P, _L, _P, _L, _P]),
    "b200clip_text_forward_train": (C.c_int, [C.POINTER(TowerCfg), C.POINTER(TextWeights), _P, _P, _I, _I, _I, _P, _L, _P, _L, _P]),
    "b200clip_vit_backward": (C.c_int, [C.POINTER(TowerCfg), C.POINTER(VitWeights), _P, _P, _I, _I, _P, C.POINTER(VitGrads), _P, _L, _P]),
    "b200clip_text_backward": (C.c_int, [C.POINTER(TowerCfg), C.POINTER(TextWeights), _P, _P, _I, _I, _I, _P, C.POINTER(TextGrads), _P, _L, _P]),
    "b200clip_adamw_chunk": (C.c_int, []),
    "b200clip_multi_cast": (C.c_int, [_P, _P, _P, _I, _P]),
    "b200clip_adamw_step": (C.c_int, [_P, _P, _P, _I, _F, _F, _F, _F, _F, _I, _F, _P]),
}
# not part of the public header: test hook that forces the GEMM N-tile
_EXTRA = {
    "b200clip_gemm_tile": (C.c_int, [_I, _P, _L, _P, _L, _P, _P, _L, _P, _L, _I, _I, _I, _I, _P, _I, _I, _I, _P]),
}

_lib = None


def load() -> C.CDLL:
    """Load libb200clip.so; raises B200ClipError (never falls back) when it is absent or incomplete."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise B200ClipError(
            f"{LIB_PATH} is missing: build it with `python -m understanding_clip_ood_b200.build` "
            "(there is no CPU / PyTorch fallback for this path)")
    lib = C.CDLL(str(LIB_PATH))
    for name, (res, args) in {**SIGNATURES, **_EXTRA}.items():
        try:
            fn = getattr(lib, name)
        except AttributeError as e:
            raise B200ClipError(f"{LIB_PATH} does not export {name}") from e
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().b200clip_last_error().decode(errors="replace")
        raise B200ClipError(f"{what} failed (rc={rc}): {msg}")


def dtype_code(t: torch.dtype) -> int:
    try:
        return _DTYPE_CODE[t]
    except KeyError:
        raise B200ClipError(f"unsupported dtype {t}") from None


def ptr(t) -> int | None:
    return None if t is None else t.data_ptr()


def stream_ptr() -> int:
    """Raw cudaStream_t of torch's current stream on the current device (the C-level query: `torch.cuda.current_stream()`
    costs ~10 us of Python per call, which showed up as a sixth of the ClipLoss step)."""
    return torch._C._cuda_getCurrentRawStream(torch._C._cuda_getDevice())


def require_cuda(*tensors) -> None:
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise B200ClipError("b200clip kernels need CUDA tensors (no CPU fallback exists for this path)")


_replayed_launches = 0


def note_replayed(n: int) -> None:
    """Kernels executed by replaying a captured CUDA graph (the C-side counter only sees them once, at capture)."""
    global _replayed_launches
    _replayed_launches += n


def launch_count() -> int:
    """Kernels of libb200clip.so launched by this process: direct launches + kernels inside replayed CUDA graphs."""
    return int(load().b200clip_launch_count()) + _replayed_launches
