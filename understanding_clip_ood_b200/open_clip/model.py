"""Host-side CLIP module: same constructor arguments, attributes, call signatures and `state_dict` keys as the
reference's `open_clip.model.CLIP` (deps/open_clip/src/open_clip/model.py:220-315) and its `VisionTransformer`
/ `TextTransformer` (transformer.py:427-643, 661-802), but every forward is ONE call into libb200clip.so
(`b200clip_vit_forward` / `b200clip_text_forward`).  The nn.Module tree below only holds parameters — it
contains no PyTorch arithmetic and there is no CPU / eager fallback: calling an encoder with CPU tensors or
without the built library raises.

Training: a tower called in training mode with autograd enabled returns features that carry an autograd
node (`open_clip/train.py`) whose backward is ONE call into the library (`b200clip_vit_backward` / `b200clip_text_backward`,
block recompute as with --grad-checkpointing).  Precision modes fp32 / bf16 / fp16 / pure_bf16 / pure_fp16, and the
mixed-precision `amp` / `amp_bf16` / `amp_bfloat16` modes as fp32 master parameters with 16-bit compute (`compute_dtype`).
"""
from __future__ import annotations

import ctypes as C
import math
import os
from typing import Optional

import numpy as np
import warnings

import torch
from torch import nn

OPENAI_DATASET_MEAN = (0.48145466, 0.4578275, 0.40821073)     # open_clip/constants.py:1-2
OPENAI_DATASET_STD = (0.26862954, 0.26130258, 0.27577711)

from .. import _lib as L

__all__ = ["CLIP", "VisionTower", "convert_weights_to_lp", "get_cast_dtype", "get_input_dtype"]


# ------------------------------------------------------------------------------------------------
# parameter containers (names chosen so that state_dict keys equal the reference's)
# ------------------------------------------------------------------------------------------------
class _Affine(nn.Module):
    """LayerNorm parameters (ln_pre / ln_1 / ln_2 / ln_post / ln_final)."""

    def __init__(self, width: int):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(width))
        self.bias = nn.Parameter(torch.zeros(width))


class _Dense(nn.Module):
    """nn.Linear-shaped parameters: weight [out, in], bias [out]."""

    def __init__(self, fan_in: int, fan_out: int, bias: bool = True):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(fan_out, fan_in))
        self.bias = nn.Parameter(torch.empty(fan_out)) if bias else None


class _PatchConv(nn.Module):
    """conv1: weight [W, 3, P, P], no bias (transformer.py:461)."""

    def __init__(self, width: int, patch: int):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(width, 3, patch, patch))


class _TokenTable(nn.Module):
    def __init__(self, vocab: int, width: int):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(vocab, width))


class _PackedMHA(nn.Module):
    """nn.MultiheadAttention parameter layout: in_proj_weight [3W, W] rows ordered q|k|v, out_proj."""

    def __init__(self, width: int):
        super().__init__()
        self.in_proj_weight = nn.Parameter(torch.empty(3 * width, width))
        self.in_proj_bias = nn.Parameter(torch.zeros(3 * width))
        self.out_proj = _Dense(width, width)


class _MLP(nn.Module):
    def __init__(self, width: int, hidden: int):
        super().__init__()
        self.c_fc = _Dense(width, hidden)
        self.c_proj = _Dense(hidden, width)


class _Block(nn.Module):
    def __init__(self, width: int, hidden: int):
        super().__init__()
        self.ln_1 = _Affine(width)
        self.attn = _PackedMHA(width)
        self.ln_2 = _Affine(width)
        self.mlp = _MLP(width, hidden)


class _Stack(nn.Module):
    """`transformer` sub-module: resblocks.{i}.* keys; mirrors Transformer's public attributes."""

    def __init__(self, width: int, layers: int, heads: int, mlp_ratio: float):
        super().__init__()
        self.width, self.layers, self.heads = width, layers, heads
        self.mlp_width = int(width * mlp_ratio)
        self.grad_checkpointing = False
        self.resblocks = nn.ModuleList([_Block(width, self.mlp_width) for _ in range(layers)])

    def get_cast_dtype(self) -> torch.dtype:
        return self.resblocks[0].mlp.c_fc.weight.dtype


# ------------------------------------------------------------------------------------------------
# random init: same distributions, and the same draw ORDER from the global torch RNG, as constructing the
# reference modules (nn.Conv2d / nn.MultiheadAttention / nn.Linear / nn.Embedding default resets followed by
# TextTransformer.init_parameters, transformer.py:724-745), so `torch.manual_seed(s); create_model(...)` yields
# the reference's weights for the same seed on CPU.
# ------------------------------------------------------------------------------------------------
def _uniform_(p: torch.Tensor, bound: float) -> None:
    with torch.no_grad():
        p.uniform_(-bound, bound)


def _weight_bound(fan_in: int) -> float:
    """Bound of nn.Linear / nn.Conv2d's default weight reset, kaiming_uniform_(a=sqrt(5)) (~ 1/sqrt(fan_in)); written
    with torch.nn.init's own operation order so the double-precision bound is bit-identical."""
    gain = math.sqrt(2.0 / (1 + math.sqrt(5) ** 2))
    std = gain / math.sqrt(fan_in)
    return math.sqrt(3.0) * std


def _bias_bound(fan_in: int) -> float:
    return 1 / math.sqrt(fan_in)


def _xavier_bound(fan_in: int, fan_out: int) -> float:
    std = 1.0 * math.sqrt(2.0 / float(fan_in + fan_out))
    return math.sqrt(3.0) * std


def _init_stack(stack: _Stack) -> None:
    W, Hd = stack.width, stack.mlp_width
    for blk in stack.resblocks:
        # nn.MultiheadAttention.__init__: out_proj (Linear reset) is built first, then _reset_parameters()
        _uniform_(blk.attn.out_proj.weight, _weight_bound(W))
        _uniform_(blk.attn.out_proj.bias, _bias_bound(W))
        _uniform_(blk.attn.in_proj_weight, _xavier_bound(W, 3 * W))
        with torch.no_grad():
            blk.attn.in_proj_bias.zero_()
            blk.attn.out_proj.bias.zero_()
        _uniform_(blk.mlp.c_fc.weight, _weight_bound(W))
        _uniform_(blk.mlp.c_fc.bias, _bias_bound(W))
        _uniform_(blk.mlp.c_proj.weight, _weight_bound(Hd))
        _uniform_(blk.mlp.c_proj.bias, _bias_bound(Hd))


# ------------------------------------------------------------------------------------------------
# C-ABI plumbing shared by both towers
# ------------------------------------------------------------------------------------------------
class _Engine:
    """Builds and caches the ctypes weight structs + workspace for one tower; rebuilt when any parameter changes."""

    def __init__(self):
        self.sig = None
        self.static_sig = None  # the part of the signature that does not change with an optimizer step
        self.keep = []          # tensors whose storage the structs point into
        self.blocks = None
        self.weights = None
        self.cfg = None
        self.ws = None          # uint8 workspace tensor
        self.graphs = {}        # key -> (torch.cuda.CUDAGraph, static output)
        self.seen = set()

    def __deepcopy__(self, memo):      # caches are per-instance and never copied / pickled with the module
        return _Engine()

    def __getstate__(self):
        return {}

    def __setstate__(self, state):
        self.__init__()

    @staticmethod
    def signature(params, *switches) -> tuple:
        """Everything the cached structs / captured graphs depend on: parameter storage, version and dtype, plus the public
        switches that are baked into TowerCfg and the folded weights (fold_layernorm, quick_gelu, ...)."""
        return _Engine.signatures(params, *switches)[0]

    @staticmethod
    def signatures(params, *switches) -> tuple:
        """-> (signature, static part).  The static part (storages, dtypes, switches) does not change with an optimizer step, the
        full signature adds the version counters.  One pass per attribute over the parameter list (this runs on EVERY encode call:
        ~50 us for the 150 tensors of a tower instead of ~125 us for per-parameter tuples built twice)."""
        static_sig = (tuple([p.data_ptr() for p in params]), tuple([p.dtype for p in params])) + tuple(switches)
        return (tuple([p._version for p in params]),) + static_sig, static_sig

    def try_refresh(self, params, static_sig: tuple) -> bool:
        """Same storages, dtypes and switches as at build time, only newer values (an optimizer step): update the converted
        copies and the derived operands in place — the structs, and any captured graph, stay valid.  Not for folded
        LayerNorms (their operands are re-derived by a full rebuild)."""
        keep = self.keep
        if self.static_sig != static_sig or not isinstance(keep, _Keep) or getattr(self.cfg, "fold_ln", 0):
            return False
        with torch.no_grad():
            if keep.pairs:
                keep.cast_pairs()
            for fn in keep.refresh:
                fn()
        self.sig = (tuple([p._version for p in params]),) + static_sig
        return True

    def workspace(self, nbytes: int, device) -> torch.Tensor:
        if self.ws is None or self.ws.numel() < nbytes or self.ws.device != device:
            self.ws = torch.empty(nbytes, dtype=torch.uint8, device=device)
            self.graphs.clear()
        return self.ws

    # A tower forward only enqueues kernels on the caller's stream (no allocation, no synchronisation).  It is issued in
    # three stages (include/b200clip.h, B200CLIP_STAGE_*): the input stage (the one kernel that reads the caller's batch)
    # and the output stage (projection + normalise, the kernels that write the result tensor) are launched directly, and
    # the ~85 launches in between — which touch nothing but the workspace — are captured into a CUDA graph the SECOND
    # time a (batch, sequence length, workspace) shape is seen and replayed afterwards.  The graph therefore does not
    # depend on the input or output pointers: a DataLoader loop that hands over a fresh tensor per batch
    # (scripts/evaluate_domainnet_lso_openai.py:18-36) replays it like a loop over one static buffer does, and the
    # result lands in a fresh tensor without a copy.
    MAX_GRAPHS = 8

    def run_staged(self, key, enqueue, out_shape, dtype, device) -> torch.Tensor:
        """enqueue(stages, out): issue the given stage mask on the current stream.  -> fresh output tensor."""
        out = torch.empty(out_shape, dtype=dtype, device=device)
        entry = self.graphs.get(key)
        if entry is None:
            if key not in self.seen:
                if len(self.seen) > 64:
                    self.seen.clear()
                self.seen.add(key)
                enqueue(L.STAGE_INPUT | L.STAGE_BODY | L.STAGE_OUTPUT, out)
                return out
            if len(self.graphs) >= self.MAX_GRAPHS:
                self.graphs.pop(next(iter(self.graphs)))
            graph = torch.cuda.CUDAGraph()
            torch.cuda.synchronize(device)
            n0 = L.launch_count()
            with torch.cuda.graph(graph, capture_error_mode="thread_local"):
                enqueue(L.STAGE_BODY, None)
            entry = (graph, L.launch_count() - n0)    # capture records the kernels without running them
            L.note_replayed(-entry[1])
            self.graphs[key] = entry
        graph, n_kernels = entry
        enqueue(L.STAGE_INPUT, None)
        graph.replay()
        L.note_replayed(n_kernels)
        enqueue(L.STAGE_OUTPUT, out)
        return out


class _Keep(list):
    """Tensors the ctypes structs point into.  `pairs` = (copy, source parameter) for every parameter that had to be converted,
    `refresh` = callables that recompute derived operands in place: after an optimizer step (same storages, new values) the
    engine re-fills the copies with one multi-tensor copy instead of rebuilding everything (see _Engine.try_refresh)."""

    def __init__(self):
        super().__init__()
        self.pairs: list = []
        self.refresh: list = []
        self._cast_table = None

    def cast_pairs(self) -> None:
        """copy <- source for every pair with ONE launch (`b200clip_multi_cast`; torch._foreach_copy_ issues one kernel per tensor
        when the dtypes differ: 199 launches per optimizer step for ViT-B/32)."""
        pairs = [(d, s) for d, s in self.pairs if d.numel() > 0]
        if not pairs:
            return
        if os.environ.get("B200CLIP_FOREACH_REFRESH") == "1":     # A/B: the per-tensor copies of torch._foreach_copy_
            torch._foreach_copy_([d for d, _ in pairs], [s_.detach() for _, s_ in pairs])
            return
        key = tuple((d.data_ptr(), s.data_ptr(), d.dtype, s.dtype, d.numel()) for d, s in pairs)
        if self._cast_table is None or self._cast_table[0] != key:
            if any(not d.is_contiguous() or not s.is_contiguous() or d.numel() != s.numel() or d.device != s.device for d, s in pairs):
                torch._foreach_copy_([d for d, _ in pairs], [s_.detach() for _, s_ in pairs])
                return
            dev = pairs[0][0].device
            chunk = int(L.load().b200clip_adamw_chunk())
            items = (L.CastTensor * len(pairs))()
            chunk_item, chunk_off = [], []
            for i, (d, s) in enumerate(pairs):
                it = items[i]
                it.src, it.dst, it.count = s.data_ptr(), d.data_ptr(), d.numel()
                it.src_dtype, it.dst_dtype = L.dtype_code(s.dtype), L.dtype_code(d.dtype)
                for off in range(0, d.numel(), chunk):
                    chunk_item.append(i)
                    chunk_off.append(off)
            raw = torch.frombuffer(bytearray(bytes(items)), dtype=torch.uint8).to(dev)
            self._cast_table = (key, raw, torch.tensor(chunk_item, dtype=torch.int32, device=dev),
                                torch.tensor(chunk_off, dtype=torch.int64, device=dev))
        _, raw, ci, co = self._cast_table
        with torch.cuda.device(raw.device):
            L.check(L.load().b200clip_multi_cast(raw.data_ptr(), ci.data_ptr(), co.data_ptr(), ci.numel(), L.stream_ptr()), "b200clip_multi_cast")


def _f32(t: torch.Tensor, keep: list) -> int:
    if t.dtype != torch.float32 or not t.is_contiguous():
        src = t
        t = t.detach().to(torch.float32).contiguous()
        if isinstance(keep, _Keep):
            keep.pairs.append((t, src))
    keep.append(t)
    return t.data_ptr()


def _as(t: Optional[torch.Tensor], dtype: torch.dtype, keep: list) -> Optional[int]:
    if t is None:
        return None
    if t.dtype != dtype or not t.is_contiguous():
        src = t
        t = t.detach().to(dtype).contiguous()
        if isinstance(keep, _Keep):
            keep.pairs.append((t, src))
    keep.append(t)
    return t.data_ptr()


def _pack_blocks(stack: _Stack, dtype: torch.dtype, keep: list, fold_ln: bool = False):
    from .. import ops
    arr = (L.BlockWeights * stack.layers)()
    for i, blk in enumerate(stack.resblocks):
        b = arr[i]
        if fold_ln:
            for dst, w, bias, ln in (("in_proj", blk.attn.in_proj_weight, blk.attn.in_proj_bias, blk.ln_1),
                                     ("fc", blk.mlp.c_fc.weight, blk.mlp.c_fc.bias, blk.ln_2)):
                wf, cs, bf = ops.fold_layernorm(w, bias, ln.weight, ln.bias, dtype)
                keep.extend((wf, cs, bf))
                setattr(b, dst + "_wf", wf.data_ptr())
                setattr(b, dst + "_c", cs.data_ptr())
                setattr(b, dst + "_bf", bf.data_ptr())
        b.ln1_g, b.ln1_b = _f32(blk.ln_1.weight, keep), _f32(blk.ln_1.bias, keep)
        b.ln2_g, b.ln2_b = _f32(blk.ln_2.weight, keep), _f32(blk.ln_2.bias, keep)
        b.in_proj_w, b.in_proj_b = _as(blk.attn.in_proj_weight, dtype, keep), _as(blk.attn.in_proj_bias, dtype, keep)
        b.out_proj_w, b.out_proj_b = _as(blk.attn.out_proj.weight, dtype, keep), _as(blk.attn.out_proj.bias, dtype, keep)
        b.fc_w, b.fc_b = _as(blk.mlp.c_fc.weight, dtype, keep), _as(blk.mlp.c_fc.bias, dtype, keep)
        b.proj_w, b.proj_b = _as(blk.mlp.c_proj.weight, dtype, keep), _as(blk.mlp.c_proj.bias, dtype, keep)
    return arr


def _resolve_compute_dtype(override, param_dtype: torch.dtype) -> torch.dtype:
    """Kernel dtype of a tower: an explicit `compute_dtype` (models created with an `amp*` precision), else the autocast dtype
    when fp32 parameters are called inside a torch.autocast context (what the reference's train loop opens), else the
    parameter dtype."""
    if override is not None:
        return override
    if param_dtype == torch.float32 and torch.is_autocast_enabled():
        return torch.get_autocast_gpu_dtype()
    return param_dtype


class _ParamSnapshot:
    """The parameter list of a module tree without walking the tree on every call.  `list(module.parameters())` costs ~0.4 ms for a
    150-tensor tower — a quarter of a 128-image forward, paid on the host on every encode call (twice: forward + engine check).
    The snapshot keeps the list together with every (`_modules` / `_parameters` dict, key, object) triple it was collected from
    and re-validates those identities (and the dict sizes) per call, ~40 us: a replaced Parameter or sub-module, or a new one, is
    seen exactly as a fresh traversal would see it; values / storages / versions are covered by the engine signature."""

    def __init__(self, root: nn.Module, buffers: bool = False):
        self.root, self.with_buffers = root, buffers
        self.params = None

    def __getstate__(self):
        return {"root": self.root, "with_buffers": self.with_buffers}

    def __setstate__(self, state):
        self.root, self.with_buffers = state["root"], state["with_buffers"]
        self.params = None

    def _collect(self) -> None:
        slots, lens = [], []
        for m in self.root.modules():
            dicts = (m._modules, m._parameters, m._buffers) if self.with_buffers else (m._modules, m._parameters)
            for d in dicts:
                lens.append((d, len(d)))
                for k, v in d.items():
                    slots.append((d, k, v))
        self._slots, self._lens = slots, lens
        self.named = list(self.root.named_parameters())
        self.params = [p for _, p in self.named]
        self.names = [n for n, _ in self.named]
        self.tensors = self.params + list(self.root.buffers()) if self.with_buffers else self.params

    def _valid(self) -> bool:
        for d, n in self._lens:
            if len(d) != n:
                return False
        for d, k, v in self._slots:
            if d.get(k) is not v:
                return False
        return True

    def refresh(self) -> "_ParamSnapshot":
        if self.params is None or not self._valid():
            self._collect()
        return self


def _snapshot(module: nn.Module, buffers: bool = False) -> _ParamSnapshot:
    snap = module.__dict__.get("_b200clip_params")
    if snap is None:
        snap = _ParamSnapshot(module, buffers)
        object.__setattr__(module, "_b200clip_params", snap)    # not a sub-module, not a parameter, not in the state_dict
    return snap.refresh()


def _wants_grad(module: nn.Module, params) -> bool:
    """Training path: the module is in training mode, autograd is recording and some parameter wants a gradient."""
    return module.training and torch.is_grad_enabled() and any(p.requires_grad for p in params)


def _check_device(t: torch.Tensor, what: str) -> None:
    if not t.is_cuda:
        raise L.B200ClipError(f"{what}: CUDA tensors required — this path has no CPU fallback (got {t.device})")


# ------------------------------------------------------------------------------------------------
# vision tower
# ------------------------------------------------------------------------------------------------
class VisionTower(nn.Module):
    """Parameters + driver of the ViT image tower; the `model.visual` object (VisionTransformer in the reference)."""

    def __init__(self, image_size: int, patch_size: int, width: int, layers: int, heads: int, mlp_ratio: float,
                 output_dim: int, quick_gelu: bool = False):
        super().__init__()
        if isinstance(image_size, (tuple, list)):
            if image_size[0] != image_size[1]:
                raise ValueError("only square images are supported")
            image_size = image_size[0]
        if image_size % patch_size != 0:
            raise ValueError(f"image_size {image_size} must be divisible by patch_size {patch_size}")
        if heads * 64 != width:
            raise ValueError(f"head_width must be 64 (width {width}, heads {heads}): the attention kernel is specialised for it")
        self.image_size = (image_size, image_size)
        self.patch_size = (patch_size, patch_size)
        self.grid_size = (image_size // patch_size, image_size // patch_size)
        self.output_dim = output_dim
        self.quick_gelu = bool(quick_gelu)
        self.pool_type = "tok"
        self.output_tokens = False
        scale = width ** -0.5
        self.conv1 = _PatchConv(width, patch_size)
        _uniform_(self.conv1.weight, _weight_bound(3 * patch_size * patch_size))
        self.class_embedding = nn.Parameter(scale * torch.randn(width))
        self.positional_embedding = nn.Parameter(scale * torch.randn(self.grid_size[0] * self.grid_size[1] + 1, width))
        self.ln_pre = _Affine(width)
        self.transformer = _Stack(width, layers, heads, mlp_ratio)
        _init_stack(self.transformer)
        self.ln_post = _Affine(width)
        self.proj = nn.Parameter(scale * torch.randn(width, output_dim))
        self._engine = _Engine()
        #: replay the forward as a CUDA graph when the same input buffer is presented again (see _Engine.run_graphed)
        self.use_cuda_graphs = True
        #: 16-bit modes: fold ln_1 / ln_2 into the QKV / c_fc GEMM epilogues (no normalised activations in HBM)
        self.fold_layernorm = True
        #: mixed precision (`amp*`): fp32 master parameters, kernels run in this dtype on 16-bit copies (None = parameter dtype)
        self.compute_dtype = None

    # -- reference API surface ---------------------------------------------------------------
    def set_grad_checkpointing(self, enable: bool = True):
        self.transformer.grad_checkpointing = enable      # the training path always recomputes per block (train.py)

    def lock(self, unlocked_groups: int = 0, freeze_bn_stats: bool = False):
        for p in self.parameters():
            p.requires_grad = False

    def _compute_dtype(self) -> torch.dtype:
        return _resolve_compute_dtype(self.compute_dtype, self.transformer.get_cast_dtype())

    def _build(self, device, for_training: bool = False, params=None) -> _Engine:
        if params is None:
            params = _snapshot(self).params
        dt = self._compute_dtype()
        # training path: no LayerNorm folding (the folded weights would have to be re-derived after every optimizer step)
        fold = bool(self.fold_layernorm) and dt != torch.float32 and not for_training
        sig, static_sig = _Engine.signatures(params, fold, bool(self.quick_gelu), dt)
        eng = self._engine
        if eng.sig == sig:
            return eng
        if eng.try_refresh(params, static_sig):
            return eng
        keep = _Keep()
        W = self.transformer.width
        P = self.patch_size[0]
        kreal = 3 * P * P
        # K of the patch GEMM: 16-byte rows in fp32 (P = 14 gives 3*P*P = 588, not a multiple of 8), whole 128-byte swizzle
        # rows in the 16-bit modes; the extra columns are zero in both operands
        kpad = (kreal + 7) // 8 * 8 if dt == torch.float32 else (kreal + 63) // 64 * 64
        conv = torch.zeros((W, kpad), dtype=dt, device=device)
        proj_t = torch.empty((self.output_dim, W), dtype=dt, device=device)

        def refresh_conv_proj():
            conv[:, :kreal].copy_(self.conv1.weight.detach().reshape(W, kreal))
            proj_t.copy_(self.proj.detach().t())

        keep.refresh.append(refresh_conv_proj)
        keep.extend((conv, proj_t))
        blocks = _pack_blocks(self.transformer, dt, keep, fold)
        w = L.VitWeights()
        w.conv1_w = conv.data_ptr()
        w.class_emb = _f32(self.class_embedding, keep)
        w.pos_emb = _f32(self.positional_embedding, keep)
        w.ln_pre_g, w.ln_pre_b = _f32(self.ln_pre.weight, keep), _f32(self.ln_pre.bias, keep)
        w.ln_post_g, w.ln_post_b = _f32(self.ln_post.weight, keep), _f32(self.ln_post.bias, keep)
        w.proj_t = proj_t.data_ptr()
        w.blocks_host = C.cast(blocks, C.c_void_p)
        if dt != torch.float32:
            # token-layout patch embedding: the class-token row of the additive table holds dtype(class_emb) + dtype(pos[0])
            # (an exact fp32 sum of two 16-bit values), the other rows are the fp32 positional embedding
            pos_cls = torch.empty_like(self.positional_embedding, dtype=torch.float32, device=device)

            def refresh_pos_cls():
                pos_cls.copy_(self.positional_embedding.detach())
                pos_cls[0] = self.class_embedding.detach().to(dt).float() + self.positional_embedding.detach()[0].to(dt).float()

            keep.refresh.append(refresh_pos_cls)
            keep.append(pos_cls)
            w.pos_cls = pos_cls.data_ptr()
        with torch.no_grad():
            for fn in keep.refresh:
                fn()
        cfg = L.TowerCfg(dtype=L.dtype_code(dt), width=W, layers=self.transformer.layers, heads=self.transformer.heads,
                         mlp_width=self.transformer.mlp_width, embed_dim=self.output_dim,
                         seq_len=self.grid_size[0] * self.grid_size[1] + 1, quick_gelu=int(self.quick_gelu),
                         image_size=self.image_size[0], patch_size=P, patch_kpad=kpad, vocab_size=0, fold_ln=int(fold))
        eng.sig, eng.static_sig, eng.keep, eng.blocks, eng.weights, eng.cfg = sig, static_sig, keep, blocks, w, cfg
        eng.graphs.clear()      # captured graphs hold the old weight pointers
        return eng

    def forward(self, image: torch.Tensor, normalize: bool = False) -> torch.Tensor:
        """[B,3,S,S] -> [B,D] (transformer.py:601-643); `normalize` fuses CLIP.encode_image's F.normalize."""
        _check_device(image, "encode_image")
        _check_device(self.proj, "encode_image (model weights)")
        dt = self._compute_dtype()
        u8 = image.dtype == torch.uint8
        if image.dtype == torch.float32 and dt != torch.float32 and self.transformer.get_cast_dtype() == torch.float32:
            image = image.to(dt)          # autocast's input cast (mixed-precision modes take fp32 batches, precision.py:5-12)
        if not u8 and image.dtype != dt:
            raise RuntimeError(f"Input type ({image.dtype}) and weight type ({dt}) should be the same "
                               f"(cast the batch with get_input_dtype(precision), as the reference requires; uint8 pixel "
                               f"batches are normalised on the GPU with visual.preprocess_cfg mean / std)")
        if image.ndim != 4 or image.shape[1] != 3 or tuple(image.shape[2:]) != self.image_size:
            raise RuntimeError(f"expected images of shape [B, 3, {self.image_size[0]}, {self.image_size[1]}], got {tuple(image.shape)}")
        image = image.contiguous()
        B = image.shape[0]
        if B == 0:
            return torch.empty((0, self.output_dim), dtype=dt, device=image.device)
        snap = _snapshot(self)
        params = snap.params
        if _wants_grad(self, params):
            if u8:
                raise RuntimeError("the training path takes batches in the tower dtype (uint8 pixel batches are an inference input)")
            from .train import VitTrainFn
            return VitTrainFn.apply(self, image, bool(normalize), snap.names, *params)
        lib = L.load()
        with torch.cuda.device(image.device):
            eng = self._build(image.device, params=params)
            nbytes = lib.b200clip_workspace_bytes(C.byref(eng.cfg), B, eng.cfg.seq_len)
            ws = eng.workspace(nbytes, image.device)

            if u8:
                # uint8 pixels: ToTensor + Normalize (open_clip/transform.py:274-392) run inside the im2col kernel
                pp = getattr(self, "preprocess_cfg", None) or {}
                mean = (C.c_float * 3)(*[float(v) for v in pp.get("mean", OPENAI_DATASET_MEAN)])
                std = (C.c_float * 3)(*[float(v) for v in pp.get("std", OPENAI_DATASET_STD)])

            else:
                mean = std = None

            def enqueue(stages: int, dst) -> None:
                rc = lib.b200clip_vit_forward_stages(C.byref(eng.cfg), C.byref(eng.weights), None if u8 else image.data_ptr(),
                                                     image.data_ptr() if u8 else None, mean, std, L.ptr(dst), B, int(normalize),
                                                     ws.data_ptr(), ws.numel(), stages, L.stream_ptr())
                L.check(rc, "b200clip_vit_forward_stages")

            if self.use_cuda_graphs and not torch.cuda.is_current_stream_capturing():
                return eng.run_staged((B, ws.data_ptr()), enqueue, (B, self.output_dim), dt, image.device)
            out = torch.empty((B, self.output_dim), dtype=dt, device=image.device)
            enqueue(L.STAGE_INPUT | L.STAGE_BODY | L.STAGE_OUTPUT, out)
        return out


# ------------------------------------------------------------------------------------------------
# CLIP
# ------------------------------------------------------------------------------------------------
class CLIP(nn.Module):
    """Drop-in for open_clip.model.CLIP (model.py:220-315) on the ViT + text-transformer path."""

    def __init__(self, embed_dim: int, vision_cfg: dict, text_cfg: dict, quick_gelu: bool = False,
                 init_logit_scale: float = np.log(1 / 0.07), init_logit_bias: Optional[float] = None,
                 cast_dtype: Optional[torch.dtype] = None, output_dict: bool = False):
        super().__init__()
        from .model_configs import TEXT_DEFAULTS, VISION_DEFAULTS
        v = {**VISION_DEFAULTS, **(vision_cfg if isinstance(vision_cfg, dict) else vars(vision_cfg))}
        t = {**TEXT_DEFAULTS, **(text_cfg if isinstance(text_cfg, dict) else vars(text_cfg))}
        if v.get("timm_model_name") or t.get("hf_model_name"):
            raise RuntimeError("only the native ViT / ModifiedResNet image towers with the native text transformer are on this hot "
                               "path (timm / HF towers: SURVEY.md §8(f))")
        resnet = isinstance(v["layers"], (tuple, list))
        for k in ("attentional_pool", "no_ln_pre", "final_ln_after_pool", "output_tokens"):
            if v.get(k):
                raise RuntimeError(f"vision_cfg.{k} is not supported on this hot path")
        if not resnet and (v.get("pool_type", "tok") != "tok" or v.get("pos_embed_type", "learnable") != "learnable"):
            raise RuntimeError("only pool_type='tok' with learnable positional embeddings is supported")
        if v.get("ls_init_value") is not None or t.get("ls_init_value") is not None:
            raise RuntimeError("LayerScale (ls_init_value) is not supported on this hot path")
        for k in ("embed_cls", "no_causal_mask", "proj_bias", "output_tokens"):
            if t.get(k):
                raise RuntimeError(f"text_cfg.{k} is not supported on this hot path")
        if t.get("pool_type", "argmax") != "argmax":
            raise RuntimeError("only text pool_type='argmax' is supported")
        self.output_dict = output_dict
        self.quick_gelu = bool(quick_gelu)

        if resnet:
            # ModifiedResNet (model.py:131-139): heads = width * 32 / head_width; quick_gelu does not reach this tower
            from .resnet import ResNetTower
            self.visual = ResNetTower(v["layers"], embed_dim, v["width"] * 32 // v["head_width"], v["image_size"], v["width"])
        else:
            self.visual = VisionTower(v["image_size"], v["patch_size"], v["width"], v["layers"], v["width"] // v["head_width"],
                                      v["mlp_ratio"], embed_dim, quick_gelu)

        # text tower; parameters are re-exported on the CLIP module like the reference does (model.py:239-248)
        Wt = t["width"]
        if t["heads"] * 64 != Wt:
            raise RuntimeError(f"text head width must be 64 (width {Wt}, heads {t['heads']})")
        self.context_length = t["context_length"]
        self.vocab_size = t["vocab_size"]
        self.text_pool_type = "argmax"
        # registration order (= state_dict key order) follows the reference: transformer, token_embedding, ln_final;
        # the RNG draws below keep the reference's construction order: nn.Embedding reset, then the blocks
        self.transformer = _Stack(Wt, t["layers"], t["heads"], t["mlp_ratio"])
        self.token_embedding = _TokenTable(self.vocab_size, Wt)
        with torch.no_grad():
            self.token_embedding.weight.normal_()                       # nn.Embedding default reset (consumes RNG)
        self.positional_embedding = nn.Parameter(torch.empty(self.context_length, Wt))
        _init_stack(self.transformer)
        self.ln_final = _Affine(Wt)
        self.text_projection = nn.Parameter(torch.empty(Wt, embed_dim))
        self.register_buffer("attn_mask", torch.full((self.context_length, self.context_length), float("-inf")).triu_(1),
                             persistent=False)
        self._init_text_parameters()

        self.logit_scale = nn.Parameter(torch.ones([]) * init_logit_scale)
        self.logit_bias = nn.Parameter(torch.ones([]) * init_logit_bias) if init_logit_bias is not None else None

        # exact causal truncation of encode_text at max(EOT)+1; off = run all `context_length`
        # positions like the reference.  Costs one small device->host read per call.
        self.truncate_text_at_eot = False
        #: 16-bit modes: fold ln_1 / ln_2 of the text blocks into the QKV / c_fc GEMM epilogues
        self.fold_layernorm = True
        self._text_engine = _Engine()
        #: replay encode_text's block stack as a CUDA graph per (batch, sequence length) (see _Engine.run_staged)
        self.use_cuda_graphs = True
        #: mixed precision (`amp*`): compute dtype of the text tower on fp32 master parameters (None = parameter dtype)
        self.compute_dtype = None

    def _init_text_parameters(self) -> None:
        """TextTransformer.init_parameters (transformer.py:724-745)."""
        st = self.transformer
        proj_std = (st.width ** -0.5) * ((2 * st.layers) ** -0.5)
        attn_std = st.width ** -0.5
        fc_std = (2 * st.width) ** -0.5
        with torch.no_grad():
            self.token_embedding.weight.normal_(std=0.02)
            self.positional_embedding.normal_(std=0.01)
            for blk in st.resblocks:
                blk.attn.in_proj_weight.normal_(std=attn_std)
                blk.attn.out_proj.weight.normal_(std=proj_std)
                blk.mlp.c_fc.weight.normal_(std=fc_std)
                blk.mlp.c_proj.weight.normal_(std=proj_std)
            self.text_projection.normal_(std=st.width ** -0.5)

    # -- reference API surface ---------------------------------------------------------------
    def lock_image_tower(self, unlocked_groups: int = 0, freeze_bn_stats: bool = False):
        self.visual.lock(unlocked_groups=unlocked_groups, freeze_bn_stats=freeze_bn_stats)

    def set_grad_checkpointing(self, enable: bool = True):
        self.visual.set_grad_checkpointing(enable)
        self.transformer.grad_checkpointing = enable

    def encode_image(self, image: torch.Tensor, normalize: bool = False) -> torch.Tensor:
        return self.visual(image, normalize=normalize)

    def _text_compute_dtype(self) -> torch.dtype:
        return _resolve_compute_dtype(self.compute_dtype, self.transformer.get_cast_dtype())

    def _text_named_parameters(self):
        """(name, parameter) of everything encode_text reads, names as in the CLIP state_dict."""
        snap = _snapshot(self.transformer)
        if snap.__dict__.get("prefixed_for") is not snap.named:      # the prefixed names are rebuilt only with the snapshot
            snap.prefixed = [("transformer." + n, p) for n, p in snap.named]
            snap.prefixed_for = snap.named
        return [("token_embedding.weight", self.token_embedding.weight), ("positional_embedding", self.positional_embedding),
                ("ln_final.weight", self.ln_final.weight), ("ln_final.bias", self.ln_final.bias), ("text_projection", self.text_projection),
                *snap.prefixed]

    def _build_text(self, device, for_training: bool = False) -> _Engine:
        params = [p for _, p in self._text_named_parameters()]
        dt = self._text_compute_dtype()
        fold = bool(self.fold_layernorm) and dt != torch.float32 and not for_training
        sig, static_sig = _Engine.signatures(params, fold, bool(self.quick_gelu), dt)
        eng = self._text_engine
        if eng.sig == sig:
            return eng
        if eng.try_refresh(params, static_sig):
            return eng
        keep = _Keep()
        proj_t = torch.empty((self.text_projection.shape[1], self.transformer.width), dtype=dt, device=device)

        def refresh_proj():
            proj_t.copy_(self.text_projection.detach().t())

        keep.refresh.append(refresh_proj)
        keep.append(proj_t)
        blocks = _pack_blocks(self.transformer, dt, keep, fold)
        w = L.TextWeights()
        w.tok_emb = _f32(self.token_embedding.weight, keep)
        w.pos_emb = _f32(self.positional_embedding, keep)
        w.ln_final_g, w.ln_final_b = _f32(self.ln_final.weight, keep), _f32(self.ln_final.bias, keep)
        w.proj_t = proj_t.data_ptr()
        w.blocks_host = C.cast(blocks, C.c_void_p)
        with torch.no_grad():
            refresh_proj()
        st = self.transformer
        cfg = L.TowerCfg(dtype=L.dtype_code(dt), width=st.width, layers=st.layers, heads=st.heads, mlp_width=st.mlp_width,
                         embed_dim=self.text_projection.shape[1], seq_len=self.context_length, quick_gelu=int(self.quick_gelu),
                         image_size=0, patch_size=0, patch_kpad=0, vocab_size=self.vocab_size, fold_ln=int(fold))
        eng.sig, eng.static_sig, eng.keep, eng.blocks, eng.weights, eng.cfg = sig, static_sig, keep, blocks, w, cfg
        eng.graphs.clear()      # captured graphs hold the old weight pointers / switches
        return eng

    def encode_text(self, text: torch.Tensor, normalize: bool = False) -> torch.Tensor:
        """[T, context_length] int64 -> [T, D] (model.py:269-284)."""
        _check_device(text, "encode_text")
        _check_device(self.text_projection, "encode_text (model weights)")
        if text.ndim != 2 or text.shape[1] != self.context_length:
            raise RuntimeError(f"expected token ids of shape [T, {self.context_length}], got {tuple(text.shape)}")
        if text.dtype != torch.int64:
            text = text.long()
        text = text.contiguous()
        T = text.shape[0]
        dt = self._text_compute_dtype()
        D = self.text_projection.shape[1]
        if T == 0:
            return torch.empty((0, D), dtype=dt, device=text.device)
        lib = L.load()
        with torch.cuda.device(text.device):
            seq_len = self.context_length
            if self.truncate_text_at_eot:
                eot = torch.empty((T,), dtype=torch.int32, device=text.device)
                L.check(lib.b200clip_eot_argmax(text.data_ptr(), self.context_length, eot.data_ptr(), T, L.stream_ptr()),
                        "b200clip_eot_argmax")
                seq_len = int(np.max(eot.cpu().numpy())) + 1
            named = self._text_named_parameters()
            if _wants_grad(self, [p for _, p in named]):
                from .train import TextTrainFn
                return TextTrainFn.apply(self, text, seq_len, bool(normalize), [n for n, _ in named], *[p for _, p in named])
            eng = self._build_text(text.device)
            nbytes = lib.b200clip_workspace_bytes(C.byref(eng.cfg), T, seq_len)
            ws = eng.workspace(nbytes, text.device)

            def enqueue(stages: int, dst) -> None:
                rc = lib.b200clip_text_forward_stages(C.byref(eng.cfg), C.byref(eng.weights), text.data_ptr(), L.ptr(dst), T, seq_len,
                                                      int(normalize), ws.data_ptr(), ws.numel(), stages, L.stream_ptr())
                L.check(rc, "b200clip_text_forward_stages")

            if self.use_cuda_graphs and not torch.cuda.is_current_stream_capturing():
                return eng.run_staged((T, seq_len, ws.data_ptr()), enqueue, (T, D), dt, text.device)
            out = torch.empty((T, D), dtype=dt, device=text.device)
            enqueue(L.STAGE_INPUT | L.STAGE_BODY | L.STAGE_OUTPUT, out)
        return out

    def get_logits(self, image, text):
        from .. import ops
        image_features = self.encode_image(image, normalize=True)
        text_features = self.encode_text(text, normalize=True)
        scale = float(self.logit_scale.detach().exp())
        logits, _, _ = ops.zeroshot(image_features, text_features, 0, normalize_img=False, want_logits=True, logit_scale=scale)
        image_logits = logits.to(image_features.dtype)
        if self.logit_bias is not None:
            image_logits = image_logits + self.logit_bias
        return image_logits, image_logits.T

    def forward(self, image: Optional[torch.Tensor] = None, text: Optional[torch.Tensor] = None):
        image_features = self.encode_image(image, normalize=True) if image is not None else None
        text_features = self.encode_text(text, normalize=True) if text is not None else None
        if self.output_dict:
            out = {"image_features": image_features, "text_features": text_features, "logit_scale": self.logit_scale.exp()}
            if self.logit_bias is not None:
                out["logit_bias"] = self.logit_bias
            return out
        if self.logit_bias is not None:
            return image_features, text_features, self.logit_scale.exp(), self.logit_bias
        return image_features, text_features, self.logit_scale.exp()


# ------------------------------------------------------------------------------------------------
# precision helpers (model.py:86-101, 396-423)
# ------------------------------------------------------------------------------------------------
def get_cast_dtype(precision: str):
    return {"bf16": torch.bfloat16, "fp16": torch.float16}.get(precision)


def get_input_dtype(precision: str):
    if precision in ("bf16", "pure_bf16"):
        return torch.bfloat16
    if precision in ("fp16", "pure_fp16"):
        return torch.float16
    return None


def convert_weights_to_lp(model: nn.Module, dtype=torch.float16):
    """Cast the GEMM operands (conv / linear / MHA weights+biases and the two projections) to `dtype`; LayerNorm
    parameters, embeddings, positional tables and logit_scale stay fp32 — same split as model.py:396-423."""

    def _convert(m):
        if isinstance(m, (_Dense, _PatchConv, nn.Conv2d, nn.Linear)):     # nn.Conv2d / nn.Linear: the ModifiedResNet tower
            m.weight.data = m.weight.data.to(dtype)
            if getattr(m, "bias", None) is not None:
                m.bias.data = m.bias.data.to(dtype)
        if isinstance(m, _PackedMHA):
            m.in_proj_weight.data = m.in_proj_weight.data.to(dtype)
            m.in_proj_bias.data = m.in_proj_bias.data.to(dtype)
        if isinstance(m, CLIP):
            m.text_projection.data = m.text_projection.data.to(dtype)
        if isinstance(m, VisionTower):
            m.proj.data = m.proj.data.to(dtype)

    model.apply(_convert)


convert_weights_to_fp16 = convert_weights_to_lp  # backwards-compatible alias, as in the reference
