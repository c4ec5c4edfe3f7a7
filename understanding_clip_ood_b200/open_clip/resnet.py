"""ModifiedResNet image tower (RN50 / RN101 / RN50x4 / x16 / x64) behind `model.visual` — SURVEY.md §8(f)4.

Parameter containers named so that `state_dict()` has the reference's keys, order and seed-0 values
(deps/open_clip/src/open_clip/modified_resnet.py:95-160: stem conv1-3 / bn1-3, layer1..4.<i>.conv1-3 / bn1-3 /
downsample.0-1, attnpool.positional_embedding / k_proj / q_proj / v_proj / c_proj), and a `forward` that is ONE
C-ABI call (`b200clip_resnet_forward_stages`, csrc/resnet.cu): NHWC activations, every convolution a tcgen05 GEMM with
the folded BatchNorm shift, ReLU and the residual add in its epilogue, the attention pool on the ViT tower's
attention kernel.

Eval mode only: BatchNorm runs on its running statistics (folded into the convolution weights when the engine is
built; re-folded whenever a parameter or a running statistic changes).  Batch statistics (training mode) and the tower's
backward are not on this path and raise.
"""
from __future__ import annotations

import ctypes as C
from collections import OrderedDict
from typing import List

import torch
from torch import nn

from .. import _lib as L


class _Bottleneck(nn.Module):
    """Parameters of one Bottleneck (modified_resnet.py:10-40); the arithmetic lives in csrc/resnet.cu."""
    expansion = 4

    def __init__(self, inplanes: int, planes: int, stride: int = 1):
        super().__init__()
        self.conv1 = nn.Conv2d(inplanes, planes, 1, bias=False)
        self.bn1 = nn.BatchNorm2d(planes)
        self.conv2 = nn.Conv2d(planes, planes, 3, padding=1, bias=False)
        self.bn2 = nn.BatchNorm2d(planes)
        self.conv3 = nn.Conv2d(planes, planes * 4, 1, bias=False)
        self.bn3 = nn.BatchNorm2d(planes * 4)
        self.stride = stride
        self.downsample = None
        if stride > 1 or inplanes != planes * 4:
            self.downsample = nn.Sequential(OrderedDict([("0", nn.Conv2d(inplanes, planes * 4, 1, stride=1, bias=False)),
                                                         ("1", nn.BatchNorm2d(planes * 4))]))


class _AttentionPool(nn.Module):
    """Parameters of AttentionPool2d (modified_resnet.py:59-67)."""

    def __init__(self, spacial_dim: int, embed_dim: int, num_heads: int, output_dim: int):
        super().__init__()
        self.positional_embedding = nn.Parameter(torch.randn(spacial_dim ** 2 + 1, embed_dim) / embed_dim ** 0.5)
        self.k_proj = nn.Linear(embed_dim, embed_dim)
        self.q_proj = nn.Linear(embed_dim, embed_dim)
        self.v_proj = nn.Linear(embed_dim, embed_dim)
        self.c_proj = nn.Linear(embed_dim, output_dim or embed_dim)
        self.num_heads = num_heads


def _fold(conv: nn.Conv2d, bn: nn.BatchNorm2d, kpad: int = 0):
    """conv + eval-mode BatchNorm as one GEMM operand: weight rows scaled by gamma / sqrt(var + eps) in K order (ky, kx, cin),
    and the per-channel shift beta - mean * gamma / sqrt(var + eps); all in fp32 (the caller rounds once to the tower dtype)."""
    w = conv.weight.detach().float()
    scale = bn.weight.detach().float() / torch.sqrt(bn.running_var.detach().float() + bn.eps)
    shift = bn.bias.detach().float() - bn.running_mean.detach().float() * scale
    w = (w * scale[:, None, None, None]).permute(0, 2, 3, 1).reshape(w.shape[0], -1)
    if kpad > w.shape[1]:
        w = torch.nn.functional.pad(w, (0, kpad - w.shape[1]))
    return w, shift


class ResNetTower(nn.Module):
    """Parameters + driver of the ModifiedResNet image tower; the `model.visual` object for the RN* configurations."""

    def __init__(self, layers: List[int], output_dim: int, heads: int, image_size: int = 224, width: int = 64):
        super().__init__()
        from .model import _Engine
        if isinstance(image_size, (tuple, list)):
            if image_size[0] != image_size[1]:
                raise ValueError("only square images are supported")
            image_size = image_size[0]
        if image_size % 32 != 0:
            raise ValueError(f"image_size {image_size} must be a multiple of 32")
        if heads * 64 != width * 32:
            raise ValueError(f"attention-pool head width must be 64 (embed {width * 32}, heads {heads})")
        if (width // 2) % 8 != 0:
            raise ValueError(f"width {width}: width / 2 must be a multiple of 8 (16-byte channel vectors)")
        self.output_dim = output_dim
        self.image_size = (image_size, image_size)
        self.width = width
        self.layers = list(layers)
        # registration order = the reference's construction order (same RNG draws, same state_dict order)
        self.conv1 = nn.Conv2d(3, width // 2, kernel_size=3, stride=2, padding=1, bias=False)
        self.bn1 = nn.BatchNorm2d(width // 2)
        self.conv2 = nn.Conv2d(width // 2, width // 2, kernel_size=3, padding=1, bias=False)
        self.bn2 = nn.BatchNorm2d(width // 2)
        self.conv3 = nn.Conv2d(width // 2, width, kernel_size=3, padding=1, bias=False)
        self.bn3 = nn.BatchNorm2d(width)
        inplanes = width
        for i, (mult, n) in enumerate(zip((1, 2, 4, 8), self.layers)):
            planes = width * mult
            blocks = [_Bottleneck(inplanes, planes, 1 if i == 0 else 2)]
            inplanes = planes * 4
            blocks += [_Bottleneck(inplanes, planes) for _ in range(1, n)]
            setattr(self, f"layer{i + 1}", nn.Sequential(*blocks))
        self.attnpool = _AttentionPool(image_size // 32, width * 32, heads, output_dim)
        self.init_parameters()
        self._engine = _Engine()
        self.use_cuda_graphs = True
        #: mixed precision (`amp*`): kernels run in this dtype on 16-bit copies of the fp32 parameters (None = parameter dtype)
        self.compute_dtype = None

    def init_parameters(self):
        """modified_resnet.py:135-146."""
        std = self.attnpool.c_proj.in_features ** -0.5
        for lin in (self.attnpool.q_proj, self.attnpool.k_proj, self.attnpool.v_proj, self.attnpool.c_proj):
            nn.init.normal_(lin.weight, std=std)
        for blk in self.bottlenecks():
            nn.init.zeros_(blk.bn3.weight)

    def bottlenecks(self):
        for i in range(4):
            yield from getattr(self, f"layer{i + 1}")

    # -- reference API surface ---------------------------------------------------------------
    def lock(self, unlocked_groups: int = 0, freeze_bn_stats: bool = False):
        assert unlocked_groups == 0, "partial locking not currently supported for this model"
        for p in self.parameters():
            p.requires_grad = False
        if freeze_bn_stats:
            for m in self.modules():
                if isinstance(m, nn.BatchNorm2d):
                    m.eval()

    def set_grad_checkpointing(self, enable: bool = True):
        pass

    def _compute_dtype(self) -> torch.dtype:
        from .model import _resolve_compute_dtype
        return _resolve_compute_dtype(self.compute_dtype, self.conv1.weight.dtype)

    def _build(self, device):
        from .model import _Engine, _Keep, _snapshot
        tensors = _snapshot(self, buffers=True).tensors
        dt = self._compute_dtype()
        sig, static_sig = _Engine.signatures(tensors, dt)
        eng = self._engine
        if eng.sig == sig:
            return eng
        if eng.try_refresh(tensors, static_sig):
            return eng
        keep = _Keep()
        es = 4 if dt == torch.float32 else 2
        stem_kpad = 32            # 27 taps -> whole 16-byte vectors in every dtype
        ap = self.attnpool
        E = self.width * 32

        # (destination weight, destination shift, conv, bn, kpad): filled by refresh() below, also after an in-place checkpoint load
        jobs = []

        def operand(conv, bn, kpad=0):
            cout = conv.weight.shape[0]
            k = max(kpad, conv.weight[0].numel())
            wt = torch.empty((cout, k), dtype=dt, device=device)
            sh = torch.empty((cout,), dtype=dt, device=device)
            jobs.append((wt, sh, conv, bn, kpad))
            keep.extend((wt, sh))
            return wt.data_ptr(), sh.data_ptr()

        w = L.ResnetWeights()
        for i, (conv, bn) in enumerate(((self.conv1, self.bn1), (self.conv2, self.bn2), (self.conv3, self.bn3))):
            w.stem_w[i], w.stem_b[i] = operand(conv, bn, stem_kpad if i == 0 else 0)
        blocks = list(self.bottlenecks())
        arr = (L.ResnetBlock * len(blocks))()
        for b, blk in zip(arr, blocks):
            b.conv1_w, b.conv1_b = operand(blk.conv1, blk.bn1)
            b.conv2_w, b.conv2_b = operand(blk.conv2, blk.bn2)
            b.conv3_w, b.conv3_b = operand(blk.conv3, blk.bn3)
            if blk.downsample is not None:
                b.down_w, b.down_b = operand(blk.downsample[0], blk.downsample[1])
            b.cin, b.planes, b.stride = blk.conv1.weight.shape[1], blk.conv1.weight.shape[0], blk.stride
        qkv_w = torch.empty((3 * E, E), dtype=dt, device=device)
        qkv_b = torch.empty((3 * E,), dtype=dt, device=device)
        cw = torch.empty((self.output_dim, E), dtype=dt, device=device)
        cb = torch.empty((self.output_dim,), dtype=dt, device=device)
        pos = torch.empty_like(ap.positional_embedding, dtype=torch.float32, device=device)

        def refresh():
            for wt, sh, conv, bn, kpad in jobs:
                fw, fs = _fold(conv, bn, kpad)
                wt.copy_(fw)
                sh.copy_(fs)
            for i, lin in enumerate((ap.q_proj, ap.k_proj, ap.v_proj)):
                qkv_w[i * E:(i + 1) * E].copy_(lin.weight.detach())
                qkv_b[i * E:(i + 1) * E].copy_(lin.bias.detach())
            cw.copy_(ap.c_proj.weight.detach())
            cb.copy_(ap.c_proj.bias.detach())
            pos.copy_(ap.positional_embedding.detach())

        keep.refresh.append(refresh)
        keep.extend((qkv_w, qkv_b, cw, cb, pos))
        with torch.no_grad():
            refresh()
        w.blocks_host = C.cast(arr, C.c_void_p)
        w.pos, w.qkv_w, w.qkv_b, w.c_proj_w, w.c_proj_b = pos.data_ptr(), qkv_w.data_ptr(), qkv_b.data_ptr(), cw.data_ptr(), cb.data_ptr()
        cfg = L.ResnetCfg(dtype=L.dtype_code(dt), image_size=self.image_size[0], width=self.width, embed_dim=self.output_dim,
                          heads=ap.num_heads, n_blocks=len(blocks), stem_kpad=stem_kpad)
        eng.sig, eng.static_sig, eng.keep, eng.blocks, eng.weights, eng.cfg = sig, static_sig, keep, arr, w, cfg
        eng.graphs.clear()
        return eng

    def forward(self, image: torch.Tensor, normalize: bool = False) -> torch.Tensor:
        """[B,3,S,S] -> [B,D] (modified_resnet.py:171-181); `normalize` fuses CLIP.encode_image's F.normalize."""
        from .model import _check_device, _snapshot, _wants_grad
        _check_device(image, "encode_image")
        _check_device(self.conv1.weight, "encode_image (model weights)")
        if any(m.training for m in self.modules() if isinstance(m, nn.BatchNorm2d)):
            raise RuntimeError("ModifiedResNet: BatchNorm batch statistics (training mode) are not on this path — call model.eval() "
                               "or visual.lock(freeze_bn_stats=True); the tower runs on BatchNorm's running statistics")
        if _wants_grad(self, _snapshot(self, buffers=True).params):
            raise RuntimeError("ModifiedResNet: the tower backward is not on this path (inference / zero-shot evaluation only)")
        dt = self._compute_dtype()
        if image.dtype == torch.float32 and dt != torch.float32 and self.conv1.weight.dtype == torch.float32:
            image = image.to(dt)          # autocast's input cast
        if image.dtype != dt:
            raise RuntimeError(f"Input type ({image.dtype}) and weight type ({dt}) should be the same "
                               f"(cast the batch with get_input_dtype(precision), as the reference requires)")
        if image.ndim != 4 or image.shape[1] != 3 or tuple(image.shape[2:]) != self.image_size:
            raise RuntimeError(f"expected images of shape [B, 3, {self.image_size[0]}, {self.image_size[1]}], got {tuple(image.shape)}")
        image = image.contiguous()
        B = image.shape[0]
        if B == 0:
            return torch.empty((0, self.output_dim), dtype=dt, device=image.device)
        lib = L.load()
        with torch.cuda.device(image.device):
            eng = self._build(image.device)
            nbytes = lib.b200clip_resnet_workspace_bytes(C.byref(eng.cfg), C.byref(eng.weights), B)
            if nbytes < 0:
                L.check(-1, "b200clip_resnet_workspace_bytes")
            ws = eng.workspace(nbytes, image.device)

            def enqueue(stages: int, dst) -> None:
                rc = lib.b200clip_resnet_forward_stages(C.byref(eng.cfg), C.byref(eng.weights), image.data_ptr(), L.ptr(dst), B,
                                                        int(normalize), ws.data_ptr(), ws.numel(), stages, L.stream_ptr())
                L.check(rc, "b200clip_resnet_forward_stages")

            if self.use_cuda_graphs and not torch.cuda.is_current_stream_capturing():
                return eng.run_staged((B, ws.data_ptr()), enqueue, (B, self.output_dim), dt, image.device)
            out = torch.empty((B, self.output_dim), dtype=dt, device=image.device)
            enqueue(L.STAGE_INPUT | L.STAGE_BODY | L.STAGE_OUTPUT, out)
        return out
