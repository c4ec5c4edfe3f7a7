"""Autograd nodes of the two towers for the training path: what `loss.backward()` reaches when the reference's
train loop (deps/open_clip/src/training/train.py:115-183) runs on this package.

Forward = `b200clip_vit_forward_train` / `b200clip_text_forward_train`: the inference kernels plus a copy of the residual stream
at every block boundary (activation policy of --grad-checkpointing, transformer.py:353-355).  Backward =
`b200clip_vit_backward` / `b200clip_text_backward`: one C-ABI call that recomputes each block and fills the gradient of every
tower parameter.  PyTorch only owns the buffers and hands gradients to autograd; there is no PyTorch arithmetic on this path
apart from the dtype casts of the mixed-precision (`amp*`) modes, where the fp32 master parameters receive the 16-bit
gradients up-cast — exactly what autocast's cast-backward does in the reference.
"""
from __future__ import annotations

import ctypes as C

import torch

from .. import _lib as L

_BLOCK_FIELDS = (("ln_1.weight", "ln1_g"), ("ln_1.bias", "ln1_b"), ("ln_2.weight", "ln2_g"), ("ln_2.bias", "ln2_b"),
                 ("attn.in_proj_weight", "in_proj_w"), ("attn.in_proj_bias", "in_proj_b"),
                 ("attn.out_proj.weight", "out_proj_w"), ("attn.out_proj.bias", "out_proj_b"),
                 ("mlp.c_fc.weight", "fc_w"), ("mlp.c_fc.bias", "fc_b"), ("mlp.c_proj.weight", "proj_w"), ("mlp.c_proj.bias", "proj_b"))
_F32_FIELDS = {"ln1_g", "ln1_b", "ln2_g", "ln2_b"}


class _Arena:
    """Gradient buffers of one backward as TWO allocations (one per dtype) carved into views: ~150 tensors per tower would
    otherwise cost one allocator call each, and — in the mixed-precision modes — one up-cast launch each."""

    def __init__(self, device, dt: torch.dtype):
        self.device, self.dt = device, dt
        self.plan: list = []          # (name, shape, is_f32, offset)
        self.n16 = self.n32 = 0

    def add(self, name: str, shape, f32: bool) -> None:
        n = 1
        for d in shape:
            n *= d
        n_al = (n + 63) // 64 * 64    # 256-byte aligned in either dtype
        if f32 or self.dt == torch.float32:
            self.plan.append((name, tuple(shape), True, self.n32))
            self.n32 += n_al
        else:
            self.plan.append((name, tuple(shape), False, self.n16))
            self.n16 += n_al

    def allocate(self) -> dict:
        self.flat32 = torch.empty(max(self.n32, 1), dtype=torch.float32, device=self.device)
        self.flat16 = torch.empty(max(self.n16, 1), dtype=self.dt, device=self.device) if self.dt != torch.float32 else None
        out = {}
        for name, shape, f32, off in self.plan:
            n = 1
            for d in shape:
                n *= d
            out[name] = (self.flat32 if f32 else self.flat16)[off:off + n].view(shape)
        return out

    def upcast(self, grads: dict) -> dict:
        """All 16-bit gradients as fp32 views of ONE converted buffer (fp32 master parameters of the `amp*` modes)."""
        if self.flat16 is None:
            return grads
        wide = self.flat16.float()
        out = dict(grads)
        for name, shape, f32, off in self.plan:
            if not f32:
                n = 1
                for d in shape:
                    n *= d
                out[name] = wide[off:off + n].view(shape)
        return out


def _plan_blocks(arena: _Arena, stack, prefix: str) -> None:
    named = dict(stack.named_parameters())
    for i in range(stack.layers):
        for pname, field in _BLOCK_FIELDS:
            arena.add(f"{prefix}resblocks.{i}.{pname}", named[f"resblocks.{i}.{pname}"].shape, field in _F32_FIELDS)


def _block_grads(stack, prefix: str, grads: dict):
    """-> ctypes array of BlockGrads pointing at the arena's views."""
    arr = (L.BlockGrads * stack.layers)()
    for i in range(stack.layers):
        for pname, field in _BLOCK_FIELDS:
            setattr(arr[i], field, grads[f"{prefix}resblocks.{i}.{pname}"].data_ptr())
    return arr


def _bwd_workspace(eng, lib, B: int, L_: int, device) -> torch.Tensor:
    n = lib.b200clip_backward_workspace_bytes(C.byref(eng.cfg), B, L_)
    ws = getattr(eng, "bws", None)
    if ws is None or ws.numel() < n or ws.device != device:
        ws = torch.empty(n, dtype=torch.uint8, device=device)
        eng.bws = ws
    return ws


def _capture_engine(ctx, eng, params) -> None:
    """The backward must see exactly the weight structs of the forward (and the tensors they point into), whatever the tower's
    engine cache does in between (an evaluation call rebuilds it with folded LayerNorms)."""
    ctx.eng, ctx.cfg, ctx.weights, ctx.blocks, ctx.keep = eng, eng.cfg, eng.weights, eng.blocks, eng.keep
    ctx.versions = tuple(p._version for p in params)


def _captured_engine(ctx, what: str):
    if tuple(p._version for p in ctx.params) != ctx.versions:
        raise RuntimeError(f"b200clip: parameters of the {what} were modified in place between forward and backward")
    return ctx.eng, ctx.cfg, ctx.weights


def _hand_over(grads: dict, names, params):
    """Gradients in the order / dtype / shape autograd expects for `params`."""
    out = []
    for n, p in zip(names, params):
        g = grads.get(n)
        if g is None or not p.requires_grad:
            out.append(None)
            continue
        if g.shape != p.shape:
            g = g.reshape(p.shape)
        out.append(g if g.dtype == p.dtype else g.to(p.dtype))
    return out


class VitTrainFn(torch.autograd.Function):
    """features = VisionTower(image) with gradients to every parameter of the tower (not to the image)."""

    @staticmethod
    def forward(ctx, tower, image, normalize, names, *params):
        lib = L.load()
        dev = image.device
        with torch.cuda.device(dev):
            eng = tower._build(dev, for_training=True)
            cfg = eng.cfg
            B = image.shape[0]
            ws = eng.workspace(lib.b200clip_workspace_bytes(C.byref(cfg), B, cfg.seq_len), dev)
            saved = torch.empty(lib.b200clip_train_saved_bytes(C.byref(cfg), B, cfg.seq_len), dtype=torch.uint8, device=dev)
            ctx.dt = tower._compute_dtype()
            out = torch.empty((B, tower.output_dim), dtype=ctx.dt, device=dev)
            L.check(lib.b200clip_vit_forward_train(C.byref(cfg), C.byref(eng.weights), image.data_ptr(), out.data_ptr(), B, int(normalize),
                                                   saved.data_ptr(), saved.numel(), ws.data_ptr(), ws.numel(), L.stream_ptr()),
                    "b200clip_vit_forward_train")
        ctx.tower, ctx.image, ctx.normalize, ctx.names, ctx.saved = tower, image, bool(normalize), names, saved
        ctx.params = params
        _capture_engine(ctx, eng, params)
        return out

    @staticmethod
    def backward(ctx, d_out):
        tower, image = ctx.tower, ctx.image
        lib = L.load()
        dev = image.device
        with torch.cuda.device(dev):
            eng, cfg, weights = _captured_engine(ctx, "vision tower")
            dt = ctx.dt
            B, Lq, W = image.shape[0], cfg.seq_len, cfg.width
            arena = _Arena(dev, dt)
            _plan_blocks(arena, tower.transformer, "transformer.")
            arena.add("conv1.padded", (W, cfg.patch_kpad), False)
            arena.add("class_embedding", (W,), True)
            arena.add("positional_embedding", (Lq, W), True)
            for n in ("ln_pre", "ln_post"):
                arena.add(f"{n}.weight", (W,), True)
                arena.add(f"{n}.bias", (W,), True)
            arena.add("proj", (W, tower.output_dim), False)
            grads = arena.allocate()
            blocks = _block_grads(tower.transformer, "transformer.", grads)
            g = L.VitGrads()
            conv = grads["conv1.padded"]
            g.conv1_w, g.class_emb, g.pos_emb = conv.data_ptr(), grads["class_embedding"].data_ptr(), grads["positional_embedding"].data_ptr()
            g.ln_pre_g, g.ln_pre_b = grads["ln_pre.weight"].data_ptr(), grads["ln_pre.bias"].data_ptr()
            g.ln_post_g, g.ln_post_b = grads["ln_post.weight"].data_ptr(), grads["ln_post.bias"].data_ptr()
            g.proj = grads["proj"].data_ptr()
            g.blocks_host = C.cast(blocks, C.c_void_p)
            d_out = d_out.to(dt).contiguous()
            bws = _bwd_workspace(eng, lib, B, Lq, dev)
            L.check(lib.b200clip_vit_backward(C.byref(cfg), C.byref(weights), image.data_ptr(), d_out.data_ptr(), B, int(ctx.normalize),
                                              ctx.saved.data_ptr(), C.byref(g), bws.data_ptr(), bws.numel(), L.stream_ptr()),
                    "b200clip_vit_backward")
            P = tower.patch_size[0]
            if any(p.dtype == torch.float32 for p in ctx.params) and dt != torch.float32:
                grads = arena.upcast(grads)
            grads["conv1.weight"] = grads["conv1.padded"][:, :3 * P * P].reshape(W, 3, P, P)
        ctx.saved = None
        return (None, None, None, None, *_hand_over(grads, ctx.names, ctx.params))


class TextTrainFn(torch.autograd.Function):
    """features = CLIP.encode_text(text) with gradients to every parameter of the text tower."""

    @staticmethod
    def forward(ctx, model, text, seq_len, normalize, names, *params):
        lib = L.load()
        dev = text.device
        with torch.cuda.device(dev):
            eng = model._build_text(dev, for_training=True)
            cfg = eng.cfg
            T = text.shape[0]
            dt = model._text_compute_dtype()
            ws = eng.workspace(lib.b200clip_workspace_bytes(C.byref(cfg), T, seq_len), dev)
            saved = torch.empty(lib.b200clip_train_saved_bytes(C.byref(cfg), T, seq_len), dtype=torch.uint8, device=dev)
            out = torch.empty((T, model.text_projection.shape[1]), dtype=dt, device=dev)
            L.check(lib.b200clip_text_forward_train(C.byref(cfg), C.byref(eng.weights), text.data_ptr(), out.data_ptr(), T, seq_len, int(normalize),
                                                    saved.data_ptr(), saved.numel(), ws.data_ptr(), ws.numel(), L.stream_ptr()),
                    "b200clip_text_forward_train")
        ctx.model, ctx.text, ctx.seq_len, ctx.normalize, ctx.names, ctx.saved = model, text, seq_len, bool(normalize), names, saved
        ctx.params = params
        _capture_engine(ctx, eng, params)
        ctx.dt = dt
        return out

    @staticmethod
    def backward(ctx, d_out):
        model, text = ctx.model, ctx.text
        lib = L.load()
        dev = text.device
        with torch.cuda.device(dev):
            eng, cfg, weights = _captured_engine(ctx, "text tower")
            dt = ctx.dt
            T, W, D = text.shape[0], cfg.width, model.text_projection.shape[1]
            arena = _Arena(dev, dt)
            _plan_blocks(arena, model.transformer, "transformer.")
            arena.add("token_embedding.weight", (model.vocab_size, W), True)
            arena.add("positional_embedding", (model.context_length, W), True)
            arena.add("ln_final.weight", (W,), True)
            arena.add("ln_final.bias", (W,), True)
            arena.add("text_projection", (W, D), False)
            grads = arena.allocate()
            blocks = _block_grads(model.transformer, "transformer.", grads)
            g = L.TextGrads()
            g.tok_emb, g.pos_emb = grads["token_embedding.weight"].data_ptr(), grads["positional_embedding"].data_ptr()
            g.ln_final_g, g.ln_final_b = grads["ln_final.weight"].data_ptr(), grads["ln_final.bias"].data_ptr()
            g.proj = grads["text_projection"].data_ptr()
            g.blocks_host = C.cast(blocks, C.c_void_p)
            d_out = d_out.to(dt).contiguous()
            bws = _bwd_workspace(eng, lib, T, ctx.seq_len, dev)
            L.check(lib.b200clip_text_backward(C.byref(cfg), C.byref(weights), text.data_ptr(), d_out.data_ptr(), T, ctx.seq_len,
                                               int(ctx.normalize), ctx.saved.data_ptr(), C.byref(g), bws.data_ptr(), bws.numel(), L.stream_ptr()),
                    "b200clip_text_backward")
            if any(p.dtype == torch.float32 for p in ctx.params) and dt != torch.float32:
                grads = arena.upcast(grads)
        ctx.saved = None
        return (None, None, None, None, None, *_hand_over(grads, ctx.names, ctx.params))
