"""ClipLoss with the reference's constructor / forward signature (deps/open_clip/src/open_clip/loss.py:66-131),
computed by the fused forward+backward CUDA path (`b200clip_cliploss`).  For world_size > 1 the feature exchange of
--local-loss --gather-with-grad (the reference issues two all-gathers, loss.py:49-50, whose autograd backward is two
reduce-scatters) runs over peer memory: NVLink stores + flags inside our own kernels (`_PeerLocalClipLoss`, csrc/p2p.cu).
Where symmetric memory is unavailable (or B200CLIP_P2P=0) it is ONE NCCL all-gather of the concatenated img‖txt features and
one reduce-scatter in the backward (`_DistLocalClipLoss`) — the same semantics as torch.distributed.nn.all_gather.

There is no PyTorch-op fallback for the loss arithmetic: CPU feature tensors raise.
"""
from __future__ import annotations

import torch
from torch import nn

try:
    import torch.distributed as dist
    has_distributed = dist.is_available()
except ImportError:  # pragma: no cover
    dist = None
    has_distributed = False

from .. import _lib as L
from .. import ops
from . import peer

__all__ = ["ClipLoss", "gather_features", "local_labels"]


def local_labels(num_logits: int, rank: int, world_size: int, local_loss: bool, device=None) -> torch.Tensor:
    """Ground-truth column of each local row (loss.py:89-100): i, offset by n*rank for the local-loss layout."""
    labels = torch.arange(num_logits, device=device, dtype=torch.long)
    if world_size > 1 and local_loss:
        labels = labels + num_logits * rank
    return labels


class _GatherCat(torch.autograd.Function):
    """all_gather along dim 0 with gradient: forward = all_gather_into_tensor, backward = reduce_scatter(SUM)
    (torch/distributed/nn/functional.py:_AllGather semantics).  Backends without reduce_scatter (gloo, used by the
    CPU host-logic tests) fall back to all_reduce + slice, which is the same sum."""

    @staticmethod
    def forward(ctx, x: torch.Tensor, group):
        world = dist.get_world_size(group)
        ctx.group, ctx.rank, ctx.n = group, dist.get_rank(group), x.shape[0]
        x = x.contiguous()
        out = torch.empty((world * x.shape[0], *x.shape[1:]), dtype=x.dtype, device=x.device)
        if x.is_cuda:
            dist.all_gather_into_tensor(out, x, group=group)
        else:
            dist.all_gather(list(out.chunk(world, dim=0)), x, group=group)
        return out

    @staticmethod
    def backward(ctx, grad_out: torch.Tensor):
        grad_out = grad_out.contiguous()
        if grad_out.is_cuda:
            gx = torch.empty((ctx.n, *grad_out.shape[1:]), dtype=grad_out.dtype, device=grad_out.device)
            dist.reduce_scatter_tensor(gx, grad_out, op=dist.ReduceOp.SUM, group=ctx.group)
        else:
            total = grad_out.clone()
            dist.all_reduce(total, op=dist.ReduceOp.SUM, group=ctx.group)
            gx = total[ctx.rank * ctx.n:(ctx.rank + 1) * ctx.n].clone()
        return gx, None


def gather_features(image_features, text_features, local_loss=False, gather_with_grad=False, rank=0, world_size=1,
                    use_horovod=False, group=None):
    """Same contract as loss.py:19-63 (rank-major concatenation along dim 0), one fused collective for both
    modalities.  Horovod is not supported."""
    assert has_distributed, "torch.distributed did not import correctly, please use a PyTorch version with support."
    if use_horovod:
        raise NotImplementedError("horovod is not supported; use torch.distributed (NCCL)")
    D = image_features.shape[1]
    both = torch.cat([image_features, text_features], dim=1)          # [n, 2D]: one message per rank
    if gather_with_grad:
        gathered = _GatherCat.apply(both, group)
    else:
        with torch.no_grad():
            gathered = _GatherCat.apply(both.detach(), group)
        if not local_loss:
            # ensure grads for the local rank when the gathered features do not carry a gradient (loss.py:56-59)
            n = both.shape[0]
            gathered = torch.cat([gathered[:rank * n], both, gathered[(rank + 1) * n:]], dim=0)
    return gathered[:, :D], gathered[:, D:]


def _consume(ctx) -> None:
    """The backward kernels turn the saved raw logits into s * dL/dlogits IN PLACE (csrc/cliploss.cu, ce_backward_kernel), and
    the peer variant hands its ring slot back: a second backward through the same node (retain_graph=True, two losses sharing
    the node) would silently compute gradients from gradients.  Make that a loud error instead."""
    if getattr(ctx, "b200clip_consumed", False):
        raise RuntimeError("b200clip ClipLoss: backward was already run through this loss node — the fused kernels consume the "
                           "saved logits in place, so retain_graph / a second backward is not supported; call the loss again")
    ctx.b200clip_consumed = True


def _f32c(t: torch.Tensor) -> torch.Tensor:
    if t.dtype is torch.float32 and t.is_contiguous():   # the usual case; only called inside Function.forward / backward (no graph)
        return t
    t = t.detach()
    if t.dtype != torch.float32:
        t = t.float()
    return t if t.is_contiguous() else t.contiguous()


class _FusedClipLoss(torch.autograd.Function):
    """forward = `b200clip_cliploss_forward` (2 launches; keeps raw logits + row log-sum-exp in a workspace),
    backward = `b200clip_cliploss_backward` (2 launches; reads the upstream gradient from device memory)."""

    @staticmethod
    def forward(ctx, img_loc, txt_loc, all_img, all_txt, logit_scale, rank: int):
        ops_f32 = [_f32c(t) for t in (img_loc, txt_loc, all_img, all_txt)]
        scale = _f32c(logit_scale).reshape(())
        loss, ws = ops.cliploss_forward(*ops_f32, scale, rank)
        ctx.rank = rank
        ctx.dtypes = [t.dtype for t in (img_loc, txt_loc, all_img, all_txt, logit_scale)]
        ctx.scale_shape = logit_scale.shape
        ctx.save_for_backward(*ops_f32, scale, ws)
        return loss

    @staticmethod
    def backward(ctx, g):
        *ops_f32, scale, ws = ctx.saved_tensors
        needs = list(ctx.needs_input_grad[:5])
        if not any(needs):
            return (None,) * 6
        _consume(ctx)
        grads = ops.cliploss_backward(*ops_f32, scale, ctx.rank, ws, _f32c(g).reshape(()), needs)
        out = []
        for i, (gr, dt) in enumerate(zip(grads, ctx.dtypes)):
            if gr is None:
                out.append(None)
                continue
            if gr.dtype != dt:
                gr = gr.to(dt)
            out.append(gr.reshape(ctx.scale_shape) if i == 4 else gr)
        return (*out, None)


class _SingleClipLoss(torch.autograd.Function):
    """world_size == 1 as one node over (image_features, text_features, logit_scale): the features are both the row and the
    column operands of the two logit blocks, and the backward writes their TOTAL gradients with one two-segment GEMM launch
    (`b200clip_cliploss_single_backward`) instead of handing autograd a row and a column gradient per tensor to add up."""

    @staticmethod
    def forward(ctx, image_features, text_features, logit_scale):
        img, txt = _f32c(image_features), _f32c(text_features)
        scale = _f32c(logit_scale).reshape(())
        loss, ws = ops.cliploss_forward(img, txt, img, txt, scale, 0)
        ctx.meta = (image_features.dtype, text_features.dtype, logit_scale.dtype, logit_scale.shape)
        ctx.save_for_backward(img, txt, scale, ws)
        return loss

    @staticmethod
    def backward(ctx, g):
        needs = ctx.needs_input_grad
        if not (needs[0] or needs[1] or needs[2]):
            return None, None, None
        _consume(ctx)
        img, txt, scale, ws = ctx.saved_tensors
        dt_i, dt_t, dt_s, s_shape = ctx.meta
        d_i, d_t, d_s = ops.cliploss_single_backward(img, txt, scale, ws, _f32c(g).reshape(()), needs)
        if d_i is not None and dt_i is not torch.float32:
            d_i = d_i.to(dt_i)
        if d_t is not None and dt_t is not torch.float32:
            d_t = d_t.to(dt_t)
        if d_s is not None:
            d_s = (d_s if dt_s is torch.float32 else d_s.to(dt_s)).reshape(s_shape)
        return d_i, d_t, d_s


class _DistLocalClipLoss(torch.autograd.Function):
    """--local-loss --gather-with-grad on world_size > 1 as ONE autograd node: pack img | txt, one all-gather, the packed
    forward kernels; backward = packed backward kernels (local-row gradients folded into the gathered gradient) + one
    reduce-scatter.  Same values as gather_features + the generic node, without the slice / cat / add kernels in between."""

    @staticmethod
    def forward(ctx, image_features, text_features, logit_scale, rank: int, world_size: int, group):
        n, D = image_features.shape
        both = torch.cat([_f32c(image_features), _f32c(text_features)], dim=1)          # [n, 2D]
        gathered = torch.empty((world_size * n, 2 * D), dtype=torch.float32, device=both.device)
        if both.is_cuda:
            dist.all_gather_into_tensor(gathered, both, group=group)
        else:  # pragma: no cover  (the kernels below need CUDA; kept so the failure is the loud one from ops)
            dist.all_gather(list(gathered.chunk(world_size, dim=0)), both, group=group)
        scale = _f32c(logit_scale).reshape(())
        loss, ws = ops.cliploss_packed_forward(gathered, scale, rank, n)
        ctx.meta = (rank, n, D, group, image_features.dtype, text_features.dtype, logit_scale.dtype, logit_scale.shape)
        ctx.save_for_backward(gathered, scale, ws)
        return loss

    @staticmethod
    def backward(ctx, g):
        _consume(ctx)
        gathered, scale, ws = ctx.saved_tensors
        rank, n, D, group, dt_i, dt_t, dt_s, s_shape = ctx.meta
        d_g, d_s = ops.cliploss_packed_backward(gathered, scale, rank, n, ws, _f32c(g).reshape(()), ctx.needs_input_grad[2])
        d_both = torch.empty((n, 2 * D), dtype=torch.float32, device=d_g.device)
        dist.reduce_scatter_tensor(d_both, d_g, op=dist.ReduceOp.SUM, group=group)
        d_i = d_both[:, :D].to(dt_i) if ctx.needs_input_grad[0] else None
        d_t = d_both[:, D:].to(dt_t) if ctx.needs_input_grad[1] else None
        d_sc = d_s.to(dt_s).reshape(s_shape) if d_s is not None else None
        return d_i, d_t, d_sc, None, None, None


class _PeerLocalClipLoss(torch.autograd.Function):
    """The same node over peer memory (csrc/p2p.cu, open_clip/peer.py): no NCCL collective on the step.  forward: every
    rank stores its fp32 img | txt rows into all gather buffers and waits on flags (one launch) + the packed forward kernels;
    backward: the gradient GEMM stores each rank's block straight into that rank's receive buffer, one launch publishes /
    waits / sums.  Values are those of `_DistLocalClipLoss` (the sum over ranks is taken in rank order instead of NCCL's)."""

    @staticmethod
    def forward(ctx, image_features, text_features, logit_scale, rank: int, world_size: int, ex):
        n, D = image_features.shape
        img, txt = image_features.detach(), text_features.detach()
        if img.dtype not in (torch.float32, torch.bfloat16, torch.float16):
            img = img.float()
        if txt.dtype != img.dtype:
            txt = txt.to(img.dtype)
        hold = any(ctx.needs_input_grad[:3])
        gathered, slot = ex.all_gather(img.contiguous(), txt.contiguous(), hold=hold)
        scale = _f32c(logit_scale).reshape(())
        loss, ws = ops.cliploss_packed_forward(gathered, scale, rank, n, ws=ex.ws[slot])
        ctx.meta = (rank, n, D, slot, image_features.dtype, text_features.dtype, logit_scale.dtype, logit_scale.shape)
        ctx.ex = ex
        ctx.token = peer._SlotToken(ex, slot) if hold else None
        ctx.save_for_backward(scale)
        return loss

    @staticmethod
    def backward(ctx, g):
        _consume(ctx)
        scale, = ctx.saved_tensors
        rank, n, D, slot, dt_i, dt_t, dt_s, s_shape = ctx.meta
        ex = ctx.ex
        ws = ex.ws[slot]
        d_s = ops.cliploss_packed_backward_p2p(ex.gathered_view(slot), scale, rank, n, ws, _f32c(g).reshape(()), ex.rs_dst[slot],
                                               ctx.needs_input_grad[2])
        d_both = ex.reduce_scatter_finish(slot)   # [2, n, D]: dense halves
        ctx.token = None                       # the ring slot may be reused by a later forward
        d_i = d_both[0].to(dt_i) if ctx.needs_input_grad[0] else None
        d_t = d_both[1].to(dt_t) if ctx.needs_input_grad[1] else None
        d_sc = d_s.to(dt_s).reshape(s_shape) if d_s is not None else None
        return d_i, d_t, d_sc, None, None, None


class ClipLoss(nn.Module):
    def __init__(self, local_loss=False, gather_with_grad=False, cache_labels=False, rank=0, world_size=1, use_horovod=False):
        super().__init__()
        self.local_loss = local_loss
        self.gather_with_grad = gather_with_grad
        self.cache_labels = cache_labels
        self.rank = rank
        self.world_size = world_size
        self.use_horovod = use_horovod
        # cache state (kept for API compatibility; the fused kernel derives labels from (n, rank) itself)
        self.prev_num_logits = 0
        self.labels = {}
        self._nccl = None      # backend of the default process group, looked up once

    def get_ground_truth(self, device, num_logits) -> torch.Tensor:
        if self.prev_num_logits != num_logits or device not in self.labels:
            labels = local_labels(num_logits, self.rank, self.world_size, self.local_loss, device)
            if self.cache_labels:
                self.labels[device] = labels
                self.prev_num_logits = num_logits
        else:
            labels = self.labels[device]
        return labels

    def _operands(self, image_features, text_features):
        """-> (img_rows, txt_rows, all_img, all_txt, rank_offset): the row / column operands of the two logit blocks."""
        if self.world_size > 1:
            all_img, all_txt = gather_features(image_features, text_features, self.local_loss, self.gather_with_grad,
                                               self.rank, self.world_size, self.use_horovod)
            if self.local_loss:
                return image_features, text_features, all_img, all_txt, self.rank
            return all_img, all_txt, all_img, all_txt, 0
        return image_features, text_features, image_features, text_features, 0

    def get_logits(self, image_features, text_features, logit_scale):
        """Materialised logits (API compatibility; the loss itself never goes through this)."""
        rows_i, rows_t, all_img, all_txt, _ = self._operands(image_features, text_features)
        s = float(logit_scale)
        li, _, _ = ops.zeroshot(rows_i.float(), all_txt.float(), 0, normalize_img=False, logit_scale=s)
        lt, _, _ = ops.zeroshot(rows_t.float(), all_img.float(), 0, normalize_img=False, logit_scale=s)
        return li, lt

    def forward(self, image_features, text_features, logit_scale, output_dict=False):
        if not image_features.is_cuda:
            raise L.B200ClipError("ClipLoss: CUDA feature tensors required — this path has no CPU fallback")
        if not torch.is_tensor(logit_scale):
            logit_scale = torch.tensor(float(logit_scale), device=image_features.device)
        if self.world_size > 1 and self.local_loss and self.gather_with_grad and not self.use_horovod:
            ex = None
            if self._nccl is None:
                self._nccl = dist.get_backend() == "nccl"
            if self._nccl:
                n, D = image_features.shape
                ex = peer.get_exchange(n, D, self.rank, self.world_size, image_features.device)
            if ex is not None:
                total_loss = _PeerLocalClipLoss.apply(image_features, text_features, logit_scale, self.rank, self.world_size, ex)
            else:
                total_loss = _DistLocalClipLoss.apply(image_features, text_features, logit_scale, self.rank, self.world_size, None)
        elif self.world_size == 1:
            total_loss = _SingleClipLoss.apply(image_features, text_features, logit_scale)
        else:
            rows_i, rows_t, all_img, all_txt, rank = self._operands(image_features, text_features)
            total_loss = _FusedClipLoss.apply(rows_i, rows_t, all_img, all_txt, logit_scale, rank)
        return {"contrastive_loss": total_loss} if output_dict else total_loss
