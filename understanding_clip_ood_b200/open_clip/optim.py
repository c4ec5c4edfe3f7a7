"""Fused AdamW for the training path: a `torch.optim.Optimizer` with torch.optim.AdamW's constructor and update rule
(decoupled weight decay, bias-corrected moments; the optimizer the reference's train loop builds,
deps/open_clip/src/training/main.py:299-326) whose `step()` is ONE kernel launch per parameter group
(`b200clip_adamw_step`, csrc/optim.cu) instead of a dozen element-wise launches per parameter.

Moments are fp32 whatever the parameter dtype.  No CPU fallback: parameters must live on a CUDA device.
"""
from __future__ import annotations

import ctypes as C

import torch

from .. import _lib as L

__all__ = ["AdamW"]


class AdamW(torch.optim.Optimizer):
    def __init__(self, params, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 1e-2):
        if lr < 0 or eps < 0 or not (0 <= betas[0] < 1 and 0 <= betas[1] < 1) or weight_decay < 0:
            raise ValueError("invalid AdamW hyper-parameters")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        self._tables: dict = {}

    def _table(self, gi: int, group, params):
        """Device tables of one group.  The chunk lists and the parameter / moment pointers are rebuilt only when the set of
        parameters changes; the GRADIENT pointers are re-pointed whenever they differ from the last step's (fresh gradient buffers
        per backward are the rule, and the caching allocator does not promise the same addresses): an in-place update of the host
        copy + one asynchronous upload from pinned memory — no synchronisation, no rebuild of the 37 k-entry chunk lists."""
        skey = tuple((p.data_ptr(), p.dtype, p.numel()) for p in params)
        dkey = tuple((p.grad.data_ptr(), p.grad.dtype) for p in params)
        cached = self._tables.get(gi)
        dev = params[0].device
        if cached is None or cached["skey"] != skey:
            chunk = int(L.load().b200clip_adamw_chunk())
            items = (L.AdamWTensor * len(params))()
            chunk_item, chunk_off = [], []
            for i, p in enumerate(params):
                st = self.state[p]
                if not st:
                    st["step"] = 0
                    st["exp_avg"] = torch.zeros(p.shape, dtype=torch.float32, device=dev)
                    st["exp_avg_sq"] = torch.zeros(p.shape, dtype=torch.float32, device=dev)
                it = items[i]
                it.param = p.data_ptr()
                it.exp_avg, it.exp_avg_sq = st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr()
                it.count, it.param_dtype = p.numel(), L.dtype_code(p.dtype)
                for off in range(0, p.numel(), chunk):
                    chunk_item.append(i)
                    chunk_off.append(off)
            cached = {"skey": skey, "dkey": None, "items": items,
                      "raw": torch.empty(C.sizeof(items), dtype=torch.uint8, device=dev),
                      "ci": torch.tensor(chunk_item, dtype=torch.int32, device=dev),
                      "co": torch.tensor(chunk_off, dtype=torch.int64, device=dev)}
            self._tables[gi] = cached
        if cached["dkey"] != dkey:
            items = cached["items"]
            for i, p in enumerate(params):
                items[i].grad, items[i].grad_dtype = p.grad.data_ptr(), L.dtype_code(p.grad.dtype)
            # a fresh pinned staging buffer per upload (PyTorch's host allocator recycles it once the copy has run)
            host = torch.empty(C.sizeof(items), dtype=torch.uint8, pin_memory=True)
            host.copy_(torch.frombuffer(items, dtype=torch.uint8))
            cached["raw"].copy_(host, non_blocking=True)
            cached["dkey"] = dkey
        return cached["raw"], cached["ci"], cached["co"]

    @torch.no_grad()
    def step(self, closure=None, grad_scale: float = 1.0):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        lib = L.load()
        for gi, group in enumerate(self.param_groups):
            params = [p for p in group["params"] if p.grad is not None]
            if not params:
                continue
            for p in params:
                if not p.is_cuda:
                    raise L.B200ClipError("AdamW: CUDA parameters required (no CPU fallback)")
                if not p.is_contiguous() or not p.grad.is_contiguous():
                    raise L.B200ClipError("AdamW: parameters and gradients must be contiguous")
            step = self.state[params[0]].get("step", 0) + 1
            with torch.cuda.device(params[0].device):
                raw, ci, co = self._table(gi, group, params)
                for p in params:
                    self.state[p]["step"] = step
                L.check(lib.b200clip_adamw_step(raw.data_ptr(), ci.data_ptr(), co.data_ptr(), ci.numel(), float(group["lr"]),
                                                float(group["betas"][0]), float(group["betas"][1]), float(group["eps"]),
                                                float(group["weight_decay"]), int(step), float(grad_scale), L.stream_ptr()),
                        "b200clip_adamw_step")
            # the kernel updated the parameters in place through raw pointers: tell autograd (and the towers' engines, which key
            # their 16-bit weight copies on the version counter) without launching anything
            torch._C._autograd._unsafe_set_version_counter(params, [p._version + 1 for p in params])
        return loss
