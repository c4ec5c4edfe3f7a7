"""GPU half of the evaluation preprocessing: decoded uint8 image -> bicubic shortest-side resize -> centre crop
-> [3, S, S] uint8, which `encode_image` / `ZeroShotClassifier.predict` take directly (ToTensor + Normalize run inside the
im2col kernel).  Replaces, bit for bit, the PIL pipeline the reference builds in `image_transform(..., is_train=False)`
(deps/open_clip/src/open_clip/transform.py:372-392): torchvision `Resize(S, interpolation=BICUBIC)` + `CenterCrop(S)`.

The arithmetic of Pillow's resampler is integer (8-bit pixels, 22-bit fixed-point coefficients), so parity is exact.  This
module restates how Pillow derives the coefficient tables (Pillow `src/libImaging/Resample.c`: `precompute_coeffs`,
`normalize_coeffs_8bpc`, bicubic filter with a = -0.5, support 2 x max(scale, 1)); the two convolution passes run in
`b200clip_resize_crop_u8` (csrc/preprocess.cu).  JPEG decoding stays on the host (or `torchvision.io.decode_jpeg(device='cuda')`).
"""
from __future__ import annotations

import functools
import math

import numpy as np
import torch

from .. import _lib as L

PRECISION_BITS = 32 - 8 - 2


def _bicubic(x: np.ndarray) -> np.ndarray:
    a = -0.5
    x = np.abs(x)
    near = ((a + 2.0) * x - (a + 3.0)) * x * x + 1
    far = (((x - 5) * x + 8) * x - 4) * a
    return np.where(x < 1.0, near, np.where(x < 2.0, far, 0.0))


def pillow_coeffs(in_size: int, out_size: int):
    """Pillow's `precompute_coeffs` + `normalize_coeffs_8bpc` for a full-axis resize (box = [0, in_size]):
    -> (bounds int32 [out, 2] = (first tap, taps), coeffs int32 [out, ksize])."""
    scale = in_size / out_size
    filterscale = max(scale, 1.0)
    support = 2.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), dtype=np.int32)
    kk = np.zeros((out_size, ksize), dtype=np.float64)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        xmin = int(center - support + 0.5)
        xmin = max(xmin, 0)
        xmax = int(center + support + 0.5)
        xmax = min(xmax, in_size) - xmin
        x = np.arange(xmax, dtype=np.float64)
        w = _bicubic((x + xmin - center + 0.5) * ss)
        ww = 0.0
        for v in w:                       # sequential sum, as the C loop does
            ww += float(v)
        if ww != 0.0:
            w = w / ww
        kk[xx, :xmax] = w
        bounds[xx] = (xmin, xmax)
    fixed = np.where(kk < 0, np.trunc(-0.5 + kk * (1 << PRECISION_BITS)), np.trunc(0.5 + kk * (1 << PRECISION_BITS))).astype(np.int32)
    return bounds, fixed


def resized_size(h: int, w: int, size: int):
    """torchvision `Resize(size: int)`: the shorter side becomes `size`, the longer int(size * long / short)."""
    short, long = (w, h) if w <= h else (h, w)
    new_short, new_long = size, int(size * long / short)
    return (new_long, new_short) if w <= h else (new_short, new_long)      # (new_h, new_w)


@functools.lru_cache(maxsize=256)
def _tables(h: int, w: int, size: int):
    """Tables of the CROP WINDOW of the resized image, as numpy arrays + the source row window the vertical pass touches."""
    new_h, new_w = resized_size(h, w, size)
    if new_h < size or new_w < size:
        raise L.B200ClipError(f"image {h}x{w} resizes to {new_h}x{new_w}: smaller than the {size}x{size} crop")
    top = int(round((new_h - size) / 2.0))          # torchvision CenterCrop
    left = int(round((new_w - size) / 2.0))

    def axis(n_in, n_out, first):
        if n_in == n_out:                           # Pillow skips the pass: identity taps reproduce the bytes
            b = np.stack([np.arange(first, first + size), np.ones(size)], axis=1).astype(np.int32)
            return b, np.full((size, 1), 1 << PRECISION_BITS, dtype=np.int32)
        b, k = pillow_coeffs(n_in, n_out)
        return np.ascontiguousarray(b[first:first + size]), np.ascontiguousarray(k[first:first + size])

    hb, hk = axis(w, new_w, left)
    vb, vk = axis(h, new_h, top)
    y0 = int(vb[:, 0].min())
    y1 = int((vb[:, 0] + vb[:, 1]).max())
    return hb, hk, vb, vk, y0, y1 - y0


_device_tables: dict = {}


def resize_center_crop(image: torch.Tensor, size: int) -> torch.Tensor:
    """image: uint8 CUDA tensor [H, W, 3] (a decoded RGB image) -> uint8 [3, size, size]."""
    if not image.is_cuda:
        raise L.B200ClipError("resize_center_crop: CUDA tensors required (no CPU fallback; the reference's PIL pipeline is the CPU path)")
    if image.dtype != torch.uint8 or image.ndim != 3 or image.shape[2] != 3:
        raise L.B200ClipError(f"resize_center_crop: expected a uint8 [H, W, 3] image, got {image.dtype} {tuple(image.shape)}")
    if image.stride(2) != 1 or image.stride(1) != 3:
        image = image.contiguous()
    h, w = int(image.shape[0]), int(image.shape[1])
    key = (h, w, size, image.device)
    tabs = _device_tables.get(key)
    if tabs is None:
        hb, hk, vb, vk, y0, rows = _tables(h, w, size)
        tabs = tuple(torch.from_numpy(a).to(image.device) for a in (hb, hk, vb, vk)) + (y0, rows)
        if len(_device_tables) > 512:
            _device_tables.clear()
        _device_tables[key] = tabs
    hb, hk, vb, vk, y0, rows = tabs
    tmp = torch.empty((rows, size, 3), dtype=torch.uint8, device=image.device)
    out = torch.empty((3, size, size), dtype=torch.uint8, device=image.device)
    with torch.cuda.device(image.device):
        rc = L.load().b200clip_resize_crop_u8(image.data_ptr(), h, w, image.stride(0), hb.data_ptr(), hk.data_ptr(), hk.shape[1], vb.data_ptr(),
                                              vk.data_ptr(), vk.shape[1], y0, rows, tmp.data_ptr(), out.data_ptr(), size, size, L.stream_ptr())
    L.check(rc, "b200clip_resize_crop_u8")
    return out


class GpuEvalTransform:
    """Callable with the role of `preprocess_val` for already-decoded images: uint8 HWC (numpy array, PIL image or tensor) ->
    uint8 [3, S, S] CUDA tensor; stack the results and hand the batch to `encode_image` / `predict`."""

    def __init__(self, image_size: int, device="cuda"):
        self.size = int(image_size[0] if isinstance(image_size, (tuple, list)) else image_size)
        self.device = torch.device(device)

    def __call__(self, img) -> torch.Tensor:
        if not torch.is_tensor(img):
            arr = np.asarray(img.convert("RGB") if hasattr(img, "convert") else img)
            img = torch.from_numpy(np.ascontiguousarray(arr))
        return resize_center_crop(img.to(self.device, non_blocking=True), self.size)


def emulate_numpy(image: np.ndarray, size: int) -> np.ndarray:
    """The two integer passes in numpy (what the CUDA kernels compute) — used by the CPU tests to pin the TABLES against PIL."""
    h, w = image.shape[:2]
    hb, hk, vb, vk, y0, rows = _tables(h, w, size)
    src = image.astype(np.int64)
    tmp = np.zeros((rows, size, 3), dtype=np.uint8)
    for x in range(size):
        xmin, xs = hb[x]
        acc = (1 << (PRECISION_BITS - 1)) + np.tensordot(src[y0:y0 + rows, xmin:xmin + xs, :], hk[x, :xs].astype(np.int64), axes=([1], [0]))
        tmp[:, x, :] = np.clip(acc >> PRECISION_BITS, 0, 255)
    out = np.zeros((3, size, size), dtype=np.uint8)
    t64 = tmp.astype(np.int64)
    for y in range(size):
        ymin, ys = vb[y]
        acc = (1 << (PRECISION_BITS - 1)) + np.tensordot(vk[y, :ys].astype(np.int64), t64[ymin - y0:ymin - y0 + ys], axes=([0], [0]))
        out[:, y, :] = np.clip(acc >> PRECISION_BITS, 0, 255).T
    return out
