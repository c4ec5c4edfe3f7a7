"""`OpenCLIP`: the thin CLIP wrapper the zero-shot classifiers and evaluation scripts of the reference are written against
(xclip/open_clip/model.py:11-56), here over the B200-native `CLIP` module.

Contract kept from the reference: `encode_image` / `encode_text` delegate to the wrapped model, `logit_scale` is
`exp(param)` clamped to [0, 100], and `from_pretrained(model_name, ckpt_path=None, **kw)` returns
`(wrapper, preprocess_train, preprocess_val)` with **fp16** as the default precision; a checkpoint may be a bare state
dict or a training checkpoint (`{"state_dict": ...}`, optionally with DistributedDataParallel's `module.` prefix), and a
stored `logit_bias` switches the bias parameter on.
"""
from __future__ import annotations

from typing import Any, Mapping, Optional

import torch

from ..open_clip import CLIP, create_model_and_transforms
from .utils import AbstractCLIP

_DDP_PREFIX = "module."


def _load_checkpoint_weights(path: str) -> dict[str, torch.Tensor]:
    """Weights of a checkpoint file as a plain `{parameter name: tensor}` dict (CPU)."""
    blob: Any = torch.load(path, map_location="cpu", weights_only=False)
    weights: Mapping[str, torch.Tensor] = blob.get("state_dict", blob) if isinstance(blob, Mapping) else blob
    first_key = next(iter(weights), "")
    if first_key.startswith("module"):            # saved from inside DistributedDataParallel
        weights = {name[len(_DDP_PREFIX):]: value for name, value in weights.items()}
    return dict(weights)


class OpenCLIP(AbstractCLIP):
    def __init__(self, clip: CLIP) -> None:
        super().__init__()
        self.clip = clip

    # ---- AbstractCLIP ---------------------------------------------------------------------------
    def encode_image(self, image: torch.Tensor, normalize: bool = False) -> torch.Tensor:
        return self.clip.encode_image(image, normalize=normalize)

    def encode_text(self, text: torch.Tensor, normalize: bool = False) -> torch.Tensor:
        return self.clip.encode_text(text, normalize=normalize)

    @property
    def logit_scale(self) -> torch.Tensor:
        return torch.clamp(torch.exp(self.clip.logit_scale), min=0, max=100)

    @property
    def vocab_size(self) -> int:
        return self.clip.vocab_size

    # ---- construction ---------------------------------------------------------------------------
    @classmethod
    def from_pretrained(cls, model_name: str, ckpt_path: Optional[str] = None, **model_kwargs):
        """-> (OpenCLIP, preprocess_train, preprocess_val)."""
        model_kwargs.setdefault("precision", "fp16")
        weights = _load_checkpoint_weights(ckpt_path) if ckpt_path else None
        if weights is not None and "logit_bias" in weights:
            model_kwargs["init_logit_bias"] = weights["logit_bias"]
        clip, preprocess_train, preprocess_val = create_model_and_transforms(model_name, **model_kwargs)
        if weights:
            clip.load_state_dict(weights)
        return cls(clip), preprocess_train, preprocess_val
