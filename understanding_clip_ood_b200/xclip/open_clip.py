"""`OpenCLIP` wrapper of the reference (xclip/open_clip/model.py:11-56) over the B200-native CLIP module."""
from __future__ import annotations

import torch

from ..open_clip import CLIP, create_model_and_transforms
from .utils import AbstractCLIP


class OpenCLIP(AbstractCLIP):
    def __init__(self, clip: CLIP) -> None:
        super().__init__()
        self.clip = clip

    def encode_image(self, image: torch.Tensor, normalize: bool = False) -> torch.Tensor:
        return self.clip.encode_image(image, normalize=normalize)

    def encode_text(self, text: torch.Tensor, normalize: bool = False) -> torch.Tensor:
        return self.clip.encode_text(text, normalize=normalize)

    @property
    def logit_scale(self) -> torch.Tensor:
        return self.clip.logit_scale.exp().clamp(0, 100)

    @property
    def vocab_size(self) -> int:
        return self.clip.vocab_size

    @classmethod
    def from_pretrained(cls, model_name: str, ckpt_path: str | None = None, **model_kwargs):
        """-> (OpenCLIP, preprocess_train, preprocess_val); fp16 is the default precision, as in the reference (:35)."""
        model_kwargs["precision"] = model_kwargs.get("precision", "fp16")
        state_dict = None
        if ckpt_path:
            state_dict = torch.load(ckpt_path, map_location="cpu", weights_only=False)
            state_dict = state_dict["state_dict"] if "state_dict" in state_dict else state_dict
            if next(iter(state_dict.items()))[0].startswith("module"):
                state_dict = {k[len("module."):]: v for k, v in state_dict.items()}
            if "logit_bias" in state_dict:
                model_kwargs["init_logit_bias"] = state_dict["logit_bias"]
        clip, preprocess_train, preprocess_val = create_model_and_transforms(model_name, **model_kwargs)
        if state_dict:
            clip.load_state_dict(state_dict)
        return cls(clip), preprocess_train, preprocess_val
