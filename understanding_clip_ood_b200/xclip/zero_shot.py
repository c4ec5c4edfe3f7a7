"""Zero-shot classifiers with the reference's class names, constructor arguments and method signatures
(xclip/zero_shot.py:11-109, 112-240), running on the fused B200 kernels:

  * prompt embedding : ALL prompts go through `encode_text` in a few large batches (the reference issues one
    86-prompt call per class, :226-236), the normalise -> template-mean -> renormalise tail is one kernel
    (`b200clip_class_mean`);
  * similarity stage : L2-normalise + image x class-prompt logits + arg-max / top-k in ONE kernel
    (`b200clip_zeroshot`), ties resolved to the lower class index like `torch.argmax`.

`predict_from_features(...)[“pred”]` returns int64 class indices (or the raw cosine logits in the feature
dtype with `return_scores=True`), exactly as the reference does.
"""
from __future__ import annotations

import json
from abc import ABC, abstractmethod
from pathlib import Path
from typing import Callable

import torch

from .. import _lib as L
from .. import ops
from .utils import AbstractCLIP, identity

_TEMPLATES_PATH = Path(__file__).resolve().parent.parent / "data" / "openai_templates.json"


def _openai_templates() -> list[str]:
    return json.loads(_TEMPLATES_PATH.read_text())


def _features_2d(x: torch.Tensor) -> torch.Tensor:
    return x.reshape(-1, x.shape[-1])


class AbstractZeroShotClassifier(ABC):
    #: prompts per encode_text call when embedding the class prompts
    text_batch_size = 8192

    def __init__(self, clip: AbstractCLIP, prompts: torch.Tensor) -> None:
        self.clip = clip
        self.clip.eval()
        if not torch.cuda.is_available():
            raise L.B200ClipError("the B200-native zero-shot classifier needs a CUDA device (no CPU fallback)")
        self.device = "cuda"
        self.clip.to(self.device)
        if getattr(self.clip, "uses_one_hot_encoding", False):
            raise NotImplementedError("one-hot text encoders are not part of this hot path")
        self.prompts = prompts
        feature_shapes = self.prompts.shape[:-1]
        input_ids = self.prompts.reshape(feature_shapes.numel(), self.prompts.shape[-1])
        with torch.inference_mode():
            txt_feat = self._encode_text_batched(input_ids)
            assert txt_feat.ndim == 2
            txt_feat = ops.normalize(txt_feat)
            txt_feat = txt_feat.reshape(*feature_shapes, txt_feat.size(-1))
        self.prompt_feat = txt_feat

    def _encode_text_batched(self, input_ids: torch.Tensor) -> torch.Tensor:
        chunks = []
        for i in range(0, input_ids.shape[0], self.text_batch_size):
            chunks.append(self.clip.encode_text(input_ids[i:i + self.text_batch_size].to(self.device)))
        return chunks[0] if len(chunks) == 1 else torch.cat(chunks, dim=0)

    @torch.inference_mode()
    def _compute_img_feat(self, img: torch.Tensor) -> torch.Tensor:
        """Encode images and L2-normalise (xclip/zero_shot.py:42-52)."""
        assert img.ndim in [3, 4]
        img = img.unsqueeze(0) if img.ndim == 3 else img
        img_feat = self.clip.encode_image(img.to(self.device), normalize=True)   # normalise fused into the tower call
        assert img_feat.ndim == 2
        return img_feat

    def _similarity(self, img_feat: torch.Tensor, k: int, want_logits: bool, normalize_img: bool = False):
        img_feat = img_feat.to(self.device)
        assert img_feat.ndim == 2
        prompt = _features_2d(self.prompt_feat)
        if prompt.dtype != img_feat.dtype:
            prompt = prompt.to(img_feat.dtype)
        return ops.zeroshot(img_feat, prompt, k, normalize_img=normalize_img, want_logits=want_logits)

    @torch.inference_mode()
    def _compute_logits(self, img_feat: torch.Tensor) -> torch.Tensor:
        """(batch, embed) x (embed, *features) -> (batch, *features); no logit scale (xclip/zero_shot.py:54-60)."""
        logits, _, _ = self._similarity(img_feat, 0, True)
        return logits.to(img_feat.dtype).reshape(img_feat.shape[0], *self.prompt_feat.shape[:-1])

    @torch.inference_mode()
    def _compute_scores(self, img_feat: torch.Tensor) -> torch.Tensor:
        """softmax(logit_scale * logits) over all feature dims (xclip/zero_shot.py:62-67); not on the hot path."""
        logits = self.clip.logit_scale * self._compute_logits(img_feat)
        return torch.softmax(logits.flatten(1), dim=1).reshape_as(logits)

    @abstractmethod
    def variance_from_features(self, img_feat: torch.Tensor) -> dict[str, torch.Tensor]:
        pass

    @abstractmethod
    def predict_from_features(self, img_feat: torch.Tensor, return_scores: bool = False) -> dict[str, torch.Tensor]:
        pass

    def predict(self, img: torch.Tensor, return_scores: bool = False) -> dict[str, torch.Tensor]:
        return self.predict_from_features(self._compute_img_feat(img), return_scores=return_scores)


class ZeroShotClassifier(AbstractZeroShotClassifier):
    def __init__(self, clip: AbstractCLIP, tokenizer, idx2class: dict[int, str] | list[str],
                 prompt_fn: Callable[[str], str] = identity) -> None:
        prompts = tokenizer([prompt_fn(idx2class[idx]) for idx in range(len(idx2class))])
        super().__init__(clip, prompts)

    def variance_from_features(self, img_feat: torch.Tensor) -> dict[str, torch.Tensor]:
        scores = self._compute_logits(img_feat.to(self.device))
        return {"variance": scores.var()}

    @torch.inference_mode()
    def predict_from_features(self, img_feat: torch.Tensor, return_scores: bool = False) -> dict[str, torch.Tensor]:
        if return_scores:
            return {"pred": self._compute_logits(img_feat.to(self.device))}
        _, idx, _ = self._similarity(img_feat, 1, False)
        return {"pred": idx[:, 0]}

    @torch.inference_mode()
    def predict_topk_from_features(self, img_feat: torch.Tensor, k: int = 5) -> dict[str, torch.Tensor]:
        """Extension: top-k class indices / logits (training/zero_shot.py:11-14 semantics), k <= 8."""
        _, idx, val = self._similarity(img_feat, k, False)
        return {"pred": idx, "scores": val}


class OpenAIZeroShotClassifier(ZeroShotClassifier):
    """Prompt-ensemble classifier: 86 OpenAI templates per class (xclip/zero_shot.py:112-240)."""

    templates = _openai_templates() if _TEMPLATES_PATH.exists() else []

    def __init__(self, clip: AbstractCLIP, tokenizer, idx2class: dict[int, str] | list[str],
                 domain_invariant: bool = False) -> None:
        self.clip = clip
        self.clip.eval()
        if not torch.cuda.is_available():
            raise L.B200ClipError("the B200-native zero-shot classifier needs a CUDA device (no CPU fallback)")
        self.device = "cuda"
        self.clip.to(self.device)
        if domain_invariant:
            self.templates = [t for t in self.templates
                              if any(d in t for d in ["clipart", "infograph", "painting", "quickdraw", "sketch"])]
        classnames = [idx2class[idx] for idx in range(len(idx2class))]
        texts = [template.format(c) for c in classnames for template in self.templates]   # class-major, template-minor
        with torch.inference_mode():
            tokens = tokenizer(texts)
            self.prompts = tokens.reshape(len(classnames), len(self.templates), -1)
            emb = self._encode_text_batched(tokens)
            self.prompt_feat = ops.class_mean(emb, len(classnames), len(self.templates))

    @classmethod
    def from_tokens(cls, clip: AbstractCLIP, tokens: torch.Tensor, classes: int, templates: int):
        """Build from pre-tokenised prompts [classes*templates, ctx] (class-major); used where no BPE table is available."""
        self = cls.__new__(cls)
        self.clip = clip
        self.clip.eval()
        self.device = "cuda"
        self.clip.to(self.device)
        with torch.inference_mode():
            self.prompts = tokens.reshape(classes, templates, -1)
            emb = self._encode_text_batched(tokens)
            self.prompt_feat = ops.class_mean(emb, classes, templates)
        return self
