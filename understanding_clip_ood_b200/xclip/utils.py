"""Interface types of the `xclip` layer: what a CLIP wrapper must offer to the zero-shot classifiers.

Same contract as the reference's `xclip/utils.py:9-31` (names, argument meaning, `uses_one_hot_encoding` default), written
for this package: the wrapper in `xclip/open_clip.py` implements it on top of the B200 towers.
"""
from __future__ import annotations

import abc
from typing import Protocol, Sequence, TypeVar, Union, runtime_checkable

import torch
from torch import nn

__all__ = ["AbstractCLIP", "TokenizerLike", "identity"]

_X = TypeVar("_X")


def _not_provided(obj: object, member: str) -> NotImplementedError:
    return NotImplementedError(f"{type(obj).__name__} must provide `{member}` to be used as a CLIP wrapper")


class AbstractCLIP(nn.Module, abc.ABC):
    """A model with an image encoder, a text encoder and a (clamped, exponentiated) logit scale.

    `encode_image(image [B,3,S,S], normalize)` and `encode_text(token ids [T,ctx], normalize)` return `[*, D]` features,
    L2-normalised when `normalize` is true; `logit_scale` is a 0-dim tensor."""

    @abc.abstractmethod
    def encode_image(self, image: torch.Tensor, normalize: bool = False) -> torch.Tensor:
        raise _not_provided(self, "encode_image")

    @abc.abstractmethod
    def encode_text(self, text: torch.Tensor, normalize: bool = False) -> torch.Tensor:
        raise _not_provided(self, "encode_text")

    @property
    @abc.abstractmethod
    def logit_scale(self) -> torch.Tensor:
        raise _not_provided(self, "logit_scale")

    @property
    def uses_one_hot_encoding(self) -> bool:
        """Text input is one-hot vectors instead of token ids (never the case for the OpenCLIP towers)."""
        return False


@runtime_checkable
class TokenizerLike(Protocol):
    """Anything that maps a prompt or a list of prompts to int64 token ids `[n, context_length]`."""

    def __call__(self, text: Union[str, Sequence[str]]) -> torch.Tensor: ...


def identity(x: _X) -> _X:
    """Default `prompt_fn` of ZeroShotClassifier: the class name is the prompt."""
    return x
