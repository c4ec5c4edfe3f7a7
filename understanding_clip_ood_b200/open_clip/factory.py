"""`create_model` / `create_model_and_transforms` / `create_loss` / `get_tokenizer` with the reference's
signatures (deps/open_clip/src/open_clip/factory.py:84-125,180-429), building the B200-native CLIP module.
Network-dependent branches (HF hub, pretrained tag download, OpenAI JIT archives) are out of scope: a
`pretrained` argument must be a local checkpoint path.
"""
from __future__ import annotations

import logging
import os
from typing import Any, Dict, Optional, Tuple, Union

import torch

from .loss import ClipLoss
from .model import CLIP, convert_weights_to_lp, get_cast_dtype
from .model_configs import get_model_config, list_models

OPENAI_DATASET_MEAN = (0.48145466, 0.4578275, 0.40821073)
OPENAI_DATASET_STD = (0.26862954, 0.26130258, 0.27577711)

__all__ = ["create_model", "create_model_and_transforms", "create_model_from_pretrained", "create_loss", "get_tokenizer",
           "load_state_dict", "load_checkpoint", "image_transform"]


def load_state_dict(checkpoint_path: str, map_location="cpu"):
    """factory.py:128-140: unwrap {'state_dict': ...} and strip a DDP 'module.' prefix."""
    checkpoint = torch.load(checkpoint_path, map_location=map_location, weights_only=False)
    if isinstance(checkpoint, dict) and "state_dict" in checkpoint:
        state_dict = checkpoint["state_dict"]
    else:
        state_dict = checkpoint
    if next(iter(state_dict.items()))[0].startswith("module"):
        state_dict = {k[7:]: v for k, v in state_dict.items()}
    return state_dict


def load_checkpoint(model, checkpoint_path, strict=True):
    state_dict = load_state_dict(checkpoint_path)
    if "logit_bias" not in state_dict and getattr(model, "logit_bias", None) is not None:
        state_dict["logit_bias"] = torch.zeros_like(state_dict["logit_scale"])
    return model.load_state_dict(state_dict, strict=strict)


def create_model(
        model_name: str,
        pretrained: Optional[str] = None,
        precision: str = "fp32",
        device: Union[str, torch.device] = "cpu",
        jit: bool = False,
        force_quick_gelu: bool = False,
        force_custom_text: bool = False,
        force_patch_dropout: Optional[float] = None,
        force_image_size: Optional[Union[int, Tuple[int, int]]] = None,
        force_preprocess_cfg: Optional[Dict[str, Any]] = None,
        pretrained_image: bool = False,
        pretrained_hf: bool = True,
        cache_dir: Optional[str] = None,
        output_dict: Optional[bool] = None,
        require_pretrained: bool = False,
        **model_kwargs,
):
    if jit:
        raise RuntimeError("jit=True is not supported: the forward is a C-ABI call, not TorchScript")
    if force_custom_text or pretrained_image:
        raise RuntimeError("custom text towers / pretrained timm image towers are outside this hot path")
    if force_patch_dropout:
        raise RuntimeError("patch dropout is a training-time augmentation of the eager model; not supported")
    model_name = model_name.replace("/", "-")
    model_cfg = get_model_config(model_name)
    if model_cfg is None:
        logging.error(f"Model config for {model_name} not found; available models {list_models()}.")
        raise RuntimeError(f"Model config for {model_name} not found.")
    if force_quick_gelu:
        model_cfg["quick_gelu"] = True
    if force_image_size is not None:
        model_cfg["vision_cfg"]["image_size"] = force_image_size
    if isinstance(device, str):
        device = torch.device(device)
    amp_dtype = {"amp": torch.float16, "amp_bf16": torch.bfloat16, "amp_bfloat16": torch.bfloat16}.get(precision)
    if amp_dtype is None and precision not in ("fp32", "bf16", "fp16", "pure_bf16", "pure_fp16"):
        raise RuntimeError(f"unknown precision {precision!r}")

    cast_dtype = get_cast_dtype(precision)
    model_cfg = dict(model_cfg, **model_kwargs)  # kwargs override cfg (factory.py:260)
    model = CLIP(**model_cfg, cast_dtype=cast_dtype)

    if precision in ("fp16", "bf16"):
        model.to(device=device)
        convert_weights_to_lp(model, dtype=torch.float16 if precision == "fp16" else torch.bfloat16)
    elif precision in ("pure_fp16", "pure_bf16"):
        model.to(device=device, dtype=torch.float16 if "fp16" in precision else torch.bfloat16)
    else:
        model.to(device=device)
    if amp_dtype is not None:
        # The reference's training default (training/params.py:201-206): fp32 master parameters, 16-bit arithmetic.  There
        # the model is plain fp32 and the train loop opens torch.autocast (training/precision.py:5-12); here the towers run
        # their kernels in the autocast dtype on 16-bit copies of the parameters — under an autocast context, or always when
        # the model was created with an `amp*` precision — and hand fp32 gradients to the master parameters.
        model.visual.compute_dtype = amp_dtype
        model.compute_dtype = amp_dtype

    pretrained_loaded = False
    if pretrained:
        if not os.path.exists(pretrained):
            raise RuntimeError(f"Pretrained weights ({pretrained}) not found for model {model_name}: only local checkpoint "
                               "paths are supported (no network)")
        logging.info(f"Loading pretrained {model_name} weights ({pretrained}).")
        load_checkpoint(model, pretrained)
        pretrained_loaded = True
    if require_pretrained and not pretrained_loaded:
        raise RuntimeError(f"Pretrained weights were required for (model: {model_name}, pretrained: {pretrained}) but not loaded.")

    if output_dict and hasattr(model, "output_dict"):
        model.output_dict = True

    size = model.visual.image_size
    preprocess_cfg = {"size": size, "mode": "RGB", "mean": OPENAI_DATASET_MEAN, "std": OPENAI_DATASET_STD,
                      "interpolation": "bicubic", "resize_mode": "shortest", "fill_color": 0}
    preprocess_cfg.update(force_preprocess_cfg or {})
    model.visual.preprocess_cfg = preprocess_cfg
    return model


def image_transform(image_size, is_train: bool, mean=OPENAI_DATASET_MEAN, std=OPENAI_DATASET_STD):
    """CPU preprocessing pipeline with the reference's defaults (transform.py:274-392: RandomResizedCrop(0.9,1) for
    training; bicubic shortest-side resize + center crop for eval; RGB; OpenAI mean/std).  Input pipeline is outside
    the accelerated path; this exists so `create_model_and_transforms` is a drop-in."""
    from torchvision import transforms as T
    from torchvision.transforms import InterpolationMode

    if isinstance(image_size, (tuple, list)) and image_size[0] == image_size[1]:
        image_size = image_size[0]

    def _to_rgb(img):
        return img.convert("RGB")

    if is_train:
        head = [T.RandomResizedCrop(image_size, scale=(0.9, 1.0), interpolation=InterpolationMode.BICUBIC)]
    else:
        head = [T.Resize(image_size, interpolation=InterpolationMode.BICUBIC), T.CenterCrop(image_size)]
    return T.Compose([*head, _to_rgb, T.ToTensor(), T.Normalize(mean=mean, std=std)])


def create_model_and_transforms(model_name: str, pretrained: Optional[str] = None, precision: str = "fp32",
                                device: Union[str, torch.device] = "cpu", jit: bool = False, force_quick_gelu: bool = False,
                                force_custom_text: bool = False, force_patch_dropout: Optional[float] = None,
                                force_image_size: Optional[Union[int, Tuple[int, int]]] = None,
                                image_mean: Optional[Tuple[float, ...]] = None, image_std: Optional[Tuple[float, ...]] = None,
                                image_interpolation: Optional[str] = None, image_resize_mode: Optional[str] = None,
                                aug_cfg=None, pretrained_image: bool = False, pretrained_hf: bool = True,
                                cache_dir: Optional[str] = None, output_dict: Optional[bool] = None, **model_kwargs):
    force_preprocess_cfg = {k: v for k, v in (("mean", image_mean), ("std", image_std), ("interpolation", image_interpolation),
                                              ("resize_mode", image_resize_mode)) if v is not None}
    model = create_model(model_name, pretrained, precision=precision, device=device, jit=jit, force_quick_gelu=force_quick_gelu,
                         force_custom_text=force_custom_text, force_patch_dropout=force_patch_dropout,
                         force_image_size=force_image_size, force_preprocess_cfg=force_preprocess_cfg,
                         pretrained_image=pretrained_image, pretrained_hf=pretrained_hf, cache_dir=cache_dir,
                         output_dict=output_dict, **model_kwargs)
    pp = model.visual.preprocess_cfg
    preprocess_train = image_transform(pp["size"], True, pp["mean"], pp["std"])
    preprocess_val = image_transform(pp["size"], False, pp["mean"], pp["std"])
    return model, preprocess_train, preprocess_val


def create_model_from_pretrained(model_name: str, pretrained: Optional[str] = None, precision: str = "fp32",
                                 device: Union[str, torch.device] = "cpu", return_transform: bool = True, **kwargs):
    model = create_model(model_name, pretrained, precision=precision, device=device, require_pretrained=True, **kwargs)
    if not return_transform:
        return model
    pp = model.visual.preprocess_cfg
    return model, image_transform(pp["size"], False, pp["mean"], pp["std"])


def create_loss(args):
    """factory.py:338-372 for the ClipLoss branch (distillation / CoCa / SigLIP objectives are out of scope)."""
    if getattr(args, "distill", False) or getattr(args, "siglip", False) or "coca" in getattr(args, "model", "").lower():
        raise RuntimeError("only the ClipLoss objective is on this hot path")
    return ClipLoss(local_loss=args.local_loss, gather_with_grad=args.gather_with_grad, cache_labels=True, rank=args.rank,
                    world_size=args.world_size, use_horovod=getattr(args, "horovod", False))


def get_tokenizer(model_name: str = "", context_length: Optional[int] = None, **kwargs):
    """factory.py:84-125 for the native (non-HF) models: a SimpleTokenizer with the model's context length."""
    from .tokenizer import SimpleTokenizer
    cfg = get_model_config(model_name.replace("/", "-")) or {}
    if context_length is None:
        context_length = cfg.get("text_cfg", {}).get("context_length", 77)
    return SimpleTokenizer(context_length=context_length, **kwargs)
