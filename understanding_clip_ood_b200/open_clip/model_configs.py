"""Architecture registry for the ViT and ModifiedResNet CLIP families on the hot path.

Same schema as the reference's JSON configs (deps/open_clip/src/open_clip/model_configs/ViT-B-32.json etc.:
`embed_dim`, `vision_cfg`, `text_cfg`, optional `quick_gelu`).  Only the ViT towers this framework
accelerates are registered; `create_model` also accepts `embed_dim=` / `vision_cfg=` / `text_cfg=`
keyword overrides exactly like the reference (factory.py:260), which is how arbitrary widths / depths /
patch sizes are built.
"""
from __future__ import annotations

import copy


def _vit(embed_dim, image_size, v_layers, v_width, patch, t_width, t_heads, t_layers=12, quick_gelu=False):
    cfg = {
        "embed_dim": embed_dim,
        "vision_cfg": {"image_size": image_size, "layers": v_layers, "width": v_width, "patch_size": patch},
        "text_cfg": {"context_length": 77, "vocab_size": 49408, "width": t_width, "heads": t_heads, "layers": t_layers},
    }
    if quick_gelu:
        cfg["quick_gelu"] = True
    return cfg


_MODEL_CONFIGS = {
    "ViT-B-32": _vit(512, 224, 12, 768, 32, 512, 8),
    "ViT-B-32-quickgelu": _vit(512, 224, 12, 768, 32, 512, 8, quick_gelu=True),
    "ViT-B-32-256": _vit(512, 256, 12, 768, 32, 512, 8),
    "ViT-B-16": _vit(512, 224, 12, 768, 16, 512, 8),
    "ViT-B-16-quickgelu": _vit(512, 224, 12, 768, 16, 512, 8, quick_gelu=True),
    "ViT-L-14": _vit(768, 224, 24, 1024, 14, 768, 12),
    "ViT-L-14-quickgelu": _vit(768, 224, 24, 1024, 14, 768, 12, quick_gelu=True),
    "ViT-L-14-336": _vit(768, 336, 24, 1024, 14, 768, 12),
    "ViT-L-16": _vit(768, 224, 24, 1024, 16, 768, 12),
}

def _rn(embed_dim, image_size, v_layers, v_width, t_width, t_heads, quick_gelu=False):
    cfg = {
        "embed_dim": embed_dim,
        "vision_cfg": {"image_size": image_size, "layers": list(v_layers), "width": v_width, "patch_size": None},
        "text_cfg": {"context_length": 77, "vocab_size": 49408, "width": t_width, "heads": t_heads, "layers": 12},
    }
    if quick_gelu:
        cfg["quick_gelu"] = True
    return cfg


# ModifiedResNet towers (model_configs/RN50.json ...): the model the paper trains (slurm/train-clip.sh:114)
_MODEL_CONFIGS.update({
    "RN50": _rn(1024, 224, (3, 4, 6, 3), 64, 512, 8),
    "RN50-quickgelu": _rn(1024, 224, (3, 4, 6, 3), 64, 512, 8, quick_gelu=True),
    "RN101": _rn(512, 224, (3, 4, 23, 3), 64, 512, 8),
    "RN101-quickgelu": _rn(512, 224, (3, 4, 23, 3), 64, 512, 8, quick_gelu=True),
    "RN50x4": _rn(640, 288, (4, 6, 10, 6), 80, 640, 10),
    "RN50x16": _rn(768, 384, (6, 8, 18, 8), 96, 768, 12),
    "RN50x64": _rn(1024, 448, (3, 15, 36, 10), 128, 1024, 16),
})

VISION_DEFAULTS = {"layers": 12, "width": 768, "head_width": 64, "mlp_ratio": 4.0, "patch_size": 16, "image_size": 224}
TEXT_DEFAULTS = {"context_length": 77, "vocab_size": 49408, "width": 512, "heads": 8, "layers": 12, "mlp_ratio": 4.0}


def list_models():
    return sorted(_MODEL_CONFIGS)


def get_model_config(model_name: str):
    """Deep copy of the config dict, or None when the name is unknown (reference: factory.py:70-74)."""
    cfg = _MODEL_CONFIGS.get(model_name)
    return copy.deepcopy(cfg) if cfg is not None else None
