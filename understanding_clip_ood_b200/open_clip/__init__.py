"""B200-native drop-in for the part of `open_clip` that the reference's CLIP hot path uses
(deps/open_clip/src/open_clip/__init__.py exports the same names)."""
from .factory import (OPENAI_DATASET_MEAN, OPENAI_DATASET_STD, create_loss, create_model, create_model_and_transforms,
                      create_model_from_pretrained, get_tokenizer, image_transform, load_checkpoint, load_state_dict)
from .loss import ClipLoss, gather_features
from .optim import AdamW
from .model import CLIP, VisionTower, convert_weights_to_fp16, convert_weights_to_lp, get_cast_dtype, get_input_dtype
from .model_configs import get_model_config, list_models
from .tokenizer import SimpleTokenizer, tokenize

__version__ = "2.24.0+b200"
