"""Namespace mirror of `label_anything.models` for the LabelAnything path (label_anything/models/__init__.py:9-60):
the reference's callers do `from label_anything.models import model_registry, LabelAnything, ...`; pointing that
import at this module swaps the implementation without touching any caller."""
from .build_encoder import ENCODERS, ViTModelWrapper, build_encoder, build_vit_b, build_vit_h, build_vit_l  # noqa: F401
from .build_lam import (LabelAnything, build_lam, build_lam_no_vit, build_lam_vit_b, build_lam_vit_b_imagenet_i21k,  # noqa: F401
                        build_lam_vit_h, build_lam_vit_l, build_lam_vit_mae_b, build_mask_decoder, model_registry)
from .image_encoder import ImageEncoderViT  # noqa: F401
from .lam import Lam  # noqa: F401
from .mask_decoder import MLP, MaskDecoderLam  # noqa: F401
from .prompt_encoder import PromptEncoder, PromptImageEncoder, RandomMatrixEncoder  # noqa: F401
from .transformer import TwoWayTransformer  # noqa: F401
