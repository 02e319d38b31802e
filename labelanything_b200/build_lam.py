"""Model builders and the `LabelAnything` hub wrapper, with the reference's names and keyword arguments.

Mirror of label_anything/models/build_lam.py (`_build_lam` :96-235, `build_mask_decoder` :238-297,
`LabelAnythingConfig` :402-464, `LabelAnything` :467-508) and models/hfhub.py (`has_config` :50-67).  Options that
select model variants outside the hot path (BinaryLam, pyramids, Affinity decoders, TokenPool prompt encoder, one-way
/ identity fusion) raise NotImplementedError instead of silently building something else.
"""
from __future__ import annotations

import inspect

import torch.nn as nn
from huggingface_hub import PyTorchModelHubMixin

from .build_encoder import (ENCODERS, build_encoder, build_vit_b, build_vit_b_imagenet_i21k, build_vit_b_mae,
                            build_vit_h, build_vit_l)
from .common import SAM_EMBED_DIM, LayerNorm2d
from .lam import Lam
from .mask_decoder import MaskDecoderLam
from .prompt_encoder import PromptImageEncoder, RandomMatrixEncoder
from .transformer import TwoWayTransformer
from .utils import load_state_dict, torch_dict_load

_CLASS_ENCODERS = {"RandomMatrixEncoder": RandomMatrixEncoder}


def has_config(func):
    """Store the call's arguments (defaults included) as `self.config`; accept `config=` (hfhub.py:50-67)."""
    signature = inspect.signature(func)

    def wrapper(self, *args, **kwargs):
        if "config" in kwargs:
            config = kwargs.pop("config")
            kwargs.update(**config)
        self.config = {k: v.default if (i - 1) >= len(args) else args[i - 1]
                       for i, (k, v) in enumerate(signature.parameters.items())
                       if v.default is not inspect.Parameter.empty}
        self.config.update(**kwargs)
        func(self, **kwargs)

    return wrapper


def build_mask_decoder(embed_dim, decoder_attention_downsample_rate, few_type="Prototype",
                       fusion_transformer="TwoWayTransformer", segment_example_logits=False, spatial_convs=None,
                       classification_layer_downsample_rate=8, conv_upsample_stride=2, transformer_feature_size=None,
                       dropout=0.0, class_fusion="sum", prototype_merge=False, classification_levels=1,
                       conv_classification=False, transformer_keys_are_images=True):
    """build_lam.py:238-297"""
    if few_type != "Prototype":
        raise NotImplementedError(f"few_type {few_type!r} (Affinity decoders) is outside the native hot path")
    if fusion_transformer != "TwoWayTransformer":
        raise NotImplementedError(f"fusion_transformer {fusion_transformer!r} is outside the native hot path")
    transformer = TwoWayTransformer(depth=2, embedding_dim=embed_dim, mlp_dim=2048, num_heads=8,
                                    attention_downsample_rate=decoder_attention_downsample_rate, dropout=dropout)
    return MaskDecoderLam(transformer_dim=embed_dim, spatial_convs=spatial_convs, transformer=transformer,
                          segment_example_logits=segment_example_logits,
                          classification_layer_downsample_rate=classification_layer_downsample_rate,
                          conv_upsample_stride=conv_upsample_stride, classification_levels=classification_levels,
                          dropout=dropout, conv_classification=conv_classification)


def _build_lam(build_vit, checkpoint=None, use_sam_checkpoint=False, use_vit_sam_neck=True,
               ignore_encoder_checkpoint=False, use_vit=True, image_embed_dim=SAM_EMBED_DIM, embed_dim=SAM_EMBED_DIM,
               image_size=1024, vit_patch_size=16, class_attention=False, example_attention=False,
               example_class_attention=True, class_embedding_dim=None, spatial_convs=None,
               encoder_attention_downsample_rate: int = 2, decoder_attention_downsample_rate: int = 2,
               classification_layer_downsample_rate: int = 8, conv_classification=False,
               use_support_features_in_prompt_encoder: bool = True, fusion_transformer="TwoWayTransformer",
               classification_levels=1, few_type="Prototype", class_fusion="sum", prompt_encoder=None,
               transformer_keys_are_images=True, transformer_feature_size=None, class_encoder=None,
               segment_example_logits=False, embeddings_per_example=None, embedding_extraction=None,
               dropout: float = 0.0, binary=False, custom_preprocess=True, is_pyramids=False,
               intermediate_channel_sizes=None):
    """build_lam.py:96-235 (same keyword arguments)."""
    if binary or is_pyramids or prompt_encoder == "TokenPool":
        raise NotImplementedError("BinaryLam / pyramid necks / TokenPool prompt encoders are outside the native hot path")
    image_embedding_size = image_size // vit_patch_size
    vit = build_vit(project_last_hidden=use_vit_sam_neck) if use_vit else None
    if class_encoder is not None:
        params = {k: v for k, v in class_encoder.items() if k != "name"}
        if class_encoder["name"] not in _CLASS_ENCODERS:
            raise NotImplementedError(f"class encoder {class_encoder['name']!r} has no native path")
        class_encoder = _CLASS_ENCODERS[class_encoder["name"]](**params)
    else:
        class_encoder = None  # the reference's identity lambda (build_lam.py:143)
    if segment_example_logits or embeddings_per_example:
        raise NotImplementedError("segment_example_logits / embeddings_per_example are outside the native hot path")
    neck = None
    if image_embed_dim != embed_dim:
        neck = nn.Sequential(nn.Conv2d(image_embed_dim, embed_dim, kernel_size=1, bias=False), LayerNorm2d(embed_dim),
                             nn.Conv2d(embed_dim, embed_dim, kernel_size=3, padding=1, bias=False),
                             LayerNorm2d(embed_dim))
    lam = Lam(
        image_size=image_size, image_encoder=vit, neck=neck,
        prompt_encoder=PromptImageEncoder(
            embed_dim=embed_dim, image_embedding_size=(image_embedding_size, image_embedding_size),
            input_image_size=(image_size, image_size), mask_in_chans=16, class_attention=class_attention,
            example_attention=example_attention, example_class_attention=example_class_attention,
            class_embedding_dim=class_embedding_dim, dropout=dropout,
            use_support_features=use_support_features_in_prompt_encoder,
            transformer=TwoWayTransformer(depth=2, embedding_dim=embed_dim, mlp_dim=2048,
                                          attention_downsample_rate=encoder_attention_downsample_rate, num_heads=8,
                                          dropout=dropout),
            class_encoder=class_encoder, embeddings_per_example=embeddings_per_example,
            embedding_extraction=embedding_extraction),
        mask_decoder=build_mask_decoder(
            embed_dim=embed_dim, spatial_convs=spatial_convs, segment_example_logits=segment_example_logits,
            fusion_transformer=fusion_transformer, decoder_attention_downsample_rate=decoder_attention_downsample_rate,
            classification_layer_downsample_rate=classification_layer_downsample_rate,
            transformer_feature_size=transformer_feature_size, dropout=dropout, few_type=few_type,
            class_fusion=class_fusion, classification_levels=classification_levels,
            conv_classification=conv_classification, transformer_keys_are_images=transformer_keys_are_images),
        custom_preprocess=custom_preprocess)
    lam.eval()
    if checkpoint is not None:
        state_dict = torch_dict_load(checkpoint)
        if use_sam_checkpoint:
            lam.init_pretrained_weights(state_dict)
        else:
            lam = load_state_dict(lam, state_dict, ignore_encoder_missing_keys=ignore_encoder_checkpoint)
    return lam


build_lam = _build_lam


def build_lam_vit_h(**kwargs):
    return _build_lam(build_vit_h, **kwargs)


def build_lam_vit_l(**kwargs):
    return _build_lam(build_vit_l, **kwargs)


def build_lam_vit_b(**kwargs):
    return _build_lam(build_vit_b, **kwargs)


def build_lam_vit_mae_b(**kwargs):
    return _build_lam(build_vit_b_mae, **kwargs)


def build_lam_vit_b_imagenet_i21k(**kwargs):
    return _build_lam(build_vit_b_imagenet_i21k, **kwargs)


def build_lam_no_vit(**kwargs):
    return _build_lam(build_vit=None, use_vit=False, **kwargs)


class LabelAnything(nn.Module, PyTorchModelHubMixin):
    """Hub wrapper: `LabelAnything(encoder="vit_b", **cfg).model` is a `Lam`; `from_pretrained` / `save_pretrained`
    come from huggingface_hub (config.json + model.safetensors), build_lam.py:467-508."""

    @has_config
    def __init__(self, encoder, checkpoint=None, use_sam_checkpoint=False, use_vit_sam_neck=True, use_vit=True,
                 image_embed_dim=SAM_EMBED_DIM, embed_dim=SAM_EMBED_DIM, image_size=1024, vit_patch_size=16,
                 class_attention=False, example_attention=False, example_class_attention=True,
                 class_embedding_dim=None, spatial_convs=None, encoder_attention_downsample_rate: int = 2,
                 decoder_attention_downsample_rate: int = 2, classification_layer_downsample_rate: int = 8,
                 use_support_features_in_prompt_encoder: bool = True, fusion_transformer="TwoWayTransformer",
                 few_type="Prototype", class_fusion="sum", transformer_keys_are_images=True,
                 transformer_feature_size=None, class_encoder=None, segment_example_logits=False,
                 dropout: float = 0.0, binary=False, custom_preprocess=True):
        super().__init__()
        config = self.config.copy()
        config.pop("encoder")
        # offline use: a builder callable instead of a registry name (then not serialisable to config.json)
        config["build_vit"] = encoder if callable(encoder) else ENCODERS[encoder]
        if callable(encoder):
            self.config["encoder"] = getattr(encoder, "__name__", "custom")
        self.model = build_lam(**config)

    def forward(self, *args, **kwargs):
        return self.model(*args, **kwargs)


model_registry = {  # label_anything/models/__init__.py:33-60 (entries on the LabelAnything path)
    "lam": build_lam,
    "lam_no_vit": build_lam_no_vit,
    "lam_h": build_lam_vit_h,
    "lam_l": build_lam_vit_l,
    "lam_b": build_lam_vit_b,
    "lam_mae_b": build_lam_vit_mae_b,
    "lam_b_imagenet_i21k": build_lam_vit_b_imagenet_i21k,
    **ENCODERS,
}

__all__ = ["LabelAnything", "Lam", "build_lam", "build_lam_no_vit", "build_lam_vit_b", "build_lam_vit_l",
           "build_lam_vit_h", "build_lam_vit_mae_b", "build_mask_decoder", "model_registry", "ENCODERS",
           "build_encoder", "has_config"]
