"""Enums and state-dict helpers the hot path shares with its callers.

Mirror of the pieces of label_anything/utils/utils.py (ResultDict :356-364, torch_dict_load :91-100,
load_state_dict :119-142) and label_anything/data/utils.py (Label :25-28, BatchKeys :43-58,
get_preprocess_shape :441-449) that label_anything/models/lam.py imports.  String values are identical so the
reference's experiment / demo code can index our result dicts with its own enums.
"""
from __future__ import annotations

from enum import Enum, IntEnum

import torch


class _StrEnum(str, Enum):
    def __str__(self) -> str:  # behaves like the reference's StrEnum in f-strings and dict lookups
        return str(self.value)


class ResultDict(_StrEnum):
    CLASS_EMBS = "class_embeddings"
    MASK_EMBEDDINGS = "mask_embeddings"
    LOGITS = "logits"
    EXAMPLES_CLASS_EMBS = "class_examples_embeddings"
    EXAMPLES_CLASS_SRC = "class_examples_src"
    LOSS = "loss"
    LAST_HIDDEN_STATE = "last_hidden_state"
    LAST_BLOCK_STATE = "last_block_state"


class Label(IntEnum):
    POSITIVE = 1
    NULL = 0
    NEGATIVE = -1


class BatchKeys(_StrEnum):
    IMAGES = "images"
    EMBEDDINGS = "embeddings"
    PROMPT_MASKS = "prompt_masks"
    FLAG_MASKS = "flag_masks"
    PROMPT_POINTS = "prompt_points"
    FLAG_POINTS = "flag_points"
    PROMPT_BBOXES = "prompt_bboxes"
    FLAG_BBOXES = "flag_bboxes"
    FLAG_EXAMPLES = "flag_examples"
    DIMS = "dims"
    CLASSES = "classes"
    INTENDED_CLASSES = "intended_classes"
    IMAGE_IDS = "image_ids"
    GROUND_TRUTHS = "ground_truths"
    CLIP_EMBEDDINGS = "clip_embeddings"


def get_preprocess_shape(oldh: int, oldw: int, long_side_length: int):
    """Size of the un-padded model input for an (oldh, oldw) original.  data/utils.py:441-449"""
    scale = long_side_length * 1.0 / max(oldh, oldw)
    return int(oldh * scale + 0.5), int(oldw * scale + 0.5)


def torch_dict_load(file_path: str):
    """utils/utils.py:91-100"""
    if file_path.endswith((".pth", ".pt", ".bin")):
        return torch.load(file_path, map_location="cpu")
    if file_path.endswith(".safetensors"):
        from safetensors import safe_open

        with safe_open(file_path, framework="pt") as f:
            return {k: f.get_tensor(k) for k in f.keys()}
    raise ValueError("File extension not supported")


def torch_dict_save(data, file_path: str) -> None:
    """utils/utils.py:102-108"""
    if file_path.endswith((".pth", ".pt", ".bin")):
        torch.save(data, file_path)
    elif file_path.endswith(".safetensors"):
        from safetensors.torch import save_file

        save_file(data, file_path)
    else:
        raise ValueError("File extension not supported")


def _keys_check(res) -> None:
    missing = [k for k in res.missing_keys if "image_encoder" not in k]
    if missing:
        raise RuntimeError(f"Missing keys: {missing}")
    if res.unexpected_keys:
        raise RuntimeError(f"Unexpected keys: {res.unexpected_keys}")


def load_state_dict(model, state_dict, strict: bool = True, ignore_encoder_missing_keys: bool = False):
    """Load, retrying with the `model.` and then the `module.` prefix removed.  utils/utils.py:119-142"""
    if ignore_encoder_missing_keys:
        strict = False
    last = None
    for prefix in (None, "model.", "module."):
        if prefix is not None:
            state_dict = {k.replace(prefix, ""): v for k, v in state_dict.items()}
        try:
            res = model.load_state_dict(state_dict, strict=strict)
            if ignore_encoder_missing_keys:
                _keys_check(res)
            return model
        except RuntimeError as e:  # noqa: PERF203 - mirrors the reference's retry ladder
            last = e
    raise last
