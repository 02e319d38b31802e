"""CPU oracle of the input preprocessing -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

numpy restatement of what the reference's data pipeline does to one image and its prompts before `Lam.forward`:

  * `CustomResize` (label_anything/data/transforms.py:14-24): `torchvision.transforms.functional.resize(PIL image,
    get_preprocess_shape(h, w, S))` = `PIL.Image.resize(..., BILINEAR)`.  The arithmetic is Pillow's (third party,
    Pillow 12.2.0 in this image; src/libImaging/Resample.c): an antialiased separable triangle filter with 22-bit
    fixed-point coefficients, horizontal pass rounded to 8 bits, then the vertical pass.  Restated here from the
    published algorithm and pinned against Pillow itself (tests/test_preprocess_cpu.py, fixture tests/golden/preprocess_f3.pt).
  * `ToTensor` + `CustomNormalize` (transforms.py:27-46): x / 255, (x - mean) / std in fp32, zero padding to S x S.
  * the non-custom pipeline `Resize((S, S))` + `ToTensor` + `Normalize` (label_anything/data/__init__.py:33-61).
  * `PromptsProcessor.apply_masks / apply_coords / apply_boxes` (transforms.py:159-224): OR of the instance masks, nearest
    resize to the preprocess shape, zero pad to S, nearest resize to 256 x 256 (torchvision on uint8 tensors ->
    F.interpolate(mode="nearest"): src = min(floor(dst * float32(in / out)), in - 1)); coordinates scaled in float64.

Only tests/ may import this module.
"""
from __future__ import annotations

import math
from typing import Sequence, Tuple

import numpy as np

PRECISION_BITS = 32 - 8 - 2


def get_preprocess_shape(oldh: int, oldw: int, long_side_length: int) -> Tuple[int, int]:
    """label_anything/data/utils.py:441-449"""
    scale = long_side_length * 1.0 / max(oldh, oldw)
    return int(oldh * scale + 0.5), int(oldw * scale + 0.5)


def pil_bilinear_coeffs(in_size: int, out_size: int):
    """Resample.c precompute_coeffs + normalize_coeffs_8bpc for the BILINEAR (triangle, support 1) filter over the whole
    axis: -> (ksize, bounds int32 [out, 2] = (xmin, count), kk int32 [out, ksize])."""
    scale = float(np.float32(in_size) - np.float32(0.0)) / out_size        # box = (0, in_size) as C floats
    filterscale = scale if scale >= 1.0 else 1.0
    support = 1.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), dtype=np.int32)
    kk = np.zeros((out_size, ksize), dtype=np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = 0.0 + (xx + 0.5) * scale
        xmin = int(center - support + 0.5)
        if xmin < 0:
            xmin = 0
        xmax = int(center + support + 0.5)
        if xmax > in_size:
            xmax = in_size
        xmax -= xmin
        ws = []
        ww = 0.0
        for x in range(xmax):
            a = (x + xmin - center + 0.5) * ss
            if a < 0.0:
                a = -a
            w = 1.0 - a if a < 1.0 else 0.0
            ws.append(w)
            ww += w
        for x in range(xmax):
            k = ws[x] / ww if ww != 0.0 else ws[x]
            kk[xx, x] = int(-0.5 + k * (1 << PRECISION_BITS)) if k < 0 else int(0.5 + k * (1 << PRECISION_BITS))
        bounds[xx] = (xmin, xmax)
    return ksize, bounds, kk


def _pass_8bpc(img: np.ndarray, out_size: int, axis: int) -> np.ndarray:
    """One separable pass over `axis` of a uint8 [H, W, C] image (ImagingResampleHorizontal/Vertical_8bpc)."""
    in_size = img.shape[axis]
    if out_size == in_size:
        return img
    _, bounds, kk = pil_bilinear_coeffs(in_size, out_size)
    src = np.moveaxis(img, axis, 0).astype(np.int64)
    out = np.empty((out_size,) + src.shape[1:], dtype=np.uint8)
    for xx in range(out_size):
        xmin, cnt = bounds[xx]
        acc = np.full(src.shape[1:], 1 << (PRECISION_BITS - 1), dtype=np.int64)
        for x in range(cnt):
            acc += src[xmin + x] * int(kk[xx, x])
        out[xx] = np.clip(acc >> PRECISION_BITS, 0, 255).astype(np.uint8)
    return np.moveaxis(out, 0, axis)


def pil_resize_bilinear_u8(img: np.ndarray, out_h: int, out_w: int) -> np.ndarray:
    """uint8 [H, W, C] -> uint8 [out_h, out_w, C], bit-identical to PIL.Image.resize((out_w, out_h), BILINEAR)."""
    return _pass_8bpc(_pass_8bpc(img, out_w, 1), out_h, 0)


def preprocess_image(img: np.ndarray, size: int, mean: Sequence[float], std: Sequence[float],
                     custom_preprocess: bool = True) -> np.ndarray:
    """uint8 [H, W, 3] -> fp32 [3, size, size] (transforms.py:14-46 / data/__init__.py:33-61)."""
    h, w = img.shape[:2]
    nh, nw = get_preprocess_shape(h, w, size) if custom_preprocess else (size, size)
    r = pil_resize_bilinear_u8(img, nh, nw)
    x = r.transpose(2, 0, 1).astype(np.float32) / np.float32(255.0)                       # ToTensor
    x = (x - np.asarray(mean, np.float32).reshape(3, 1, 1)) / np.asarray(std, np.float32).reshape(3, 1, 1)
    out = np.zeros((3, size, size), dtype=np.float32)
    out[:, :nh, :nw] = x
    return out


def _nearest_index(dst: np.ndarray, in_size: int, out_size: int) -> np.ndarray:
    """ATen nearest (legacy 'nearest' mode): identity when sizes match, else min(floor(dst * float32(in/out)), in-1)."""
    if in_size == out_size:
        return dst.astype(np.int64)
    scale = np.float32(in_size) / np.float32(out_size)
    return np.minimum(np.floor(dst.astype(np.float32) * scale).astype(np.int64), in_size - 1)


def rasterize_masks(masks: np.ndarray, long_side: int = 1024, out_side: int = 256,
                    custom_preprocess: bool = True) -> np.ndarray:
    """PromptsProcessor.apply_masks (transforms.py:196-224): masks bool/uint8 [n, H, W] -> uint8 [out_side, out_side]."""
    if len(masks) == 0:
        return np.zeros((out_side, out_side), dtype=np.uint8)
    m = np.logical_or.reduce(np.asarray(masks) != 0).astype(np.uint8)
    H, W = m.shape
    if custom_preprocess:
        nh, nw = get_preprocess_shape(H, W, long_side)
        ys = _nearest_index(np.arange(nh), H, nh)
        xs = _nearest_index(np.arange(nw), W, nw)
        small = m[ys][:, xs]
        m = np.zeros((long_side, long_side), dtype=np.uint8)
        m[:nh, :nw] = small
    ys = _nearest_index(np.arange(out_side), m.shape[0], out_side)
    xs = _nearest_index(np.arange(out_side), m.shape[1], out_side)
    return m[ys][:, xs]


def apply_coords(coords: np.ndarray, original_size: Tuple[int, int], long_side: int = 1024,
                 custom_preprocess: bool = True) -> np.ndarray:
    """PromptsProcessor.apply_coords (transforms.py:159-174): float64 scaling of (x, y) pairs."""
    old_h, old_w = original_size
    new_h, new_w = get_preprocess_shape(old_h, old_w, long_side) if custom_preprocess else (long_side, long_side)
    out = np.array(coords, dtype=np.float64, copy=True)
    out[..., 0] = out[..., 0] * (new_w / old_w)
    out[..., 1] = out[..., 1] * (new_h / old_h)
    return out


def apply_boxes(boxes: np.ndarray, original_size: Tuple[int, int], long_side: int = 1024,
                custom_preprocess: bool = True) -> np.ndarray:
    """transforms.py:186-194"""
    return apply_coords(np.asarray(boxes).reshape(-1, 2, 2), original_size, long_side, custom_preprocess).reshape(-1, 4)
