"""Input preprocessing on the GPU with the reference's names.

Mirror of label_anything/data/transforms.py (`CustomResize` :14-24, `CustomNormalize` :27-46, `PromptsProcessor`
`apply_masks / apply_coords / apply_boxes` :159-224) and of `get_preprocessing` (label_anything/data/__init__.py:33-61):
the reference runs these per image on the host CPU (PIL + torch); here the raw uint8 pixels / instance masks go to the
device once and three kernels (csrc/la_preprocess.cu) produce exactly the tensors `Lam.forward` consumes -- the same
bytes as the reference's pipeline (tests/test_preprocess_gpu.py against tests/golden/preprocess_f3.pt).

The only host-side arithmetic is Pillow's coefficient table per (source size, target size) pair
(`pil_bilinear_coeffs`, a few thousand doubles, cached): it must be computed in double precision with Pillow's exact
expression order, which is what makes the integer passes on the device bit-identical to `PIL.Image.resize`.
"""
from __future__ import annotations

import math
from functools import lru_cache
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import ops
from .utils import get_preprocess_shape

PRECISION_BITS = 32 - 8 - 2   # Pillow, src/libImaging/Resample.c
DEFAULT_MEAN, DEFAULT_STD = (0.485, 0.456, 0.406), (0.229, 0.224, 0.225)


@lru_cache(maxsize=256)
def pil_bilinear_coeffs(in_size: int, out_size: int) -> Tuple[int, np.ndarray, np.ndarray]:
    """Pillow's precompute_coeffs + normalize_coeffs_8bpc for the BILINEAR filter (triangle, support 1.0) over a whole
    axis -> (ksize, bounds int32 [out, 2] = (first source index, count), kk int32 [out, ksize]).  Vectorised over the
    output coordinate; every per-element operation is the same IEEE double operation Pillow's C code performs."""
    scale = float(np.float32(in_size) - np.float32(0.0)) / out_size       # box = (0, in_size) held as C floats
    filterscale = scale if scale >= 1.0 else 1.0
    support = 1.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    xx = np.arange(out_size, dtype=np.float64)
    center = 0.0 + (xx + 0.5) * scale
    xmin = np.trunc(center - support + 0.5).astype(np.int64)               # (int) of a double: truncation
    xmin = np.maximum(xmin, 0)
    xmax = np.minimum(np.trunc(center + support + 0.5).astype(np.int64), in_size)
    cnt = xmax - xmin
    ss = 1.0 / filterscale
    x = np.arange(ksize, dtype=np.int64)[None, :]
    a = np.abs(((x + xmin[:, None]).astype(np.float64) - center[:, None] + 0.5) * ss)
    w = np.where(a < 1.0, 1.0 - a, 0.0)
    w = np.where(x < cnt[:, None], w, 0.0)
    ww = np.zeros(out_size, dtype=np.float64)
    for j in range(ksize):                                                  # left-to-right sum, like the C loop
        ww = ww + w[:, j]
    k = np.where(ww[:, None] != 0.0, w / np.where(ww[:, None] != 0.0, ww[:, None], 1.0), w)
    kk = np.where(k < 0, np.trunc(-0.5 + k * (1 << PRECISION_BITS)), np.trunc(0.5 + k * (1 << PRECISION_BITS)))
    kk = np.where(x < cnt[:, None], kk, 0).astype(np.int32)
    bounds = np.stack([xmin, cnt], axis=1).astype(np.int32)
    return ksize, np.ascontiguousarray(bounds), np.ascontiguousarray(kk)


def _as_u8_hwc(img) -> torch.Tensor:
    """PIL image / numpy array / tensor -> contiguous uint8 [H, W, 3] tensor (host or device)."""
    if isinstance(img, torch.Tensor):
        t = img
    else:
        t = torch.from_numpy(np.ascontiguousarray(np.asarray(img)))
    if t.dtype != torch.uint8 or t.dim() != 3 or t.shape[-1] != 3:
        raise ValueError(f"expected a uint8 RGB image [H, W, 3], got {tuple(t.shape)} {t.dtype}")
    return t.contiguous()


class ImagePreprocessor:
    """`Compose([CustomResize(size), ToTensor(), CustomNormalize(size, mean, std)])` (custom_preprocess=True) or
    `Compose([Resize((size, size)), ToTensor(), Normalize(mean, std)])` on the GPU.  Call with one image or a list;
    returns fp32 [3, size, size] / [n, 3, size, size] on `device`."""

    def __init__(self, size: int = 1024, mean: Sequence[float] = DEFAULT_MEAN, std: Sequence[float] = DEFAULT_STD,
                 custom_preprocess: bool = True, device="cuda") -> None:
        self.size = size
        # the reference holds mean / std as float32 tensors (transforms.py:31-32)
        self.mean = [float(v) for v in torch.tensor(list(mean), dtype=torch.float32)]
        self.std = [float(v) for v in torch.tensor(list(std), dtype=torch.float32)]
        self.custom_preprocess = custom_preprocess
        self.device = torch.device(device)
        self._tables: dict = {}

    def _table(self, in_size: int, out_size: int):
        key = (in_size, out_size)
        hit = self._tables.get(key)
        if hit is None:
            ksize, bounds, kk = pil_bilinear_coeffs(in_size, out_size)
            hit = (ksize, torch.from_numpy(bounds).to(self.device), torch.from_numpy(kk).to(self.device))
            self._tables[key] = hit
        return hit

    def target_shape(self, h: int, w: int) -> Tuple[int, int]:
        return get_preprocess_shape(h, w, self.size) if self.custom_preprocess else (self.size, self.size)

    def one(self, img, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        src = _as_u8_hwc(img).to(self.device, non_blocking=True)
        H, W = src.shape[:2]
        nh, nw = self.target_shape(H, W)
        if out is None:
            out = torch.empty((3, self.size, self.size), dtype=torch.float32, device=self.device)
        kx = bx = kkx = ky = by = kky = tmp = None
        if nw != W:
            kx, bx, kkx = self._table(W, nw)
            tmp = torch.empty((H, nw, 3), dtype=torch.uint8, device=self.device)
        if nh != H:
            ky, by, kky = self._table(H, nh)
        ops.preprocess_image_u8(src, nh, nw, self.size, bx, kkx, kx or 0, by, kky, ky or 0, tmp, self.mean, self.std, out)
        return out

    def __call__(self, images):
        if isinstance(images, (list, tuple)):
            out = torch.empty((len(images), 3, self.size, self.size), dtype=torch.float32, device=self.device)
            for i, im in enumerate(images):
                self.one(im, out[i])
            return out
        return self.one(images)


def get_preprocessing(params: dict, device="cuda") -> ImagePreprocessor:
    """label_anything/data/__init__.py:33-61 (`mean` / `std` given as lists; the named presets of get_mean_std are the
    ImageNet defaults)."""
    common = params.get("common", {})
    size = common.get("image_size", 1024)
    custom = common.get("custom_preprocess", True)
    pre = common.get("preprocess", {})
    mean = pre.get("mean", "default")
    std = pre.get("std", "default")
    mean = DEFAULT_MEAN if mean == "default" else mean
    std = DEFAULT_STD if std == "default" else std
    return ImagePreprocessor(size, mean, std, custom, device)


class PromptsProcessor:
    """GPU counterpart of the tensor-producing methods of the reference's PromptsProcessor (transforms.py:68-224):
    `apply_masks`, `apply_coords` / `torch_apply_coords`, `apply_boxes`.  (RLE / polygon decoding and prompt sampling are
    dataset code and stay on the host.)"""

    def __init__(self, long_side_length: int = 1024, masks_side_length: int = 256, custom_preprocess: bool = True,
                 device="cuda") -> None:
        self.long_side_length = long_side_length
        self.masks_side_length = masks_side_length
        self.custom_preprocess = custom_preprocess
        self.device = torch.device(device)

    def _new_shape(self, h: int, w: int) -> Tuple[int, int]:
        return get_preprocess_shape(h, w, self.long_side_length) if self.custom_preprocess else \
            (self.long_side_length, self.long_side_length)

    def apply_masks(self, masks, out: Optional[torch.Tensor] = None, flag: Optional[torch.Tensor] = None) -> torch.Tensor:
        """masks: [n, H, W] bool / uint8 (array or tensor; n may be 0) -> fp32 {0, 1} [side, side]; `flag` (uint8 scalar
        tensor on the device, optional) is raised when the result is not empty."""
        side = self.masks_side_length
        if out is None:
            out = torch.empty((side, side), dtype=torch.float32, device=self.device)
        m = masks if isinstance(masks, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(np.asarray(masks)))
        n = int(m.shape[0]) if m.dim() == 3 else 0
        if n == 0:
            ops.rasterize_masks_u8(None, 0, 0, 0, 0, 0, self.long_side_length, side, out, flag)
            return out
        m = (m != 0).to(torch.uint8).contiguous().to(self.device, non_blocking=True)
        H, W = int(m.shape[1]), int(m.shape[2])
        nh, nw = self._new_shape(H, W) if self.custom_preprocess else (0, 0)
        ops.rasterize_masks_u8(m, n, H, W, nh, nw, self.long_side_length, side, out, flag)
        return out

    def apply_coords(self, coords, original_size: Tuple[int, int], out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """coords [..., 2] (x, y) float64 -> fp32 [..., 2] in the resized frame."""
        c = coords if isinstance(coords, torch.Tensor) else torch.from_numpy(np.asarray(coords, dtype=np.float64))
        c = c.to(torch.float64).contiguous().to(self.device, non_blocking=True)
        old_h, old_w = original_size
        new_h, new_w = self._new_shape(old_h, old_w)
        if out is None:
            out = torch.empty(c.shape, dtype=torch.float32, device=self.device)
        ops.scale_coords_f64(c, new_w / old_w, new_h / old_h, out)
        return out

    torch_apply_coords = apply_coords

    def apply_boxes(self, boxes, original_size: Tuple[int, int]) -> torch.Tensor:
        """boxes [n, 4] (x1, y1, x2, y2) -> fp32 [n, 4]."""
        b = boxes if isinstance(boxes, torch.Tensor) else torch.from_numpy(np.asarray(boxes, dtype=np.float64))
        return self.apply_coords(b.reshape(-1, 2, 2), original_size).reshape(-1, 4)


def preprocess_images(images: List, size: int = 1024, mean=DEFAULT_MEAN, std=DEFAULT_STD, custom_preprocess: bool = True,
                      device="cuda") -> Tuple[torch.Tensor, torch.Tensor]:
    """A list of RGB images -> (fp32 [n, 3, size, size] on the device, dims int64 [n, 2] = original (H, W)): the
    `images` / `dims` entries of the reference's batch (demo/preprocess.py:123-211)."""
    pre = ImagePreprocessor(size, mean, std, custom_preprocess, device)
    srcs = [_as_u8_hwc(im) for im in images]
    dims = torch.tensor([[int(s.shape[0]), int(s.shape[1])] for s in srcs], dtype=torch.int64)
    return pre(srcs), dims
