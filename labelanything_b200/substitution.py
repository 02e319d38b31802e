"""Iterative prompting on the GPU: corrective points from prediction errors.

Mirror of `generate_points_from_errors` and `Substitutor.generate_new_points`
(label_anything/experiment/substitution.py:17-96, 161-197).  The reference builds two one-hot [B, C, H, W] int64
tensors, their difference and its full `torch.nonzero` list to draw ONE pixel per (episode, class); here two kernels
(la_error_points, csrc/la_metrics.cu) read the logits twice and materialise nothing.

Randomness: the reference draws `torch.randint(0, count, (num_points,))` per (b, c) group inside the function; here
the draw is an explicit argument (`rand`, int64 [B, C, num_points]; the kernel uses rand mod count), by default filled
with `torch.randint` from an optional generator -- same distribution, reproducible, and the unit tests can pin it.

Two reference quirks, stated rather than copied: (1) the reference orders its rows with `argsort(b * B + c)` (sic: B,
not C) using an unstable sort, which scrambles (episode, class) rows whenever C > B; the intended (b, c) order is
produced here, identical to the reference whenever its keys are unique (B = 1, or C <= B).  (2) with num_points > 1
the reference raises for any class without errors (one padding row per class, n expected); here such a class gets n
padding points.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from . import _native, ops
from .utils import BatchKeys, get_preprocess_shape


def generate_points_from_errors(prediction: torch.Tensor, ground_truth: torch.Tensor, num_points: int,
                                ignore_index: int = -100, rand: Optional[torch.Tensor] = None,
                                generator: Optional[torch.Generator] = None,
                                scale_xy: Optional[Tuple[torch.Tensor, torch.Tensor]] = None):
    """prediction: logits fp32 [B, C, H, W]; ground_truth int64 [B, H, W] -> (points fp32 [B, C, n, 2] as (x, y),
    labels fp32 [B, C, n] in {+1 false negative, -1 false positive, 0}).  scale_xy = (sx [B], sy [B]) fp32 multiplies
    the coordinates (Substitutor.generate_new_points applies torch_apply_coords right after)."""
    ops._require_cuda(prediction, ground_truth, rand)
    B, C, H, W = prediction.shape
    dev = prediction.device
    if rand is None:
        rand = torch.randint(0, 2 ** 31 - 1, (B, C, num_points), device=dev, generator=generator, dtype=torch.int64)
    assert rand.dtype == torch.int64 and tuple(rand.shape) == (B, C, num_points)
    if scale_xy is None:
        sx = sy = torch.ones(B, dtype=torch.float32, device=dev)
    else:
        sx, sy = (t.to(device=dev, dtype=torch.float32).contiguous() for t in scale_xy)
    return ops.error_points(prediction.float().contiguous(), ground_truth.contiguous(), rand.contiguous(), sx, sy,
                            ignore_index)


class Substitutor:
    """The prompt-refinement half of the reference's Substitutor (substitution.py:98-197): `reset(batch)` then
    `generate_new_points(prediction, ground_truth)` appends one corrective point per class to the query-side prompts of
    `batch` (prompt_points [B, M, C, P, 2] -> P + n, flag_points likewise).  Query/example rotation (`__next__`) is
    dataset bookkeeping and stays with the caller."""

    def __init__(self, threshold: Optional[float] = None, num_points: int = 1, substitute: bool = True,
                 long_side_length: int = 1024, custom_preprocess: bool = True) -> None:
        self.num_points = num_points
        self.substitute = substitute and threshold is None
        self.long_side_length = long_side_length
        self.custom_preprocess = custom_preprocess
        self.batch = None
        self.ground_truths = None

    def reset(self, batch) -> None:
        self.batch, self.ground_truths = batch

    def _scales(self, dims: torch.Tensor):
        """new_w / old_w, new_h / old_h of every episode's query image (substitution.py:166-171, transforms.py:176-184)."""
        sx, sy = [], []
        for old_h, old_w in dims[:, 0].tolist():
            new_h, new_w = get_preprocess_shape(old_h, old_w, self.long_side_length) if self.custom_preprocess else \
                (self.long_side_length, self.long_side_length)
            sx.append(new_w / old_w)
            sy.append(new_h / old_h)
        return torch.tensor(sx, dtype=torch.float32), torch.tensor(sy, dtype=torch.float32)

    def generate_new_points(self, prediction: torch.Tensor, ground_truth: torch.Tensor,
                            rand: Optional[torch.Tensor] = None, generator: Optional[torch.Generator] = None) -> None:
        if not (self.substitute and self.num_points > 0):
            return
        b = self.batch
        pts, labels = generate_points_from_errors(prediction, ground_truth, self.num_points, rand=rand,
                                                  generator=generator, scale_xy=self._scales(b[BatchKeys.DIMS]))
        old_p, old_f = b[BatchKeys.PROMPT_POINTS], b[BatchKeys.FLAG_POINTS]
        B, M, C, P = old_f.shape
        n = self.num_points
        new_p = torch.zeros((B, M, C, P + n, 2), dtype=old_p.dtype, device=old_p.device)
        new_f = torch.zeros((B, M, C, P + n), dtype=old_f.dtype, device=old_f.device)
        new_p[:, :, :, :P] = old_p
        new_f[:, :, :, :P] = old_f
        new_p[:, 0, :, P:] = pts.to(old_p.dtype)          # the query slot; the other examples get zero padding
        new_f[:, 0, :, P:] = labels.to(old_f.dtype)
        b[BatchKeys.PROMPT_POINTS], b[BatchKeys.FLAG_POINTS] = new_p, new_f
