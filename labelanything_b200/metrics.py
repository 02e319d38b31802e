"""The step right after the hot path in the reference's train / validation loops:

    preds = logits.argmax(dim=1)
    glob_preds, glob_gt = to_global_multiclass(classes, categories, preds, gt)
    metrics.update(glob_preds, glob_gt)            # StrictMeanIoU / MeanIoU, ignore_index=-100

(label_anything/experiment/run.py:520-541,696-704; data/utils.py:567-590; utils/metrics.py:28-42).  Same names and
argument meaning as the reference; the arithmetic runs in one CUDA kernel (`la_label_confusion`): logits are read
once, the label maps are written once, the confusion matrix is accumulated with integer atomics.

The IoU reduction over the tiny [G, G] matrix follows torchmetrics 1.7.1 (`_jaccard_index_reduce`, average="macro";
third party, pinned in the reference's uv.lock:2672-2673, NOT installed in this image: restated from its published
algorithm) and runs on the host in fp32 like torchmetrics does.
"""
from __future__ import annotations

from typing import Optional, Sequence

import torch

from . import ops

__all__ = ["chain_label_map", "to_global_multiclass", "MeanIoU", "StrictMeanIoU"]


def chain_label_map(classes: Sequence[Sequence[Sequence[int]]], categories: dict, compact: bool = True,
                    map_len: Optional[int] = None) -> torch.Tensor:
    """int64 [B, map_len] table T with T[b][v] = the value to_global_multiclass leaves in a pixel of episode b that
    held v.  The reference substitutes `t == j + 1 -> value_j` for j = 0, 1, ... IN SEQUENCE on the same tensor
    (data/utils.py:583-589), so an already substituted pixel is substituted again when its new value equals a later
    j + 1; composing the steps per start value reproduces that exactly.  Values outside [0, map_len) are never
    touched by the reference (map_len > number of episode classes), nor by the kernel."""
    cats_map = {k: i + 1 for i, k in enumerate(categories.keys())}
    per_item = [sorted(set(sum((list(c) for c in classes[i]), []))) for i in range(len(classes))]
    need = 1 + max((len(c) for c in per_item), default=0)
    map_len = need if map_len is None else map_len
    assert map_len >= need, "map_len must cover every episode-local label"
    table = torch.arange(map_len, dtype=torch.int64).repeat(len(classes), 1)
    for i, longest in enumerate(per_item):
        values = [cats_map[v] if compact else v for v in longest]
        for start in range(map_len):
            x = start
            for j, value in enumerate(values):
                if x == j + 1:
                    x = value
            table[i, start] = x
    return table


def to_global_multiclass(classes, categories: dict, *tensors: torch.Tensor, compact: bool = True) -> list[torch.Tensor]:
    """Drop-in for label_anything/data/utils.py:567-590 on CUDA int64 label tensors [B, ...]."""
    if not tensors:
        return []
    table = chain_label_map(classes, categories, compact).to(tensors[0].device)
    out = []
    for t in tensors:
        if t.dtype != torch.int64:
            raise TypeError("to_global_multiclass: label tensors must be int64 (torch.long), like argmax outputs")
        mapped, _ = ops.label_confusion(None, t.contiguous(), None, table, want_gt=False)
        out.append(mapped)
    return out


class MeanIoU:
    """Multiclass Jaccard index, macro average (the reference's `MeanIoU(MulticlassJaccardIndex)`,
    utils/metrics.py:39-40; constructor as used in experiment/run.py:451-457,657-668)."""

    def __init__(self, num_classes: int, ignore_index: Optional[int] = None, average: str = "macro",
                 device: torch.device | str = "cuda", validate_args: bool = True, **_ignored):
        if average != "macro":
            raise NotImplementedError("only average='macro' (the reference's setting) is implemented")
        self.num_classes = int(num_classes)
        self.ignore_index = ignore_index
        self.validate_args = validate_args
        self.device = torch.device(device)
        self.confmat = torch.zeros(self.num_classes, self.num_classes, dtype=torch.int64, device=self.device)
        self._invalid = torch.zeros(1, dtype=torch.int64, device=self.device)

    # ---- state -------------------------------------------------------------------------------------------------
    def reset(self) -> None:
        self.confmat.zero_()
        self._invalid.zero_()

    def to(self, device) -> "MeanIoU":
        self.device = torch.device(device)
        self.confmat = self.confmat.to(self.device)
        self._invalid = self._invalid.to(self.device)
        return self

    def sync(self, group=None) -> None:
        """Sum the state over the ranks (torchmetrics: dist_reduce_fx="sum" on `confmat`)."""
        import torch.distributed as dist

        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            dist.all_reduce(self.confmat, op=dist.ReduceOp.SUM, group=group)
            dist.all_reduce(self._invalid, op=dist.ReduceOp.SUM, group=group)

    # ---- updates -----------------------------------------------------------------------------------------------
    def _ignore(self) -> int:
        # no target can equal this sentinel when ignore_index is None
        return self.ignore_index if self.ignore_index is not None else -(1 << 62)

    def update(self, preds: torch.Tensor, target: torch.Tensor) -> None:
        """preds / target: int64 label maps [B, ...] (already global), like torchmetrics' update."""
        if preds.is_floating_point():      # torchmetrics takes argmax over dim 1 of float predictions
            ops.label_confusion(preds.float().contiguous(), None, target.contiguous(), None, self.confmat,
                                self._invalid, self._ignore(), want_preds=False, want_gt=False)
        else:
            ops.label_confusion(None, preds.contiguous(), target.contiguous(), None, self.confmat, self._invalid,
                                self._ignore(), want_preds=False, want_gt=False)

    def update_from_logits(self, logits: torch.Tensor, gt: torch.Tensor, classes, categories: dict,
                           compact: bool = True, want_labels: bool = False):
        """Fused `argmax -> to_global_multiclass -> update` (run.py:520-541): one pass over the logits.
        Returns (glob_preds, glob_gt) when want_labels (the reference logs them), else (None, None)."""
        table = chain_label_map(classes, categories, compact, map_len=max(logits.shape[1], 1 + max(
            (len(set(sum((list(c) for c in cl), []))) for cl in classes), default=0))).to(logits.device)
        return ops.label_confusion(logits.contiguous(), None, gt.contiguous(), table, self.confmat, self._invalid,
                                   self._ignore(), want_preds=want_labels, want_gt=want_labels)

    # ---- value -------------------------------------------------------------------------------------------------
    def _checked_confmat(self) -> torch.Tensor:
        conf = self.confmat.cpu()
        if self.validate_args and int(self._invalid.cpu()) != 0:
            raise RuntimeError(f"Detected {int(self._invalid.cpu())} label(s) outside [0, {self.num_classes}) in "
                               "`preds` / `target` (torchmetrics raises for them under validate_args)")
        return conf

    @staticmethod
    def _macro_jaccard(confmat: torch.Tensor, ignore_index: Optional[int]) -> torch.Tensor:
        conf = confmat.float()
        num = torch.diag(conf)
        denom = conf.sum(0) + conf.sum(1) - num
        jaccard = torch.where(denom != 0, num / torch.where(denom != 0, denom, torch.ones_like(denom)),
                              torch.zeros_like(num))                  # _safe_divide, zero_division = 0
        weights = torch.ones_like(jaccard)
        if ignore_index is not None and 0 <= ignore_index < conf.shape[0]:
            weights[ignore_index] = 0.0
        weights[conf.sum(1) + conf.sum(0) == 0] = 0.0
        return ((weights * jaccard) / weights.sum()).sum()

    def compute(self) -> torch.Tensor:
        return self._macro_jaccard(self._checked_confmat(), self.ignore_index)

    def __call__(self, preds: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
        self.update(preds, target)
        return self.compute()


class StrictMeanIoU(MeanIoU):
    """utils/metrics.py:28-36: the macro mean with the background IoU taken out."""

    def compute(self) -> torch.Tensor:
        conf_i = self._checked_confmat()
        metric = self._macro_jaccard(conf_i, self.ignore_index)
        conf = conf_i   # the reference divides the integer state: true division -> fp32
        bg_iou = conf[0, 0] / (conf[0, 0] + conf[0, 1:].sum() + conf[1:, 0].sum())
        return (metric * self.num_classes - bg_iou) / (self.num_classes - 1)
