"""The reference's default training / validation loss (first piece): focal loss with
label-frequency class weighting, `parameters/trainval/coco/mael.yaml:24-28`.

Same names and argument meaning as label_anything/loss/__init__.py:30-116, loss/focal.py:8-25 and
loss/utils.py:17-42.  The arithmetic runs in `la_focal_loss` / `la_label_class_weights` (csrc/la_loss.cu): the value
in one pass over the logits, the gradient w.r.t. the logits in one more pass (nothing but the logits and the labels is
kept for the backward pass); the other reference components (dice, rmi, prompt / embedding contrastive losses) are not
built and raise.
"""
from __future__ import annotations

import torch
from torch import nn

from . import ops
from .utils import _StrEnum

__all__ = ["FocalLoss", "LabelAnythingLoss", "LossDict", "get_weight_matrix_from_labels"]


class LossDict(_StrEnum):          # label_anything/utils/utils.py:367-369
    VALUE = "value"
    COMPONENTS = "components"


class _ClassWeighted:
    """Stands for `class_weights[target]` without materialising the [B, H, W] map (the kernel gathers it)."""


FROM_CLASS_WEIGHTS = _ClassWeighted()


def get_weight_matrix_from_labels(labels: torch.Tensor, num_classes: int, ignore_index: int = -100):
    """loss/utils.py:17-42 -> (wtarget fp32 like labels, class_weights fp32 [num_classes])."""
    labels = labels.contiguous()
    class_w, hist = ops.label_class_weights(labels, num_classes, ignore_index)
    _, _, wt = ops.focal_loss(None, labels, class_w, 0.0, ignore_index, want_loss=False, want_wtarget=True)
    return wt, class_w


class _FocalFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, target, class_w, gamma, ignore_index, mean):
        ctx.in_dtype = x.dtype
        xc = x.float().contiguous()       # autocast / bf16 logits: the kernel computes in fp32 like the reference's CE
        loss, _, _ = ops.focal_loss(xc, target, class_w, gamma, ignore_index, mean)
        ctx.save_for_backward(xc, target, class_w if class_w is not None else torch.empty(0, device=x.device))
        ctx.cfg = (gamma, ignore_index, mean, class_w is not None)
        return loss

    @staticmethod
    def backward(ctx, grad_out):
        xc, target, class_w = ctx.saved_tensors
        gamma, ignore_index, mean, has_w = ctx.cfg
        _, grad, _ = ops.focal_loss(xc, target, class_w if has_w else None, gamma, ignore_index, mean, want_loss=False,
                                    want_grad=True, grad_scale=grad_out.float().contiguous())
        return grad.to(ctx.in_dtype), None, None, None, None, None


class FocalLoss(nn.Module):
    """loss/focal.py:8-25.  `weight_matrix` must be the class-weight map of `class_weights` (what
    LabelAnythingLoss.logits_loss passes); arbitrary per-pixel maps are not supported by the fused kernel."""

    def __init__(self, gamma: float = 2.0, reduction: str = "mean", ignore_index: int = -100, **kwargs):
        super().__init__()
        if reduction not in ("mean", "sum"):
            raise NotImplementedError(f"Invalid reduction mode for the fused loss: {reduction}")
        self.gamma = float(gamma)
        self.reduction = reduction
        self.ignore_index = ignore_index

    def forward(self, x, target, weight_matrix=None, class_weights=None, **kwargs):
        if weight_matrix is not None and class_weights is None:
            raise NotImplementedError("FocalLoss: pass class_weights (the fused kernel gathers class_weights[target])")
        cw = class_weights if weight_matrix is not None else None
        return _FocalFn.apply(x, target.contiguous(), cw, self.gamma, self.ignore_index, self.reduction == "mean")


LOGITS_LOSSES = {"focal": FocalLoss}


class LabelAnythingLoss(nn.Module):
    """loss/__init__.py:30-116 with the focal component.  Like the reference, the component weight is applied twice
    to the summed value (`loss_res = w * loss(...)`, then `w * loss_value`, lines 76-88) and once to the logged one."""

    def __init__(self, components, class_weighting=None):
        super().__init__()
        components = {k: dict(v) for k, v in components.items()}
        self.weights = {k: v.pop("weight") for k, v in components.items()}
        unknown = set(components) - set(LOGITS_LOSSES)
        if unknown:
            raise NotImplementedError(f"loss components not built natively: {sorted(unknown)}")
        self.components = nn.ModuleDict([[k, LOGITS_LOSSES[k](**v)] for k, v in components.items()])
        self.class_weighting = class_weighting

    def logits_loss(self, logits, target):
        weight_matrix, class_weights = None, None
        if self.class_weighting:
            class_weights, _ = ops.label_class_weights(target.contiguous(), logits.shape[1])
            weight_matrix = FROM_CLASS_WEIGHTS
        loss_values, loss_dict = [], {}
        loss_value = 0
        for k, loss in self.components.items():
            loss_res = self.weights[k] * loss(logits, target, weight_matrix=weight_matrix, class_weights=class_weights)
            loss_dict[k] = loss_res.item()
            loss_values.append(self.weights[k] * loss_res)
            loss_value = sum(loss_values)
        return {LossDict.VALUE: loss_value, LossDict.COMPONENTS: loss_dict}

    def forward(self, result, target):
        logits = result if isinstance(result, torch.Tensor) else result["logits"]
        return self.logits_loss(logits, target)
