"""CPU restatement (numpy, float32) of the reference's default loss, SURVEY.md §8 row f1 (first piece).

TEST INFRASTRUCTURE: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this file.

Follows:
  * get_weight_matrix_from_labels        label_anything/loss/utils.py:17-42
  * FocalLoss.__call__                   label_anything/loss/focal.py:17-25
  * LabelAnythingLoss.logits_loss        label_anything/loss/__init__.py:67-92 (component weight applied to the logged
                                         value once and to the summed value twice)
  * the gradient is the analytic derivative of the same expression (what torch autograd computes for it).
Pinned by tests/golden/loss_f1.pt: values, gradients, weight maps and class weights of the UNMODIFIED reference
(oracle/make_golden.py loss).
"""
from __future__ import annotations

import numpy as np

F32 = np.float32


def get_weight_matrix_from_labels(labels: np.ndarray, num_classes: int, ignore_index: int = -100):
    there_is_ignore = bool((labels == ignore_index).any())
    if there_is_ignore:
        wl = labels + 1
        wl[wl == ignore_index + 1] = 0
        n_w = num_classes + 1
    else:
        wl, n_w = labels, num_classes
    weights = np.ones(n_w, dtype=F32)
    classes, counts = np.unique(wl, return_counts=True)
    weights[classes] = F32(1) / np.log(F32(1.1) + counts.astype(F32) / F32(counts.sum()), dtype=F32)
    if there_is_ignore:
        weights[0] = 0
        class_weights = weights[1:]
    else:
        class_weights = weights
    return weights[wl], class_weights


def focal_terms(x: np.ndarray, target: np.ndarray, gamma: float, weight_matrix=None, ignore_index: int = -100):
    """per-pixel focal loss [B, *spatial] (float32) and the softmax / target bookkeeping for the gradient."""
    x = x.astype(F32)
    m = x.max(axis=1, keepdims=True)
    e = np.exp(x - m, dtype=F32)
    lse = m + np.log(e.sum(axis=1, keepdims=True), dtype=F32)
    logp = x - lse
    valid = target != ignore_index
    t = np.where(valid, target, 0)
    ce = -np.take_along_axis(logp, t[:, None], axis=1)[:, 0]
    ce = np.where(valid, ce, F32(0)).astype(F32)
    pt = np.exp(-ce, dtype=F32)
    w = np.ones_like(ce) if weight_matrix is None else weight_matrix.astype(F32)
    focal = np.power(F32(1) - pt, F32(gamma), dtype=F32) * w * ce
    return focal.astype(F32), (logp, valid, t, ce, pt, w)


def focal_loss(x, target, gamma=2.0, weight_matrix=None, reduction="mean"):
    focal, _ = focal_terms(x, target, gamma, weight_matrix)
    return F32(focal.mean(dtype=np.float64)) if reduction == "mean" else F32(focal.sum(dtype=np.float64))


def focal_loss_grad(x, target, gamma=2.0, weight_matrix=None, reduction="mean", upstream=1.0):
    _, (logp, valid, t, ce, pt, w) = focal_terms(x, target, gamma, weight_matrix)
    om = (F32(1) - pt).astype(np.float64)
    g = float(gamma)
    dpow = np.where(om > 0, g * np.power(np.where(om > 0, om, 1.0), g - 1.0), 0.0 if g > 1 else g)
    coef = w * (dpow * pt * (-ce) - np.power(om, g)) * valid            # d focal / d log-softmax path
    p = np.exp(logp.astype(np.float64))
    onehot = np.zeros_like(p)
    np.put_along_axis(onehot, t[:, None], 1.0, axis=1)
    grad = coef[:, None] * (onehot - p)
    n = target.size if reduction == "mean" else 1
    return (grad * (upstream / n)).astype(F32)


def label_anything_loss(x, target, gamma=2.0, component_weight=1.0, class_weighting=True):
    wm = get_weight_matrix_from_labels(target.copy(), x.shape[1])[0] if class_weighting else None
    res = F32(component_weight) * focal_loss(x, target, gamma, wm)
    return {"value": F32(component_weight) * res, "components": {"focal": float(res)}}
