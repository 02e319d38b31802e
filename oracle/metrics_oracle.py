"""CPU restatement (numpy) of the reference's post-logits step, SURVEY.md §8 row f4.

TEST INFRASTRUCTURE: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this file; it is
the checker, never the product path.

Follows, line by line:
  * `outputs.argmax(dim=1)`                         label_anything/experiment/run.py:521,697
  * `to_global_multiclass`                           label_anything/data/utils.py:567-590 (sequential torch.where)
  * confusion-matrix update of `MulticlassJaccardIndex` (torchmetrics 1.7.1, uv.lock:2672-2673; NOT vendored in
    /root/reference and not installed here): targets equal to `ignore_index` are removed, then
    confmat = bincount(target * C + preds, minlength=C*C).reshape(C, C)
    (torchmetrics/functional/classification/confusion_matrix.py `_multiclass_confusion_matrix_format/_update`)
  * `_jaccard_index_reduce(confmat, average="macro", ignore_index)` (torchmetrics/functional/classification/jaccard.py)
  * `StrictMeanIoU.compute`                          label_anything/utils/metrics.py:28-36

Pinning: `to_global_multiclass` + argmax are pinned by tests/golden/metrics_f4.pt, generated from the UNMODIFIED
reference function (oracle/make_golden.py metrics).  The torchmetrics part has no golden vectors in the reference and
the package is absent: PARITY UNPINNED for the jaccard reduction (restated from the published algorithm).
"""
from __future__ import annotations

import numpy as np


def argmax_dim1(logits: np.ndarray) -> np.ndarray:
    """torch.argmax(dim=1): first maximal value; NaN counts as the maximum (first NaN wins)."""
    nan = np.isnan(logits)
    x = np.where(nan, np.inf, logits)
    arg = np.argmax(x, axis=1)                         # numpy: first occurrence
    has_nan = nan.any(axis=1)
    first_nan = np.argmax(nan, axis=1)
    return np.where(has_nan, first_nan, arg).astype(np.int64)


def to_global_multiclass(classes, categories: dict, *arrays: np.ndarray, compact: bool = True) -> list[np.ndarray]:
    """data/utils.py:567-590 — note the substitutions run in sequence on the same array."""
    out = [a.copy() for a in arrays]
    cats_map = {k: i + 1 for i, k in enumerate(categories.keys())}
    for i in range(len(classes)):
        longest = sorted(list(set(sum([list(c) for c in classes[i]], []))))
        for j, v in enumerate(longest):
            for a in out:
                value = cats_map[v] if compact else v
                a[i] = np.where(a[i] == j + 1, value, a[i])
    return out


def confusion_matrix(preds: np.ndarray, target: np.ndarray, num_classes: int, ignore_index=None) -> np.ndarray:
    p, t = preds.reshape(-1), target.reshape(-1)
    if ignore_index is not None:
        keep = t != ignore_index
        p, t = p[keep], t[keep]
    if ((t < 0) | (t >= num_classes) | (p < 0) | (p >= num_classes)).any():
        raise RuntimeError("labels outside [0, num_classes)")
    return np.bincount(t * num_classes + p, minlength=num_classes ** 2).reshape(num_classes, num_classes).astype(np.int64)


def macro_jaccard(confmat: np.ndarray, ignore_index=None) -> np.float32:
    conf = confmat.astype(np.float32)
    num = np.diag(conf)
    denom = conf.sum(0) + conf.sum(1) - num
    jac = np.where(denom != 0, num / np.where(denom != 0, denom, 1), 0).astype(np.float32)
    w = np.ones_like(jac)
    if ignore_index is not None and 0 <= ignore_index < conf.shape[0]:
        w[ignore_index] = 0
    w[conf.sum(1) + conf.sum(0) == 0] = 0
    return np.float32(((w * jac) / w.sum()).sum())


def strict_mean_iou(confmat: np.ndarray, ignore_index=None) -> np.float32:
    n = confmat.shape[0]
    metric = macro_jaccard(confmat, ignore_index)
    c = confmat.astype(np.float32)
    bg = c[0, 0] / (c[0, 0] + c[0, 1:].sum() + c[1:, 0].sum())
    return np.float32((metric * n - bg) / (n - 1))


def generate_points_from_errors(logits: np.ndarray, gt: np.ndarray, rand: np.ndarray, ignore_index: int = -100,
                                scale_xy=None):
    """label_anything/experiment/substitution.py:17-96 with the random draw made explicit: rand int64 [B, C, n] stands
    in for `torch.randint(0, count, (n,))` of each (b, c) group (index = rand mod count), rows in the intended (b, c)
    order (the reference's `argsort(b * B + c)` agrees whenever its keys are unique).  scale_xy = (sx [B], sy [B]) fp32
    is PromptsProcessor.torch_apply_coords of Substitutor.generate_new_points (:166-171).
    -> points fp32 [B, C, n, 2] (x, y), labels fp32 [B, C, n]."""
    B, C, H, W = logits.shape
    n = rand.shape[2]
    t = np.where(gt == ignore_index, 0, gt)
    pred = argmax_dim1(logits)
    points = np.zeros((B, C, n, 2), dtype=np.float32)
    labels = np.zeros((B, C, n), dtype=np.float32)
    for b in range(B):
        for c in range(C):
            err = (t[b] == c).astype(np.int64) - (pred[b] == c).astype(np.int64)      # one_hot(gt) - one_hot(pred)
            ys, xs = np.nonzero(err)                                                     # row-major, like torch.nonzero
            if len(ys) == 0:
                continue                                                                 # the (0, 0) / label 0 padding row
            for i in range(n):
                k = int(abs(int(rand[b, c, i])) % len(ys))
                points[b, c, i] = (xs[k], ys[k])                                         # x / y swapped (:66-68)
                labels[b, c, i] = err[ys[k], xs[k]]
    labels[:, 0] = 0                                                                     # ignore background (:94-95)
    if scale_xy is not None:
        sx, sy = (np.asarray(v, dtype=np.float32) for v in scale_xy)
        points[..., 0] = points[..., 0] * sx[:, None, None]
        points[..., 1] = points[..., 1] * sy[:, None, None]
    return points, labels
