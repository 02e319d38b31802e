"""Post-logits step on CPU: the numpy oracle against the fixture generated from the UNMODIFIED
reference (`to_global_multiclass` + torch.argmax), the host-side label table, the IoU reductions, the C-ABI argument
checks and the 2-rank state reduction."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "oracle"))
import metrics_oracle as mo  # noqa: E402

from labelanything_b200 import _native  # noqa: E402
from labelanything_b200.metrics import MeanIoU, StrictMeanIoU, chain_label_map  # noqa: E402

GOLD = torch.load(ROOT / "tests" / "golden" / "metrics_f4.pt", weights_only=False)


@pytest.mark.parametrize("case", range(len(GOLD["cases"])))
def test_oracle_matches_the_reference_fixture(case):
    c = GOLD["cases"][case]
    preds = mo.argmax_dim1(c["logits"].numpy())
    assert np.array_equal(preds, c["preds"].numpy())                      # ties, -inf planes, NaN
    gp, gg = mo.to_global_multiclass(c["classes"], GOLD["categories"], preds, c["gt"].numpy())
    assert np.array_equal(gp, c["glob_preds"].numpy()) and np.array_equal(gg, c["glob_gt"].numpy())
    conf = mo.confusion_matrix(gp, gg, c["num_classes"], ignore_index=-100)
    assert np.array_equal(conf, c["confmat"].numpy())
    assert conf.sum() == int((c["gt"] != -100).sum())


@pytest.mark.parametrize("case", range(len(GOLD["cases"])))
def test_chained_label_table_reproduces_the_sequential_substitutions(case):
    c = GOLD["cases"][case]
    table = chain_label_map(c["classes"], GOLD["categories"], map_len=c["logits"].shape[1] + 2)
    for name in ("preds", "gt"):
        local, glob = c[name], c["glob_" + name]
        inside = (local >= 0) & (local < table.shape[1])
        idx = local.clamp(0, table.shape[1] - 1)
        b = torch.arange(local.shape[0]).view(-1, 1, 1).expand_as(local)
        mapped = torch.where(inside, table[b, idx], local)
        assert torch.equal(mapped, glob), name


def test_chain_differs_from_a_one_shot_lookup():
    # classes {3, 7} of categories (1, 3, 5, 7): local 1 -> 2 -> 4 because the second substitution sees the first
    t = chain_label_map([[[7, 3], [3]]], {1: {}, 3: {}, 5: {}, 7: {}})
    assert t.tolist() == [[0, 4, 4]]
    assert chain_label_map([[[7, 3]]], {1: {}, 3: {}, 5: {}, 7: {}}, compact=False).tolist() == [[0, 3, 7]]


def test_iou_reductions_match_the_oracle_and_a_hand_example():
    conf = torch.tensor([[50, 2, 0, 0], [3, 20, 0, 0], [0, 0, 0, 0], [4, 0, 0, 10]])
    j = [50 / 59, 20 / 25, 0.0, 10 / 14]                                  # class 2 is absent: weight 0
    want = (j[0] + j[1] + j[3]) / 3
    assert abs(float(MeanIoU._macro_jaccard(conf, -100)) - want) < 1e-6
    assert abs(float(mo.macro_jaccard(conf.numpy(), -100)) - want) < 1e-6
    m = StrictMeanIoU(num_classes=4, ignore_index=-100, device="cpu")
    m.confmat.copy_(conf)
    strict = (want * 4 - 50 / 59) / 3                                     # utils/metrics.py:31-35
    assert abs(float(m.compute()) - strict) < 1e-6
    assert abs(float(mo.strict_mean_iou(conf.numpy(), -100)) - strict) < 1e-6
    m._invalid += 1
    with pytest.raises(RuntimeError, match="outside"):
        m.compute()


def test_label_confusion_argument_checks():
    lib = _native.lib()
    assert lib.la_label_confusion(None, None, None, None, None, None, None, None, None, 1, 2, 16, 0, 0, -100) == -1
    assert b"nothing to read" in lib.la_last_error()
    assert lib.la_label_confusion(None, None, None, None, None, None, None, None, None, 0, 2, 16, 0, 0, -100) == -1
    assert b"empty problem" in lib.la_last_error()


def test_updates_refuse_cpu_tensors():
    m = MeanIoU(num_classes=3, ignore_index=-100, device="cpu")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m.update(torch.zeros(1, 4, 4, dtype=torch.int64), torch.zeros(1, 4, 4, dtype=torch.int64))


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        m = StrictMeanIoU(num_classes=3, ignore_index=-100, device="cpu")
        m.confmat += torch.tensor([[5, 1, 0], [0, 4, 2], [1, 0, 3]]) * (rank + 1)
        m.sync()
        if rank == 0:
            q.put((m.confmat.tolist(), float(m.compute())))
    finally:
        dist.destroy_process_group()


def test_two_rank_state_reduction():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    conf, value = q.get()
    assert conf == [[15, 3, 0], [0, 12, 6], [3, 0, 9]]
    want = mo.strict_mean_iou(np.array(conf), -100)
    assert abs(value - float(want)) < 1e-6


def test_chained_table_equals_sequential_substitution_on_random_episodes():
    """Property check against the oracle's literal restatement of data/utils.py:583-589 (itself pinned by the
    reference fixture): random category sets, random class lists (with repeats and several examples), both `compact`
    settings, labels that include untouched values."""
    rng = np.random.default_rng(7)
    for trial in range(60):
        n_cat = int(rng.integers(2, 25))
        cat_ids = sorted(rng.choice(np.arange(1, 91), size=n_cat, replace=False).tolist())
        categories = {int(k): {} for k in cat_ids}
        B = int(rng.integers(1, 4))
        classes = [[rng.choice(cat_ids, size=int(rng.integers(1, min(6, n_cat) + 1)), replace=False).tolist()
                    for _ in range(int(rng.integers(1, 4)))] for _ in range(B)]
        compact = bool(trial % 2)
        labels = rng.integers(-1, 9, size=(B, 5, 6)).astype(np.int64)
        labels[:, 0, 0] = -100
        want = mo.to_global_multiclass(classes, categories, labels, compact=compact)[0]
        table = chain_label_map(classes, categories, compact=compact, map_len=32).numpy()
        inside = (labels >= 0) & (labels < table.shape[1])
        got = np.where(inside, table[np.arange(B)[:, None, None], np.clip(labels, 0, table.shape[1] - 1)], labels)
        assert np.array_equal(got, want), (classes, cat_ids, compact)


def test_error_points_oracle_matches_the_reference_fixture():
    """generate_points_from_errors (substitution.py:17-96): the oracle with the recorded draws reproduces the UNMODIFIED
    reference function (whose torch.randint was pinned to the same draws) exactly -- coordinates, labels, padding rows,
    the all-correct early return."""
    import torch

    g = torch.load(ROOT / "tests" / "golden" / "points_f4.pt", weights_only=False)
    for c in g["cases"]:
        pts, labels = mo.generate_points_from_errors(c["logits"].numpy(), c["gt"].numpy(), c["rand"].numpy())
        assert np.array_equal(pts, c["points"].numpy()), tuple(c["logits"].shape)
        assert np.array_equal(labels, c["labels"].numpy())


def test_macro_jaccard_known_answer_from_the_torchmetrics_documentation():
    """torchmetrics is not installed here, so its reduce is restated; this pins it to the worked example of the
    MulticlassJaccardIndex docstring of torchmetrics 1.7.1 (the version of the reference's uv.lock:2672-2673):
        target = [2, 1, 0, 0], preds = [2, 1, 0, 1], num_classes = 3  ->  tensor(0.6667)
    (IoU per class 1/2, 1/2, 1) and to the absent-class rule of `_jaccard_index_reduce` (a class with no prediction and
    no target gets weight 0 in the macro average)."""
    target, preds = np.array([2, 1, 0, 0]), np.array([2, 1, 0, 1])
    conf = mo.confusion_matrix(preds, target, 3) if hasattr(mo, "confusion_matrix") else None
    if conf is None:
        conf = np.zeros((3, 3), dtype=np.int64)
        for t, p in zip(target, preds):
            conf[t, p] += 1
    val = MeanIoU._macro_jaccard(torch.from_numpy(conf), None)
    assert abs(float(val) - 2.0 / 3.0) < 1e-6 and f"{float(val):.4f}" == "0.6667"
    conf4 = np.zeros((4, 4), dtype=np.int64)
    conf4[:3, :3] = conf                                               # class 3 never appears
    assert abs(float(MeanIoU._macro_jaccard(torch.from_numpy(conf4), None)) - 2.0 / 3.0) < 1e-6
    # ignore_index inside the class range: that class is dropped from the average
    assert abs(float(MeanIoU._macro_jaccard(torch.from_numpy(conf), 2)) - 0.5) < 1e-6
