"""Post-logits step on the GPU, through the C ABI: argmax + global-label mapping + confusion
matrix of `la_label_confusion` against the numpy oracle and the fixture generated from the unmodified reference.
Integer work: every comparison is bit-exact."""
import sys
from pathlib import Path

import numpy as np
import pytest
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "oracle"))
pytestmark = pytest.mark.gpu

GOLD = torch.load(ROOT / "tests" / "golden" / "metrics_f4.pt", weights_only=False)


def _mo():
    import metrics_oracle

    return metrics_oracle


@pytest.mark.parametrize("case", range(len(GOLD["cases"])))
def test_fused_update_matches_the_reference_fixture(case):
    from labelanything_b200.metrics import StrictMeanIoU, to_global_multiclass

    c = GOLD["cases"][case]
    m = StrictMeanIoU(num_classes=c["num_classes"], ignore_index=-100)
    gp, gg = m.update_from_logits(c["logits"].cuda(), c["gt"].cuda(), c["classes"], GOLD["categories"], want_labels=True)
    assert torch.equal(gp.cpu(), c["glob_preds"]) and torch.equal(gg.cpu(), c["glob_gt"])
    assert torch.equal(m.confmat.cpu(), c["confmat"])
    want = _mo().strict_mean_iou(c["confmat"].numpy(), -100)
    assert abs(float(m.compute()) - float(want)) < 1e-6
    # the unfused route of the reference: argmax'd labels -> to_global_multiclass -> update
    m2 = StrictMeanIoU(num_classes=c["num_classes"], ignore_index=-100)
    g2p, g2g = to_global_multiclass(c["classes"], GOLD["categories"], c["preds"].cuda(), c["gt"].cuda())
    assert torch.equal(g2p.cpu(), c["glob_preds"]) and torch.equal(g2g.cpu(), c["glob_gt"])
    m2.update(g2p, g2g)
    m2.update(g2p, g2g)                                                   # the state accumulates
    assert torch.equal(m2.confmat.cpu(), 2 * c["confmat"])


@pytest.mark.parametrize("B,C,H,W,G", [
    (2, 6, 64, 64, 21),        # vector path, shared-memory histogram
    (3, 2, 33, 17, 3),         # odd pixel count: scalar path
    (1, 21, 128, 96, 81),      # COCO-sized label space
    (2, 4, 16, 16, 200),       # label space too large for shared memory: global atomics
])
def test_label_confusion_matches_the_oracle(B, C, H, W, G):
    from labelanything_b200 import ops

    mo = _mo()
    g = torch.Generator().manual_seed(B * 1000 + C)
    logits = torch.randn(B, C, H, W, generator=g)
    logits[:, :, ::3] = torch.round(logits[:, :, ::3])                    # ties -> first index
    logits[0, 0, 0, 0] = float("nan")
    gt = torch.randint(0, C, (B, H, W), generator=g)
    gt[:, :2] = -100
    table = torch.stack([torch.randperm(G, generator=g)[: C + 1] for _ in range(B)]).to(torch.int64)
    conf = torch.zeros(G, G, dtype=torch.int64, device="cuda")
    bad = torch.zeros(1, dtype=torch.int64, device="cuda")
    p, t = ops.label_confusion(logits.cuda(), None, gt.cuda(), table.cuda(), conf, bad)
    preds = mo.argmax_dim1(logits.numpy())
    bi = np.arange(B).reshape(-1, 1, 1)
    want_p = table.numpy()[bi, preds]
    gtn = gt.numpy()
    want_t = np.where(gtn >= 0, table.numpy()[bi, np.clip(gtn, 0, C)], gtn)
    assert np.array_equal(p.cpu().numpy(), want_p) and np.array_equal(t.cpu().numpy(), want_t)
    assert np.array_equal(conf.cpu().numpy(), mo.confusion_matrix(want_p, want_t, G, -100))
    assert int(bad.cpu()) == 0


def test_out_of_range_labels_are_counted_not_binned():
    from labelanything_b200.metrics import MeanIoU

    m = MeanIoU(num_classes=3, ignore_index=-100)
    preds = torch.tensor([[[0, 1, 2, 5]]], device="cuda")
    gt = torch.tensor([[[0, 1, 7, 2]]], device="cuda")
    m.update(preds, gt)
    assert m.confmat.cpu().tolist() == [[1, 0, 0], [0, 1, 0], [0, 0, 0]]
    with pytest.raises(RuntimeError, match="2 label"):
        m.compute()


def test_full_size_checksum_properties():
    """BASELINE size (8 x 6 x 1024 x 1024): every non-ignored pixel lands in exactly one bin, row sums equal the
    target histogram, and the matrix is the sum of the matrices of any split of the batch."""
    from labelanything_b200 import ops

    B, C, H, W = 8, 6, 1024, 1024
    g = torch.Generator(device="cuda").manual_seed(0)
    logits = torch.randn(B, C, H, W, device="cuda", generator=g)
    gt = torch.randint(0, C, (B, H, W), device="cuda", generator=g)
    gt[:, :, :7] = -100
    conf = torch.zeros(C, C, dtype=torch.int64, device="cuda")
    bad = torch.zeros(1, dtype=torch.int64, device="cuda")
    p, _ = ops.label_confusion(logits, None, gt, None, conf, bad, want_gt=False)
    assert torch.equal(p, logits.argmax(dim=1))
    assert int(conf.sum()) == int((gt != -100).sum()) and int(bad) == 0
    assert torch.equal(conf.sum(1), torch.bincount(gt[gt != -100], minlength=C))
    parts = torch.zeros_like(conf)
    for b0 in (0, 3):
        b1 = 3 if b0 == 0 else B
        ops.label_confusion(logits[b0:b1], None, gt[b0:b1], None, parts, bad, want_preds=False, want_gt=False)
    assert torch.equal(parts, conf)


def test_error_points_match_the_reference_fixture_and_the_oracle():
    """la_error_points (generate_points_from_errors, substitution.py:17-96) through the C ABI: exact coordinates and
    labels against the fixture of the unmodified reference (same pinned draws), against the oracle on a large ragged
    case with several points per class, and the Substitutor glue."""
    from labelanything_b200.substitution import Substitutor, generate_points_from_errors

    g = torch.load(ROOT / "tests" / "golden" / "points_f4.pt", weights_only=False)
    for c in g["cases"]:
        pts, labels = generate_points_from_errors(c["logits"].cuda(), c["gt"].cuda(), c["num_points"],
                                                  rand=c["rand"].cuda())
        assert torch.equal(pts.cpu(), c["points"]) and torch.equal(labels.cpu(), c["labels"])
    gen = torch.Generator().manual_seed(1)
    B, C, H, W, n = 4, 21, 300, 517, 3
    logits = torch.randn(B, C, H, W, generator=gen)
    logits[:, :, : H // 2] = torch.round(logits[:, :, : H // 2])           # ties -> first maximum
    gt = torch.randint(0, C, (B, H, W), generator=gen)
    gt[:, -5:] = -100
    gt[0][gt[0] == 7] = 0
    logits[0, 7] = -60.0                                                    # class 7 of episode 0: no error -> padding
    rand = torch.randint(0, 2 ** 31 - 1, (B, C, n), generator=gen)
    sx, sy = torch.rand(B, generator=gen) + 0.5, torch.rand(B, generator=gen) + 0.5
    pts, labels = generate_points_from_errors(logits.cuda(), gt.cuda(), n, rand=rand.cuda(), scale_xy=(sx, sy))
    mo = _mo()
    rp, rl = mo.generate_points_from_errors(logits.numpy(), gt.numpy(), rand.numpy(), scale_xy=(sx.numpy(), sy.numpy()))
    assert np.array_equal(pts.cpu().numpy(), rp) and np.array_equal(labels.cpu().numpy(), rl)
    assert float(labels[0, 7].abs().sum()) == 0 and float(pts[0, 7].abs().sum()) == 0 and float(labels[:, 0].abs().sum()) == 0
    # Substitutor.generate_new_points: one more point slot on the query example, zero padding on the others
    from labelanything_b200.utils import BatchKeys

    M, P = 3, 2
    batch = {BatchKeys.PROMPT_POINTS: torch.rand(B, M, C, P, 2).cuda(), BatchKeys.FLAG_POINTS: torch.ones(B, M, C, P).cuda(),
             BatchKeys.DIMS: torch.tensor([[[H, W]] * M] * B)}
    sub = Substitutor(num_points=1, long_side_length=1024)
    sub.reset((batch, None))
    sub.generate_new_points(logits.cuda(), gt.cuda(), rand=rand[:, :, :1].contiguous().cuda())
    assert batch[BatchKeys.PROMPT_POINTS].shape == (B, M, C, P + 1, 2) and batch[BatchKeys.FLAG_POINTS].shape == (B, M, C, P + 1)
    assert float(batch[BatchKeys.FLAG_POINTS][:, 1:, :, P:].abs().sum()) == 0
    sxs, sys_ = sub._scales(batch[BatchKeys.DIMS])
    rp1, rl1 = mo.generate_points_from_errors(logits.numpy(), gt.numpy(), rand[:, :, :1].numpy(),
                                              scale_xy=(sxs.numpy(), sys_.numpy()))
    assert np.array_equal(batch[BatchKeys.PROMPT_POINTS][:, 0, :, P:].cpu().numpy(), rp1)
    assert np.array_equal(batch[BatchKeys.FLAG_POINTS][:, 0, :, P:].cpu().numpy(), rl1)
