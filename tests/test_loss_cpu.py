"""Focal loss + class weighting on CPU: the numpy oracle against the fixture generated from the
UNMODIFIED reference loss (values, autograd gradients, weight maps), and the C-ABI argument checks.
Floating point: 1e-5 relative (fp32 arithmetic, summation order differs)."""
import sys
from pathlib import Path

import numpy as np
import pytest
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "oracle"))
import loss_oracle as lo  # noqa: E402

from labelanything_b200 import _native  # noqa: E402

GOLD = torch.load(ROOT / "tests" / "golden" / "loss_f1.pt", weights_only=False)


@pytest.mark.parametrize("case", range(len(GOLD["cases"])))
def test_oracle_matches_the_reference_fixture(case):
    c = GOLD["cases"][case]
    x, t = c["logits"].numpy(), c["target"].numpy()
    wt, cw = lo.get_weight_matrix_from_labels(t.copy(), x.shape[1])
    np.testing.assert_allclose(wt, c["wtarget"].numpy(), rtol=1e-6)
    np.testing.assert_allclose(cw, c["class_weights"].numpy(), rtol=1e-6)
    out = lo.label_anything_loss(x, t, c["gamma"], c["component_weight"], c["class_weighting"])
    assert abs(out["value"] - float(c["value"])) <= 1e-5 * abs(float(c["value"]))
    assert abs(out["components"]["focal"] - c["component"]) <= 1e-5 * abs(c["component"])
    g = lo.focal_loss_grad(x, t, c["gamma"], wt if c["class_weighting"] else None, upstream=c["component_weight"] ** 2)
    np.testing.assert_allclose(g, c["grad"].numpy(), rtol=2e-4, atol=1e-7)


def test_argument_checks_and_workspace():
    lib = _native.lib()
    assert lib.la_focal_loss_workspace_bytes() >= 148 * 8 * 8
    assert lib.la_focal_loss(None, None, None, None, None, None, None, None, None, 1, 2, 16, 2.0, -100, 1) == -1
    assert b"null pointer" in lib.la_last_error()
    assert lib.la_label_class_weights(None, None, 4, 2, -100, None, None) == -1


def test_loss_module_mirrors_the_reference_constructor():
    from labelanything_b200.loss import FocalLoss, LabelAnythingLoss

    loss = LabelAnythingLoss({"focal": {"weight": 0.5, "gamma": 1.5}}, class_weighting=True)
    assert loss.weights == {"focal": 0.5} and isinstance(loss.components["focal"], FocalLoss)
    assert loss.components["focal"].gamma == 1.5 and loss.class_weighting
    with pytest.raises(NotImplementedError, match="dice"):
        LabelAnythingLoss({"dice": {"weight": 1.0}})
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        loss(torch.zeros(1, 2, 4, 4), torch.zeros(1, 4, 4, dtype=torch.int64))


def test_oracle_matches_torch_autograd_on_random_inputs():
    """Independent of the fixture: the same formula written with torch ops (F.cross_entropy, exp, pow, mean) and
    differentiated by autograd, on random shapes / gammas / ignore patterns."""
    import torch.nn.functional as F

    g = torch.Generator().manual_seed(3)
    for B, C, H, W, gamma in [(1, 2, 5, 7, 2.0), (3, 7, 4, 4, 1.0), (2, 9, 6, 5, 2.5), (1, 21, 3, 3, 2.0)]:
        x = (torch.randn(B, C, H, W, generator=g) * 2).requires_grad_(True)
        t = torch.randint(0, C, (B, H, W), generator=g)
        t[0, 0, :2] = -100
        wt_np, _ = lo.get_weight_matrix_from_labels(t.numpy().copy(), C)
        wt = torch.from_numpy(wt_np)
        ce = F.cross_entropy(x, t, reduction="none")
        loss = torch.mean(torch.pow(1 - torch.exp(-ce), gamma) * wt * ce)
        loss.backward()
        assert abs(float(lo.focal_loss(x.detach().numpy(), t.numpy(), gamma, wt_np)) - float(loss)) <= 1e-5 * abs(float(loss))
        np.testing.assert_allclose(lo.focal_loss_grad(x.detach().numpy(), t.numpy(), gamma, wt_np), x.grad.numpy(),
                                   rtol=2e-4, atol=1e-7)
