"""Focal loss + class weighting on the GPU, through the C ABI, against the fixture generated
from the unmodified reference and the numpy oracle.  fp32 arithmetic: values 1e-5 relative, gradients 2e-4 relative
+ 1e-7 absolute (exp / log of the device math library vs the host's)."""
import sys
from pathlib import Path

import numpy as np
import pytest
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "oracle"))
pytestmark = pytest.mark.gpu

GOLD = torch.load(ROOT / "tests" / "golden" / "loss_f1.pt", weights_only=False)


@pytest.mark.parametrize("case", range(len(GOLD["cases"])))
def test_loss_value_and_gradient_match_the_reference_fixture(case):
    from labelanything_b200.loss import LabelAnythingLoss, get_weight_matrix_from_labels

    c = GOLD["cases"][case]
    x = c["logits"].cuda().requires_grad_(True)
    t = c["target"].cuda()
    loss = LabelAnythingLoss({"focal": {"weight": c["component_weight"], "gamma": c["gamma"]}},
                             class_weighting=c["class_weighting"])
    out = loss(x, t)
    assert abs(float(out["value"]) - float(c["value"])) <= 1e-5 * abs(float(c["value"]))
    assert abs(out["components"]["focal"] - c["component"]) <= 1e-5 * abs(c["component"])
    out["value"].backward()
    np.testing.assert_allclose(x.grad.cpu().numpy(), c["grad"].numpy(), rtol=2e-4, atol=1e-7)
    wt, cw = get_weight_matrix_from_labels(t, x.shape[1])
    np.testing.assert_allclose(wt.cpu().numpy(), c["wtarget"].numpy(), rtol=1e-6)
    np.testing.assert_allclose(cw.cpu().numpy(), c["class_weights"].numpy(), rtol=1e-6)


def test_full_size_properties():
    """BASELINE size (8 x 6 x 1024 x 1024): deterministic value, sum == mean * N, gradient rows sum to zero, ignored
    pixels get exactly zero gradient, and the value agrees with torch's own ops on the same tensors."""
    import torch.nn.functional as F

    from labelanything_b200 import ops

    B, C, H, W = 8, 6, 1024, 1024
    g = torch.Generator(device="cuda").manual_seed(1)
    x = torch.randn(B, C, H, W, device="cuda", generator=g) * 2
    t = torch.randint(0, C, (B, H, W), device="cuda", generator=g)
    t[:, :, :5] = -100
    cw, hist = ops.label_class_weights(t, C)
    assert int(hist.sum()) == t.numel() and int(hist[C]) == int((t == -100).sum()) and int(hist[C + 1]) == 0
    v1, grad, _ = ops.focal_loss(x, t, cw, 2.0, want_grad=True)
    v2, _, _ = ops.focal_loss(x, t, cw, 2.0)
    vs, _, _ = ops.focal_loss(x, t, cw, 2.0, mean=False)
    assert float(v1) == float(v2)                                           # fixed summation order
    assert abs(float(vs) - float(v1) * t.numel()) <= 1e-5 * abs(float(vs))
    ce = F.cross_entropy(x, t, reduction="none")
    wt = torch.where(t == -100, torch.zeros((), device="cuda"), cw[t.clamp(min=0)])
    ref = ((1 - torch.exp(-ce)) ** 2 * wt * ce).double().mean()
    assert abs(float(v1) - float(ref)) <= 1e-5 * abs(float(ref))
    assert float(grad.sum(dim=1).abs().max()) < 1e-9                        # softmax gradient rows sum to zero
    assert float(grad[:, :, :, :5].abs().max()) == 0.0


def test_out_of_range_target_poisons_the_loss():
    from labelanything_b200 import ops

    x = torch.zeros(1, 3, 4, 4, device="cuda")
    t = torch.zeros(1, 4, 4, dtype=torch.int64, device="cuda")
    t[0, 1, 1] = 7
    v, _, _ = ops.focal_loss(x, t, None, 2.0)
    assert torch.isnan(v)
