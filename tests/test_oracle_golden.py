"""Pin the CPU oracle (oracle/lam_oracle.py) against fixtures produced by the UNMODIFIED reference
(oracle/make_golden.py -> tests/golden/*.pt).  CPU-only; runs in seconds."""
import sys
from pathlib import Path

import pytest
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "oracle"))
import lam_oracle as O  # noqa: E402

GOLD = ROOT / "tests" / "golden"


def _load(name):
    p = GOLD / name
    if not p.exists():
        pytest.skip(f"{name} not generated")
    return torch.load(p, weights_only=False)


def _cmp(a, b, tol):
    assert a.shape == b.shape
    fin = torch.isfinite(b)
    assert torch.equal(torch.isfinite(a), fin), "-inf padding pattern differs"
    assert torch.equal(a[~fin], b[~fin])
    err = (a[fin] - b[fin]).abs().max().item() if fin.any() else 0.0
    assert err <= tol, f"max abs err {err} > {tol}"


@pytest.mark.parametrize("name", ["tiny_sam_lam.pt", "tiny_sam_lam_masks_only.pt", "tiny_mae_lam.pt"])
def test_tiny_lam_matches_reference(name):
    g = _load(name)
    with torch.no_grad():
        out = O.lam_forward(g["state_dict"], g["cfg"], dict(g["episode"]), class_rows=g["class_rows"],
                            return_intermediates=True)
    if "encoder_out" in g:
        _cmp(out["encoder_out"], g["encoder_out"], 2e-5)
    _cmp(out["class_examples_embeddings"], g["class_examples_embeddings"], 2e-5)
    _cmp(out["logits"], g["logits"], 5e-5)


def test_rel_pos_table_interpolates_like_reference():
    # get_rel_pos resizes the table linearly when its length mismatches (image_encoder.py:321-328)
    t = torch.randn(11, 8)
    full = O.rel_pos_table(4, 4, t)  # 2*4-1 = 7 != 11 -> interpolation path
    assert full.shape == (4, 4, 8)
    t7 = torch.nn.functional.interpolate(t.t()[None], size=7, mode="linear")[0].t()
    assert torch.allclose(full[0, 0], t7[3]) and torch.allclose(full[3, 0], t7[6]) and torch.allclose(full[0, 3], t7[0])


def test_prepare_prompts_drops_all_zero_types():
    ep = {"prompt_points": torch.zeros(1, 1, 2, 1, 2), "flag_points": torch.zeros(1, 1, 2, 1),
          "prompt_masks": torch.zeros(1, 1, 2, 8, 8), "flag_masks": torch.ones(1, 1, 2),
          "flag_examples": torch.ones(1, 1, 2)}
    p, b, m, fe = O.prepare_prompts(ep)
    assert p is None and b is None and m is not None


def test_missing_inputs_raise_like_reference():
    with pytest.raises(ValueError, match="Either 'images' or 'embeddings'"):
        O.lam_forward({}, {"image_size": 64}, {"dims": torch.zeros(1, 1, 2)})


@pytest.mark.parametrize("name", ["tiny_sam_lam.pt", "tiny_sam_lam_masks_only.pt", "tiny_mae_lam.pt"])
def test_bf16_matched_oracle_is_the_same_algorithm(name):
    """oracle/lam_oracle_bf16.py restates the path with the native rounding points (and the native path's exact
    re-associations: positional tables projected separately, the 1x1 mask conv after the resize, the single-key
    softmax).  With the roundings switched off it must BE the reference algorithm: pinned against the same fixtures of
    the unmodified reference, at fp32 re-association tolerance.  With them on, it must stay within bf16 drift of it."""
    import lam_oracle_bf16 as OB

    g = _load(name)
    OB.EXACT = True
    try:
        with torch.no_grad():
            out = OB.lam_forward(g["state_dict"], g["cfg"], dict(g["episode"]), class_rows=g["class_rows"])
    finally:
        OB.EXACT = False
    _cmp(out["class_examples_embeddings"], g["class_examples_embeddings"], 5e-5)
    _cmp(out["logits"], g["logits"], 1e-4)
    with torch.no_grad():
        rounded = OB.lam_forward(g["state_dict"], g["cfg"], dict(g["episode"]), class_rows=g["class_rows"])
    fin = torch.isfinite(g["logits"])
    assert torch.equal(torch.isfinite(rounded["logits"]), fin)
    err = (rounded["logits"][fin] - g["logits"][fin]).abs()
    std = g["logits"][fin].std()
    assert 0 < err.max() < 0.35 * std and err.mean() < 0.03 * std, (err.max().item(), err.mean().item(), std.item())
