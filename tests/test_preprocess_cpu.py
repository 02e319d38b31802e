"""Pin the preprocessing oracle (oracle/preprocess_oracle.py): bit-exact against the fixture produced
by the reference's own CustomResize / ToTensor / CustomNormalize / PromptsProcessor (oracle/make_golden.py preprocess ->
tests/golden/preprocess_f3.pt) and, where Pillow is importable, against Pillow itself on more shapes."""
import sys
from pathlib import Path

import numpy as np
import pytest
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "oracle"))
import preprocess_oracle as P  # noqa: E402

GOLD = torch.load(ROOT / "tests" / "golden" / "preprocess_f3.pt", weights_only=False)


@pytest.mark.parametrize("i", range(len(GOLD["images"])))
def test_image_pipeline_is_bit_exact_with_the_reference(i):
    c = GOLD["images"][i]
    out = P.preprocess_image(c["image"].numpy(), c["size"], GOLD["mean"], GOLD["std"], c["custom_preprocess"])
    assert out.shape == c["out_shape"]
    st = c["stride"]
    assert np.array_equal(out[:, ::st, ::st], c["out"].numpy())            # fp32, bit for bit
    assert abs(out.astype(np.float64).sum() - c["out_sum"]) <= 1e-9 * max(1.0, abs(c["out_sum"]))


@pytest.mark.parametrize("i", range(len(GOLD["prompts"])))
def test_prompt_rasterisation_is_bit_exact_with_the_reference(i):
    c = GOLD["prompts"][i]
    m = P.rasterize_masks(c["masks"].numpy(), 1024, 256, c["custom_preprocess"])
    assert np.array_equal(m, c["mask_out"].numpy())
    if c["custom_preprocess"]:
        assert np.array_equal(P.apply_coords(c["points"].numpy(), c["original_size"]), c["points_out"].numpy())
        assert np.array_equal(P.apply_boxes(c["boxes"].numpy(), c["original_size"]), c["boxes_out"].numpy())


def test_resize_restatement_matches_pillow_on_more_shapes():
    Image = pytest.importorskip("PIL.Image")
    rng = np.random.default_rng(0)
    for h, w, oh, ow in [(37, 53, 20, 30), (480, 640, 768, 1024), (300, 200, 1024, 683), (64, 64, 64, 32), (5, 7, 11, 3),
                         (256, 256, 256, 256), (1001, 751, 512, 384)]:
        img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        ref = np.asarray(Image.fromarray(img).resize((ow, oh), Image.BILINEAR))
        assert np.array_equal(P.pil_resize_bilinear_u8(img, oh, ow), ref), (h, w, oh, ow)
