"""GPU input preprocessing through the C ABI: bit-exact against the fixture produced by the
reference's own CustomResize / ToTensor / CustomNormalize / PromptsProcessor (tests/golden/preprocess_f3.pt) and against
the CPU oracle (oracle/preprocess_oracle.py, pinned to Pillow) on more shapes; full-size properties."""
import sys
from pathlib import Path

import numpy as np
import pytest
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "oracle"))
pytestmark = pytest.mark.gpu
GOLD = torch.load(ROOT / "tests" / "golden" / "preprocess_f3.pt", weights_only=False)


@pytest.mark.parametrize("i", range(len(GOLD["images"])))
def test_image_pipeline_is_bit_exact_with_the_reference(i):
    from labelanything_b200.transforms import ImagePreprocessor

    c = GOLD["images"][i]
    pre = ImagePreprocessor(c["size"], GOLD["mean"], GOLD["std"], c["custom_preprocess"])
    out = pre(c["image"]).cpu()
    assert tuple(out.shape) == c["out_shape"]
    st = c["stride"]
    assert torch.equal(out[:, ::st, ::st], c["out"])                       # fp32, bit for bit
    assert abs(out.double().sum().item() - c["out_sum"]) <= 1e-9 * max(1.0, abs(c["out_sum"]))


@pytest.mark.parametrize("i", range(len(GOLD["prompts"])))
def test_prompt_rasterisation_is_bit_exact_with_the_reference(i):
    from labelanything_b200.transforms import PromptsProcessor

    c = GOLD["prompts"][i]
    pp = PromptsProcessor(1024, 256, c["custom_preprocess"])
    flag = torch.zeros(1, dtype=torch.uint8, device="cuda")
    m = pp.apply_masks(c["masks"], flag=flag).cpu()
    assert torch.equal(m, c["mask_out"].float())
    assert int(flag.item()) == int(c["mask_out"].sum() > 0)
    if c["custom_preprocess"]:
        assert torch.equal(pp.apply_coords(c["points"], c["original_size"]).cpu(), c["points_out"].float())
        assert torch.equal(pp.apply_boxes(c["boxes"], c["original_size"]).cpu(), c["boxes_out"].float())


def test_against_the_oracle_on_more_shapes_and_a_batch():
    import preprocess_oracle as P

    from labelanything_b200.transforms import preprocess_images

    rng = np.random.default_rng(3)
    imgs = [rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
            for h, w in [(1333, 2000), (480, 640), (1024, 1024), (1024, 700), (33, 1500), (2048, 1365)]]
    out, dims = preprocess_images(imgs, size=1024)
    assert out.shape == (6, 3, 1024, 1024) and dims.tolist() == [[im.shape[0], im.shape[1]] for im in imgs]
    mean, std = (0.485, 0.456, 0.406), (0.229, 0.224, 0.225)
    for i in (0, 3, 4):                                                      # the oracle is slow python: three of them
        ref = P.preprocess_image(imgs[i], 1024, mean, std)
        assert np.array_equal(out[i].cpu().numpy(), ref), i
    # properties at full size: the padding is exactly zero, the un-padded region is never the pad value by accident
    for i, im in enumerate(imgs):
        nh, nw = P.get_preprocess_shape(im.shape[0], im.shape[1], 1024)
        assert float(out[i, :, nh:, :].abs().sum()) == 0.0 and float(out[i, :, :, nw:].abs().sum()) == 0.0
        assert bool(torch.isfinite(out[i]).all())
    # idempotence of the identity case: a 1024 x 1024 image only gets normalised
    ident = (torch.from_numpy(imgs[2]).permute(2, 0, 1).float() / 255.0 - torch.tensor(mean).view(3, 1, 1)) / \
        torch.tensor(std).view(3, 1, 1)
    assert torch.equal(out[2].cpu(), ident)


def test_bad_arguments_fail_loudly():
    from labelanything_b200.transforms import ImagePreprocessor

    with pytest.raises(ValueError, match="uint8 RGB"):
        ImagePreprocessor(256)(torch.zeros(10, 10, 3))
