"""GPU parity of the native image encoders against golden tensors produced by the unmodified reference
(tests/golden/sam512_vit_1img.pt, mae256_1w1s.pt; weights are the deterministic synthetic ones)."""
from pathlib import Path

import pytest
import torch

GOLD = Path(__file__).resolve().parent / "golden"
pytestmark = pytest.mark.gpu


def _stats(a, b):
    err = (a.float() - b.float()).abs()
    return err.max().item(), err.mean().item(), b.float().abs().mean().item()


def test_sam_vit_b_1024_matches_reference_golden():
    from labelanything_b200.build_encoder import build_vit_b
    from labelanything_b200.synthetic import load_synth_weights, make_episode, synth_tensor
    from labelanything_b200.vit_engine import pack_neck, run_neck

    g = torch.load(GOLD / "sam512_vit_1img.pt", weights_only=False)
    vit = build_vit_b(project_last_hidden=False)
    sd = vit.state_dict()
    vit.load_state_dict({k: synth_tensor("image_encoder." + k, tuple(v.shape), 0) for k, v in sd.items()})
    vit = vit.cuda()
    img = make_episode(**g["episode_args"])["images"][0, :1].cuda()
    with torch.no_grad():
        out = vit(img)
    assert out.shape == (1, 768, 64, 64)
    mx, mean, ref_mag = _stats(out[0, ::16].cpu(), g["encoder_out_sub"])
    print(f"SAM ViT-B encoder: max_abs_err={mx:.4f} mean_abs_err={mean:.5f} ref_mean_abs={ref_mag:.4f}")
    # bf16 operands / fp32 accumulate through 12 blocks against an fp32 reference: end-to-end drift bound
    assert mean < 0.02 * max(ref_mag, 1e-3) + 2e-3 and mx < 0.25


def test_hf_vit_b_480_matches_reference_golden():
    from labelanything_b200.build_encoder import build_vit_from_config
    from labelanything_b200.synthetic import make_episode, synth_tensor

    g = torch.load(GOLD / "mae256_1w1s.pt", weights_only=False)
    vit = build_vit_from_config()
    sd = vit.state_dict()
    vit.load_state_dict({k: synth_tensor("image_encoder." + k, tuple(v.shape), 0) for k, v in sd.items()})
    vit = vit.cuda()
    img = make_episode(**g["episode_args"])["images"][0, :1].cuda()
    with torch.no_grad():
        out = vit(img)
    assert out.shape == (1, 768, 30, 30)
    mx, mean, ref_mag = _stats(out[0, ::16].cpu(), g["encoder_out_sub"])
    print(f"HF ViT-B encoder: max_abs_err={mx:.4f} mean_abs_err={mean:.5f} ref_mean_abs={ref_mag:.4f}")
    assert mean < 0.02 * max(ref_mag, 1e-3) + 2e-3 and mx < 0.25
