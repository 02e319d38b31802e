"""GPU parity of the native image encoders against golden tensors produced by the unmodified reference
(tests/golden/sam512_vit_1img.pt, sam_vit_neck_1img.pt, mae256_1w1s.pt, mael256_2w5s.pt; weights are the deterministic
synthetic ones).  The native path computes with bf16 operands / fp32 accumulation against an fp32 reference, so the
bounds are END-TO-END drift bounds, stated relative to the standard deviation of the reference tensor: max |err| <=
MAX_REL * std, mean |err| <= MEAN_REL * std (bf16 carries 8 significant bits: 0.4 % per rounding, a dozen roundings per
block, 12-24 blocks).  Per-kernel parity at kernel tolerance lives in tests/test_kernels_gpu.py; the comparison with
the bf16-matched oracle in tests/test_lam_gpu.py."""
from pathlib import Path

import pytest
import torch

GOLD = Path(__file__).resolve().parent / "golden"
pytestmark = pytest.mark.gpu


MAX_REL, MEAN_REL = 0.06, 0.009     # measured (round 2): 0.023-0.039 / 0.0042-0.0061


def _check(name, a, b):
    a, b = a.float().cpu(), b.float().cpu()
    assert a.shape == b.shape, (name, a.shape, b.shape)
    assert bool(torch.isfinite(a).all()), f"{name}: non-finite output"
    err = (a - b).abs()
    std = b.std().item()
    mx, mean = err.max().item() / std, err.mean().item() / std
    print(f"{name}: max_abs_err={err.max().item():.4f} mean_abs_err={err.mean().item():.5f} ref_std={std:.4f} "
          f"-> max/std={mx:.4f} mean/std={mean:.5f}")
    assert mx < MAX_REL and mean < MEAN_REL, f"{name}: max/std {mx:.4f} (< {MAX_REL}), mean/std {mean:.5f} (< {MEAN_REL})"


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
    _check("SAM ViT-B encoder", out[0, ::16], g["encoder_out_sub"])


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
    _check("HF ViT-B encoder", out[0, ::16], g["encoder_out_sub"])


def test_hf_vit_l_480_matches_reference_golden():
    """The encoder of BASELINE config 4 (MAE-L: 1024 wide, 24 layers, 16 heads, MLP 4096)."""
    from labelanything_b200.build_encoder import build_vit_from_config
    from labelanything_b200.synthetic import make_episode, synth_tensor

    g = torch.load(GOLD / "mael256_2w5s.pt", weights_only=False)
    vit = build_vit_from_config(hidden_size=1024, num_hidden_layers=24, num_attention_heads=16, intermediate_size=4096)
    sd = vit.state_dict()
    vit.load_state_dict({k: synth_tensor("image_encoder." + k, tuple(v.shape), 0) for k, v in sd.items()})
    vit = vit.cuda()
    img = make_episode(**g["episode_args"])["images"][0, :1].cuda()
    with torch.no_grad():
        out = vit(img)
    assert out.shape == (1, 1024, 30, 30)
    _check("HF ViT-L encoder", out[0, ::16], g["encoder_out_sub"])


def test_sam_vit_with_neck_and_last_block_state_matches_reference_golden():
    """`forward(x, return_last_block_state=True)` of the SAM encoder built with its own neck: the dict the embedding
    extraction writes (label_anything/preprocess.py:160-162, image_encoder.py:120-131)."""
    from labelanything_b200.build_encoder import build_vit_b
    from labelanything_b200.synthetic import load_synth_weights, make_episode

    g = torch.load(GOLD / "sam_vit_neck_1img.pt", weights_only=False)
    vit = build_vit_b(project_last_hidden=True)
    assert {k: tuple(v.shape) for k, v in vit.state_dict().items()} == g["shapes"]
    load_synth_weights(vit, seed=g["weights_seed"])
    vit = vit.cuda()
    img = make_episode(**g["episode_args"])["images"][0, :1].cuda()
    with torch.no_grad():
        out = vit(img, return_last_block_state=True)
        plain = vit(img)
    assert sorted(out.keys()) == g["keys"] == ["last_block_state", "last_hidden_state"]
    assert out["last_hidden_state"].shape == (1, 256, 64, 64) and out["last_block_state"].shape == (1, 768, 64, 64)
    assert torch.equal(plain, out["last_hidden_state"])
    _check("SAM ViT-B last_block_state", out["last_block_state"][0, ::16], g["last_block_state_sub"])
    _check("SAM ViT-B + neck last_hidden_state", out["last_hidden_state"][0, ::8], g["last_hidden_state_sub"])


def test_chunk_ready_hook_is_called_once_per_encoder_chunk_and_changes_nothing():
    """ImageEncoderViT.chunk_ready (the input pipeline's per-slice wait, bench.py e2e): called with the image range of
    every balanced chunk before its first launch; the features are those of the un-hooked call."""
    from functools import partial

    from labelanything_b200.image_encoder import ImageEncoderViT
    from labelanything_b200.synthetic import load_synth_weights

    vit = ImageEncoderViT(depth=2, embed_dim=128, img_size=1024, mlp_ratio=4, norm_layer=partial(torch.nn.LayerNorm, eps=1e-6),
                          num_heads=2, patch_size=16, qkv_bias=True, use_rel_pos=True, global_attn_indexes=[1],
                          project_last_hidden=False, window_size=14, out_chans=256)
    load_synth_weights(vit, seed=3)
    vit = vit.cuda()
    x = torch.randn(5, 3, 1024, 1024, device="cuda", generator=torch.Generator(device="cuda").manual_seed(1))
    vit.max_images_per_chunk = 2
    with torch.no_grad():
        ref = vit(x)
        calls = []
        vit.chunk_ready = lambda first, n: calls.append((first, n))
        got = vit(x)
        vit.chunk_ready = None
    assert calls == [(0, 2), (2, 2), (4, 1)]
    assert torch.equal(got, ref)
