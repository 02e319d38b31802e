"""Host-side logic of the drop-in boundary (no GPU): module surface, state-dict keys, config capture, registry,
pickling, error behaviour — mirroring what the reference's callers rely on."""
import pickle
from pathlib import Path

import pytest
import torch

from labelanything_b200 import models
from labelanything_b200.build_encoder import build_vit_from_config
from labelanything_b200.synthetic import load_synth_weights, make_episode
from labelanything_b200.utils import BatchKeys, ResultDict, get_preprocess_shape, load_state_dict

GOLD = Path(__file__).resolve().parent / "golden"


def _mae256(**kw):
    return models.build_lam(build_vit=lambda project_last_hidden: build_vit_from_config(), image_embed_dim=768,
                            embed_dim=256, image_size=480, spatial_convs=3,
                            class_encoder={"name": "RandomMatrixEncoder", "bank_size": 100, "embed_dim": 256},
                            custom_preprocess=False, **kw)


def test_state_dict_keys_equal_the_reference():
    g = torch.load(GOLD / "mae256_1w1s.pt", weights_only=False)
    assert {k: tuple(v.shape) for k, v in _mae256().state_dict().items()} == g["shapes"]
    g = torch.load(GOLD / "sam512_vit_1img.pt", weights_only=False)
    lam = models.build_lam_vit_b(image_embed_dim=768, embed_dim=512, image_size=1024, use_vit_sam_neck=False,
                                 spatial_convs=3, example_attention=True, example_class_attention=False,
                                 class_encoder={"name": "RandomMatrixEncoder", "bank_size": 100, "embed_dim": 512})
    assert {k: tuple(v.shape) for k, v in lam.state_dict().items()} == g["shapes"]


def test_registry_and_submodule_attributes():
    for key in ("lam", "lam_no_vit", "lam_b", "lam_l", "lam_h", "lam_mae_b", "vit_b", "vit_l", "vit_h", "vit_b_mae"):
        assert key in models.model_registry
    lam = models.model_registry["lam_no_vit"](image_embed_dim=384, embed_dim=256, image_size=256, spatial_convs=3)
    assert lam.image_encoder is None and lam.neck is not None and lam.class_embeddings is None
    # attributes reached from outside the model in the reference (explainer.py, run.py, lam.py:254-302)
    pe, md = lam.prompt_encoder, lam.mask_decoder
    assert pe.transformer.layers[-1].cross_attn_image_to_token.q_proj.weight.shape == (128, 256)
    assert pe.transformer.attention_downsample_rate == 2 and hasattr(pe, "pe_layer")
    assert md.class_mlp.layers[2].weight.shape == (32, 256) and len(md.output_upscaling) == 4
    assert md.spatial_convs is not None and hasattr(md, "_get_pe_result")
    assert tuple(lam.get_dense_pe().shape) == (1, 256, 16, 16)


def test_vit_h_is_refused_at_build_time():
    """ViT-H has head_dim 80; the native attention kernels are head_dim 64.  The registry keeps the reference's keys,
    but building must fail loudly (before any checkpoint is loaded), not at the first forward."""
    import pytest

    with pytest.raises(NotImplementedError, match="head_dim 80"):
        models.model_registry["vit_h"]()
    with pytest.raises(NotImplementedError, match="head_dim 80"):
        models.model_registry["lam_h"]()


def test_hf_vit_with_another_activation_is_refused():
    import pytest
    from transformers import ViTConfig

    vit = models.ViTModelWrapper(ViTConfig(hidden_size=64, num_hidden_layers=1, num_attention_heads=1,
                                           intermediate_size=128, image_size=32, patch_size=16,
                                           hidden_act="gelu_new"))
    with pytest.raises(NotImplementedError, match="hidden_act"):
        vit._spec(2, 2)


def test_labelanything_wrapper_captures_config_and_pickles():
    m = models.LabelAnything(encoder=lambda project_last_hidden: build_vit_from_config(), image_embed_dim=768,
                             embed_dim=256, image_size=480, spatial_convs=3, custom_preprocess=False)
    assert m.config["embed_dim"] == 256 and m.config["image_size"] == 480 and m.config["use_vit"] is True
    assert isinstance(m.model, models.Lam)
    lam = m.model
    lam.packed("dummy", lambda: torch.zeros(1), lam.neck[0].weight)   # a packed-weight cache entry must not be pickled
    clone = pickle.loads(pickle.dumps(lam))
    assert "_la_cache" not in clone.__dict__
    assert clone.state_dict().keys() == lam.state_dict().keys()


def test_packed_cache_follows_parameter_updates():
    lam = models.build_lam_no_vit(image_embed_dim=384, embed_dim=256, image_size=256, spatial_convs=3)
    w = lam.neck[0].weight
    a = lam.packed("k", lambda: w.detach().clone(), w)
    assert lam.packed("k", lambda: None, w) is a
    with torch.no_grad():
        w.add_(1.0)
    b = lam.packed("k", lambda: w.detach().clone(), w)
    assert b is not a and torch.equal(b, w)


def test_load_state_dict_prefix_fallbacks_and_learnable_params():
    lam = _mae256()
    sd = {"model." + k: v for k, v in lam.state_dict().items()}
    load_state_dict(lam, sd)
    params = lam.get_learnable_params({"freeze_backbone": True})
    assert all(not p.requires_grad for p in lam.image_encoder.parameters())
    n_enc = sum(1 for _ in lam.image_encoder.parameters())
    assert len(params) == sum(1 for _ in lam.parameters()) - n_enc
    with pytest.raises(ValueError, match="Cannot freeze the backbone"):
        lam.get_learnable_params({"freeze_backbone": True, "backbone_lr": 1e-5})


def test_errors_match_the_reference_conventions():
    lam = models.build_lam_no_vit(image_embed_dim=384, embed_dim=256, image_size=256, spatial_convs=3)
    with pytest.raises(ValueError, match="Either 'images' or 'embeddings' must be provided."):
        lam({"dims": torch.zeros(1, 2, 2)})
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        lam(make_episode(1, 1, 1, 256, embeddings=(384, 16)))
    with pytest.raises(NotImplementedError):
        models.build_lam_no_vit(few_type="Affinity")
    assert get_preprocess_shape(480, 640, 1024) == (768, 1024)
    assert str(ResultDict.LOGITS) == "logits" and BatchKeys.FLAG_EXAMPLES == "flag_examples"


def test_prepare_prompts_drops_all_zero_types_like_the_reference():
    lam = models.build_lam_no_vit(image_embed_dim=384, embed_dim=256, image_size=256, spatial_convs=3)
    ep = make_episode(1, 2, 1, 256, embeddings=(384, 16))
    pts, bxs, msk, fe = lam.prepare_prompts(ep)
    assert pts is None and bxs is None and msk is not None and fe is ep["flag_examples"]
    ep = make_episode(1, 2, 1, 256, prompts="mixed", embeddings=(384, 16))
    pts, bxs, msk, _ = lam.prepare_prompts(ep)
    assert pts is not None and bxs is not None and msk is not None


def test_synthetic_weights_are_order_independent():
    a, b = _mae256(), _mae256()
    load_synth_weights(a, seed=3)
    load_synth_weights(b.prompt_encoder, seed=3)   # different traversal: only a sub-module, keys lose their prefix
    ka = a.state_dict()["prompt_encoder.no_mask_embed.weight"]
    load_synth_weights(b, seed=3)
    assert torch.equal(ka, b.state_dict()["prompt_encoder.no_mask_embed.weight"])
