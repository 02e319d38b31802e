"""Generate the committed golden fixtures under tests/golden/ from the UNMODIFIED reference.

TEST INFRASTRUCTURE.  Runs only in the build container (needs /root/reference); the fixtures it writes are
what travels.  Usage:  python oracle/make_golden.py [tiny] [mae256] [samvit] [sam512] [mael256] [sam20w] [samragged] [samneck]
        [preprocess] [points] [metrics] [loss] [train]

Every fixture stores: the oracle `cfg`, the inputs, the reference outputs and either the full state dict
(tiny models) or the synthetic-weight seed (real-size models; weights are a pure function of
(name, shape, seed), see labelanything_b200/synthetic.py).  Version info of the generating stack is recorded.
"""
from __future__ import annotations

import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "oracle"))

import torch

import ref_import
from labelanything_b200.synthetic import load_synth_weights, make_episode

GOLD = ROOT / "tests" / "golden"
GOLD.mkdir(parents=True, exist_ok=True)


def _meta():
    import transformers

    return {"torch": torch.__version__, "transformers": transformers.__version__,
            "reference_commit": "6bb2c5a", "generator": "oracle/make_golden.py"}


def _pin_rows(lam, rows):
    ce = lam.prompt_encoder.class_encoder
    if hasattr(ce, "sample_rows"):
        ce.sample_rows = lambda C, device=None: rows[:C]


def _tiny_episode(B, M, C, S, mask_hw, n_points, n_boxes, seed, with_points=True, with_boxes=True, dims=None):
    g = torch.Generator().manual_seed(seed)
    ep = {"images": torch.randn(B, M + 1, 3, S, S, generator=g),
          "prompt_masks": (torch.rand(B, M, C, mask_hw, mask_hw, generator=g) > 0.5).float(),
          "flag_masks": (torch.rand(B, M, C, generator=g) > 0.2).to(torch.uint8),
          "flag_examples": (torch.rand(B, M, C, generator=g) > 0.3).to(torch.uint8)}
    ep["flag_examples"][:, :, 0] = 1
    if with_points:
        ep["prompt_points"] = torch.rand(B, M, C, n_points, 2, generator=g) * S
        ep["flag_points"] = torch.randint(-1, 2, (B, M, C, n_points), generator=g).float()
        ep["flag_points"][0, 0, 0, 0] = 1
    if with_boxes:
        xy = torch.rand(B, M, C, n_boxes, 2, generator=g) * S / 2
        ep["prompt_bboxes"] = torch.cat([xy, xy + S / 4], dim=-1)
        ep["flag_bboxes"] = torch.randint(-1, 2, (B, M, C, n_boxes), generator=g).float()
        ep["flag_bboxes"][0, 0, 0, 0] = 1
    ep["dims"] = dims if dims is not None else torch.full((B, M + 1, 2), S, dtype=torch.int64)
    return ep


def tiny_sam(models):
    """Tiny SAM-style Lam: windowed (pad 4->6) + global blocks with rel-pos, neck, all prompt types,
    RandomMatrixEncoder with pinned rows, example_attention + class_attention + class_example_attention,
    non-square original sizes with custom_preprocess."""
    from label_anything.models.build_lam import build_mask_decoder
    from label_anything.models.common import LayerNorm2d

    torch.manual_seed(0)
    S, D, Ce = 64, 32, 48
    vit = models.ImageEncoderViT(img_size=S, patch_size=16, embed_dim=Ce, depth=3, num_heads=3, use_rel_pos=True,
                                 window_size=3, global_attn_indexes=(1,), out_chans=D, project_last_hidden=False,
                                 norm_layer=lambda d: torch.nn.LayerNorm(d, eps=1e-6))
    neck = torch.nn.Sequential(torch.nn.Conv2d(Ce, D, 1, bias=False), LayerNorm2d(D),
                               torch.nn.Conv2d(D, D, 3, padding=1, bias=False), LayerNorm2d(D))
    pe = models.PromptImageEncoder(embed_dim=D, image_embedding_size=(4, 4), input_image_size=(S, S),
                                   mask_in_chans=16, class_attention=True, example_attention=True,
                                   example_class_attention=True,
                                   transformer=models.TwoWayTransformer(depth=2, embedding_dim=D, mlp_dim=64,
                                                                        num_heads=8),
                                   class_encoder=models.RandomMatrixEncoder(bank_size=10, embed_dim=D))
    dec = build_mask_decoder(embed_dim=D, decoder_attention_downsample_rate=2, spatial_convs=3)
    lam = models.Lam(image_encoder=vit, prompt_encoder=pe, mask_decoder=dec, neck=neck, image_size=S,
                     custom_preprocess=True).eval()
    load_synth_weights(lam, seed=11)
    rows = torch.tensor([0, 4, 2, 7])
    _pin_rows(lam, rows)
    B, M, C = 2, 2, 3
    dims = torch.tensor([[[50, 64], [64, 64], [64, 64]], [[64, 40], [64, 64], [64, 64]]], dtype=torch.int64)
    ep = _tiny_episode(B, M, C, S, 16, 2, 2, seed=5, dims=dims)
    ep["flag_gts"] = torch.tensor([[True, True, False], [True, True, True]])
    with torch.no_grad():
        out = lam(ep)
        enc = vit(ep["images"].flatten(0, 1))
    cfg = {"image_size": S, "image_embedding_size": (4, 4), "has_neck": True, "spatial_convs": 3,
           "class_attention": True, "example_attention": True, "example_class_attention": True,
           "custom_preprocess": True,
           "encoder": {"kind": "sam", "num_heads": 3, "depth": 3, "global_attn": [1], "window": 3}}
    torch.save({"meta": _meta(), "cfg": cfg, "state_dict": lam.state_dict(), "episode": ep, "class_rows": rows,
                "logits": out["logits"], "class_examples_embeddings": out["class_examples_embeddings"],
                "encoder_out": enc}, GOLD / "tiny_sam_lam.pt")
    print("tiny_sam_lam.pt", out["logits"].shape)

    # variant: masks only (points/boxes flags all zero -> dropped), no class encoder, no custom preprocess
    torch.manual_seed(1)
    pe2 = models.PromptImageEncoder(embed_dim=D, image_embedding_size=(4, 4), input_image_size=(S, S),
                                    mask_in_chans=16, class_attention=False, example_attention=False,
                                    example_class_attention=True,
                                    transformer=models.TwoWayTransformer(depth=2, embedding_dim=D, mlp_dim=64,
                                                                         num_heads=8),
                                    class_encoder=lambda x, y: (x, y))
    lam2 = models.Lam(image_encoder=vit, prompt_encoder=pe2, mask_decoder=dec, neck=neck, image_size=S,
                      custom_preprocess=False).eval()
    load_synth_weights(lam2, seed=12)
    ep2 = _tiny_episode(B, M, C, S, 16, 1, 1, seed=6)
    ep2["flag_points"].zero_()
    ep2["flag_bboxes"].zero_()
    with torch.no_grad():
        out2 = lam2(ep2)
    cfg2 = dict(cfg, class_attention=False, example_attention=False, custom_preprocess=False)
    torch.save({"meta": _meta(), "cfg": cfg2, "state_dict": lam2.state_dict(), "episode": ep2, "class_rows": None,
                "logits": out2["logits"], "class_examples_embeddings": out2["class_examples_embeddings"]},
               GOLD / "tiny_sam_lam_masks_only.pt")
    print("tiny_sam_lam_masks_only.pt", out2["logits"].shape)


def _hf_wrapper(hidden, layers, heads, inter, pretrain_size):
    from label_anything.models.build_encoder import ViTModelWrapper
    from transformers import ViTConfig

    return ViTModelWrapper(ViTConfig(hidden_size=hidden, num_hidden_layers=layers, num_attention_heads=heads,
                                     intermediate_size=inter, image_size=pretrain_size, patch_size=16))


def tiny_mae(models):
    """Tiny HF-ViT Lam: bicubic pos-emb resize (2x2 -> 3x3 grid), CLS dropped, 16x16 prompt masks -> 4x4 -> bilinear 3x3."""
    from label_anything.models.build_lam import _build_lam

    torch.manual_seed(0)
    S, D, Ce = 48, 32, 48
    lam = _build_lam(build_vit=lambda project_last_hidden: _hf_wrapper(Ce, 2, 3, 96, 32), image_embed_dim=Ce,
                     embed_dim=D, image_size=S, spatial_convs=3, class_attention=False, example_attention=False,
                     example_class_attention=True, custom_preprocess=False).eval()
    load_synth_weights(lam, seed=21)
    B, M, C = 2, 1, 2
    ep = _tiny_episode(B, M, C, S, 16, 1, 1, seed=8, with_points=False, with_boxes=False)
    with torch.no_grad():
        out = lam(ep)
        enc = lam.image_encoder(ep["images"].flatten(0, 1))
    cfg = {"image_size": S, "image_embedding_size": (3, 3), "has_neck": True, "spatial_convs": 3,
           "class_attention": False, "example_attention": False, "example_class_attention": True,
           "custom_preprocess": False, "encoder": {"kind": "hf", "num_heads": 3, "depth": 2}}
    torch.save({"meta": _meta(), "cfg": cfg, "state_dict": lam.state_dict(), "episode": ep, "class_rows": None,
                "logits": out["logits"], "class_examples_embeddings": out["class_examples_embeddings"],
                "encoder_out": enc}, GOLD / "tiny_mae_lam.pt")
    print("tiny_mae_lam.pt", out["logits"].shape)


MAE256_CFG = {"image_size": 480, "image_embedding_size": (30, 30), "has_neck": True, "spatial_convs": 3,
              "class_attention": False, "example_attention": False, "example_class_attention": True,
              "custom_preprocess": False, "encoder": {"kind": "hf", "num_heads": 12, "depth": 12}}


def mae256(models):
    """BASELINE config 1: MAE-256 (HF ViT-B, 480 px), 1-way 1-shot, B=1 — the reference's CPU-runnable case
    (parameters/trainval/coco20i/mae_noembs.yaml:42-55).  Weights = synthetic seed 0; rows pinned."""
    from label_anything.models.build_lam import _build_lam

    lam = _build_lam(build_vit=lambda project_last_hidden: _hf_wrapper(768, 12, 12, 3072, 224),
                     image_embed_dim=768, embed_dim=256, image_size=480, spatial_convs=3, class_attention=False,
                     example_attention=False, example_class_attention=True,
                     class_encoder={"name": "RandomMatrixEncoder", "bank_size": 100, "embed_dim": 256},
                     custom_preprocess=False).eval()
    load_synth_weights(lam, seed=0)
    rows = torch.arange(2)
    _pin_rows(lam, rows)
    ep = make_episode(1, 1, 1, 480, seed=0)
    t0 = time.time()
    with torch.no_grad():
        out = lam(ep)
        enc = lam.image_encoder(ep["images"][0, :1])
    print(f"mae256 reference forward {time.time() - t0:.1f}s")
    torch.save({"meta": _meta(), "cfg": MAE256_CFG, "weights_seed": 0, "episode_args": dict(batch=1, n_ways=1, k_shots=1, image_size=480, seed=0),
                "class_rows": rows, "logits_sub3": out["logits"][..., ::3, ::3].clone(),
                "class_examples_embeddings": out["class_examples_embeddings"],
                "encoder_out_sub": enc[0, ::16].clone(),
                "shapes": {k: tuple(v.shape) for k, v in lam.state_dict().items()}}, GOLD / "mae256_1w1s.pt")
    print("mae256_1w1s.pt", out["logits"].shape)


SAM512_CFG = {"image_size": 1024, "image_embedding_size": (64, 64), "has_neck": True, "spatial_convs": 3,
              "class_attention": False, "example_attention": True, "example_class_attention": False,
              "custom_preprocess": True,
              "encoder": {"kind": "sam", "num_heads": 12, "depth": 12, "global_attn": [2, 5, 8, 11], "window": 14}}


def _sam512(models):
    from label_anything.models.build_lam import _build_lam
    from label_anything.models.build_encoder import build_vit_b

    lam = _build_lam(build_vit=build_vit_b, image_embed_dim=768, embed_dim=512, image_size=1024,
                     use_vit_sam_neck=False, spatial_convs=3, class_attention=False, example_attention=True,
                     example_class_attention=False,
                     class_encoder={"name": "RandomMatrixEncoder", "bank_size": 100, "embed_dim": 512},
                     custom_preprocess=True).eval()
    load_synth_weights(lam, seed=0)
    return lam


def samvit(models):
    """SAM ViT-B (1024 px) encoder alone on one synthetic image, synthetic weights seed 0
    (parameters/trainval/other/COCO_vit.yaml:47-63 encoder)."""
    lam = _sam512(models)
    ep = make_episode(1, 1, 1, 1024, seed=0)
    img = ep["images"][0, :1]
    t0 = time.time()
    with torch.no_grad():
        enc = lam.image_encoder(img)
        feat = lam.neck(enc)
    print(f"samvit reference forward {time.time() - t0:.1f}s")
    torch.save({"meta": _meta(), "cfg": SAM512_CFG, "weights_seed": 0,
                "episode_args": dict(batch=1, n_ways=1, k_shots=1, image_size=1024, seed=0),
                "encoder_out_sub": enc[0, ::16].clone(), "neck_out_sub": feat[0, ::16].clone(),
                "shapes": {k: tuple(v.shape) for k, v in lam.state_dict().items()}}, GOLD / "sam512_vit_1img.pt")
    print("sam512_vit_1img.pt", enc.shape)


def sam512(models):
    """BASELINE config 3 shape at B=1: SAM-512 5-way 5-shot (26 images, 150 sequences).  Encoder fed 2 images
    per call and the result handed over through the `embeddings` key (identical arithmetic, lam.py:139-146)."""
    lam = _sam512(models)
    rows = torch.arange(6)
    _pin_rows(lam, rows)
    ep = make_episode(1, 5, 5, 1024, seed=0)
    imgs = ep.pop("images").flatten(0, 1)
    t0 = time.time()
    with torch.no_grad():
        feats = torch.cat([lam.image_encoder(imgs[i:i + 2]) for i in range(0, imgs.shape[0], 2)])
        ep["embeddings"] = feats.unsqueeze(0)
        out = lam(ep)
    print(f"sam512 5w5s reference forward {time.time() - t0:.1f}s")
    torch.save({"meta": _meta(), "cfg": SAM512_CFG, "weights_seed": 0,
                "episode_args": dict(batch=1, n_ways=5, k_shots=5, image_size=1024, seed=0), "class_rows": rows,
                "logits_sub8": out["logits"][..., ::8, ::8].clone(),
                "class_examples_embeddings": out["class_examples_embeddings"]}, GOLD / "sam512_5w5s.pt")
    print("sam512_5w5s.pt", out["logits"].shape)


MAEL256_CFG = {"image_size": 480, "image_embedding_size": (30, 30), "has_neck": True, "spatial_convs": 3,
               "class_attention": False, "example_attention": False, "example_class_attention": True,
               "custom_preprocess": False, "encoder": {"kind": "hf", "num_heads": 16, "depth": 24}}


def mael256(models):
    """BASELINE config 4's model forward: MAE-L-256 (HF ViT-L 1024/24/16/4096, 480 px, embed 256, NO class encoder;
    parameters/trainval/coco/mael.yaml:42-51), 2-way 5-shot, B=1 through the `images` key, plus the ViT-L encoder output
    of the query image."""
    from label_anything.models.build_lam import _build_lam

    lam = _build_lam(build_vit=lambda project_last_hidden: _hf_wrapper(1024, 24, 16, 4096, 224),
                     image_embed_dim=1024, embed_dim=256, image_size=480, spatial_convs=3, class_attention=False,
                     example_attention=False, example_class_attention=True, custom_preprocess=False).eval()
    load_synth_weights(lam, seed=0)
    ep = make_episode(1, 2, 5, 480, seed=4)
    t0 = time.time()
    with torch.no_grad():
        out = lam(ep)
        enc = lam.image_encoder(ep["images"][0, :1])
    print(f"mael256 reference forward {time.time() - t0:.1f}s")
    torch.save({"meta": _meta(), "cfg": MAEL256_CFG, "weights_seed": 0,
                "episode_args": dict(batch=1, n_ways=2, k_shots=5, image_size=480, seed=4), "class_rows": None,
                "logits_sub3": out["logits"][..., ::3, ::3].clone(),
                "class_examples_embeddings": out["class_examples_embeddings"],
                "encoder_out_sub": enc[0, ::16].clone(),
                "shapes": {k: tuple(v.shape) for k, v in lam.state_dict().items()}}, GOLD / "mael256_2w5s.pt")
    print("mael256_2w5s.pt", out["logits"].shape)


def sam20w(models):
    """BASELINE config 5's episode shape at B=1 -- 20-way 5-shot: M = 100 support images, C = 21 classes, S = 2100
    prompt sequences -- through the `embeddings` key of the SAM-512 head (neck 768 -> 512, example_attention,
    RandomMatrixEncoder with 21 pinned rows).  The reference materialises S x D x T fp32 tensors (17.6 GB each at
    T = 4096): the fixture uses a 512-px model (T = 32 x 32), which the reference fits in host memory;
    the native test lowers max_rows_per_pass so that the prompt encoder's chunking is exercised all the same."""
    from label_anything.models.build_lam import build_lam_no_vit

    lam = build_lam_no_vit(image_embed_dim=768, embed_dim=512, image_size=512, spatial_convs=3, class_attention=False,
                           example_attention=True, example_class_attention=False,
                           class_encoder={"name": "RandomMatrixEncoder", "bank_size": 100, "embed_dim": 512},
                           custom_preprocess=True).eval()
    load_synth_weights(lam, seed=0)
    rows = torch.arange(21)
    _pin_rows(lam, rows)
    ep = make_episode(1, 20, 5, 512, seed=5, embeddings=(768, 32))
    t0 = time.time()
    with torch.no_grad():
        out = lam(ep)
    print(f"sam 20-way 5-shot reference forward {time.time() - t0:.1f}s")
    cfg = dict(SAM512_CFG, image_size=512, image_embedding_size=(32, 32))
    cfg.pop("encoder")
    torch.save({"meta": _meta(), "cfg": cfg, "weights_seed": 0,
                "model_args": dict(image_embed_dim=768, embed_dim=512, image_size=512, spatial_convs=3,
                                   class_attention=False, example_attention=True, example_class_attention=False,
                                   class_encoder={"name": "RandomMatrixEncoder", "bank_size": 100, "embed_dim": 512},
                                   custom_preprocess=True),
                "episode_args": dict(batch=1, n_ways=20, k_shots=5, image_size=512, seed=5, embeddings=(768, 32)),
                "class_rows": rows, "logits_sub4": out["logits"][..., ::4, ::4].clone(),
                "class_examples_embeddings": out["class_examples_embeddings"]}, GOLD / "sam512head_20w5s.pt")
    print("sam512head_20w5s.pt", out["logits"].shape)


def samragged(models):
    """SAM-512, 1-way 1-shot, B = 2 through the `images` key with RAGGED original sizes (custom_preprocess crop +
    per-item resize + -inf padding to the batch maximum, lam.py:383-453)."""
    lam = _sam512(models)
    rows = torch.arange(2)
    _pin_rows(lam, rows)
    ep = make_episode(2, 1, 1, 1024, seed=6)
    ep["dims"] = torch.tensor([[[683, 1024], [1024, 1024]], [[960, 771], [500, 375]]], dtype=torch.int64)
    t0 = time.time()
    with torch.no_grad():
        out = lam(ep)
    print(f"sam ragged B=2 reference forward {time.time() - t0:.1f}s")
    torch.save({"meta": _meta(), "cfg": SAM512_CFG, "weights_seed": 0,
                "episode_args": dict(batch=2, n_ways=1, k_shots=1, image_size=1024, seed=6), "dims": ep["dims"],
                "class_rows": rows, "logits_sub4": out["logits"][..., ::4, ::4].clone(),
                "logits_shape": tuple(out["logits"].shape),
                "class_examples_embeddings": out["class_examples_embeddings"]}, GOLD / "sam512_b2_ragged.pt")
    print("sam512_b2_ragged.pt", out["logits"].shape)


def samneck(models):
    """SAM ViT-B with its own neck (project_last_hidden=True, out_chans 256) and return_last_block_state=True
    (image_encoder.py:110-131; used by preprocess.py:160-162)."""
    from label_anything.models.build_encoder import build_vit_b

    vit = build_vit_b(project_last_hidden=True).eval()
    load_synth_weights(vit, seed=0)
    img = make_episode(1, 1, 1, 1024, seed=0)["images"][0, :1]
    t0 = time.time()
    with torch.no_grad():
        out = vit(img, return_last_block_state=True)
        plain = vit(img)
    print(f"sam vit + neck reference forward {time.time() - t0:.1f}s", {k: tuple(v.shape) for k, v in out.items()})
    assert torch.equal(plain, out["last_hidden_state"])
    torch.save({"meta": _meta(), "weights_seed": 0,
                "episode_args": dict(batch=1, n_ways=1, k_shots=1, image_size=1024, seed=0),
                "keys": sorted(str(getattr(k, "value", k)) for k in out.keys()), "last_hidden_state_sub": out["last_hidden_state"][0, ::8].clone(),
                "last_block_state_sub": out["last_block_state"][0, ::16].clone(),
                "shapes": {k: tuple(v.shape) for k, v in vit.state_dict().items()}}, GOLD / "sam_vit_neck_1img.pt")
    print("sam_vit_neck_1img.pt")


def preprocess_f3(models):
    """Input preprocessing through the reference's OWN classes: CustomResize + ToTensor + CustomNormalize
    (and the non-custom Resize + ToTensor + Normalize) on PIL images, PromptsProcessor.apply_masks / apply_coords /
    apply_boxes.  Inputs are seeded uint8 arrays; outputs are stored whole or strided."""
    import numpy as np
    from PIL import Image
    from torchvision.transforms import Compose, Resize, ToTensor

    from label_anything.data.transforms import CustomNormalize, CustomResize, Normalize, PromptsProcessor

    rng = np.random.default_rng(12)
    mean, std = [0.485, 0.456, 0.406], [0.229, 0.224, 0.225]
    images = []
    for (h, w, size, custom, stride) in [(300, 200, 256, True, 1), (123, 457, 256, True, 1), (256, 256, 256, True, 1),
                                         (97, 64, 128, False, 1), (420, 630, 384, True, 1), (120, 160, 512, True, 2)]:
        img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        # smooth half of the image so the fixture is not pure noise (gradients exercise the rounding differently)
        yy, xx = np.mgrid[0:h, 0:w]
        img[: h // 2] = np.stack([(xx[: h // 2] * 255 // max(w - 1, 1)), (yy[: h // 2] * 255 // max(h - 1, 1)),
                                  ((xx[: h // 2] + yy[: h // 2]) % 256)], axis=-1).astype(np.uint8)
        tf = (Compose([CustomResize(size), ToTensor(), CustomNormalize(size, mean, std)]) if custom
              else Compose([Resize(size=(size, size)), ToTensor(), Normalize(mean, std)]))
        out = tf(Image.fromarray(img))
        images.append({"image": torch.from_numpy(img), "size": size, "custom_preprocess": custom, "stride": stride,
                       "out": out[:, ::stride, ::stride].clone(), "out_shape": tuple(out.shape),
                       "out_sum": out.double().sum().item()})
    prompts = []
    for (n, h, w, custom) in [(3, 300, 200, True), (1, 123, 457, True), (2, 640, 480, False), (0, 50, 60, True),
                              (2, 420, 630, True)]:
        pp = PromptsProcessor(long_side_length=1024, masks_side_length=256, custom_preprocess=custom)
        masks = np.zeros((n, h, w), dtype=np.uint8)
        for i in range(n):
            y0, x0 = rng.integers(0, h // 2), rng.integers(0, w // 2)
            masks[i, y0:y0 + rng.integers(1, h // 2), x0:x0 + rng.integers(1, w // 2)] = 1
            masks[i] |= (rng.random((h, w)) > 0.97).astype(np.uint8)                 # isolated pixels: nearest sampling
        out_mask = torch.as_tensor(np.asarray(pp.apply_masks(masks if n else np.array([]))))
        pts = rng.random((4, 2)) * np.array([w, h])
        boxes = np.concatenate([rng.random((3, 2)) * np.array([w, h]) / 2, rng.random((3, 2)) * np.array([w, h]) / 2 + np.array([w, h]) / 2], axis=1)
        prompts.append({"masks": torch.from_numpy(masks), "custom_preprocess": custom, "mask_out": out_mask.reshape(256, 256),
                        "points": torch.from_numpy(pts), "points_out": torch.from_numpy(pp.apply_coords(pts, (h, w))),
                        "boxes": torch.from_numpy(boxes), "boxes_out": torch.from_numpy(pp.apply_boxes(boxes, (h, w))),
                        "original_size": (h, w)})
    import PIL
    import torchvision

    meta = dict(_meta(), pillow=PIL.__version__, torchvision=torchvision.__version__)
    torch.save({"meta": meta, "mean": mean, "std": std, "images": images, "prompts": prompts}, GOLD / "preprocess_f3.pt")
    print("preprocess_f3.pt", [tuple(i["out"].shape) for i in images], [tuple(p["mask_out"].shape) for p in prompts])


def points_f4(models):
    """Iterative prompting: the reference's own generate_points_from_errors with torch.randint replaced,
    for the duration of the call, by a recorded stand-in (index = fixed draw mod count) -- the same device the rows of
    RandomMatrixEncoder are pinned with.  B = 1 and C <= B cases only: for C > B the reference's `argsort(b * B + c)` has
    duplicate keys and an unstable sort decides the row order."""
    import label_anything.experiment.substitution as sub

    g = torch.Generator().manual_seed(6)
    cases = []
    for B, C, H, W, n in [(1, 4, 37, 53, 1), (1, 6, 64, 48, 1), (3, 3, 40, 40, 1), (1, 3, 32, 32, 1), (1, 5, 24, 24, 1)]:
        logits = torch.randn(B, C, H, W, generator=g)
        gt = torch.randint(0, C, (B, H, W), generator=g)
        gt[:, -3:, :] = -100
        if C == 3 and B == 1:                      # a class with no error at all -> the padding row
            gt[gt == 2] = 0
            logits[:, 2] = -50.0
        if C == 5:                                 # perfect prediction: the early-return branch
            gt = logits.argmax(dim=1)
        draws = torch.randint(0, 2 ** 31 - 1, (B, C, n), generator=g)
        order = iter(draws.flatten().tolist())

        def fake_randint(low, high, size, device=None, _order=None):
            raise RuntimeError

        real = torch.randint
        # groups are visited in sorted (b, c) order; classes without errors are skipped by the reference
        pred = logits.argmax(dim=1)
        t = gt.clone()
        t[t == -100] = 0
        has_err = [[bool(((t[b] == c).long() - (pred[b] == c).long()).abs().sum() > 0) for c in range(C)] for b in range(B)]
        seq = [int(draws[b, c, i]) for b in range(B) for c in range(C) if has_err[b][c] for i in range(n)]
        it = iter(seq)

        def pinned_randint(low, high, size, device=None, **kw):
            return torch.tensor([next(it) % int(high) for _ in range(size[0])], dtype=torch.int64)

        torch.randint = pinned_randint
        try:
            pts, labels = sub.generate_points_from_errors(logits, gt, n)
        finally:
            torch.randint = real
        cases.append({"logits": logits, "gt": gt, "rand": draws, "points": pts.float(), "labels": labels.float(),
                      "num_points": n})
    torch.save({"meta": _meta(), "cases": cases}, GOLD / "points_f4.pt")
    print("points_f4.pt", [tuple(c["points"].shape) for c in cases])


def metrics_f4(models):
    """Post-logits step: torch.argmax + the reference's own to_global_multiclass on seeded
    inputs with ties, -inf planes, NaNs and ignore_index targets; the confusion matrix is torch.bincount of the
    reference's global labels (torchmetrics itself is not installed: its reduce stays unpinned)."""
    from label_anything.data.utils import to_global_multiclass

    g = torch.Generator().manual_seed(4)
    cases = []
    categories = {k: {"name": str(k)} for k in (1, 2, 3, 5, 7, 8, 11, 13, 17, 20)}       # 10 categories -> G = 11
    for B, C, H, W, classes in [
        (3, 5, 24, 20, [[[7, 3], [3, 13]], [[1, 2, 3, 5]], [[20, 17], [17], [8]]]),       # chained substitutions
        (2, 3, 7, 9, [[[2]], [[11, 5]]]),                                                  # odd sizes (scalar path)
        (1, 6, 32, 32, [[[1, 2, 3, 5, 7]]]),
    ]:
        logits = torch.randn(B, C, H, W, generator=g)
        logits[:, :, : H // 3] = torch.round(logits[:, :, : H // 3])                      # ties
        logits[0, C - 1] = float("-inf")                                                   # absent class
        logits[0, :, -1, -1] = float("-inf")                                               # all -inf -> 0
        logits[-1, 1, 0, :3] = float("nan")
        gt = torch.randint(0, C, (B, H, W), generator=g)
        gt[:, -2:, :] = -100
        preds = logits.argmax(dim=1)
        glob_preds, glob_gt = to_global_multiclass(classes, categories, preds, gt)
        G = len(categories) + 1
        keep = glob_gt != -100
        conf = torch.bincount(glob_gt[keep] * G + glob_preds[keep], minlength=G * G).reshape(G, G)
        cases.append({"logits": logits, "gt": gt, "classes": classes, "preds": preds, "glob_preds": glob_preds,
                      "glob_gt": glob_gt, "confmat": conf, "num_classes": G})
    torch.save({"meta": _meta(), "categories": categories, "cases": cases}, GOLD / "metrics_f4.pt")
    print("metrics_f4.pt", [tuple(c["logits"].shape) for c in cases])


def loss_f1(models):
    """Focal loss + class weighting: values and autograd gradients of the UNMODIFIED
    label_anything.loss.LabelAnythingLoss / get_weight_matrix_from_labels on seeded inputs."""
    from label_anything.loss import LabelAnythingLoss
    from label_anything.loss.utils import get_weight_matrix_from_labels

    g = torch.Generator().manual_seed(9)
    cases = []
    for B, C, H, W, gamma, weighting, ignore, comp_w in [
        (2, 3, 7, 9, 2.0, True, True, 1.0),        # odd sizes (scalar path), MAE-L training classes
        (2, 6, 16, 12, 2.0, True, True, 0.5),      # 5-way + background, component weight != 1 (applied twice)
        (1, 6, 8, 8, 2.0, True, False, 1.0),       # no ignored pixels: the other branch of the weight function
        (2, 11, 8, 12, 1.5, True, True, 1.0),      # more planes than the register window, non-integer gamma
        (2, 4, 8, 8, 2.0, False, True, 1.0),       # no class weighting
    ]:
        logits = (torch.randn(B, C, H, W, generator=g) * 3).requires_grad_(True)
        target = torch.randint(0, C, (B, H, W), generator=g)
        target[target == C - 1] = 0 if C > 4 else C - 1                                    # an absent class for C > 4
        if ignore:
            target[:, -2:, :] = -100
        with torch.no_grad():
            logits[0, 1 if target[0, 0, 0] != 1 else 2, 0, 0] = float("-inf")              # padded-pixel style -inf
        loss = LabelAnythingLoss({"focal": {"weight": comp_w, "gamma": gamma}}, class_weighting=weighting)
        out = loss(logits, target)
        out["value"].backward()
        wt, cw = get_weight_matrix_from_labels(target, C)
        cases.append({"logits": logits.detach().clone(), "target": target, "gamma": gamma, "class_weighting": weighting,
                      "component_weight": comp_w, "value": out["value"].detach().clone(),
                      "component": out["components"]["focal"], "grad": logits.grad.clone(), "wtarget": wt,
                      "class_weights": cw})
    torch.save({"meta": _meta(), "cases": cases}, GOLD / "loss_f1.pt")
    print("loss_f1.pt", [(tuple(c["logits"].shape), float(c["value"])) for c in cases])


def scale_matrices(module, gain):
    """Multiply every weight matrix / convolution kernel (>= 2-D, more than one output row) by `gain`."""
    if gain != 1.0:
        with torch.no_grad():
            for p in module.parameters():
                if p.dim() >= 2 and p.shape[0] > 1:
                    p.mul_(gain)


def grad_sample(g, n=4096):
    """{"norm", "values"}: the whole tensor when it has at most n entries, else the n evenly strided entries at
    `sample_index(numel, n)` (tests/test_training_*.py compare the same positions)."""
    flat = g.reshape(-1)
    return {"norm": flat.double().norm().float(), "values": flat[sample_index(flat.numel(), n)].clone()}


def sample_index(numel, n=4096):
    return torch.arange(numel) if numel <= n else torch.linspace(0, numel - 1, n).long()


def train_f1(models):
    """Training step of the pre-computed-embeddings configuration: loss value and the autograd
    gradient of EVERY parameter of the unmodified reference `Lam` (neck + prompt encoder + mask decoder, `embeddings`
    key) under the unmodified `LabelAnythingLoss` (focal + class weighting, parameters/trainval/coco/mael.yaml:24-28).
    Case `mixed`: all prompt types, pinned RandomMatrixEncoder rows, all three merge attentions, mask resize, ragged
    original sizes with the un-pad crop, flag_gts.  Case `masks_only`: mael.yaml's shape of model (no class encoder,
    class_example_attention only), mask prompts only (one sparse token per sequence)."""
    from label_anything.loss import LabelAnythingLoss
    from label_anything.models.build_lam import build_lam_no_vit

    S, D, Ce, g = 128, 128, 64, 8
    cases = {}
    for name in ("mixed", "masks_only", "mixed_scaled"):
        mixed = name != "masks_only"
        # `mixed_scaled`: the mixed model with every weight matrix halved.  At gain 1 this random-weight model is
        # ill-conditioned (one bf16 rounding of the weights moves the fp32 gradients by 60 %); at gain 0.5 by 7 %, which
        # makes the bf16 training path's gradient error a meaningful number.
        gain = 0.5 if name == "mixed_scaled" else 1.0
        torch.manual_seed(3 if mixed else 4)
        build = dict(image_embed_dim=Ce, embed_dim=D, image_size=S, spatial_convs=3, class_attention=mixed,
                     example_attention=mixed, example_class_attention=True, custom_preprocess=mixed,
                     class_encoder={"name": "RandomMatrixEncoder", "bank_size": 10, "embed_dim": D} if mixed else None)
        lam = build_lam_no_vit(**build).train()          # the reference's own builder (models/build_lam.py:81-86)
        load_synth_weights(lam, seed=21 if mixed else 22)
        scale_matrices(lam, gain)
        rows = torch.tensor([0, 4, 2, 7]) if mixed else None
        if mixed:
            _pin_rows(lam, rows)
        B, M, C = 2, 2, 3
        dims = (torch.tensor([[[100, 128], [128, 128], [128, 128]], [[128, 80], [128, 128], [128, 128]]],
                             dtype=torch.int64) if mixed else None)
        ep = _tiny_episode(B, M, C, S, 48 if mixed else 32, 2, 1, seed=31 if mixed else 32, with_points=mixed,
                           with_boxes=mixed, dims=dims)
        del ep["images"]
        gen = torch.Generator().manual_seed(33)
        ep["embeddings"] = torch.randn(B, M + 1, Ce, g, g, generator=gen)
        if mixed:
            ep["flag_gts"] = torch.tensor([[True, True, False], [True, True, True]])
        out = lam(ep)
        logits = out["logits"]
        Hm, Wm = logits.shape[-2:]
        gt = torch.randint(0, C, (B, Hm, Wm), generator=gen)
        if mixed:
            gt[0][gt[0] == 2] = 1                       # class 2 is absent from episode 0 (flag_gts)
            for b in range(B):                          # padded pixels are ignored, as the data pipeline pads the gt
                oh, ow = (int(v) for v in dims[b, 0])
                gt[b, oh:, :] = -100
                gt[b, :, ow:] = -100
        loss_fn = LabelAnythingLoss({"focal": {"weight": 1.0, "gamma": 2.0}}, class_weighting=True)
        loss = loss_fn(out, gt)
        loss["value"].backward()
        # every gradient in full would be 30 MB per case: small tensors are stored whole, large ones as their norm
        # + 4096 evenly strided entries (`grad_sample`); the weights are a pure function of (name, shape, seed)
        grads = {k: (None if p.grad is None else grad_sample(p.grad.detach())) for k, p in lam.named_parameters()}
        # how much the reference's OWN fp32 gradients move when its weight matrices and the input embeddings are rounded
        # to bf16 once (nothing else changes): the scale of "bf16 rounding noise" for this model and episode, against
        # which the bf16 training path is judged (the fp32-accurate bf16x3 mode is judged against the gradients)
        full = {k: p.grad.detach().clone() for k, p in lam.named_parameters() if p.grad is not None}
        with torch.no_grad():
            for p in lam.parameters():
                if p.dim() >= 2:
                    p.copy_(p.to(torch.bfloat16).float())
        lam.zero_grad(set_to_none=True)
        ep_r = dict(ep, embeddings=ep["embeddings"].to(torch.bfloat16).float())
        loss_fn(lam(ep_r), gt)["value"].backward()
        rels, num, den = [], 0.0, 0.0
        for k, p in lam.named_parameters():
            if k in full and float(full[k].double().norm()) > 1e-6:
                d2 = float((p.grad.double() - full[k].double()).pow(2).sum())
                r2 = float(full[k].double().pow(2).sum())
                rels.append((d2 / r2) ** 0.5)
                num, den = num + d2, den + r2
        rels.sort()
        sens = {"total": (num / den) ** 0.5, "median": rels[len(rels) // 2], "p90": rels[int(0.9 * len(rels))],
                "worst": rels[-1]}
        print(f"train_f1[{name}]: gradient change under one bf16 rounding of weights + inputs: {sens}")
        # the reference's OWN mixed-precision mode (torch.autocast(bfloat16), what `accelerate` gives its training loop):
        # fp32 weights, bf16 matmul / conv operands in forward and backward -- the like-for-like yardstick of the native
        # bf16 path, measured against the same fp32 gradients
        load_synth_weights(lam, seed=21 if mixed else 22)
        scale_matrices(lam, gain)
        lam.zero_grad(set_to_none=True)
        with torch.autocast(device_type="cpu", dtype=torch.bfloat16):
            out_ac = lam(ep)
        loss_fn({"logits": out_ac["logits"].float()}, gt)["value"].backward()
        rels, num, den = [], 0.0, 0.0
        for k, p in lam.named_parameters():
            if k in full and float(full[k].double().norm()) > 1e-6 and p.grad is not None:
                d2 = float((p.grad.double() - full[k].double()).pow(2).sum())
                r2 = float(full[k].double().pow(2).sum())
                rels.append((d2 / r2) ** 0.5)
                num, den = num + d2, den + r2
        rels.sort()
        fin_ = torch.isfinite(logits)
        autocast = {"total": (num / den) ** 0.5, "median": rels[len(rels) // 2], "p90": rels[int(0.9 * len(rels))],
                    "worst": rels[-1],
                    "logits_mean_over_std": float((out_ac["logits"].float()[fin_] - logits[fin_]).abs().mean()
                                                  / logits[fin_].std())}
        print(f"train_f1[{name}]: gradient error of the reference under torch.autocast(bfloat16): {autocast}")
        cfg = {"image_size": S, "image_embedding_size": (g, g), "has_neck": True, "spatial_convs": 3,
               "class_attention": mixed, "example_attention": mixed, "example_class_attention": True,
               "custom_preprocess": mixed}
        cases[name] = {"bf16_sensitivity": sens, "bf16_autocast_error": autocast, "cfg": cfg, "weights_seed": 21 if mixed else 22, "weight_gain": gain, "episode": ep, "class_rows": rows, "gt": gt, "logits": logits.detach().clone(),
                       "loss": loss["value"].detach().clone(), "grads": grads,
                       "build": build}
        used = sum(v is not None for v in grads.values())
        print(f"train_f1[{name}]: loss {float(loss['value']):.6f}, {used}/{len(grads)} parameters with a gradient, "
              f"logits {tuple(logits.shape)}")
    torch.save({"meta": _meta(), "cases": cases}, GOLD / "train_f1.pt")


if __name__ == "__main__":
    which = sys.argv[1:] or ["tiny", "mae256", "samvit", "metrics", "loss"]
    models = ref_import.import_reference()
    torch.set_num_threads(8)
    if "tiny" in which:
        tiny_sam(models)
        tiny_mae(models)
    if "mae256" in which:
        mae256(models)
    if "samvit" in which:
        samvit(models)
    if "sam512" in which:
        sam512(models)
    if "mael256" in which:
        mael256(models)
    if "sam20w" in which:
        sam20w(models)
    if "samragged" in which:
        samragged(models)
    if "samneck" in which:
        samneck(models)
    if "preprocess" in which:
        preprocess_f3(models)
    if "points" in which:
        points_f4(models)
    if "metrics" in which:
        metrics_f4(models)
    if "loss" in which:
        loss_f1(models)
    if "train" in which:
        train_f1(models)
