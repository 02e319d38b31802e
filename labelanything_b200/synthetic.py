"""Deterministic synthetic weights and episodes shared by bench.py, the tests and the
golden-fixture generator.  No checkpoints or datasets are reachable offline, so every measurement and parity
check runs on these: weights are a pure function of (parameter name, shape, seed) — independent of module
construction order — and episodes are seeded CPU-generator draws.
"""
from __future__ import annotations

import zlib
from typing import Dict, Mapping, Optional, Tuple

import torch


def _gen(key: str, seed: int) -> torch.Generator:
    return torch.Generator().manual_seed((zlib.crc32(key.encode()) ^ (seed * 0x9E3779B1)) & 0x7FFFFFFF)


def synth_tensor(key: str, shape: Tuple[int, ...], seed: int = 0) -> torch.Tensor:
    """Synthetic value for one named parameter/buffer (fp32, CPU)."""
    g = _gen(key, seed)
    leaf = key.rsplit(".", 1)[-1]
    r = torch.randn(shape, generator=g, dtype=torch.float32)
    if leaf == "positional_encoding_gaussian_matrix":
        return r  # N(0,1) like the reference buffer (prompt_encoder.py:196-199)
    if leaf in ("pos_embed", "position_embeddings", "cls_token", "pos_embedding"):
        return 0.02 * r
    if leaf in ("rel_pos_h", "rel_pos_w"):
        return 0.1 * r  # zero-init in the reference; non-zero here so the rel-pos path is exercised
    if leaf == "bias":
        return 0.02 * r
    if leaf == "weight":
        if len(shape) == 1:
            return 1.0 + 0.1 * r  # norm scales
        if len(shape) == 2 and shape[0] == 1:
            return 0.5 * r  # nn.Embedding(1, D) tokens
        fan_in = 1
        for d in shape[1:]:
            fan_in *= d
        if ".output_upscaling." in key and len(shape) == 4:
            fan_in = shape[0]  # ConvTranspose2d weight is [Cin, Cout, k, k]
        return r / max(fan_in, 1) ** 0.5
    return 0.02 * r


def synth_state_dict(shapes: Mapping[str, Tuple[int, ...]], seed: int = 0) -> Dict[str, torch.Tensor]:
    return {k: synth_tensor(k, tuple(s), seed) for k, s in shapes.items()}


def load_synth_weights(module: torch.nn.Module, seed: int = 0) -> None:
    """Overwrite every floating-point parameter/buffer of `module` with its synthetic value."""
    sd = module.state_dict()
    new = {}
    for k, v in sd.items():
        new[k] = synth_tensor(k, tuple(v.shape), seed).to(v.dtype) if v.is_floating_point() else v
    module.load_state_dict(new)


def make_episode(batch: int, n_ways: int, k_shots: int, image_size: int, *, seed: int = 0,
                 prompts: str = "mask", n_points: int = 5, n_boxes: int = 2,
                 embeddings: Optional[Tuple[int, int]] = None, diagonal: bool = False) -> Dict[str, torch.Tensor]:
    """A seeded synthetic batch in the reference's `batched_input` format (label_anything/data/utils.py:43-58).

    M = n_ways * k_shots support images, C = n_ways + 1 classes (background first).
    prompts: "mask" (validation default: one rectangle mask per (m, c); points/boxes flagged off) or
             "mixed" (masks + n_points points + n_boxes boxes).
    embeddings: (channels, hw) -> provide precomputed `embeddings` instead of `images`.
    diagonal: flag_examples marks each support image positive for background + one class only.
    """
    g = torch.Generator().manual_seed(1000 + seed)
    B, M, C, S = batch, n_ways * k_shots, n_ways + 1, image_size
    out: Dict[str, torch.Tensor] = {}
    if embeddings is None:
        out["images"] = torch.randn(B, M + 1, 3, S, S, generator=g)
    else:
        ce, hw = embeddings
        out["embeddings"] = torch.randn(B, M + 1, ce, hw, hw, generator=g)
    # one random axis-aligned rectangle per (b, m, c) on the 256x256 prompt-mask grid
    x0 = torch.randint(0, 128, (B, M, C), generator=g)
    y0 = torch.randint(0, 128, (B, M, C), generator=g)
    ww = torch.randint(16, 128, (B, M, C), generator=g)
    hh = torch.randint(16, 128, (B, M, C), generator=g)
    ys = torch.arange(256).view(1, 1, 1, 256, 1)
    xs = torch.arange(256).view(1, 1, 1, 1, 256)
    masks = ((ys >= y0[..., None, None]) & (ys < (y0 + hh)[..., None, None]) &
             (xs >= x0[..., None, None]) & (xs < (x0 + ww)[..., None, None])).float()
    out["prompt_masks"] = masks
    out["flag_masks"] = torch.ones(B, M, C, dtype=torch.uint8)
    if prompts == "mixed":
        P, Bx = n_points, n_boxes
        out["prompt_points"] = torch.rand(B, M, C, P, 2, generator=g) * S
        out["flag_points"] = (torch.randint(0, 2, (B, M, C, P), generator=g) * 2 - 1).float()
        xy = torch.rand(B, M, C, Bx, 2, generator=g) * (S / 2)
        out["prompt_bboxes"] = torch.cat([xy, xy + S / 4], dim=-1)
        out["flag_bboxes"] = torch.ones(B, M, C, Bx)
    else:
        out["prompt_points"] = torch.zeros(B, M, C, 1, 2)
        out["flag_points"] = torch.zeros(B, M, C, 1)
        out["prompt_bboxes"] = torch.zeros(B, M, C, 1, 4)
        out["flag_bboxes"] = torch.zeros(B, M, C, 1)
    if diagonal:
        fe = torch.zeros(B, M, C, dtype=torch.uint8)
        fe[:, :, 0] = 1
        for m in range(M):
            fe[:, m, 1 + (m // k_shots) % n_ways] = 1
        out["flag_examples"] = fe
    else:
        out["flag_examples"] = torch.ones(B, M, C, dtype=torch.uint8)
    out["dims"] = torch.full((B, M + 1, 2), S, dtype=torch.int64)
    return out
