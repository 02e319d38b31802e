"""CPU oracle for the LabelAnything hot path — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A functional, fp32, torch-CPU restatement of the reference algorithm
(pasqualedem/LabelAnything @ 6bb2c5a): image encoder -> neck -> prompt encoder -> mask decoder ->
postprocess.  It operates on a flat state dict that uses the reference's own parameter names
(`image_encoder.*`, `neck.*`, `prompt_encoder.*`, `mask_decoder.*`), so weights produced by the
reference (or by labelanything_b200's modules, which keep the same keys) can be fed to it unchanged.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this
module, and only as the checker / the timed CPU baseline.  The product path (labelanything_b200/) never
imports it.

Parity pin: tests/test_oracle_golden.py checks every function here against tensors produced by the
UNMODIFIED reference imported in the build container (oracle/make_golden.py -> tests/golden/*.pt).
The reference itself ships no tests or golden vectors, so those fixtures are the pin.

Each function cites the reference file:line it restates (paths relative to the reference repo root).
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
SD = Dict[str, Tensor]

POSITIVE, NULL, NEGATIVE = 1, 0, -1  # label_anything/data/utils.py:25-28 (Label)


# ----------------------------------------------------------------------------------------------
# small building blocks
# ----------------------------------------------------------------------------------------------
def linear(sd: SD, name: str, x: Tensor) -> Tensor:
    return F.linear(x, sd[name + ".weight"], sd.get(name + ".bias"))


def layer_norm(sd: SD, name: str, x: Tensor, eps: float) -> Tensor:
    return F.layer_norm(x, (x.shape[-1],), sd[name + ".weight"], sd[name + ".bias"], eps)


def layer_norm_2d(sd: SD, name: str, x: Tensor, eps: float = 1e-6) -> Tensor:
    """Channel LayerNorm on NCHW with biased variance.  label_anything/models/common.py:42-54"""
    mu = x.mean(dim=1, keepdim=True)
    var = ((x - mu) ** 2).mean(dim=1, keepdim=True)
    xn = (x - mu) / torch.sqrt(var + eps)
    return xn * sd[name + ".weight"].view(1, -1, 1, 1) + sd[name + ".bias"].view(1, -1, 1, 1)


def mlp_block(sd: SD, name: str, x: Tensor, act: str) -> Tensor:
    """lin1 -> act -> lin2.  label_anything/models/common.py:19-37"""
    h = linear(sd, name + ".lin1", x)
    h = F.gelu(h) if act == "gelu" else F.relu(h)
    return linear(sd, name + ".lin2", h)


def sam_attention(sd: SD, name: str, q: Tensor, k: Tensor, v: Tensor, num_heads: int) -> Tensor:
    """q/k/v projections, per-head softmax(QK^T/sqrt(dh))V, out projection.

    label_anything/models/common.py:97-148.  key_mask / attn_mask are no-ops in the reference
    (`score_mask` is created all-False and never filled, common.py:117-139), so none is applied here.
    """
    qp, kp, vp = linear(sd, name + ".q_proj", q), linear(sd, name + ".k_proj", k), linear(sd, name + ".v_proj", v)
    b, nq, ci = qp.shape
    dh = ci // num_heads
    qh = qp.view(b, nq, num_heads, dh).transpose(1, 2)
    kh = kp.view(b, -1, num_heads, dh).transpose(1, 2)
    vh = vp.view(b, -1, num_heads, dh).transpose(1, 2)
    att = torch.softmax(qh @ kh.transpose(-1, -2) / math.sqrt(dh), dim=-1)
    o = (att @ vh).transpose(1, 2).reshape(b, nq, ci)
    return linear(sd, name + ".out_proj", o)


def attention_mlp_block(sd: SD, name: str, x: Tensor, num_heads: int = 8) -> Tensor:
    """Self-attention + GELU MLP, the SAME LayerNorm applied twice.  common.py:151-184"""
    a = layer_norm(sd, name + ".norm", sam_attention(sd, name + ".attn", x, x, x, num_heads) + x, 1e-5)
    return layer_norm(sd, name + ".norm", mlp_block(sd, name + ".mlp", a, "gelu") + a, 1e-5)


# ----------------------------------------------------------------------------------------------
# SAM ViT image encoder — label_anything/models/image_encoder.py
# ----------------------------------------------------------------------------------------------
def rel_pos_table(q_size: int, k_size: int, rel_pos: Tensor) -> Tensor:
    """[q_size, k_size, C] gather of the (optionally linearly resized) table.  image_encoder.py:307-337"""
    span = 2 * max(q_size, k_size) - 1
    if rel_pos.shape[0] != span:
        rel_pos = F.interpolate(rel_pos.t().unsqueeze(0), size=span, mode="linear").squeeze(0).t()
    qc = torch.arange(q_size, dtype=torch.float32)[:, None] * max(k_size / q_size, 1.0)
    kc = torch.arange(k_size, dtype=torch.float32)[None, :] * max(q_size / k_size, 1.0)
    idx = (qc - kc + (k_size - 1) * max(q_size / k_size, 1.0)).long()
    return rel_pos[idx]


def vit_attention(sd: SD, name: str, x: Tensor, num_heads: int, use_rel_pos: bool) -> Tensor:
    """x [B,H,W,C] -> [B,H,W,C]; decomposed rel-pos bias uses the UNSCALED q.  image_encoder.py:239-255,340-376"""
    B, H, W, C = x.shape
    dh = C // num_heads
    qkv = linear(sd, name + ".qkv", x).view(B, H * W, 3, num_heads, dh).permute(2, 0, 3, 1, 4)
    q, k, v = qkv[0], qkv[1], qkv[2]  # [B, heads, HW, dh]
    att = (q * dh ** -0.5) @ k.transpose(-1, -2)
    if use_rel_pos:
        Rh = rel_pos_table(H, H, sd[name + ".rel_pos_h"])
        Rw = rel_pos_table(W, W, sd[name + ".rel_pos_w"])
        q5 = q.reshape(B, num_heads, H, W, dh)
        bias_h = torch.einsum("bnhwc,hkc->bnhwk", q5, Rh)
        bias_w = torch.einsum("bnhwc,wkc->bnhwk", q5, Rw)
        att = (att.view(B, num_heads, H, W, H, W) + bias_h[..., :, None] + bias_w[..., None, :]).view(
            B, num_heads, H * W, H * W)
    att = att.softmax(dim=-1)
    o = (att @ v).transpose(1, 2).reshape(B, H, W, C)
    return linear(sd, name + ".proj", o)


def vit_block(sd: SD, name: str, x: Tensor, num_heads: int, window: int, use_rel_pos: bool, eps: float) -> Tensor:
    """Pre-LN block; windowed blocks zero-pad AFTER norm1 and crop after attention.  image_encoder.py:181-197,258-304"""
    B, H, W, C = x.shape
    y = layer_norm(sd, name + ".norm1", x, eps)
    if window > 0:
        ph, pw = (-H) % window, (-W) % window
        y = F.pad(y, (0, 0, 0, pw, 0, ph))
        Hp, Wp = H + ph, W + pw
        y = y.view(B, Hp // window, window, Wp // window, window, C).permute(0, 1, 3, 2, 4, 5)
        y = y.reshape(-1, window, window, C)
        y = vit_attention(sd, name + ".attn", y, num_heads, use_rel_pos)
        y = y.view(B, Hp // window, Wp // window, window, window, C).permute(0, 1, 3, 2, 4, 5)
        y = y.reshape(B, Hp, Wp, C)[:, :H, :W]
    else:
        y = vit_attention(sd, name + ".attn", y, num_heads, use_rel_pos)
    x = x + y
    return x + mlp_block(sd, name + ".mlp", layer_norm(sd, name + ".norm2", x, eps), "gelu")


def neck(sd: SD, name: str, x: Tensor) -> Tensor:
    """conv1x1 (no bias) -> LN2d -> conv3x3 pad 1 (no bias) -> LN2d.
    image_encoder.py:92-108 (SAM neck) and build_lam.py:150-171 (Lam.neck)."""
    x = F.conv2d(x, sd[name + ".0.weight"])
    x = layer_norm_2d(sd, name + ".1", x)
    x = F.conv2d(x, sd[name + ".2.weight"], padding=1)
    return layer_norm_2d(sd, name + ".3", x)


def sam_vit(sd: SD, name: str, images: Tensor, *, num_heads: int, depth: int, global_attn: Sequence[int],
            window: int = 14, use_rel_pos: bool = True, project_last_hidden: bool = False,
            eps: float = 1e-6) -> Tensor:
    """images [I,3,S,S] -> [I,C,h,w].  image_encoder.py:110-131 (forward), 402-410 (patch embed)."""
    w = sd[name + ".patch_embed.proj.weight"]
    x = F.conv2d(images, w, sd[name + ".patch_embed.proj.bias"], stride=w.shape[-1]).permute(0, 2, 3, 1)
    if name + ".pos_embed" in sd:
        x = x + sd[name + ".pos_embed"]
    for i in range(depth):
        x = vit_block(sd, f"{name}.blocks.{i}", x, num_heads, 0 if i in global_attn else window, use_rel_pos, eps)
    x = x.permute(0, 3, 1, 2)
    if project_last_hidden:
        x = neck(sd, name + ".neck", x)
    return x


# ----------------------------------------------------------------------------------------------
# HuggingFace ViT (MAE encoders) — transformers/models/vit/modeling_vit.py (5.5.0 in this image; the
# reference pins 4.51.3, uv.lock:2750) via label_anything/models/build_encoder.py:83-100 (ViTModelWrapper)
# ----------------------------------------------------------------------------------------------
def hf_vit(sd: SD, name: str, images: Tensor, *, num_heads: int, depth: int, patch: int = 16,
           eps: float = 1e-12) -> Tensor:
    """images [I,3,H,W] -> [I,C,H/16,W/16]: CLS kept through all layers, dropped at the end.

    modeling_vit.py: embeddings 43-129 (bicubic pos-emb resize when the grid differs), layer 315-346
    (pre-LN attention + pre-LN MLP with exact GELU), final layernorm 416; wrapper: build_encoder.py:94-100.
    """
    p = name + ("." if name else "")
    I, _, H, W = images.shape
    x = F.conv2d(images, sd[p + "embeddings.patch_embeddings.projection.weight"],
                 sd[p + "embeddings.patch_embeddings.projection.bias"], stride=patch)
    gh, gw = x.shape[-2:]
    x = x.flatten(2).transpose(1, 2)
    x = torch.cat([sd[p + "embeddings.cls_token"].expand(I, -1, -1), x], dim=1)
    pos = sd[p + "embeddings.position_embeddings"]
    n_pos = pos.shape[1] - 1
    if not (gh * gw == n_pos and H == W):
        side = int(n_pos ** 0.5)
        grid = pos[:, 1:].reshape(1, side, side, -1).permute(0, 3, 1, 2)
        grid = F.interpolate(grid, size=(gh, gw), mode="bicubic", align_corners=False)
        pos = torch.cat([pos[:, :1], grid.permute(0, 2, 3, 1).reshape(1, gh * gw, -1)], dim=1)
    x = x + pos
    C = x.shape[-1]
    dh = C // num_heads
    for i in range(depth):
        lp = f"{p}encoder.layer.{i}."
        y = layer_norm(sd, lp + "layernorm_before", x, eps)
        q = linear(sd, lp + "attention.attention.query", y).view(I, -1, num_heads, dh).transpose(1, 2)
        k = linear(sd, lp + "attention.attention.key", y).view(I, -1, num_heads, dh).transpose(1, 2)
        v = linear(sd, lp + "attention.attention.value", y).view(I, -1, num_heads, dh).transpose(1, 2)
        att = torch.softmax(q @ k.transpose(-1, -2) * dh ** -0.5, dim=-1)
        o = (att @ v).transpose(1, 2).reshape(I, -1, C)
        x = x + linear(sd, lp + "attention.output.dense", o)
        y = layer_norm(sd, lp + "layernorm_after", x, eps)
        y = F.gelu(linear(sd, lp + "intermediate.dense", y))
        x = x + linear(sd, lp + "output.dense", y)
    x = layer_norm(sd, p + "layernorm", x, eps)
    return x[:, 1:].transpose(1, 2).reshape(I, C, gh, gw).contiguous()


# ----------------------------------------------------------------------------------------------
# two-way transformer — label_anything/models/transformer.py:157-329
# ----------------------------------------------------------------------------------------------
def two_way_transformer(sd: SD, name: str, image_embedding: Tensor, image_pe: Tensor, tokens: Tensor, *,
                        depth: int = 2, num_heads: int = 8) -> Tuple[Tensor, Tensor]:
    """image_embedding [S,D,h,w], image_pe [S or 1,D,h,w], tokens [S,n,D] -> (tokens', image tokens' [S,hw,D]).

    Per layer (transformer.py:298-329): token self-attn (layer 0 REPLACES the tokens, no residual, no PE),
    tokens->image cross-attn, ReLU MLP, image->tokens cross-attn; each followed by LayerNorm(eps 1e-5).
    Then a final tokens->image attention + LayerNorm (transformer.py:245-252).
    """
    keys = image_embedding.flatten(2).transpose(1, 2)
    key_pe = image_pe.flatten(2).transpose(1, 2)
    queries, query_pe = tokens, tokens
    for i in range(depth):
        lp = f"{name}.layers.{i}"
        if i == 0:
            queries = sam_attention(sd, lp + ".self_attn", queries, queries, queries, num_heads)
        else:
            q = queries + query_pe
            queries = queries + sam_attention(sd, lp + ".self_attn", q, q, queries, num_heads)
        queries = layer_norm(sd, lp + ".norm1", queries, 1e-5)
        q, k = queries + query_pe, keys + key_pe
        queries = queries + sam_attention(sd, lp + ".cross_attn_token_to_image", q, k, keys, num_heads)
        queries = layer_norm(sd, lp + ".norm2", queries, 1e-5)
        queries = layer_norm(sd, lp + ".norm3", queries + mlp_block(sd, lp + ".mlp", queries, "relu"), 1e-5)
        q, k = queries + query_pe, keys + key_pe
        keys = keys + sam_attention(sd, lp + ".cross_attn_image_to_token", k, q, queries, num_heads)
        keys = layer_norm(sd, lp + ".norm4", keys, 1e-5)
    q, k = queries + query_pe, keys + key_pe
    queries = queries + sam_attention(sd, name + ".final_attn_token_to_image", q, k, keys, num_heads)
    queries = layer_norm(sd, name + ".norm_final_attn", queries, 1e-5)
    return queries, keys


# ----------------------------------------------------------------------------------------------
# prompt encoder — label_anything/models/prompt_encoder.py
# ----------------------------------------------------------------------------------------------
def fourier_pe(gauss: Tensor, coords01: Tensor) -> Tensor:
    """coords in [0,1]^2 (x, y) -> [sin | cos](2*pi*(2c-1) @ G).  prompt_encoder.py:201-211"""
    c = (2.0 * coords01 - 1.0) @ gauss
    c = 2.0 * math.pi * c
    return torch.cat([torch.sin(c), torch.cos(c)], dim=-1)


def dense_pe(gauss: Tensor, h: int, w: int) -> Tensor:
    """[1, D, h, w] PE of pixel centres.  prompt_encoder.py:213-224 and 72-81"""
    ys = (torch.arange(h, dtype=torch.float32) + 0.5) / h
    xs = (torch.arange(w, dtype=torch.float32) + 0.5) / w
    grid = torch.stack([xs[None, :].expand(h, w), ys[:, None].expand(h, w)], dim=-1)
    return fourier_pe(gauss, grid).permute(2, 0, 1).unsqueeze(0)


def embed_points(sd: SD, name: str, coords: Tensor, labels: Tensor, pad: bool, image_size: int) -> Tensor:
    """coords [S,P,2] (x,y px), labels [S,P] in {1,0,-1} -> [S,P(+1),D].  prompt_encoder.py:83-103,648-654.

    The padding point appended when there are no boxes carries label -1, which in this code base is
    Label.NEGATIVE (not "null" as in SAM) -- it therefore receives PE + point_embeddings[0]; restated as is."""
    gauss = sd[name + ".pe_layer.positional_encoding_gaussian_matrix"]
    pts = coords + 0.5
    if pad:
        pts = torch.cat([pts, torch.zeros(pts.shape[0], 1, 2)], dim=1)
        labels = torch.cat([labels, -torch.ones(labels.shape[0], 1)], dim=1)
    emb = fourier_pe(gauss, pts / float(image_size))
    is_null = (labels == NULL).unsqueeze(-1)
    emb = torch.where(is_null, sd[name + ".not_a_point_embed.weight"].expand_as(emb), emb)
    emb = emb + (labels == NEGATIVE).unsqueeze(-1) * sd[name + ".point_embeddings.0.weight"]
    emb = emb + (labels == POSITIVE).unsqueeze(-1) * sd[name + ".point_embeddings.1.weight"]
    return emb


def embed_boxes(sd: SD, name: str, boxes: Tensor, flags: Tensor, image_size: int) -> Tensor:
    """boxes [B,M,C,n,4], flags [B,M,C,n] -> [B*M*C, 2n, D].  prompt_encoder.py:105-114,656-669.

    Note the reference's null-box overwrite indexes the (n xy)-flattened corner axis with
    `flags.repeat(1,1,1,2)`, i.e. corner slot j is governed by flags[..., j % n] — restated as is."""
    gauss = sd[name + ".pe_layer.positional_encoding_gaussian_matrix"]
    B, M, C, n, _ = boxes.shape
    corners = (boxes + 0.5).reshape(B * M * C * n, 2, 2)
    emb = fourier_pe(gauss, corners / float(image_size))
    emb[:, 0] = emb[:, 0] + sd[name + ".point_embeddings.2.weight"]
    emb[:, 1] = emb[:, 1] + sd[name + ".point_embeddings.3.weight"]
    emb = emb.reshape(B, M, C, 2 * n, -1)
    is_null = (flags.repeat(1, 1, 1, 2) == NULL).unsqueeze(-1)
    emb = torch.where(is_null, sd[name + ".not_a_point_embed.weight"].expand_as(emb), emb)
    return emb.reshape(B * M * C, 2 * n, -1)


def embed_masks(sd: SD, name: str, masks: Tensor, flags: Tensor) -> Tensor:
    """masks [B,M,C,Hm,Wm], flags [B,M,C] -> [B,M,C,D,Hm/4,Wm/4].  prompt_encoder.py:61-69,516-540"""
    B, M, C, Hm, Wm = masks.shape
    md = name + ".mask_downscaling"
    x = masks.reshape(B * M * C, 1, Hm, Wm)
    x = F.conv2d(x, sd[md + ".0.weight"], sd[md + ".0.bias"], stride=2)
    x = F.gelu(layer_norm_2d(sd, md + ".1", x))
    x = F.conv2d(x, sd[md + ".3.weight"], sd[md + ".3.bias"], stride=2)
    x = F.gelu(layer_norm_2d(sd, md + ".4", x))
    x = F.conv2d(x, sd[md + ".6.weight"], sd[md + ".6.bias"])
    x = x.view(B, M, C, -1, x.shape[-2], x.shape[-1])
    is_null = (flags == NULL).view(B, M, C, 1, 1, 1)
    return torch.where(is_null, sd[name + ".not_a_mask_embed.weight"].view(1, 1, 1, -1, 1, 1).expand_as(x), x)


def prompt_encoder(sd: SD, name: str, cfg: dict, image_embeddings: Tensor,
                   points: Optional[Tuple[Tensor, Tensor]], boxes: Optional[Tuple[Tensor, Tensor]],
                   masks: Optional[Tuple[Tensor, Tensor]], flag_examples: Tensor,
                   class_rows: Optional[Tensor] = None) -> Dict[str, Tensor]:
    """PromptImageEncoder.forward.  prompt_encoder.py:752-827 (+ 564-646, 671-750).

    image_embeddings [B,M,D,h,w]; returns class embeddings [B,C,D] and per-example embeddings [B,M,C,D].
    `class_rows` pins RandomMatrixEncoder.sample_rows (prompt_encoder.py:245-264); None = no class encoder.
    """
    image_size = cfg["image_size"]
    src_any = points[0] if points is not None else boxes[0] if boxes is not None else masks[0] if masks is not None else None
    if src_any is None:
        raise ValueError("No prompts provided")  # prompt_encoder.py:562
    B, M, C = src_any.shape[:3]
    S = B * M * C
    D = sd[name + ".no_mask_embed.weight"].shape[1]

    parts = []
    if points is not None:
        coords, labels = points
        parts.append(embed_points(sd, name, coords.reshape(S, -1, 2), labels.reshape(S, -1).float(),
                                  pad=(boxes is None), image_size=image_size))
    if boxes is not None:
        parts.append(embed_boxes(sd, name, boxes[0], boxes[1], image_size))
    if parts:
        sparse = torch.cat(parts, dim=1)
    else:
        sparse = sd[name + ".no_sparse_embedding.weight"].view(1, 1, D).expand(S, 1, D)
    n = sparse.shape[1]
    # attention over all (class, token) sparse embeddings of one support image.  prompt_encoder.py:613-629
    sparse = attention_mlp_block(sd, name + ".sparse_embedding_attention", sparse.reshape(B * M, C * n, D))
    sparse = sparse.reshape(B, M, C, n, D)

    h, w = image_embeddings.shape[-2:]
    if masks is not None:
        dense = embed_masks(sd, name, masks[0], masks[1]).flatten(0, 2)
        if dense.shape[-2:] != (h, w):
            dense = F.interpolate(dense, size=(h, w), mode="bilinear", align_corners=False)
    else:
        dense = sd[name + ".no_mask_embed.weight"].view(1, D, 1, 1).expand(S, D, h, w)
    src = image_embeddings.unsqueeze(2).expand(B, M, C, D, h, w).reshape(S, D, h, w) + dense
    gauss = sd[name + ".pe_layer.positional_encoding_gaussian_matrix"]
    gh, gw = cfg["image_embedding_size"]
    pos = dense_pe(gauss, gh, gw)

    if class_rows is not None:  # RandomMatrixEncoder.forward_with_rows, prompt_encoder.py:250-264
        code = sd[name + ".class_encoder.pos_embedding"][0, 0, class_rows[:C]]  # [C, D]
        src = (src.view(B, M, C, D, h, w) + code.view(1, 1, C, D, 1, 1)).reshape(S, D, h, w)
        sparse = sparse + code.view(1, 1, C, 1, D)

    _, fused = two_way_transformer(sd, name + ".transformer", src, pos, sparse.reshape(S, n, D))
    fused = fused.transpose(1, 2).reshape(S, D, h, w)
    emb = fused.mean(dim=(2, 3)).view(B, M, C, D)  # prompt_encoder.py:733-735

    # prompt_class_information_merge, prompt_encoder.py:696-717 (masks passed there are no-ops)
    if cfg.get("class_attention", False):
        emb = attention_mlp_block(sd, name + ".class_attention", emb.reshape(B * M, C, D)).view(B, M, C, D)
    if cfg.get("example_attention", False):
        e = emb.permute(0, 2, 1, 3).reshape(B * C, M, D)
        emb = attention_mlp_block(sd, name + ".example_attention", e).view(B, C, M, D).permute(0, 2, 1, 3)
    if cfg.get("example_class_attention", True):
        emb = attention_mlp_block(sd, name + ".class_example_attention", emb.reshape(B, M * C, D)).view(B, M, C, D)

    fe = flag_examples.to(emb.dtype)
    norm = fe.sum(dim=1).unsqueeze(-1)
    norm = torch.where(norm == 0, torch.ones_like(norm), norm)
    class_emb = (emb * fe.unsqueeze(-1)).sum(dim=1) / norm  # prompt_encoder.py:738-745
    return {"class_embeddings": class_emb, "class_examples_embeddings": emb, "flag_examples": flag_examples,
            "class_examples_src": fused}


# ----------------------------------------------------------------------------------------------
# mask decoder — label_anything/models/mask_decoder.py:169-363, 776-804
# ----------------------------------------------------------------------------------------------
def mask_decoder(sd: SD, name: str, cfg: dict, query_embeddings: Tensor, image_pe: Tensor,
                 class_embeddings: Tensor) -> Tensor:
    """query [B,D,h,w], class embeddings [B,C,D] -> logits [B,C,4h,4w].  mask_decoder.py:316-363"""
    B, D, h, w = query_embeddings.shape
    cls, keys = two_way_transformer(sd, name + ".transformer", query_embeddings, image_pe, class_embeddings)
    feat = keys.transpose(1, 2).reshape(B, D, h, w)
    # class_mlp: 3-layer ReLU MLP.  mask_decoder.py:223-229,776-804
    c = cls
    for i in range(3):
        c = linear(sd, f"{name}.class_mlp.layers.{i}", c)
        if i < 2:
            c = F.relu(c)
    # output_upscaling.  mask_decoder.py:206-222
    up = name + ".output_upscaling"
    x = F.conv_transpose2d(feat, sd[up + ".0.weight"], sd[up + ".0.bias"], stride=2)
    x = F.gelu(layer_norm_2d(sd, up + ".1", x))
    x = F.conv_transpose2d(x, sd[up + ".3.weight"], sd[up + ".3.bias"], stride=2)
    # spatial_convs: conv3x3 (+LN2d+GELU between).  mask_decoder.py:236-255
    n_sc = cfg.get("spatial_convs") or 0
    for i in range(n_sc):
        sc = f"{name}.spatial_convs.{3 * i}"
        x = F.conv2d(x, sd[sc + ".weight"], sd[sc + ".bias"], padding=1)
        if i < n_sc - 1:
            x = F.gelu(layer_norm_2d(sd, f"{name}.spatial_convs.{3 * i + 1}", x))
    b, d, H, W = x.shape
    return (c @ x.view(b, d, H * W)).view(b, -1, H, W)  # mask_decoder.py:309


# ----------------------------------------------------------------------------------------------
# Lam — label_anything/models/lam.py
# ----------------------------------------------------------------------------------------------
def preprocess_shape(oh: int, ow: int, long_side: int) -> Tuple[int, int]:
    """label_anything/data/utils.py:441-449"""
    scale = long_side * 1.0 / max(oh, ow)
    return int(oh * scale + 0.5), int(ow * scale + 0.5)


def postprocess_masks(logits: Tensor, dims: Tensor, image_size: int, custom_preprocess: bool) -> Tensor:
    """Bilinear to image_size, optional un-pad crop, bilinear to each query's original size, pad to the
    batch max with -inf (background channel padding -> 0).  lam.py:383-453.  dims [B, M+1, 2] (H, W)."""
    max_h, max_w = (int(v) for v in dims.view(-1, 2).max(dim=0).values)
    q_sizes = dims[:, 0, :]
    x = F.interpolate(logits, (image_size, image_size), mode="bilinear", align_corners=False)
    outs = []
    for i in range(x.shape[0]):
        oh, ow = int(q_sizes[i, 0]), int(q_sizes[i, 1])
        xi = x[i:i + 1]
        if custom_preprocess:
            ih, iw = preprocess_shape(oh, ow, image_size)
            xi = xi[:, :, :ih, :iw]
        xi = F.interpolate(xi, (oh, ow), mode="bilinear", align_corners=False)
        xi = F.pad(xi, (0, max_w - ow, 0, max_h - oh), value=float("-inf"))
        outs.append(xi)
    out = torch.cat(outs)
    bg = out[:, 0]
    bg[bg == float("-inf")] = 0
    return out


def encode_images(sd: SD, cfg: dict, images: Tensor, chunk: int = 2) -> Tensor:
    """images [I,3,S,S] -> encoder features [I,C,h,w] (before Lam.neck).  Processes `chunk` images per call:
    arithmetic is identical to one big call but never materialises I x heads x T x T."""
    enc = cfg["encoder"]
    outs = []
    for i in range(0, images.shape[0], chunk):
        x = images[i:i + chunk]
        if enc["kind"] == "sam":
            outs.append(sam_vit(sd, "image_encoder", x, num_heads=enc["num_heads"], depth=enc["depth"],
                                global_attn=enc["global_attn"], window=enc.get("window", 14),
                                project_last_hidden=enc.get("project_last_hidden", False)))
        else:
            outs.append(hf_vit(sd, "image_encoder", x, num_heads=enc["num_heads"], depth=enc["depth"]))
    return torch.cat(outs)


def prepare_prompts(batch: dict):
    """Drop prompt types whose flags are all zero.  lam.py:214-239"""
    pts = bxs = msk = None
    if "prompt_points" in batch and not bool((batch["flag_points"] == 0).all()):
        pts = (batch["prompt_points"], batch["flag_points"])
    if "prompt_bboxes" in batch and not bool((batch["flag_bboxes"] == 0).all()):
        bxs = (batch["prompt_bboxes"], batch["flag_bboxes"])
    if "prompt_masks" in batch and not bool((batch["flag_masks"] == 0).all()):
        msk = (batch["prompt_masks"], batch["flag_masks"])
    return pts, bxs, msk, batch["flag_examples"]


def lam_forward(sd: SD, cfg: dict, batch: dict, class_rows: Optional[Tensor] = None,
                return_intermediates: bool = False) -> Dict[str, Tensor]:
    """Lam.forward: lam.py:57-170.  `batch` holds either `images` [B,M+1,3,S,S] or `embeddings` [B,M+1,Ce,h,w]."""
    if "embeddings" in batch:
        emb = batch["embeddings"]
        B, N1 = emb.shape[:2]
        feats = emb.flatten(0, 1)
    elif "images" in batch:
        B, N1 = batch["images"].shape[:2]
        feats = encode_images(sd, cfg, batch["images"].flatten(0, 1))
    else:
        raise ValueError("Either 'images' or 'embeddings' must be provided.")  # lam.py:165
    enc_out = feats
    if cfg.get("has_neck", False):
        feats = neck(sd, "neck", feats)
    feats = feats.view(B, N1, *feats.shape[1:])
    query, support = feats[:, 0], feats[:, 1:]
    pts, bxs, msk, flag_examples = prepare_prompts(batch)
    pe = prompt_encoder(sd, "prompt_encoder", cfg, support, pts, bxs, msk, flag_examples, class_rows)
    gauss = sd["prompt_encoder.pe_layer.positional_encoding_gaussian_matrix"]
    gh, gw = cfg["image_embedding_size"]
    low = mask_decoder(sd, "mask_decoder", cfg, query, dense_pe(gauss, gh, gw), pe["class_embeddings"])
    seg = postprocess_masks(low, batch["dims"], cfg["image_size"], cfg.get("custom_preprocess", True))
    if "flag_gts" in batch:
        seg[batch["flag_gts"].logical_not()] = float("-inf")  # lam.py:92-93
    out = {"logits": seg, "class_examples_embeddings": pe["class_examples_embeddings"]}
    if return_intermediates:
        out.update(encoder_out=enc_out, features=feats, class_embs=pe["class_embeddings"], low_res_logits=low)
    return out
