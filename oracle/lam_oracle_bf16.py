"""bf16-matched CPU oracle for the LabelAnything hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

The fp32 oracle (oracle/lam_oracle.py, pinned to the unmodified reference) answers "is this the reference's
algorithm?".  This module answers the other half of the parity question: "is the native path that algorithm evaluated
with bf16 tensor-core operands, and nothing else?".  It restates the SAME reference algorithm (every function cites the
reference file:line through its fp32 sibling) but rounds exactly where the native path rounds:

  * every GEMM / attention operand is a bf16 value, products are accumulated in fp32 (torch CPU fp32 matmul of
    bf16-rounded operands; only the summation order differs from the tensor core's);
  * a GEMM result is rounded to bf16 where the native path stores it as bf16 (q / k / v, attention outputs, the
    attention branch before its residual add, MLP hidden activations, `src` of the prompt encoder ...) and kept in fp32
    where it keeps fp32 (residual streams, LayerNorm statistics, the lin2 residual accumulate, neck conv outputs,
    logits);
  * softmax probabilities are rounded to bf16 before the PV product of the ViT attention (the row sum stays fp32);
    the CUDA-core token attention of the two-way transformer keeps them in fp32;
  * the rel-pos tables of the global blocks are fp16, those of the windowed blocks fp32; the 16 -> D mask-embedding
    projection has TF32 operands.

What it can and cannot certify (measured, tests/test_lam_gpu.py): two bf16 evaluations of this network
that differ by as little as the fp32 summation order end up a full bf16 rounding noise apart after a handful of
GEMM -> round stages (a relative perturbation d of a layer's inputs flips ~d / 2^-8 of the next roundings, which
perturbs the following layer by ~sqrt(2^-8 d)), so native (N), this oracle (M) and the fp32 reference (F) form an
almost equilateral triangle: mean |N-F| 0.015-0.027, |M-F| 0.012-0.023, |N-M| 0.013-0.024 at logit std 1.0-1.8.  The
tests assert exactly that: the native drift is no larger than this independent bf16 evaluation's (x1.5) and N - M is
what two independent noise realisations give -- no error component beyond bf16 rounding.  north_star's 1e-3 max-abs is
held per kernel on identical inputs (tests/test_kernels_gpu.py), where it is meaningful.

Only tests/ may import this module.  Pin: with rounding switched off (`exact=True`) every function reduces to its
fp32 sibling in lam_oracle.py, which tests/test_oracle_golden.py checks against the reference's golden tensors.
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

import lam_oracle as O

Tensor = torch.Tensor
SD = Dict[str, Tensor]

EXACT = False   # True: no rounding anywhere (reduces to lam_oracle.py; used by the pin test)


def r16(x: Tensor) -> Tensor:
    """Round to bf16 (RNE), keep computing in fp32."""
    return x if EXACT else x.to(torch.bfloat16).float()


def rh(x: Tensor) -> Tensor:
    return x if EXACT else x.to(torch.float16).float()


def tf32(x: Tensor) -> Tensor:
    """cvt.rna.tf32.f32: 10 explicit mantissa bits, round to nearest, ties away from zero."""
    if EXACT:
        return x
    i = x.contiguous().view(torch.int32)
    return ((i + 0x1000) & ~0x1FFF).view(torch.float32)


def lin(sd: SD, name: str, x16: Tensor, *, out16: bool = True, act: Optional[str] = None) -> Tensor:
    """la_gemm_bf16: x16 holds bf16 values; weight rounded to bf16; fp32 accumulate + fp32 bias; activation; output
    rounded to bf16 unless the native GEMM writes fp32."""
    y = F.linear(x16, r16(sd[name + ".weight"]), sd.get(name + ".bias"))
    if act == "gelu":
        y = F.gelu(y)
    elif act == "relu":
        y = F.relu(y)
    return r16(y) if out16 else y


def ln(sd: SD, name: str, x: Tensor, eps: float) -> Tensor:
    return F.layer_norm(x, (x.shape[-1],), sd[name + ".weight"], sd[name + ".bias"], eps)


# ----------------------------------------------------------------------------------------------
# ViT attention with bf16 P (la_attention_bf16 / la_attention_window_bf16)
# ----------------------------------------------------------------------------------------------
def _vit_attention_core(q: Tensor, k: Tensor, v: Tensor, bias: Optional[Tensor]) -> Tensor:
    """q, k, v [n, heads, L, 64] bf16 values -> bf16 values.  softmax in fp32, P rounded to bf16 for the PV product,
    normalisation by the UNROUNDED fp32 row sum (image_encoder.py:246-254)."""
    s = (q @ k.transpose(-1, -2)) * (q.shape[-1] ** -0.5)
    if bias is not None:
        s = s + bias
    p = torch.exp(s - s.max(dim=-1, keepdim=True).values)
    return r16((r16(p) @ v) / p.sum(dim=-1, keepdim=True))


def _rel_bias(q: Tensor, rel_h: Tensor, rel_w: Tensor, g: int, fp16_tables: bool) -> Tensor:
    """q [n, heads, g*g, 64] (bf16 values, UNSCALED) -> decomposed rel-pos bias [n, heads, g*g, g*g]
    (image_encoder.py:340-376); the table operand is bf16, the products fp32 (windows) or fp16 (global blocks)."""
    Rh = r16(O.rel_pos_table(g, g, rel_h))
    Rw = r16(O.rel_pos_table(g, g, rel_w))
    q5 = q.reshape(q.shape[0], q.shape[1], g, g, q.shape[-1])
    bh = torch.einsum("bnhwc,hkc->bnhwk", q5, Rh)
    bw = torch.einsum("bnhwc,wkc->bnhwk", q5, Rw)
    if fp16_tables:
        bh, bw = rh(bh), rh(bw)
    return (bh[..., :, None] + bw[..., None, :]).reshape(q.shape[0], q.shape[1], g * g, g * g)


def vit_block(sd: SD, name: str, x: Tensor, num_heads: int, window: int, eps: float, accumulate: bool,
              chunk: int = 2) -> Tensor:
    """One SAM ViT block on the fp32 residual stream x [B, H, W, C] (image_encoder.py:181-197) with the native
    rounding points (labelanything_b200/vit_engine.py::run_vit)."""
    B, H, W, C = x.shape
    dh = C // num_heads
    y = r16(ln(sd, name + ".norm1", x, eps))
    if window > 0:
        ph, pw = (-H) % window, (-W) % window
        y = F.pad(y, (0, 0, 0, pw, 0, ph))
        Hp, Wp = H + ph, W + pw
        y = y.view(B, Hp // window, window, Wp // window, window, C).permute(0, 1, 3, 2, 4, 5).reshape(-1, window, window, C)
        g = window
    else:
        g = H
    n = y.shape[0]
    qkv = lin(sd, name + ".attn.qkv", y.reshape(n, g * g, C))            # bf16 q | k | v (two GEMMs natively; same values)
    qkv = qkv.view(n, g * g, 3, num_heads, dh).permute(2, 0, 3, 1, 4)
    outs = []
    for s in range(0, n, chunk if window == 0 else 512):
        e = s + (chunk if window == 0 else 512)
        q, k, v = qkv[0, s:e], qkv[1, s:e], qkv[2, s:e]
        bias = _rel_bias(q, sd[name + ".attn.rel_pos_h"], sd[name + ".attn.rel_pos_w"], g, fp16_tables=(window == 0)) \
            if name + ".attn.rel_pos_h" in sd else None
        outs.append(_vit_attention_core(q, k, v, bias))
    o = torch.cat(outs).transpose(1, 2).reshape(n, g, g, C)
    if window > 0:
        o = o.view(B, Hp // window, Wp // window, window, window, C).permute(0, 1, 3, 2, 4, 5).reshape(B, Hp, Wp, C)
        o = o[:, :H, :W]
    x = x + lin(sd, name + ".attn.proj", o)                                # bf16 branch added to the fp32 stream
    y2 = r16(ln(sd, name + ".norm2", x, eps))
    h = lin(sd, name + ".mlp.lin1", y2, act="gelu")
    return x + lin(sd, name + ".mlp.lin2", h, out16=not accumulate)        # fp32 accumulate epilogue when supported


def _accumulate_supported(rows: int, d: int) -> bool:
    """labelanything_b200/ops.py::gemm_accumulate_supported"""
    return rows >= 2048 and d >= 256 and d % 8 == 0


def sam_vit(sd: SD, name: str, images: Tensor, *, num_heads: int, depth: int, global_attn: Sequence[int],
            window: int = 14, eps: float = 1e-6, out16: bool = True) -> Tensor:
    """images [I, 3, S, S] -> [I, C, h, w] (no SAM neck).  image_encoder.py:110-131, 402-410."""
    w = sd[name + ".patch_embed.proj.weight"]
    I = images.shape[0]
    p = w.shape[-1]
    cols = F.unfold(r16(images), p, stride=p).transpose(1, 2)               # im2col in bf16
    x = r16(F.linear(cols, r16(w.reshape(w.shape[0], -1)), sd[name + ".patch_embed.proj.bias"]))
    g = images.shape[-1] // p
    x = x.view(I, g, g, -1)
    if name + ".pos_embed" in sd:
        x = x + sd[name + ".pos_embed"]
    acc = _accumulate_supported(I * g * g, x.shape[-1])
    for i in range(depth):
        x = vit_block(sd, f"{name}.blocks.{i}", x, num_heads, 0 if i in global_attn else window, eps, acc)
    x = x.permute(0, 3, 1, 2)
    return r16(x) if out16 else x


def hf_vit(sd: SD, name: str, images: Tensor, *, num_heads: int, depth: int, patch: int = 16, eps: float = 1e-12,
           out16: bool = True) -> Tensor:
    """HF ViT (modeling_vit.py:43-129, 315-346, 416; build_encoder.py:94-100) with the native rounding points."""
    pfx = name + ("." if name else "")
    I, _, H, W = images.shape
    wp = sd[pfx + "embeddings.patch_embeddings.projection.weight"]
    cols = F.unfold(r16(images), patch, stride=patch).transpose(1, 2)
    x = r16(F.linear(cols, r16(wp.reshape(wp.shape[0], -1)), sd[pfx + "embeddings.patch_embeddings.projection.bias"]))
    gh, gw = H // patch, W // patch
    x = torch.cat([sd[pfx + "embeddings.cls_token"].expand(I, -1, -1), x], dim=1)
    pos = sd[pfx + "embeddings.position_embeddings"]
    n_pos = pos.shape[1] - 1
    if not (gh * gw == n_pos and H == W):
        side = int(n_pos ** 0.5)
        grid = pos[:, 1:].reshape(1, side, side, -1).permute(0, 3, 1, 2)
        grid = F.interpolate(grid, size=(gh, gw), mode="bicubic", align_corners=False)
        pos = torch.cat([pos[:, :1], grid.permute(0, 2, 3, 1).reshape(1, gh * gw, -1)], dim=1)
    x = x + pos
    C = x.shape[-1]
    dh = C // num_heads
    acc = _accumulate_supported(I * x.shape[1], C)
    for i in range(depth):
        lp = f"{pfx}encoder.layer.{i}."
        y = r16(ln(sd, lp + "layernorm_before", x, eps))
        q = lin(sd, lp + "attention.attention.query", y).view(I, -1, num_heads, dh).transpose(1, 2)
        k = lin(sd, lp + "attention.attention.key", y).view(I, -1, num_heads, dh).transpose(1, 2)
        v = lin(sd, lp + "attention.attention.value", y).view(I, -1, num_heads, dh).transpose(1, 2)
        o = _vit_attention_core(q, k, v, None).transpose(1, 2).reshape(I, -1, C)
        x = x + lin(sd, lp + "attention.output.dense", o)
        y = r16(ln(sd, lp + "layernorm_after", x, eps))
        h = lin(sd, lp + "intermediate.dense", y, act="gelu")
        x = x + lin(sd, lp + "output.dense", h, out16=not acc)
    x = ln(sd, pfx + "layernorm", x, eps)
    x = x[:, 1:].transpose(1, 2).reshape(I, C, gh, gw)
    return r16(x) if out16 else x.contiguous()


def neck(sd: SD, name: str, x16: Tensor) -> Tensor:
    """conv1x1 -> LN2d -> conv3x3 -> LN2d (build_lam.py:150-171) on bf16 feature values; fp32 conv outputs."""
    t1 = F.conv2d(x16, r16(sd[name + ".0.weight"]))
    y1 = r16(O.layer_norm_2d(sd, name + ".1", t1))
    t2 = F.conv2d(y1, r16(sd[name + ".2.weight"]), padding=1)
    return O.layer_norm_2d(sd, name + ".3", t2)


# ----------------------------------------------------------------------------------------------
# token attention (la_attention_tokens: CUDA cores, fp32 softmax and PV) and the blocks built on it
# ----------------------------------------------------------------------------------------------
def _tok_attention(q: Tensor, k: Tensor, v: Tensor, num_heads: int, q_add: Optional[Tensor] = None,
                   k_add: Optional[Tensor] = None) -> Tensor:
    """q [S, nq, Di], k / v [S, nk, Di] bf16 values (+ fp32 positional tables) -> bf16 values [S, nq, Di]."""
    if q_add is not None:
        q = q + q_add
    if k_add is not None:
        k = k + k_add
    S, nq, Di = q.shape
    dh = Di // num_heads
    qh = q.view(S, nq, num_heads, dh).transpose(1, 2)
    kh = k.view(S, -1, num_heads, dh).transpose(1, 2)
    vh = v.view(S, -1, num_heads, dh).transpose(1, 2)
    att = torch.softmax(qh @ kh.transpose(-1, -2) / math.sqrt(dh), dim=-1)
    return r16((att @ vh).transpose(1, 2).reshape(S, nq, Di))


def _pe_table(sd: SD, name: str, pe: Tensor) -> Tensor:
    """pe @ W^T (fp64 -> fp32), the positional encoding's share of a projection (bias stays in the GEMM)."""
    return (pe.double() @ sd[name + ".weight"].double().t()).float()


def attention_mlp_block(sd: SD, name: str, x: Tensor, num_heads: int = 8) -> Tensor:
    """common.py:151-184 with the rounding points of transformer.py::run_attention_mlp_block.  x fp32 [n, L, D]."""
    xb = r16(x)
    a = name + ".attn"
    q, k, v = lin(sd, a + ".q_proj", xb), lin(sd, a + ".k_proj", xb), lin(sd, a + ".v_proj", xb)
    o = lin(sd, a + ".out_proj", _tok_attention(q, k, v, num_heads))
    a32 = ln(sd, name + ".norm", x + o, 1e-5)
    h = lin(sd, name + ".mlp.lin1", r16(a32), act="gelu")
    m = lin(sd, name + ".mlp.lin2", h)
    return ln(sd, name + ".norm", a32 + m, 1e-5)


def two_way(sd: SD, name: str, keys16: Tensor, keys32: Optional[Tensor], pe: Tensor, tokens: Tensor, *,
            depth: int = 2, num_heads: int = 8, want_queries: bool, pool: bool):
    """transformer.py:206-252, 298-329 with the rounding points of labelanything_b200/transformer.py::run_two_way.
    keys16 [S, T, D] bf16 values (keys32: the same in fp32 when the native path has it), pe [T, D], tokens [S, n, D]
    fp32.  Returns (queries | None, keys16 | None, pooled | None)."""
    S, n, D = tokens.shape
    qpe = tokens
    tok32 = tokens
    tok16 = tokpe16 = None
    pooled = None
    for i in range(depth):
        lp = f"{name}.layers.{i}"
        last = i == depth - 1
        sa = lp + ".self_attn"
        if i == 0:          # skip_first_layer_pe: queries REPLACED by the attention output (transformer.py:302-303)
            xb = tok16 if tok16 is not None else r16(tok32)
            q, k, v = lin(sd, sa + ".q_proj", xb), lin(sd, sa + ".k_proj", xb), lin(sd, sa + ".v_proj", xb)
            resid = 0.0
        else:
            q, k = lin(sd, sa + ".q_proj", tokpe16), lin(sd, sa + ".k_proj", tokpe16)
            v = lin(sd, sa + ".v_proj", tok16)
            resid = tok32
        o = lin(sd, sa + ".out_proj", _tok_attention(q, k, v, num_heads))
        tok32 = ln(sd, lp + ".norm1", resid + o, 1e-5)
        tok16, tokpe16 = r16(tok32), r16(tok32 + qpe)

        t2i, i2t = lp + ".cross_attn_token_to_image", lp + ".cross_attn_image_to_token"
        # (2) tokens attend to the image
        tq = lin(sd, t2i + ".q_proj", tokpe16)
        if n == 1 and num_heads <= 8 and D in (64, 128, 256, 512):
            # one query token: the native path moves the k / v projections to the query side
            # (labelanything_b200/transformer.py::_pooled_token_to_image)
            Dc = tq.shape[-1]
            dh = Dc // num_heads
            wk, wv = r16(sd[t2i + ".k_proj.weight"]), r16(sd[t2i + ".v_proj.weight"])
            qh = tq.view(S, num_heads, dh)
            u = r16(torch.einsum("shj,hjd->shd", qh, wk.view(num_heads, dh, D)))             # u_h = W_k[h]^T q_h
            e = u @ r16(pe).t()                                                                # u_h . pe_t (fp32)
            sc = (torch.einsum("shd,std->sht", u, keys16) + e) * dh ** -0.5
            pr = torch.exp(sc - sc.max(dim=-1, keepdim=True).values)
            y = r16(torch.einsum("sht,std->shd", r16(pr), keys16) / pr.sum(dim=-1, keepdim=True))
            ov = r16(torch.einsum("shd,hjd->shj", y, wv.view(num_heads, dh, D)) +
                     sd[t2i + ".v_proj.bias"].view(1, num_heads, dh))
            o = lin(sd, t2i + ".out_proj", ov.reshape(S, 1, Dc))
        else:
            pk, pv = lin(sd, t2i + ".k_proj", keys16), lin(sd, t2i + ".v_proj", keys16)
            o = lin(sd, t2i + ".out_proj", _tok_attention(tq, pk, pv, num_heads, k_add=_pe_table(sd, t2i + ".k_proj", pe)))
        tok32 = ln(sd, lp + ".norm2", tok32 + o, 1e-5)
        tok16 = r16(tok32)
        # (3) token MLP (ReLU)
        hmid = lin(sd, lp + ".mlp.lin1", tok16, act="relu")
        tok32 = ln(sd, lp + ".norm3", tok32 + lin(sd, lp + ".mlp.lin2", hmid), 1e-5)
        tok16, tokpe16 = r16(tok32), r16(tok32 + qpe)
        # (4) image attends to the tokens
        tv = lin(sd, i2t + ".v_proj", tok16)
        if n > 1:
            pq = lin(sd, i2t + ".q_proj", keys16)
            tk = lin(sd, i2t + ".k_proj", tokpe16)
            o = _tok_attention(pq, tk, tv, num_heads, q_add=_pe_table(sd, i2t + ".q_proj", pe))
            delta = lin(sd, i2t + ".out_proj", o)
        else:               # softmax over one key == 1: the branch is out_proj(v_proj(token)) for every image position
            delta = lin(sd, i2t + ".out_proj", tv, out16=False)          # [S, 1, D] fp32, broadcast over T
        x = (keys32 if keys32 is not None else keys16) + delta
        y = ln(sd, lp + ".norm4", x, 1e-5)
        if last and pool:
            pooled = y.mean(dim=1)
            keys16 = keys32 = None
        else:
            keys16 = r16(y)
            keys32 = y if ((not last) and not pool) else None
    queries = None
    if want_queries:
        fa = name + ".final_attn_token_to_image"
        k, v = lin(sd, fa + ".k_proj", keys16), lin(sd, fa + ".v_proj", keys16)
        tq = lin(sd, fa + ".q_proj", tokpe16)
        o = lin(sd, fa + ".out_proj", _tok_attention(tq, k, v, num_heads, k_add=_pe_table(sd, fa + ".k_proj", pe)))
        queries = ln(sd, name + ".norm_final_attn", tok32 + o, 1e-5)
    return queries, keys16, pooled


# ----------------------------------------------------------------------------------------------
# prompt encoder / mask decoder / Lam
# ----------------------------------------------------------------------------------------------
def _tokens(x_nchw: Tensor) -> Tensor:
    return x_nchw.flatten(2).transpose(1, 2)


def prompt_encoder(sd: SD, name: str, cfg: dict, support32: Tensor, points, boxes, masks, flag_examples: Tensor,
                   class_rows: Optional[Tensor] = None) -> Dict[str, Tensor]:
    """prompt_encoder.py:752-827 with the rounding points of labelanything_b200/prompt_encoder.py::encode.
    support32 [B, M, D, h, w] fp32 neck features."""
    image_size = cfg["image_size"]
    any_p = points[0] if points is not None else boxes[0] if boxes is not None else masks[0] if masks is not None else None
    if any_p is None:
        raise ValueError("No prompts provided")
    B, M, C = any_p.shape[:3]
    S = B * M * C
    D = sd[name + ".no_mask_embed.weight"].shape[1]
    parts = []
    if points is not None:
        parts.append(O.embed_points(sd, name, points[0].reshape(S, -1, 2), points[1].reshape(S, -1).float(),
                                    pad=(boxes is None), image_size=image_size))
    if boxes is not None:
        parts.append(O.embed_boxes(sd, name, boxes[0], boxes[1], image_size))
    sparse = torch.cat(parts, dim=1) if parts else sd[name + ".no_sparse_embedding.weight"].view(1, 1, D).expand(S, 1, D)
    n = sparse.shape[1]
    sparse = attention_mlp_block(sd, name + ".sparse_embedding_attention", sparse.reshape(B * M, C * n, D))
    sparse = sparse.reshape(B, M, C, n, D)
    code = None
    if class_rows is not None:
        code = sd[name + ".class_encoder.pos_embedding"][0, 0, class_rows[:C]]
        sparse = sparse + code.view(1, 1, C, 1, D)

    h, w = support32.shape[-2:]
    if masks is not None:
        mk, mf = masks
        md = name + ".mask_downscaling"
        x = mk.reshape(S, 1, *mk.shape[-2:])
        x = F.conv2d(x, sd[md + ".0.weight"], sd[md + ".0.bias"], stride=2)
        x = F.gelu(O.layer_norm_2d(sd, md + ".1", x))
        x = F.conv2d(x, sd[md + ".3.weight"], sd[md + ".3.bias"], stride=2)
        m16 = F.gelu(O.layer_norm_2d(sd, md + ".4", x))                  # [S, 16, Hm/4, Wm/4] fp32
        if m16.shape[-2:] != (h, w):                                       # bilinear commutes with the 1x1 conv
            m16 = F.interpolate(m16, size=(h, w), mode="bilinear", align_corners=False)
        w6 = sd[md + ".6.weight"].reshape(D, -1)
        dense = torch.einsum("schw,dc->sdhw", tf32(m16), tf32(w6)) + sd[md + ".6.bias"].view(1, D, 1, 1)   # TF32 MMA
        is_null = (mf.reshape(S) == 0).view(S, 1, 1, 1)
        dense = torch.where(is_null, sd[name + ".not_a_mask_embed.weight"].view(1, D, 1, 1).expand_as(dense), dense)
    else:
        dense = sd[name + ".no_mask_embed.weight"].view(1, D, 1, 1).expand(S, D, h, w)
    src = support32.unsqueeze(2).expand(B, M, C, D, h, w).reshape(S, D, h, w) + dense
    if code is not None:
        src = (src.view(B, M, C, D, h, w) + code.view(1, 1, C, D, 1, 1)).reshape(S, D, h, w)
    src16 = r16(_tokens(src))                                             # la_build_src writes bf16 once
    gh, gw = cfg["image_embedding_size"]
    pe = _tokens(O.dense_pe(sd[name + ".pe_layer.positional_encoding_gaussian_matrix"], gh, gw))[0]
    _, _, pooled = two_way(sd, name + ".transformer", src16, None, pe, sparse.reshape(S, n, D), want_queries=False,
                           pool=True)
    emb = pooled.view(B, M, C, D)
    if cfg.get("class_attention", False):
        emb = attention_mlp_block(sd, name + ".class_attention", emb.reshape(B * M, C, D)).view(B, M, C, D)
    if cfg.get("example_attention", False):
        e = emb.permute(0, 2, 1, 3).reshape(B * C, M, D)
        emb = attention_mlp_block(sd, name + ".example_attention", e).view(B, C, M, D).permute(0, 2, 1, 3)
    if cfg.get("example_class_attention", True):
        emb = attention_mlp_block(sd, name + ".class_example_attention", emb.reshape(B, M * C, D)).view(B, M, C, D)
    fe = (flag_examples != 0).to(emb.dtype)
    norm = fe.sum(dim=1).unsqueeze(-1)
    norm = torch.where(norm == 0, torch.ones_like(norm), norm)
    class_emb = (emb * fe.unsqueeze(-1)).sum(dim=1) / norm
    return {"class_embeddings": class_emb, "class_examples_embeddings": emb}


def mask_decoder(sd: SD, name: str, cfg: dict, query32: Tensor, class_embeddings: Tensor) -> Tensor:
    """mask_decoder.py:316-363 with the rounding points of labelanything_b200/mask_decoder.py::decode.
    query32 [B, D, h, w] fp32 neck features, class_embeddings [B, C, D] -> logits [B, C, 4h, 4w] fp32."""
    B, D, h, w = query32.shape
    gauss = sd["prompt_encoder.pe_layer.positional_encoding_gaussian_matrix"]
    pe = _tokens(O.dense_pe(gauss, h, w))[0]
    q32 = _tokens(query32)
    queries, keys16, _ = two_way(sd, name + ".transformer", r16(q32), q32, pe, class_embeddings, want_queries=True,
                                 pool=False)
    c = r16(queries)
    for i in range(3):
        c = lin(sd, f"{name}.class_mlp.layers.{i}", c, out16=i < 2, act="relu" if i < 2 else None)
    up = name + ".output_upscaling"
    feat = keys16.transpose(1, 2).reshape(B, D, h, w)
    x = r16(F.conv_transpose2d(feat, r16(sd[up + ".0.weight"]), sd[up + ".0.bias"], stride=2))
    x = r16(F.gelu(O.layer_norm_2d(sd, up + ".1", x)))
    x = r16(F.conv_transpose2d(x, r16(sd[up + ".3.weight"]), sd[up + ".3.bias"], stride=2))
    n_sc = cfg.get("spatial_convs") or 0
    for i in range(n_sc):
        sc = f"{name}.spatial_convs.{3 * i}"
        x = r16(F.conv2d(x, r16(sd[sc + ".weight"]), sd[sc + ".bias"], padding=1))
        if i < n_sc - 1:
            x = r16(F.gelu(O.layer_norm_2d(sd, f"{name}.spatial_convs.{3 * i + 1}", x)))
    b, d, H, W = x.shape
    return (c @ x.view(b, d, H * W)).view(b, -1, H, W)


def encode_images(sd: SD, cfg: dict, images: Tensor, out16: bool) -> Tensor:
    enc = cfg["encoder"]
    if enc["kind"] == "sam":
        return sam_vit(sd, "image_encoder", images, num_heads=enc["num_heads"], depth=enc["depth"],
                       global_attn=enc["global_attn"], window=enc.get("window", 14), out16=out16)
    return hf_vit(sd, "image_encoder", images, num_heads=enc["num_heads"], depth=enc["depth"], out16=out16)


def lam_forward(sd: SD, cfg: dict, batch: dict, class_rows: Optional[Tensor] = None) -> Dict[str, Tensor]:
    """Lam.forward (lam.py:57-170) with the native rounding points (labelanything_b200/lam.py)."""
    has_neck = cfg.get("has_neck", False)
    if "embeddings" in batch:
        emb = batch["embeddings"]
        B, N1 = emb.shape[:2]
        feats = emb.flatten(0, 1).float()
        feats = r16(feats) if has_neck else feats
    else:
        B, N1 = batch["images"].shape[:2]
        feats = encode_images(sd, cfg, batch["images"].flatten(0, 1), out16=has_neck)
    if has_neck:
        feats = neck(sd, "neck", feats)
    feats = feats.view(B, N1, *feats.shape[1:])
    query, support = feats[:, 0], feats[:, 1:]
    pts, bxs, msk, flag_examples = O.prepare_prompts(batch)
    pe = prompt_encoder(sd, "prompt_encoder", cfg, support, pts, bxs, msk, flag_examples, class_rows)
    low = mask_decoder(sd, "mask_decoder", cfg, query, pe["class_embeddings"])
    seg = O.postprocess_masks(low, batch["dims"], cfg["image_size"], cfg.get("custom_preprocess", True))
    if "flag_gts" in batch:
        seg[batch["flag_gts"].logical_not()] = float("-inf")
    return {"logits": seg, "class_examples_embeddings": pe["class_examples_embeddings"], "low_res_logits": low,
            "features": feats, "class_embeddings": pe["class_embeddings"]}
