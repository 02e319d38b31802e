"""Launch sequence of the ViT image encoders (SAM ViT and HuggingFace ViT) over the native kernels.

The engine is weight-layout agnostic: `image_encoder.ImageEncoderViT` and `build_encoder.ViTModelWrapper` pack
their parameters into `BlockWeights` and call `run_vit`.  Per transformer block (image_encoder.py:181-197 /
modeling_vit.py:315-346) the launches are

    add+LN1 (window partition folded in) -> Q GEMM, KV GEMM -> [rel-pos table GEMM, global blocks] -> fused attention
    (window un-partition folded in) -> proj GEMM -> add+LN2 -> lin1 GEMM (+GELU) -> lin2 GEMM (+ residual)

The fp32 residual stream is updated inside add+LN2 (attention branch) and in the epilogue of the lin2 GEMM (MLP
branch, `ops.gemm_accumulate`); every other GEMM writes bf16 with a plain bias/activation epilogue.  Images are processed in chunks to bound the workspace (HBM-resident, ~125 MB per
1024-px image).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional

import torch

from . import ops


@dataclass
class BlockWeights:
    ln1_w: torch.Tensor
    ln1_b: torch.Tensor
    wq: torch.Tensor          # [d, d] bf16
    bq: Optional[torch.Tensor]
    wkv: torch.Tensor         # [2d, d] bf16
    bkv: Optional[torch.Tensor]
    wproj: torch.Tensor
    bproj: Optional[torch.Tensor]
    ln2_w: torch.Tensor
    ln2_b: torch.Tensor
    w1: torch.Tensor
    b1: Optional[torch.Tensor]
    w2: torch.Tensor
    b2: Optional[torch.Tensor]
    window: int = 0                           # 0 = global attention
    rel_table: Optional[torch.Tensor] = None  # [2P, 64] bf16: reversed rel_pos_h (P rows, zero padded) ++ reversed rel_pos_w
    rel_pad: int = 0                          # P


@dataclass
class VitSpec:
    d: int
    heads: int
    eps: float
    blocks: List[BlockWeights]
    grid: int                                  # tokens per side (64 for 1024 px, 30 for 480 px)
    n_cls: int = 0                             # 1 for HF ViT (CLS token kept through all layers)
    final_ln_w: Optional[torch.Tensor] = None  # HF: layernorm after the last block
    final_ln_b: Optional[torch.Tensor] = None


# experiment switch: also take the attention branch's residual add in the proj GEMM epilogue (prefetching variant)
_FUSE_PROJ = bool(__import__("os").environ.get("LA_FUSE_PROJ"))
# experiment switch (import time): 0 keeps the window-partitioned projections of round 1 for A/B runs
_GRID_ROUTE = __import__("os").environ.get("LA_WINDOW_GRID_ROUTE", "1") != "0"
# element type of the global blocks' rel-pos tables q . rel_pos (the rel_w half is rounded to fp16 inside the attention
# kernel either way; fp16 halves the table traffic).  LA_REL_TABLE_F32=1 keeps the fp32 tables.
_TABLE_DTYPE = torch.float32 if __import__("os").environ.get("LA_REL_TABLE_F32") else torch.float16


def reversed_rel_table(rel_pos: torch.Tensor, size: int, pad_to: int) -> torch.Tensor:
    """rel_pos [L, 64] -> [pad_to, 64] with row i = rel_pos'[2*size-2 - i], rel_pos' = table resized (linear) to
    2*size-1 rows when L differs (get_rel_pos, image_encoder.py:319-330)."""
    span = 2 * size - 1
    t = rel_pos.detach().float()
    if t.shape[0] != span:
        t = torch.nn.functional.interpolate(t.t().unsqueeze(0), size=span, mode="linear").squeeze(0).t()
    out = torch.zeros(pad_to, t.shape[1], dtype=torch.float32, device=t.device)
    out[:span] = torch.flip(t, dims=[0])
    return out


def run_vit(spec: VitSpec, x: torch.Tensor, n_img: int, out_dtype: torch.dtype) -> torch.Tensor:
    """x: fp32 residual stream [n_img * (grid^2 + n_cls), d] (already patch-embedded + positional).
    Returns the encoder output tokens [n_img * grid^2, d] (CLS dropped) in `out_dtype`."""
    d, heads, g = spec.d, spec.heads, spec.grid
    dev = x.device
    T = g * g + spec.n_cls
    rows = n_img * T
    assert x.shape == (rows, d) and d == heads * 64, "native ViT kernels are built for head_dim 64"
    scale = 64 ** -0.5
    delta = None
    for bw in spec.blocks:
        if bw.window > 0:
            assert spec.n_cls == 0
            win = bw.window
            nwin = (g + win - 1) // win
            seq_len, n_seq = win * win, n_img * nwin * nwin
            r_att = n_seq * seq_len
            # Padded-grid route (g % 32 == 0, 14 x 14 windows): norm1 in image order, the projections store their
            # g x g tokens into a (nwin * win)^2 grid whose padding positions get the bias row -- what the reference's
            # F.pad after norm1 projects to (image_encoder.py:183-192) -- and the attention kernel fetches every window
            # as one 4-D box: 16 % fewer projected rows and no window-partition copy.
            grid_route = _GRID_ROUTE and g % 32 == 0 and win == 14 and bw.rel_table is not None
            if grid_route:
                y = torch.empty((rows, d), dtype=torch.bfloat16, device=dev)
                ops.add_layernorm(x, delta, bw.ln1_w, bw.ln1_b, spec.eps, rows=rows, d=d,
                                  x_out=x if delta is not None else None, y_out=y)
            else:
                y = torch.empty((r_att, d), dtype=torch.bfloat16, device=dev)
                ops.add_layernorm(x, delta, bw.ln1_w, bw.ln1_b, spec.eps, rows=r_att, d=d,
                                  x_out=x if delta is not None else None, y_out=y,
                                  map_mode=1, win=win, nwin=nwin, hw=g)
        else:
            grid_route = False
            win, nwin, seq_len, n_seq, r_att = 0, 0, T, n_img, rows
            y = torch.empty((rows, d), dtype=torch.bfloat16, device=dev)
            ops.add_layernorm(x, delta, bw.ln1_w, bw.ln1_b, spec.eps, rows=rows, d=d,
                              x_out=x if delta is not None else None, y_out=y)
        if grid_route:
            q = ops.gemm_to_grid(y, bw.wq, bw.bq, g, nwin * win)
            kv = ops.gemm_to_grid(y, bw.wkv, bw.bkv, g, nwin * win)
        else:
            q = ops.gemm(y, bw.wq, bw.bq)
            kv = ops.gemm(y, bw.wkv, bw.bkv)
        del y
        att = torch.empty((rows, d), dtype=torch.bfloat16, device=dev)
        if bw.rel_table is not None and win > 0:
            # windowed block: the rel-pos table products are formed inside the attention kernel
            assert win == 14, "native windowed attention is built for 14x14 windows"
            ops.attention_window(q, kv, n_seq, heads, scale, att, 0, 0, d, bw.rel_table, bw.rel_pad, out_mode=1,
                                 nwin=nwin, img_hw=g, in_pad=nwin * win if grid_route else 0)
        else:
            bias_h = bias_w = None
            grid_hw = 0
            if bw.rel_table is not None:
                P = bw.rel_pad
                tab = ops.gemm(q.view(r_att * heads, 64), bw.rel_table, None, out_dtype=_TABLE_DTYPE)
                tab = tab.view(r_att, heads, 2 * P)
                bias_h, bias_w = tab[:, :, :P], tab[:, :, P:]
                grid_hw = g
            ops.attention(q, kv, n_seq, seq_len, heads, scale, att, 0, 0, d, bias_h, bias_w, grid_hw=grid_hw,
                          out_mode=1 if win > 0 else 0, nwin=nwin, img_hw=g)
            del bias_h, bias_w
        del q, kv
        if _FUSE_PROJ and ops.gemm_accumulate_supported(rows, d):
            ops.gemm_accumulate(att, bw.wproj, bw.bproj, x)
            delta = None
        else:
            delta = ops.gemm(att, bw.wproj, bw.bproj)
        del att
        y2 = torch.empty((rows, d), dtype=torch.bfloat16, device=dev)
        ops.add_layernorm(x, delta, bw.ln2_w, bw.ln2_b, spec.eps, rows=rows, d=d, x_out=x if delta is not None else None,
                          y_out=y2)
        h = ops.gemm(y2, bw.w1, bw.b1, act=ops.ACT_GELU)
        del y2
        if ops.gemm_accumulate_supported(rows, d):
            # x += h @ W2^T + b in the GEMM epilogue (fp32, on the accumulator): lin2's mainloop is long enough to hide
            # the residual stream's read and write, and the next LayerNorm only reads x
            ops.gemm_accumulate(h, bw.w2, bw.b2, x)
            delta = None
        else:
            delta = ops.gemm(h, bw.w2, bw.b2)
        del h
    out = torch.empty((n_img * g * g, d), dtype=out_dtype, device=dev)
    if spec.n_cls:
        ops.add_layernorm(x, delta, spec.final_ln_w, spec.final_ln_b, spec.eps, rows=rows, d=d, y_out=out,
                          map_mode=2, seq_len=T)
    else:
        ops.add_layernorm(x, delta, spec.final_ln_w, spec.final_ln_b, spec.eps, rows=rows, d=d, y_out=out)
    return out


@dataclass
class NeckWeights:
    w1: torch.Tensor    # [Co, Ci] bf16 (1x1 conv, no bias)
    ln1_w: torch.Tensor
    ln1_b: torch.Tensor
    w3: torch.Tensor    # [Co, 9*Co] bf16, column = (ky*3+kx)*Co + ci
    ln2_w: torch.Tensor
    ln2_b: torch.Tensor
    eps: float = 1e-6


def pack_neck(mod, seq) -> NeckWeights:
    """seq = nn.Sequential(Conv2d 1x1, LayerNorm2d, Conv2d 3x3, LayerNorm2d)  (image_encoder.py:92-108,
    build_lam.py:150-171).  `mod` owns the cache."""
    from .common import bf16_weight, f32

    c1, n1, c3, n2 = seq[0], seq[1], seq[2], seq[3]
    w3 = mod.packed("neck.w3:" + str(id(seq)),
                    lambda: c3.weight.detach().permute(0, 2, 3, 1).reshape(c3.weight.shape[0], -1)
                    .to(torch.bfloat16).contiguous(), c3.weight)
    return NeckWeights(bf16_weight(mod, "neck.0:" + str(id(seq)), c1.weight), f32(mod, "neck.1w:" + str(id(seq)), n1.weight),
                       f32(mod, "neck.1b:" + str(id(seq)), n1.bias), w3, f32(mod, "neck.3w:" + str(id(seq)), n2.weight),
                       f32(mod, "neck.3b:" + str(id(seq)), n2.bias), n1.eps)


def run_neck(nw: NeckWeights, tokens: torch.Tensor, n_img: int, g: int, out_dtype: torch.dtype) -> torch.Tensor:
    """tokens bf16 [n_img*g*g, Ci] -> [n_img*g*g, Co]: 1x1 conv GEMM -> channel LN -> im2col 3x3 + GEMM -> channel LN."""
    assert tokens.dtype == torch.bfloat16
    rows = n_img * g * g
    co = nw.w1.shape[0]
    t1 = ops.gemm(tokens, nw.w1, None, out_dtype=torch.float32)
    y1 = torch.empty((rows, co), dtype=torch.bfloat16, device=tokens.device)
    ops.add_layernorm(t1, None, nw.ln1_w, nw.ln1_b, nw.eps, rows=rows, d=co, y_out=y1)
    del t1
    if ops.conv3x3_supported(g, g, co, nw.w3.shape[0]):
        t2 = ops.conv3x3(y1, nw.w3, None, n_img, g, g, co, out_dtype=torch.float32)   # implicit GEMM, no im2col
    else:
        col = ops.im2col_3x3(y1, n_img, g, g, co)
        t2 = ops.gemm(col, nw.w3, None, out_dtype=torch.float32)
        del col
    out = torch.empty((rows, co), dtype=out_dtype, device=tokens.device)
    ops.add_layernorm(t2, None, nw.ln2_w, nw.ln2_b, nw.eps, rows=rows, d=co, y_out=out)
    return out


def tokens_to_nchw(tokens: torch.Tensor, n_img: int, g: int) -> torch.Tensor:
    """[n_img*g*g, C] -> [n_img, C, g, g] (layout change for callers that expect the reference's NCHW output)."""
    return ops.tokens_to_nchw(tokens if tokens.dtype == torch.float32 else tokens.float(), n_img, g, g)
