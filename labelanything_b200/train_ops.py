"""Differentiable native ops of the training step.

Every class below is a `torch.autograd.Function` whose forward AND backward are launches of the C ABI
(include/labelanything_b200.h, "training step" section; csrc/la_train.cu + the tcgen05 GEMM): torch.autograd only
records the graph and routes the gradients -- what the reference gets from autograd over its torch modules
(label_anything/experiment/run.py:359-361).  Activations travel as contiguous fp32 [rows, channels] (token-major)
tensors; GEMM operands are rounded to bf16 exactly where the inference path rounds them, the accumulation is fp32.

dgrad / wgrad of a Linear y = x W^T + b over M rows:
    dx = dy W                  la_gemm_bf16(a = bf16(dy) [M, N],  w = W^T [K, N])
    dW = dy^T x                la_gemm_bf16(a = dy^T [N, M8],     w = x^T [K, M8])       (M8 = M padded to a multiple of 8)
    db = column sum of dy      la_bcast_reduce_f32
There is no CPU / torch fallback: CPU tensors raise.
"""
from __future__ import annotations

import weakref
from typing import Optional

import torch
from torch.autograd import Function

from . import ops
from .ops import ACT_GELU, ACT_NONE, ACT_RELU, DT_BF16, DT_F32, _call, _require_cuda

_once = torch.autograd.function.once_differentiable


# ----------------------------------------------------------------------------------------------------------------
# raw wrappers (allocate outputs, validate, launch)
# ----------------------------------------------------------------------------------------------------------------
def _f32c(t: torch.Tensor) -> torch.Tensor:
    assert t.dtype == torch.float32, t.dtype
    return t if t.is_contiguous() else t.contiguous()


def cast_bf16(x: torch.Tensor) -> torch.Tensor:
    _require_cuda(x)
    x = _f32c(x)
    out = torch.empty(x.shape, dtype=torch.bfloat16, device=x.device)
    _call("cast_bf16", "la_cast_bf16", x, out, x.numel())
    return out


def add_f32(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    _require_cuda(a, b)
    a, b = _f32c(a), _f32c(b)
    assert a.shape == b.shape, (a.shape, b.shape)
    out = torch.empty_like(a)
    _call("add_f32", "la_add_f32", a, b, out, a.numel())
    return out


def cast_transpose(x: torch.Tensor) -> torch.Tensor:
    """fp32 / bf16 [R, C] -> bf16 [C, R8] (R8 = R rounded up to a multiple of 8, the padding columns zero)."""
    _require_cuda(x)
    assert x.dim() == 2 and x.stride(1) == 1 and x.dtype in (torch.float32, torch.bfloat16)
    R, C = x.shape
    R8 = (R + 7) // 8 * 8
    out = torch.empty((C, R8), dtype=torch.bfloat16, device=x.device)
    _call("cast_transpose", "la_cast_transpose_bf16", x, DT_F32 if x.dtype == torch.float32 else DT_BF16, x.stride(0),
          out, R8, R, C)
    return out


def bcast_reduce(dy: torch.Tensor, row_div: int = 1, b_mod: int = 1, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """out[(r // row_div) % b_mod] (+)= dy[r]; fp32 [rows, d] -> [b_mod, d]."""
    _require_cuda(dy, out)
    dy = _f32c(dy)
    rows, d = dy.shape
    acc = out is not None
    if out is None:
        out = torch.empty((b_mod, d), dtype=torch.float32, device=dy.device)
    _call("bcast_reduce", "la_bcast_reduce_f32", dy, out, rows, d, row_div, b_mod, 1 if acc else 0)
    return out


# GEMM precision of the training path.  "bf16": operands rounded to bf16 (what the inference path does).  "bf16x3":
# every fp32 operand is split into two bf16 terms and the product is three tensor-core GEMMs with fp32 accumulation
# (hi*hi + hi*lo + lo*hi; the dropped lo*lo term is 2^-16 relative) -- fp32-accurate gradients at three times the GEMM
# work.  "bf16x6": three terms (24 mantissa bits), the six products whose orders sum to <= 2: fp32-level contractions.
# An operand is a tuple of bf16 tensors: (hi,), (hi, lo) or (hi, mid, lo).
_PRECISION = "bf16"


class precision:
    """with train_ops.precision("bf16x3"): ...   (captured at forward time; the backward of an op uses the same mode)"""

    def __init__(self, mode: str) -> None:
        assert mode in ("bf16", "bf16x3", "bf16x6"), mode
        self.mode = mode

    def __enter__(self):
        global _PRECISION
        self.prev, _PRECISION = _PRECISION, self.mode
        return self

    def __exit__(self, *exc):
        global _PRECISION
        _PRECISION = self.prev


# Operands derived from a tensor (its bf16 split, the transposes of those) are shared between the ops that consume the
# same tensor -- `keys + pe` feeds two projections per two-way layer, the input of an attention block three; each of
# them would cast it in the forward and transpose it in the backward pass again.  Entries are keyed on the storage
# address + shape + version of the (contiguous) source and dropped when the source tensor dies, so an address the
# allocator hands out again can never hit a stale entry.
_DERIVED: dict = {}


def _derived(tag: str, src: torch.Tensor, build):
    key = (tag, src.data_ptr(), tuple(src.shape), src.dtype, src._version)
    hit = _DERIVED.get(key)
    if hit is not None and hit[0]() is not None:
        return hit[1]
    val = build()
    try:
        ref = weakref.ref(src, lambda _r, k=key: _DERIVED.pop(k, None))
    except TypeError:
        return val
    _DERIVED[key] = (ref, val)
    return val


def _split(x: torch.Tensor, mode: Optional[str] = None) -> tuple:
    """fp32 tensor -> GEMM operand (tuple of bf16 tensors of the same shape)."""
    mode = mode or _PRECISION
    _require_cuda(x)
    x = _f32c(x)

    def build():
        if mode == "bf16":
            return (cast_bf16(x),)
        parts = tuple(torch.empty(x.shape, dtype=torch.bfloat16, device=x.device) for _ in range(2 if mode == "bf16x3" else 3))
        _call("split_bf16", "la_split_bf16" if mode == "bf16x3" else "la_split3_bf16", x, *parts, x.numel())
        return parts

    return _derived("split:" + mode, x, build)


def _tr(op: tuple) -> tuple:
    return tuple(_derived("T", t, lambda t=t: cast_transpose(t)) for t in op)


def _relu_f32(x: torch.Tensor) -> torch.Tensor:
    y = torch.empty_like(x)
    _call("relu", "la_relu_bwd_f32", x, x, y, x.numel())        # y = x where x > 0 else 0
    return y


def grad_prep(dy: torch.Tensor, want16: bool, want_t: bool, want_sum: bool):
    """(bf16(dy), bf16(dy)^T padded to 8 rows, column sums) of an fp32 [rows, cols] gradient in one pass."""
    _require_cuda(dy)
    dy = _f32c(dy)
    rows, cols = dy.shape
    r8 = (rows + 7) // 8 * 8
    o16 = torch.empty((rows, cols), dtype=torch.bfloat16, device=dy.device) if want16 else None
    ot = torch.empty((cols, r8), dtype=torch.bfloat16, device=dy.device) if want_t else None
    cs = torch.empty(cols, dtype=torch.float32, device=dy.device) if want_sum else None
    _call("grad_prep", "la_grad_prep_bf16", dy, dy.stride(0), o16, ot, r8, cs, rows, cols)
    return o16, ot, cs


def _backward_operands(dy: torch.Tensor, mode: str, need_dx: bool, need_dw: bool, need_db: bool):
    """-> (dy as a GEMM operand | None, its transpose as a GEMM operand | None, bias gradient | None).  bf16 mode: one
    fused pass over dy; split-operand modes: the generic split / transpose / column-sum kernels."""
    if mode == "bf16":
        o16, ot, cs = grad_prep(dy, need_dx, need_dw, need_db)
        return (None if o16 is None else (o16,)), (None if ot is None else (ot,)), cs
    dyb = _split(dy, mode) if (need_dx or need_dw) else None
    return (dyb if need_dx else None), (_tr(dyb) if need_dw else None), (bcast_reduce(dy).view(-1) if need_db else None)


def _gemm_f32(a16: torch.Tensor, w16: torch.Tensor, bias=None, act=ACT_NONE) -> torch.Tensor:
    """One fp32-output GEMM; products with a long contraction and too few output tiles to fill the GPU (the weight
    gradients over the image-token rows: [N_out, 54 000] x [K_in, 54 000]^T) go through the split-K kernel."""
    M, K = a16.shape
    N = w16.shape[0]
    tiles = -(-M // 128) * -(-N // (256 if N >= 256 else 128 if N > 64 else 64))
    # K >= 8192 singles out the weight gradients (contraction over token rows); forward and data-gradient products
    # (K <= 2304 features) stay on the unsplit kernel, whose summation order is fixed: the forward pass is reproducible
    if act == ACT_NONE and K >= 8192 and tiles <= 32:
        return ops.gemm_splitk(a16, w16, bias)
    return ops.gemm(a16, w16, bias, act=act, out_dtype=torch.float32)


def _mm(a: tuple, w: tuple, bias=None, act=ACT_NONE) -> torch.Tensor:
    """fp32 [M, N] = act(a @ w^T + bias) for operands a [M, K], w [N, K]."""
    if len(a) == 1 and len(w) == 1:
        return _gemm_f32(a[0], w[0], bias, act)
    # every product a_i w_j^T with i + j <= order (3 of 4 for two terms, 6 of 9 for three), the small ones summed first
    order = max(len(a), len(w)) - 1
    out = None
    for total in range(order, 0, -1):
        for i in range(len(a)):
            j = total - i
            if 0 <= j < len(w):
                t = _gemm_f32(a[i], w[j])
                out = t if out is None else add_f32(out, t)
    out = add_f32(_gemm_f32(a[0], w[0], bias), out)
    if act == ACT_RELU:
        out = _relu_f32(out)
    else:
        assert act == ACT_NONE
    return out


# ----------------------------------------------------------------------------------------------------------------
# Linear (+ ReLU): nn.Linear, 1x1 Conv2d, ConvTranspose2d(k = s = 2) as a per-pixel GEMM
# ----------------------------------------------------------------------------------------------------------------
class _Linear(Function):
    @staticmethod
    def forward(ctx, x, w, b, act):
        _require_cuda(x, w, b)
        assert x.dim() == 2 and w.dim() == 2 and x.shape[1] == w.shape[1], (x.shape, w.shape)
        assert act in (ACT_NONE, ACT_RELU)
        N, K = w.shape
        assert N % 8 == 0 and K % 8 == 0, "native Linear needs in / out features that are multiples of 8"
        xb, wb = _split(x), _split(w)
        y = _mm(xb, wb, None if b is None else _f32c(b), act)
        ctx.act = act
        ctx.has_bias = b is not None
        ctx.mode = _PRECISION
        ctx.save_for_backward(y if act == ACT_RELU else None, *xb, *wb)
        return y

    @staticmethod
    @_once
    def backward(ctx, dy):
        y, *parts = ctx.saved_tensors
        n = len(parts) // 2
        xb, wb = tuple(parts[:n]), tuple(parts[n:])
        dy = _f32c(dy)
        if ctx.act == ACT_RELU:
            g = torch.empty_like(dy)
            _call("relu_bwd", "la_relu_bwd_f32", dy, y, g, dy.numel())
            dy = g
        need = ctx.needs_input_grad
        dyb, dyt, db = _backward_operands(dy, ctx.mode, need[0], need[1], ctx.has_bias and need[2])
        dx = dw = None
        if need[0]:
            dx = _mm(dyb, _tr(wb))                                    # [M, N] @ [K, N]^T
        if need[1]:
            dw = _mm(dyt, _tr(xb))                                    # [N, M8] @ [K, M8]^T
        return dx, dw, db, None


def linear(x: torch.Tensor, w: torch.Tensor, b: Optional[torch.Tensor], act: int = ACT_NONE) -> torch.Tensor:
    """fp32 [M, K] x fp32 [N, K] (+ [N]) -> fp32 [M, N] = act(x w^T + b), bf16 operands / fp32 accumulation."""
    return _Linear.apply(x, w.reshape(w.shape[0], -1), b, act)


# ----------------------------------------------------------------------------------------------------------------
# 3x3 convolution (zero pad 1) on a token-major map as im2col + GEMM; backward = GEMMs + col2im
# ----------------------------------------------------------------------------------------------------------------
class _Conv3x3(Function):
    @staticmethod
    def forward(ctx, x, wmat, b, n_img, h, w):
        _require_cuda(x, wmat, b)
        c = x.shape[1]
        assert x.shape[0] == n_img * h * w and wmat.shape[1] == 9 * c and c % 8 == 0 and wmat.shape[0] % 8 == 0
        col = tuple(ops.im2col_3x3(t, n_img, h, w, c) for t in _split(x))      # im2col is linear: split first
        wb = _split(wmat)
        y = _mm(col, wb, None if b is None else _f32c(b))
        ctx.geom = (n_img, h, w, c)
        ctx.has_bias = b is not None
        ctx.mode = _PRECISION
        ctx.save_for_backward(*col, *wb)
        return y

    @staticmethod
    @_once
    def backward(ctx, dy):
        parts = ctx.saved_tensors
        n = len(parts) // 2
        col, wb = tuple(parts[:n]), tuple(parts[n:])
        n_img, h, w, c = ctx.geom
        dy = _f32c(dy)
        need = ctx.needs_input_grad
        dyb, dyt, db = _backward_operands(dy, ctx.mode, need[0], need[1], ctx.has_bias and need[2])
        dx = dw = None
        if need[0]:
            dcol = _mm(dyb, _tr(wb))                                  # [rows, 9c] fp32
            dx = torch.empty((n_img * h * w, c), dtype=torch.float32, device=dy.device)
            _call("col2im_3x3", "la_col2im_3x3_f32", dcol, dx, n_img, h, w, c)
        if need[1]:
            dw = _mm(dyt, _tr(col))
        return dx, dw, db, None, None, None


def conv3x3(x: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor], n_img: int, h: int, w: int) -> torch.Tensor:
    """x fp32 [n_img*h*w, Ci]; weight = the Conv2d parameter [Co, Ci, 3, 3] -> fp32 [n_img*h*w, Co]."""
    wmat = weight.permute(0, 2, 3, 1).reshape(weight.shape[0], -1)    # column = (ky*3 + kx)*Ci + ci, as la_im2col_3x3
    return _Conv3x3.apply(x, wmat, bias, n_img, h, w)


# ----------------------------------------------------------------------------------------------------------------
# LayerNorm (+ GELU), GELU, adds, row permutation
# ----------------------------------------------------------------------------------------------------------------
class _LayerNorm(Function):
    @staticmethod
    def forward(ctx, x, gamma, beta, eps, act):
        _require_cuda(x, gamma, beta)
        x = _f32c(x)
        rows, d = x.shape
        y = torch.empty_like(x)
        g, b = _f32c(gamma), _f32c(beta)
        _call("layernorm_f32", "la_layernorm_f32", x, g, b, float(eps), act, y, rows, d)
        ctx.cfg = (float(eps), act)
        ctx.save_for_backward(x, g, b)
        return y

    @staticmethod
    @_once
    def backward(ctx, dy):
        x, g, b = ctx.saved_tensors
        eps, act = ctx.cfg
        rows, d = x.shape
        dy = _f32c(dy)
        dx = torch.empty_like(x)
        dg = torch.zeros(d, dtype=torch.float32, device=x.device)
        db = torch.zeros(d, dtype=torch.float32, device=x.device)
        _call("layernorm_f32_bwd", "la_layernorm_f32_bwd", x, g, b, eps, act, dy, dx, dg, db, rows, d)
        return dx, dg, db, None, None


def layernorm(x, gamma, beta, eps: float, act: int = ACT_NONE) -> torch.Tensor:
    return _LayerNorm.apply(x, gamma, beta, eps, act)


class _Gelu(Function):
    @staticmethod
    def forward(ctx, x):
        _require_cuda(x)
        x = _f32c(x)
        y = torch.empty_like(x)
        _call("gelu_f32", "la_gelu_f32", x, y, x.numel())
        ctx.save_for_backward(x)
        return y

    @staticmethod
    @_once
    def backward(ctx, dy):
        (x,) = ctx.saved_tensors
        dx = torch.empty_like(x)
        _call("gelu_bwd", "la_gelu_bwd_f32", _f32c(dy), x, dx, x.numel())
        return dx


def gelu(x: torch.Tensor) -> torch.Tensor:
    return _Gelu.apply(x)


class _Add(Function):
    @staticmethod
    def forward(ctx, a, b):
        return add_f32(a, b)

    @staticmethod
    @_once
    def backward(ctx, g):
        return g, g


def add(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    return _Add.apply(a, b)


class _AddBcast(Function):
    @staticmethod
    def forward(ctx, a, b, row_div, b_mod):
        ctx.cfg = (row_div, b_mod)
        return ops.add_bcast(_f32c(a), _f32c(b), row_div, b_mod)

    @staticmethod
    @_once
    def backward(ctx, g):
        row_div, b_mod = ctx.cfg
        db = bcast_reduce(g, row_div, b_mod) if ctx.needs_input_grad[1] else None
        return g, db, None, None


def add_bcast(a: torch.Tensor, b: torch.Tensor, row_div: int, b_mod: int) -> torch.Tensor:
    """a[r] + b[(r // row_div) % b_mod]; a fp32 [rows, d], b fp32 [b_mod, d]."""
    return _AddBcast.apply(a, b, row_div, b_mod)


class _PermuteRows(Function):
    @staticmethod
    def forward(ctx, x, outer, na, nb):
        ctx.cfg = (outer, na, nb)
        return ops.permute_rows(_f32c(x), outer, na, nb)

    @staticmethod
    @_once
    def backward(ctx, g):
        outer, na, nb = ctx.cfg
        return ops.permute_rows(_f32c(g), outer, nb, na), None, None, None


def permute_rows(x: torch.Tensor, outer: int, na: int, nb: int) -> torch.Tensor:
    """rows (o, i, j) -> (o, j, i) of fp32 [outer*na*nb, d]."""
    return _PermuteRows.apply(x, outer, na, nb)


# ----------------------------------------------------------------------------------------------------------------
# attention after the projections
# ----------------------------------------------------------------------------------------------------------------
class _Attention(Function):
    @staticmethod
    def forward(ctx, q, k, v, n_seq, nq, nk, heads):
        _require_cuda(q, k, v)
        q, k, v = _f32c(q), _f32c(k), _f32c(v)
        width = q.shape[1]
        dh = width // heads
        assert q.shape == (n_seq * nq, width) and k.shape == (n_seq * nk, width) and v.shape == k.shape
        out = torch.empty_like(q)
        lse = torch.empty((n_seq, heads, nq), dtype=torch.float32, device=q.device)
        scale = dh ** -0.5
        _call(f"attention_f32.s{n_seq}.q{nq}.k{nk}.d{dh}" if ops._PROF is not None else "attention_f32", "la_attention_f32",
              q, k, v, out, lse, n_seq, nq, nk, heads, dh, scale)
        ctx.cfg = (n_seq, nq, nk, heads, dh, scale)
        ctx.save_for_backward(q, k, v, out, lse)
        return out

    @staticmethod
    @_once
    def backward(ctx, dout):
        q, k, v, out, lse = ctx.saved_tensors
        n_seq, nq, nk, heads, dh, scale = ctx.cfg
        dq, dk, dv = torch.empty_like(q), torch.empty_like(k), torch.empty_like(v)
        delta = torch.empty_like(lse)
        _call(f"attention_f32_bwd.s{n_seq}.q{nq}.k{nk}.d{dh}" if ops._PROF is not None else "attention_f32_bwd",
              "la_attention_f32_bwd", q, k, v, out, lse, _f32c(dout), delta, dq, dk, dv, n_seq, nq, nk, heads, dh, scale)
        return dq, dk, dv, None, None, None, None


def attention(q, k, v, n_seq: int, nq: int, nk: int, heads: int) -> torch.Tensor:
    """softmax(q k^T / sqrt(dh)) v per (sequence, head); fp32 [n_seq*nq | nk, heads*dh] (common.py:97-148)."""
    return _Attention.apply(q, k, v, n_seq, nq, nk, heads)


# ----------------------------------------------------------------------------------------------------------------
# prompt-encoder pieces
# ----------------------------------------------------------------------------------------------------------------
class _MaskDownscale(Function):
    @staticmethod
    def forward(ctx, masks, wpack, eps1, eps2):
        _require_cuda(masks, wpack)
        masks, wpack = _f32c(masks), _f32c(wpack)
        assert masks.dim() == 3 and wpack.numel() == 332, "mask_downscaling is built for mask_in_chans = 16"
        S, H, W = masks.shape
        out = torch.empty((S, H // 4, W // 4, 16), dtype=torch.float32, device=masks.device)
        _call("mask_downscale_dev", "la_mask_downscale_dev", masks, wpack, float(eps1), float(eps2), out, S, H, W)
        ctx.cfg = (float(eps1), float(eps2))
        ctx.save_for_backward(masks, wpack)
        return out

    @staticmethod
    @_once
    def backward(ctx, dout):
        masks, wpack = ctx.saved_tensors
        S, H, W = masks.shape
        dw = torch.empty_like(wpack)
        _call("mask_downscale_bwd", "la_mask_downscale_bwd", masks, wpack, ctx.cfg[0], ctx.cfg[1], _f32c(dout), dw, S, H, W)
        return None, dw, None, None


def mask_downscale(masks: torch.Tensor, md) -> torch.Tensor:
    """masks fp32 [S, H, W]; md = the `mask_downscaling` nn.Sequential -> fp32 [S, H/4, W/4, 16] (layers 0..5)."""
    wpack = torch.cat([md[0].weight.reshape(-1), md[0].bias, md[1].weight, md[1].bias, md[3].weight.reshape(-1),
                       md[3].bias, md[4].weight, md[4].bias])
    return _MaskDownscale.apply(masks, wpack, md[1].eps, md[4].eps)


class _ResizeBilinear(Function):
    @staticmethod
    def forward(ctx, x, out_h, out_w):
        x = _f32c(x)
        ctx.shape = tuple(x.shape)
        return ops.resize_bilinear(x, out_h, out_w)

    @staticmethod
    @_once
    def backward(ctx, dout):
        n, h, w, c = ctx.shape
        dout = _f32c(dout)
        din = torch.empty(ctx.shape, dtype=torch.float32, device=dout.device)
        _call("resize_bilinear_bwd", "la_resize_bilinear_bwd", dout, din, n, h, w, dout.shape[1], dout.shape[2], c)
        return din, None, None


def resize_bilinear(x: torch.Tensor, out_h: int, out_w: int) -> torch.Tensor:
    """token-major fp32 [n, h, w, c] -> [n, out_h, out_w, c]."""
    return _ResizeBilinear.apply(x, out_h, out_w)


class _SrcCombine(Function):
    @staticmethod
    def forward(ctx, feat, dense, alt, mflags, n_seq, T, D, C):
        _require_cuda(feat, dense, alt, mflags)
        feat, alt = _f32c(feat), _f32c(alt)
        dense = None if dense is None else _f32c(dense)
        assert feat.shape == (n_seq // C * T, D) and alt.numel() == D
        assert dense is None or dense.shape == (n_seq * T, D)
        out = torch.empty((n_seq * T, D), dtype=torch.float32, device=feat.device)
        _call("src_combine", "la_src_combine_f32", feat, dense, mflags, alt, out, n_seq, T, D, C)
        ctx.cfg = (n_seq, T, D, C, dense is not None)
        ctx.save_for_backward(mflags)
        return out

    @staticmethod
    @_once
    def backward(ctx, dsrc):
        (mflags,) = ctx.saved_tensors
        n_seq, T, D, C, has_dense = ctx.cfg
        dsrc = _f32c(dsrc)
        dev = dsrc.device
        dfeat = torch.empty((n_seq // C * T, D), dtype=torch.float32, device=dev)
        ddense = torch.empty((n_seq * T, D), dtype=torch.float32, device=dev) if has_dense else None
        dalt_rows = torch.empty_like(dfeat)
        _call("src_combine_bwd", "la_src_combine_bwd", dsrc, mflags, 1 if has_dense else 0, dfeat, ddense, dalt_rows,
              n_seq, T, D, C)
        dalt = bcast_reduce(dalt_rows).view(-1)
        return dfeat, ddense, dalt, None, None, None, None, None


def src_combine(feat, dense, alt, mflags, n_seq: int, T: int, D: int, C: int) -> torch.Tensor:
    """feat [n_seq/C*T, D] + (dense [n_seq*T, D] where mflags[s] (all if mflags is None) else alt [D])."""
    return _SrcCombine.apply(feat, dense, alt.reshape(-1), mflags, n_seq, T, D, C)


class _EmbedSparse(Function):
    @staticmethod
    def forward(ctx, tab4, nap, pts, lab, bx, bfl, gauss, n_seq, D, image_w, image_h):
        out = ops.embed_sparse(pts, lab, bx, bfl, gauss, _f32c(nap).view(-1), _f32c(tab4), n_seq, D, image_w, image_h)
        ctx.cfg = (n_seq, D, 0 if pts is None else pts.shape[1], 0 if bx is None else bx.shape[1], pts is not None)
        ctx.save_for_backward(lab, bfl)
        return out

    @staticmethod
    @_once
    def backward(ctx, dout):
        lab, bfl = ctx.saved_tensors
        n_seq, D, P, Bx, has_pts = ctx.cfg
        dout = _f32c(dout)
        dtab = torch.empty((4, D), dtype=torch.float32, device=dout.device)
        dnap = torch.empty((1, D), dtype=torch.float32, device=dout.device)
        _call("embed_sparse_bwd", "la_embed_sparse_bwd", lab, P, bfl, Bx, 1 if has_pts else 0, dout, dtab, dnap, n_seq, D)
        return dtab, dnap, None, None, None, None, None, None, None, None, None


def embed_sparse(tab4, nap, pts, lab, bx, bfl, gauss, n_seq: int, D: int, image_w: int, image_h: int) -> torch.Tensor:
    """-> fp32 [n_seq, n, D] sparse tokens; tab4 [4, D] = point_embeddings.0-3 stacked, nap [1, D] = not_a_point_embed."""
    return _EmbedSparse.apply(tab4, nap, pts, lab, bx, bfl, gauss, n_seq, D, image_w, image_h)


class _SegmentMean(Function):
    @staticmethod
    def forward(ctx, x, n_seg, seg_rows):
        x = _f32c(x)
        d = x.shape[1]
        assert x.shape[0] == n_seg * seg_rows
        out = torch.empty((n_seg, d), dtype=torch.float32, device=x.device)
        _call("segment_mean", "la_segment_mean_f32", x, out, n_seg, seg_rows, d)
        ctx.cfg = (n_seg, seg_rows, d)
        return out

    @staticmethod
    @_once
    def backward(ctx, dout):
        n_seg, seg_rows, d = ctx.cfg
        dout = _f32c(dout)
        dx = torch.empty((n_seg * seg_rows, d), dtype=torch.float32, device=dout.device)
        _call("segment_mean_bwd", "la_segment_mean_bwd", dout, dx, n_seg, seg_rows, d)
        return dx, None, None


def segment_mean(x: torch.Tensor, n_seg: int, seg_rows: int) -> torch.Tensor:
    return _SegmentMean.apply(x, n_seg, seg_rows)


class _MaskedMean(Function):
    @staticmethod
    def forward(ctx, emb, flags):
        ctx.save_for_backward(flags)
        ctx.shape = tuple(emb.shape)
        return ops.masked_mean(_f32c(emb), flags)

    @staticmethod
    @_once
    def backward(ctx, dout):
        (flags,) = ctx.saved_tensors
        B, M, C, D = ctx.shape
        dout = _f32c(dout)
        demb = torch.empty(ctx.shape, dtype=torch.float32, device=dout.device)
        _call("masked_mean_bwd", "la_masked_mean_bwd", dout, flags, demb, B, M, C, D)
        return demb, None


def masked_mean(emb: torch.Tensor, flags: torch.Tensor) -> torch.Tensor:
    """emb fp32 [B, M, C, D], flags uint8 [B, M, C] -> [B, C, D] (prompt_encoder.py:738-745)."""
    return _MaskedMean.apply(emb, flags)


# ----------------------------------------------------------------------------------------------------------------
# decoder head and post-processing
# ----------------------------------------------------------------------------------------------------------------
class _Classify(Function):
    @staticmethod
    def forward(ctx, x, cls, batch, pixels):
        x16 = _split(x)
        cls = _f32c(cls)
        ctx.cfg = (batch, pixels)
        ctx.save_for_backward(cls, *x16)
        out = ops.classify(x16[0], cls, batch, pixels)
        for t in x16[1:]:                                             # linear in x: the low-order term adds
            out = add_f32(out, ops.classify(t, cls, batch, pixels))
        return out

    @staticmethod
    @_once
    def backward(ctx, dl):
        cls, *x16 = ctx.saved_tensors
        batch, pixels = ctx.cfg
        C, dk = cls.shape[1], cls.shape[2]
        dl = _f32c(dl)
        dx = torch.empty(x16[0].shape, dtype=torch.float32, device=dl.device)
        dcls = torch.empty_like(cls)
        _call("classify_bwd", "la_classify_bwd", dl, x16[0], cls, dx, dcls, batch, pixels, C, dk)
        for t in x16[1:]:
            dx2, dcls2 = torch.empty_like(dx), torch.empty_like(cls)
            _call("classify_bwd", "la_classify_bwd", dl, t, cls, dx2, dcls2, batch, pixels, C, dk)
            dcls = add_f32(dcls, dcls2)
        return dx, dcls, None, None


def classify(x: torch.Tensor, cls: torch.Tensor, batch: int, pixels: int) -> torch.Tensor:
    """x fp32 [batch*pixels, dk], cls fp32 [batch, C, dk] -> logits fp32 [batch, C, pixels] (mask_decoder.py:309)."""
    return _Classify.apply(x, cls, batch, pixels)


class _Postprocess(Function):
    @staticmethod
    def forward(ctx, logits, sizes, fg, image_size, out_h, out_w):
        logits = _f32c(logits)
        ctx.cfg = (tuple(logits.shape), image_size, out_h, out_w)
        ctx.save_for_backward(sizes, fg)
        return ops.postprocess_masks(logits, sizes, fg, image_size, out_h, out_w)

    @staticmethod
    @_once
    def backward(ctx, dout):
        sizes, fg = ctx.saved_tensors
        (B, C, lh, lw), image_size, out_h, out_w = ctx.cfg
        dout = _f32c(dout)
        din = torch.empty((B, C, lh, lw), dtype=torch.float32, device=dout.device)
        _call("postprocess_masks_bwd", "la_postprocess_masks_bwd", dout, sizes, fg, din, B, C, lh, lw, image_size, out_h,
              out_w)
        return din, None, None, None, None, None


def postprocess_masks(logits, sizes, fg, image_size: int, out_h: int, out_w: int) -> torch.Tensor:
    return _Postprocess.apply(logits, sizes, fg, image_size, out_h, out_w)


# ----------------------------------------------------------------------------------------------------------------
# optimiser
# ----------------------------------------------------------------------------------------------------------------
def adamw_step(params: torch.Tensor, grads: torch.Tensor, exp_avg: torch.Tensor, exp_avg_sq: torch.Tensor, lr: float,
               beta1: float, beta2: float, eps: float, weight_decay: float, step: int, grad_scale: float = 1.0) -> None:
    """torch.optim.AdamW update of one flat fp32 bucket, in place."""
    _require_cuda(params, grads, exp_avg, exp_avg_sq)
    for t in (params, grads, exp_avg, exp_avg_sq):
        assert t.dtype == torch.float32 and t.is_contiguous() and t.numel() == params.numel()
    _call("adamw", "la_adamw_f32", params, grads, exp_avg, exp_avg_sq, params.numel(), float(lr), float(beta1), float(beta2),
          float(eps), float(weight_decay), int(step), float(grad_scale))


def adamw_step_dev(params: torch.Tensor, grads: torch.Tensor, exp_avg: torch.Tensor, exp_avg_sq: torch.Tensor, beta1: float,
                   beta2: float, eps: float, weight_decay: float, step_scalars: torch.Tensor, grad_scale: float = 1.0) -> None:
    """`adamw_step` with {1 - beta1^t, sqrt(1 - beta2^t), lr} read from the device tensor `step_scalars` (fp32 [3])."""
    _require_cuda(params, grads, exp_avg, exp_avg_sq, step_scalars)
    assert step_scalars.dtype == torch.float32 and step_scalars.numel() == 3 and step_scalars.is_contiguous()
    _call("adamw", "la_adamw_f32_dev", params, grads, exp_avg, exp_avg_sq, params.numel(), float(beta1), float(beta2),
          float(eps), float(weight_decay), step_scalars, float(grad_scale))
