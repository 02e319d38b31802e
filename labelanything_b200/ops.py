"""Thin torch-tensor wrappers over the C ABI (include/labelanything_b200.h).

PyTorch is used for device memory and the current stream only; every function here validates shapes/dtypes,
hands raw pointers to the native library and raises RuntimeError on failure.  Nothing falls back to torch math.
"""
from __future__ import annotations

import torch

from . import _native

ACT_NONE, ACT_GELU, ACT_RELU = 0, 1, 2
DT_BF16, DT_F32 = 0, 1


def _stream(t: torch.Tensor) -> int:
    return torch.cuda.current_stream(t.device).cuda_stream


def _require_cuda(*ts: torch.Tensor) -> None:
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError(
                "labelanything_b200 runs on CUDA (sm_100a) tensors only; got a tensor on "
                f"{t.device}. There is no CPU fallback."
            )


def gemm(a: torch.Tensor, w: torch.Tensor, bias: torch.Tensor | None = None, act: int = ACT_NONE,
         out_dtype: torch.dtype = torch.bfloat16, out: torch.Tensor | None = None) -> torch.Tensor:
    """out = act(a @ w.T + bias).  a [M,K] bf16 (row stride free), w [N,K] bf16, bias [N] fp32."""
    _require_cuda(a, w, bias, out)
    assert a.dtype == torch.bfloat16 and w.dtype == torch.bfloat16, (a.dtype, w.dtype)
    assert a.dim() == 2 and w.dim() == 2 and a.shape[1] == w.shape[1], (a.shape, w.shape)
    assert a.stride(1) == 1 and w.stride(1) == 1
    M, K = a.shape
    N = w.shape[0]
    if out is None:
        out = torch.empty((M, N), dtype=out_dtype, device=a.device)
    assert out.shape == (M, N) and out.stride(1) == 1 and out.dtype in (torch.bfloat16, torch.float32)
    if bias is not None:
        assert bias.dtype == torch.float32 and bias.numel() == N and bias.is_contiguous()
    rc = _native.lib().la_gemm_bf16(
        _stream(a), a.data_ptr(), a.stride(0), w.data_ptr(), w.stride(0),
        bias.data_ptr() if bias is not None else None, out.data_ptr(), out.stride(0),
        DT_BF16 if out.dtype == torch.bfloat16 else DT_F32, M, N, K, act)
    _native.check(rc, "gemm")
    return out


def _ptr(t):
    return t.data_ptr() if t is not None else None


def attention(q: torch.Tensor, kv: torch.Tensor, n_seq: int, seq_len: int, n_heads: int, scale: float,
              out: torch.Tensor, q_off: int, k_off: int, v_off: int, bias_h: torch.Tensor | None = None,
              bias_w: torch.Tensor | None = None, grid_hw: int = 0, out_mode: int = 0, nwin: int = 0,
              img_hw: int = 0) -> torch.Tensor:
    """Fused MHSA (head_dim 64); q [rows, ld_q], kv [rows, ld_kv] bf16 (may be one packed buffer).
    bias_h / bias_w: fp32 views [rows, heads, >=2g-1] sharing one row stride (see the C header)."""
    _require_cuda(q, kv, out, bias_h, bias_w)
    for t in (q, kv, out):
        assert t.dtype == torch.bfloat16 and t.dim() == 2 and t.stride(1) == 1
    assert q.shape[0] == kv.shape[0]
    ldb = 0
    if bias_h is not None:
        assert bias_h.dtype == torch.float32 and bias_w.dtype == torch.float32
        assert bias_h.dim() == 3 and bias_h.shape[:2] == (q.shape[0], n_heads) and bias_w.shape[:2] == bias_h.shape[:2]
        assert bias_h.stride(2) == 1 and bias_w.stride(2) == 1 and bias_h.stride() == bias_w.stride()
        ldb = bias_h.stride(1)
        assert bias_h.stride(0) == ldb * n_heads
    rc = _native.lib().la_attention_bf16(
        _stream(q), q.data_ptr(), q.stride(0), q_off, kv.data_ptr(), kv.stride(0), k_off, v_off, q.shape[0], n_seq,
        seq_len, n_heads, float(scale), _ptr(bias_h), _ptr(bias_w), ldb, grid_hw, out.data_ptr(), out.stride(0),
        out_mode, nwin, img_hw)
    _native.check(rc, "attention")
    return out


def add_layernorm(x_in: torch.Tensor | None, delta: torch.Tensor | None, gamma: torch.Tensor | None,
                  beta: torch.Tensor | None, eps: float, *, rows: int, d: int, x_out: torch.Tensor | None = None,
                  y_out: torch.Tensor | None = None, y2_out: torch.Tensor | None = None,
                  pe: torch.Tensor | None = None, ype_out: torch.Tensor | None = None, act: int = ACT_NONE,
                  x_mod: int = 0, map_mode: int = 0, seq_len: int = 0, win: int = 0, nwin: int = 0,
                  hw: int = 0) -> None:
    """x = x_in + delta (-> x_out); y = act(LN(x)) -> y_out (bf16/fp32), y2_out (fp32), ype_out (bf16, y + pe)."""
    _require_cuda(x_in, delta, gamma, beta, x_out, y_out, y2_out, pe, ype_out)
    ref = x_in if x_in is not None else delta
    for t in (x_in, x_out, gamma, beta, y2_out, pe):
        assert t is None or (t.dtype == torch.float32 and t.is_contiguous())
    for t in (delta, ype_out):
        assert t is None or (t.dtype == torch.bfloat16 and t.is_contiguous())
    assert y_out is None or (y_out.is_contiguous() and y_out.dtype in (torch.bfloat16, torch.float32))
    pe_mod = 0
    if ype_out is not None:
        assert pe is not None and pe.shape[-1] == d
        pe_mod = pe.numel() // d
    rc = _native.lib().la_add_layernorm(
        _stream(ref), _ptr(x_in), x_mod, _ptr(delta), _ptr(x_out), _ptr(gamma), _ptr(beta), float(eps), act,
        _ptr(y_out), DT_F32 if (y_out is not None and y_out.dtype == torch.float32) else DT_BF16, _ptr(y2_out),
        _ptr(pe), pe_mod, _ptr(ype_out), rows, d, map_mode, seq_len, win, nwin, hw)
    _native.check(rc, "add_layernorm")


def add_layernorm_meanpool(x_in: torch.Tensor | None, delta: torch.Tensor | None, gamma: torch.Tensor,
                           beta: torch.Tensor, eps: float, n_seq: int, rows_per_seq: int, d: int,
                           slices: int = 8) -> torch.Tensor:
    """mean over each sequence's rows of LN(x_in + delta) -> fp32 [n_seq, d]."""
    _require_cuda(x_in, delta, gamma, beta)
    ref = x_in if x_in is not None else delta
    assert x_in is None or (x_in.dtype == torch.float32 and x_in.is_contiguous())
    assert delta is None or (delta.dtype == torch.bfloat16 and delta.is_contiguous())
    ws = torch.empty((n_seq * slices, d), dtype=torch.float32, device=ref.device)
    out = torch.empty((n_seq, d), dtype=torch.float32, device=ref.device)
    rc = _native.lib().la_add_layernorm_meanpool(_stream(ref), _ptr(x_in), _ptr(delta), gamma.data_ptr(),
                                                 beta.data_ptr(), float(eps), n_seq, rows_per_seq, d,
                                                 ws.data_ptr(), slices, out.data_ptr())
    _native.check(rc, "add_layernorm_meanpool")
    return out


def embed_tokens(patch: torch.Tensor, cls: torch.Tensor | None, pos: torch.Tensor | None, x: torch.Tensor,
                 n_img: int, tokens_per_img: int, n_cls: int, d: int) -> torch.Tensor:
    _require_cuda(patch, cls, pos, x)
    assert patch.dtype == torch.bfloat16 and patch.is_contiguous() and x.dtype == torch.float32 and x.is_contiguous()
    rc = _native.lib().la_embed_tokens(_stream(x), patch.data_ptr(), _ptr(cls), _ptr(pos), x.data_ptr(), n_img,
                                       tokens_per_img, n_cls, d)
    _native.check(rc, "embed_tokens")
    return x


def im2col_patch16(images: torch.Tensor, out: torch.Tensor | None = None) -> torch.Tensor:
    """images [I, C, S, S] fp32 -> [I*(S/16)^2, C*256] bf16."""
    _require_cuda(images, out)
    assert images.dtype == torch.float32 and images.is_contiguous() and images.dim() == 4
    I, C, S, S2 = images.shape
    assert S == S2 and S % 16 == 0
    if out is None:
        out = torch.empty((I * (S // 16) ** 2, C * 256), dtype=torch.bfloat16, device=images.device)
    rc = _native.lib().la_im2col_patch16(_stream(images), images.data_ptr(), out.data_ptr(), I, C, S)
    _native.check(rc, "im2col_patch16")
    return out


def im2col_3x3(x: torch.Tensor, n_img: int, h: int, w: int, c: int, out: torch.Tensor | None = None) -> torch.Tensor:
    """token-major bf16 [n_img*h*w, c] -> [n_img*h*w, 9c] bf16 (zero pad 1)."""
    _require_cuda(x, out)
    assert x.dtype == torch.bfloat16 and x.is_contiguous() and x.numel() == n_img * h * w * c
    if out is None:
        out = torch.empty((n_img * h * w, 9 * c), dtype=torch.bfloat16, device=x.device)
    rc = _native.lib().la_im2col_3x3(_stream(x), x.data_ptr(), out.data_ptr(), n_img, h, w, c)
    _native.check(rc, "im2col_3x3")
    return out
