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
