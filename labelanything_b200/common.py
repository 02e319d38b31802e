"""Parameter-holding building blocks with the reference's names and state-dict keys.

Mirror of label_anything/models/common.py (MLPBlock :19-37, LayerNorm2d :42-54, Attention :57-148,
AttentionMLPBlock :151-184).  These modules own the fp32 parameters (so `state_dict()`, `load_state_dict`,
`from_pretrained`, `.to(device)`, DDP wrapping and pickling behave exactly like the reference's); they carry
NO arithmetic of their own.  The CUDA kernels read their weights through `NativeCache`, a per-module cache of
device-side packed copies (bf16 GEMM operands, fused / reordered tables) that is rebuilt whenever a parameter
is modified in place or moved, and is never pickled.
"""
from __future__ import annotations

from typing import Callable, Dict, Type

import torch
import torch.nn as nn

SAM_EMBED_DIM = 256  # label_anything/models/common.py:16


class NativeModule(nn.Module):
    """nn.Module + a lazily-built, non-persistent cache of packed device tensors for the native kernels."""

    def _cache(self) -> Dict[str, object]:
        c = self.__dict__.get("_la_cache")
        if c is None:
            c = {}
            self.__dict__["_la_cache"] = c
        return c

    def packed(self, key: str, build: Callable[[], object], *deps: torch.Tensor):
        """Return cache[key], rebuilding it when any tensor in `deps` changed (in-place update, .to(), load).

        Under torch.compile tracing an existing entry is used as is (a graph constant guarded by dynamo) -- the
        (data_ptr, _version) signature cannot be evaluated on traced tensors; a compiled model therefore assumes frozen
        weights between recompilations, like any weight-prepacking backend.  Missing entries are built outside the
        graph."""
        if torch.compiler.is_compiling():
            hit = self._cache().get(key)
            if hit is not None:
                return hit[1]
            return self._packed_eager(key, build, deps)
        return self._packed_checked(key, build, deps)

    @torch.compiler.disable
    def _packed_eager(self, key, build, deps):
        return self._packed_checked(key, build, deps)

    def _packed_checked(self, key, build, deps):
        sig = tuple((d.data_ptr(), d._version, str(d.device), d.dtype) for d in deps)
        c = self._cache()
        hit = c.get(key)
        if hit is not None and hit[0] == sig:
            return hit[1]
        with torch.no_grad():
            val = build()
        c[key] = (sig, val)
        return val

    def __getstate__(self):
        state = self.__dict__.copy()
        state.pop("_la_cache", None)
        return state

    def _apply(self, fn, *a, **k):
        self.__dict__.pop("_la_cache", None)
        return super()._apply(fn, *a, **k)

    def forward(self, *a, **k):  # pragma: no cover - containers are driven by their owning model
        raise RuntimeError(
            f"{type(self).__name__} is a parameter container of labelanything_b200; its arithmetic runs inside "
            "the owning model's native (CUDA, sm_100a) forward. There is no eager fallback."
        )


def bf16_weight(mod: NativeModule, name: str, w: torch.Tensor) -> torch.Tensor:
    """bf16 [N, K] copy of an nn.Linear-style weight (any trailing dims flattened into K)."""
    return mod.packed("w:" + name, lambda: w.detach().reshape(w.shape[0], -1).to(torch.bfloat16).contiguous(), w)


def f32(mod: NativeModule, name: str, t: torch.Tensor | None) -> torch.Tensor | None:
    if t is None:
        return None
    return mod.packed("f:" + name, lambda: t.detach().float().contiguous(), t)


class MLPBlock(NativeModule):
    def __init__(self, embedding_dim: int, mlp_dim: int, act: Type[nn.Module] = nn.GELU, dropout: float = 0.0):
        super().__init__()
        self.lin1 = nn.Linear(embedding_dim, mlp_dim)
        self.lin2 = nn.Linear(mlp_dim, embedding_dim)
        self.act = act()
        self.drop = nn.Dropout(dropout) if dropout > 0.0 else nn.Identity()


class LayerNorm2d(NativeModule):
    def __init__(self, num_channels: int, eps: float = 1e-6):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(num_channels))
        self.bias = nn.Parameter(torch.zeros(num_channels))
        self.eps = eps


class Attention(NativeModule):
    """q/k/v/out projections of the SAM-style attention (masks are no-ops in the reference, common.py:117-139)."""

    def __init__(self, embedding_dim: int, num_heads: int, downsample_rate: int = 1, dropout: float = 0.0):
        super().__init__()
        self.embedding_dim = embedding_dim
        self.internal_dim = embedding_dim // downsample_rate
        self.num_heads = num_heads
        self.drop = nn.Dropout(dropout) if dropout > 0.0 else nn.Identity()
        assert self.internal_dim % num_heads == 0, "num_heads must divide embedding_dim."
        self.q_proj = nn.Linear(embedding_dim, self.internal_dim)
        self.k_proj = nn.Linear(embedding_dim, self.internal_dim)
        self.v_proj = nn.Linear(embedding_dim, self.internal_dim)
        self.out_proj = nn.Linear(self.internal_dim, embedding_dim)


class AttentionMLPBlock(NativeModule):
    def __init__(self, embed_dim: int, downsample_rate: int, mlp_dim: int, num_heads: int,
                 act: Type[nn.Module] = nn.GELU, dropout: float = 0.0):
        super().__init__()
        self.norm = nn.LayerNorm(embed_dim)
        self.mlp = MLPBlock(embed_dim, mlp_dim, act, dropout=dropout)
        self.attn = Attention(embed_dim, num_heads=num_heads, downsample_rate=downsample_rate, dropout=dropout)
