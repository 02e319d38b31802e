"""Mask decoder with the reference's module surface, running on the native kernels.

Mirror of label_anything/models/mask_decoder.py: `MaskDecoderLam` (:169-363) and `MLP` (:776-804); state-dict keys
`output_upscaling.{0,1,3}`, `class_mlp.layers.0-2`, `transformer.*`, `spatial_convs.{0,1,3,4,6}`.

Launch sequence of `MaskDecoderLam.decode` (token-major throughout):
    two-way transformer (class tokens <-> query-image tokens, incl. the final token->image attention)
    -> class_mlp (3 GEMMs, ReLU epilogues)
    -> ConvTranspose2d(k=s=2) as a per-pixel GEMM [T, D] x [D, 4*D/4]; pixel shuffle + LayerNorm2d + GELU in one row kernel
    -> ConvTranspose2d as GEMM [4T, D/4] x [D/4, 4*D/8]; pixel shuffle in the cast kernel
    -> spatial 3x3 convs as im2col + GEMM (K = 9*D/8) with LayerNorm2d + GELU row kernels between
    -> hypernetwork dot product logits[b, c, p] = <class_mlp(token_c), pixel_p>.
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple, Type

import torch
import torch.nn as nn

from . import ops
from .common import LayerNorm2d, NativeModule, bf16_weight, f32
from .transformer import TwoWayTransformer, _cast_bf16, run_two_way
from .utils import BatchKeys, ResultDict


class MLP(NativeModule):
    """mask_decoder.py:776-804"""

    def __init__(self, input_dim: int, hidden_dim: int, output_dim: int, num_layers: int,
                 sigmoid_output: bool = False, dropout: float = 0.0) -> None:
        super().__init__()
        self.num_layers = num_layers
        h = [hidden_dim] * (num_layers - 1)
        self.layers = nn.ModuleList(nn.Linear(n, k) for n, k in zip([input_dim] + h, h + [output_dim]))
        self.sigmoid_output = sigmoid_output
        self.dropout = nn.Dropout(dropout) if dropout > 0 else nn.Identity()
        if sigmoid_output:
            raise NotImplementedError("sigmoid_output is not used on the LabelAnything path")

    def run(self, x: torch.Tensor) -> torch.Tensor:
        """x fp32 [rows, in] -> fp32 [rows, out]; ReLU between layers."""
        y = _cast_bf16(x)
        for i, layer in enumerate(self.layers):
            last = i == self.num_layers - 1
            y = ops.gemm(y, bf16_weight(self, f"l{i}", layer.weight), f32(self, f"l{i}.b", layer.bias),
                         act=ops.ACT_NONE if last else ops.ACT_RELU,
                         out_dtype=torch.float32 if last else torch.bfloat16)
        return y


class MaskDecoderLam(NativeModule):
    def __init__(self, *, transformer_dim: int, transformer: nn.Module, spatial_convs: Optional[int] = None,
                 activation: Type[nn.Module] = nn.GELU, segment_example_logits: bool = False,
                 classification_layer_downsample_rate: int = 8, conv_upsample_stride: int = 2,
                 classification_levels: int = 1, dropout: float = 0.0, conv_classification: bool = False) -> None:
        super().__init__()
        self.attention_dim = transformer_dim
        self.segment_example_logits = segment_example_logits
        if (segment_example_logits or classification_levels > 1 or conv_classification or conv_upsample_stride != 2
                or classification_layer_downsample_rate <= 1 or activation is not nn.GELU):
            raise NotImplementedError(
                "the native decoder covers the default LabelAnything head: two stride-2 transposed convolutions, "
                "GELU, dot-product classification (no segment_example_logits / classification_levels / "
                "conv_classification)")
        first = classification_layer_downsample_rate // 2
        self.level_reducer = None
        c1, c2 = transformer_dim // first, transformer_dim // classification_layer_downsample_rate
        self.output_upscaling = nn.Sequential(
            nn.ConvTranspose2d(transformer_dim, c1, kernel_size=2, stride=2), LayerNorm2d(c1), activation(),
            nn.ConvTranspose2d(c1, c2, kernel_size=2, stride=2))
        self.class_mlp = MLP(transformer_dim, transformer_dim, c2, 3, dropout=dropout)
        self.transformer = transformer
        self.spatial_convs = None
        if spatial_convs is not None:
            mods = []
            for i in range(spatial_convs):
                mods.append(nn.Conv2d(c2, c2, kernel_size=3, padding=1))
                if i < spatial_convs - 1:
                    mods.append(LayerNorm2d(c2))
                    mods.append(activation())
            self.spatial_convs = nn.Sequential(*mods)
        self.prototype_tconv = None

    # ------------------------------------------------------------------ helpers kept for reference callers
    def _get_pe_result(self, pe_result, flag_examples):
        """mask_decoder.py:273-287"""
        flag_examples = flag_examples if BatchKeys.FLAG_EXAMPLES not in pe_result else pe_result[BatchKeys.FLAG_EXAMPLES]
        class_embeddings = pe_result[ResultDict.CLASS_EMBS]
        embedding_mask = flag_examples.sum(dim=1).bool().int() if flag_examples is not None else None
        return class_embeddings, flag_examples, embedding_mask

    # ------------------------------------------------------------------ packed weights
    def _tconv_w(self, idx: int) -> Tuple[torch.Tensor, torch.Tensor]:
        """ConvTranspose2d(k=s=2) weight [Cin, Cout, 2, 2] -> GEMM weight [(ky, kx, co), ci] bf16 + bias repeated x4."""
        m = self.output_upscaling[idx]
        w = self.packed(f"up{idx}.w", lambda: m.weight.detach().permute(2, 3, 1, 0).reshape(-1, m.weight.shape[0])
                        .to(torch.bfloat16).contiguous(), m.weight)
        b = self.packed(f"up{idx}.b", lambda: m.bias.detach().float().repeat(4).contiguous(), m.bias)
        return w, b

    def _conv3_w(self, conv: nn.Conv2d, key: str) -> torch.Tensor:
        return self.packed(key, lambda: conv.weight.detach().permute(0, 2, 3, 1).reshape(conv.weight.shape[0], -1)
                           .to(torch.bfloat16).contiguous(), conv.weight)

    # ------------------------------------------------------------------ native forward
    def decode(self, q32: torch.Tensor, q16: torch.Tensor, pe: torch.Tensor, class_embeddings: torch.Tensor,
               B: int, h: int, w: int, pe_cached: bool = False) -> torch.Tensor:
        """q32 / q16: query-image features, token-major fp32 / bf16 [B*h*w, D]; pe fp32 [h*w, D];
        class_embeddings fp32 [B, C, D] -> low-resolution logits fp32 [B, C, 4h, 4w]."""
        assert h == w, "the native pixel-shuffle kernels expect square feature maps"
        D = self.attention_dim
        T = h * w
        C = class_embeddings.shape[1]
        assert isinstance(self.transformer, TwoWayTransformer), "only TwoWayTransformer has a native path"
        tokens = class_embeddings.float().contiguous().view(B * C, D)
        queries, keys16, _ = run_two_way(self.transformer, q16, q32, pe, tokens, B, T, C, want_queries=True,
                                         pe_cached=pe_cached)
        cls = self.class_mlp.run(queries)                                # [B*C, D/8] fp32

        up = self.output_upscaling
        w0, b0 = self._tconv_w(0)
        c1 = up[0].weight.shape[1]
        u = ops.gemm(keys16, w0, b0)                                      # [B*T, 4*c1] rows (img, y, x | ky, kx, co)
        x1 = torch.empty((B * T * 4, c1), dtype=torch.bfloat16, device=u.device)
        ops.add_layernorm(None, u.view(B * T * 4, c1), f32(self, "up1.w", up[1].weight), f32(self, "up1.b", up[1].bias),
                          up[1].eps, rows=B * T * 4, d=c1, y_out=x1, act=ops.ACT_GELU, map_mode=3, hw=h)
        w3, b3 = self._tconv_w(3)
        c2 = up[3].weight.shape[1]
        u = ops.gemm(x1, w3, b3)                                          # [B*4T, 4*c2]
        x = torch.empty((B * T * 16, c2), dtype=torch.bfloat16, device=u.device)
        ops.add_layernorm(None, u.view(B * T * 16, c2), None, None, 0.0, rows=B * T * 16, d=c2, y_out=x, map_mode=3,
                          hw=2 * h)
        del u, x1
        H4, W4 = 4 * h, 4 * w
        if self.spatial_convs is not None:
            mods = list(self.spatial_convs)
            i = 0
            while i < len(mods):
                conv = mods[i]
                col = ops.im2col_3x3(x, B, H4, W4, c2)
                t = ops.gemm(col, self._conv3_w(conv, f"sc{i}.w"), f32(self, f"sc{i}.b", conv.bias))
                del col
                if i + 1 < len(mods):                                      # LayerNorm2d + GELU follow
                    ln = mods[i + 1]
                    x = torch.empty_like(t)
                    ops.add_layernorm(None, t, f32(self, f"sc{i + 1}.w", ln.weight), f32(self, f"sc{i + 1}.b", ln.bias),
                                      ln.eps, rows=t.shape[0], d=c2, y_out=x, act=ops.ACT_GELU)
                    i += 3
                else:
                    x = t
                    i += 1
        logits = ops.classify(x, cls.view(B, C, c2), B, H4 * W4)
        return logits.view(B, C, H4, W4)

    def forward(self, query_embeddings: torch.Tensor, support_embeddings, image_pe: torch.Tensor, pe_result: Dict,
                flag_examples) -> torch.Tensor:
        """Reference signature (mask_decoder.py:316-363): query_embeddings [B, D, h, w], image_pe [1, D, h, w]."""
        B, D, h, w = query_embeddings.shape
        class_embeddings, _, _ = self._get_pe_result(pe_result, flag_examples)
        q32, q16 = ops.nchw_to_tokens(query_embeddings.float().contiguous(), want_f32=True, want_bf16=True)
        pe, _ = ops.nchw_to_tokens(image_pe[:1].float().contiguous())
        return self.decode(q32, q16, pe, class_embeddings, B, h, w)
