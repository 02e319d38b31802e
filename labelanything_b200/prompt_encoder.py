"""Visual prompt encoder with the reference's module surface, running on the native kernels.

Mirror of label_anything/models/prompt_encoder.py: `PromptEncoder` (:21-184), `PositionEmbeddingRandom`
(:187-233), `RandomMatrixEncoder` (:236-277), `PromptImageEncoder` (:396-827).  Attribute names and state-dict
keys are the reference's (`pe_layer.positional_encoding_gaussian_matrix`, `point_embeddings.0-3`,
`not_a_point_embed`, `mask_downscaling.{0,1,3,4,6}`, `no_mask_embed`, `transformer.*`,
`class_encoder.pos_embedding`, `sparse_embedding_attention.*`, `no_sparse_embedding`,
`{class,example,class_example}_attention.*`, `not_a_mask_embed`).

Launch sequence of `PromptImageEncoder.encode` (all token-major, nothing of size S x D x h x w in fp32 is ever
materialised):
    sparse tokens (Fourier PE kernel) -> AttentionMLPBlock over the (class, token) set of each support image ->
    mask downscaling kernel (conv2x2 -> LN2d -> GELU -> conv2x2 -> LN2d -> GELU, 16 channels) [-> bilinear resize]
    -> src = support features + W6 . mask16 + class code (one bf16 write) -> two-way transformer with the last
    LayerNorm fused with the spatial mean -> class / example attention blocks -> flag-masked mean over examples.
"""
from __future__ import annotations

import math
from typing import Any, Dict, Optional, Tuple, Type

import torch
import torch.nn as nn

from . import ops
from .common import AttentionMLPBlock, LayerNorm2d, NativeModule, f32
from .transformer import run_attention_mlp_block, run_two_way
from .utils import BatchKeys, ResultDict


class PositionEmbeddingRandom(NativeModule):
    """Positional encoding using random spatial frequencies (prompt_encoder.py:187-233)."""

    def __init__(self, num_pos_feats: int = 64, scale: Optional[float] = None) -> None:
        super().__init__()
        if scale is None or scale <= 0.0:
            scale = 1.0
        self.register_buffer("positional_encoding_gaussian_matrix", scale * torch.randn((2, num_pos_feats)))

    def _pe_encoding(self, coords: torch.Tensor) -> torch.Tensor:
        coords = 2 * coords.to(self.positional_encoding_gaussian_matrix.dtype) - 1
        coords = 2 * math.pi * (coords @ self.positional_encoding_gaussian_matrix)
        return torch.cat([torch.sin(coords), torch.cos(coords)], dim=-1)

    def forward(self, size: Tuple[int, int]) -> torch.Tensor:
        """Dense grid encoding [C, h, w] — a constant of the model (weight-preparation time, cached by callers)."""
        h, w = size
        dev = self.positional_encoding_gaussian_matrix.device
        ys = (torch.arange(h, device=dev, dtype=torch.float32) + 0.5) / h
        xs = (torch.arange(w, device=dev, dtype=torch.float32) + 0.5) / w
        grid = torch.stack([xs[None, :].expand(h, w), ys[:, None].expand(h, w)], dim=-1)
        return self._pe_encoding(grid).permute(2, 0, 1)


class RandomMatrixEncoder(NativeModule):
    """Bank of random class codes; the background takes row 0, foreground classes a random permutation of the
    other rows on EVERY call, train and eval (prompt_encoder.py:236-277).  `fixed_rows` (not part of
    the state dict) pins the rows for reproducible runs and parity tests."""

    def __init__(self, bank_size: int, embed_dim: int):
        super().__init__()
        self.bank_size = bank_size
        self.embed_dim = embed_dim
        self.pos_embedding = nn.Parameter(torch.zeros(1, 1, bank_size, embed_dim))
        nn.init.normal_(self.pos_embedding, std=0.02)
        self.fixed_rows: Optional[torch.Tensor] = None

    def sample_rows(self, C: int, device) -> torch.Tensor:
        if self.fixed_rows is not None:
            return self.fixed_rows[:C].to(device=device, dtype=torch.long)
        fg_rows = torch.randperm(self.bank_size - 1, device=device)[: C - 1] + 1
        return torch.cat([torch.zeros(1, device=device, dtype=torch.long), fg_rows])

    def class_codes(self, C: int, device) -> torch.Tensor:
        """[C, D] fp32 codes of this call."""
        rows = self.sample_rows(C, device)
        return self.pos_embedding.detach()[0, 0].float().index_select(0, rows).contiguous()


class PromptEncoder(NativeModule):
    def __init__(self, embed_dim: int, image_embedding_size: Tuple[int, int], input_image_size: Tuple[int, int],
                 mask_in_chans: int, activation: Type[nn.Module] = nn.GELU) -> None:
        super().__init__()
        self.embed_dim = embed_dim
        self.input_image_size = input_image_size
        self.image_embedding_size = image_embedding_size
        self.pe_layer = PositionEmbeddingRandom(embed_dim // 2)
        self.num_point_embeddings: int = 4  # pos/neg point + 2 box corners
        self.point_embeddings = nn.ModuleList([nn.Embedding(1, embed_dim) for _ in range(self.num_point_embeddings)])
        self.not_a_point_embed = nn.Embedding(1, embed_dim)
        self.mask_input_size = (4 * image_embedding_size[0], 4 * image_embedding_size[1])
        self.mask_downscaling = nn.Sequential(
            nn.Conv2d(1, mask_in_chans // 4, kernel_size=2, stride=2), LayerNorm2d(mask_in_chans // 4), activation(),
            nn.Conv2d(mask_in_chans // 4, mask_in_chans, kernel_size=2, stride=2), LayerNorm2d(mask_in_chans),
            activation(), nn.Conv2d(mask_in_chans, embed_dim, kernel_size=1))
        self.no_mask_embed = nn.Embedding(1, embed_dim)
        if mask_in_chans != 16 or activation is not nn.GELU:
            raise NotImplementedError("the native mask-downscaling kernel is built for mask_in_chans=16 with GELU")

    def get_dense_pe(self) -> torch.Tensor:
        """1 x embed_dim x h x w (prompt_encoder.py:72-81)."""
        return self.pe_layer(self.image_embedding_size).unsqueeze(0)

    def dense_pe_tokens(self, size: Optional[Tuple[int, int]] = None) -> torch.Tensor:
        """The dense positional encoding as token-major fp32 [h*w, D], cached per (matrix, size)."""
        size = tuple(size or self.image_embedding_size)
        g = self.pe_layer.positional_encoding_gaussian_matrix
        return self.packed(f"dense_pe:{size}",
                           lambda: self.pe_layer(size).permute(1, 2, 0).reshape(size[0] * size[1], -1).contiguous(), g)


class PromptImageEncoder(PromptEncoder):
    def __init__(self, embed_dim: int, image_embedding_size: Tuple[int, int], input_image_size: Tuple[int, int],
                 mask_in_chans: int, transformer: nn.Module, class_encoder: Any,
                 example_class_attention: bool = True, class_attention: bool = False,
                 class_embedding_dim: Optional[int] = None, example_attention: bool = False,
                 activation: Type[nn.Module] = nn.GELU, use_support_features: bool = True,
                 embeddings_per_example: Optional[int] = 1, embedding_extraction: Optional[str] = None,
                 dropout: float = 0.0) -> None:
        super().__init__(embed_dim, image_embedding_size, input_image_size, mask_in_chans, activation)
        num_heads, attention_downsample_rate, mlp_dim = 8, 2, 2048
        self.embeddings_per_example = embeddings_per_example
        self.transformer = transformer
        self.class_encoder = class_encoder
        self.use_support_features = use_support_features
        if embedding_extraction is not None or (embeddings_per_example and embeddings_per_example > 1):
            raise NotImplementedError("embedding_extraction / embeddings_per_example > 1 are outside the native hot path")
        if not use_support_features:
            raise NotImplementedError("use_support_features=False (proto_chooser) is outside the native hot path")
        if class_embedding_dim is not None:
            raise NotImplementedError("class_embedding_dim projectors are outside the native hot path")
        self.embedding_extraction = None
        self.sparse_embedding_attention = AttentionMLPBlock(embed_dim=embed_dim, num_heads=num_heads,
                                                            downsample_rate=1, mlp_dim=mlp_dim, act=activation,
                                                            dropout=dropout)
        self.no_sparse_embedding = nn.Embedding(1, embed_dim)
        self.class_projector_in = nn.Identity()
        self.class_projector_out = nn.Identity()

        def block():
            return AttentionMLPBlock(embed_dim=embed_dim, num_heads=num_heads,
                                     downsample_rate=attention_downsample_rate, mlp_dim=mlp_dim, act=activation,
                                     dropout=dropout)

        self.class_attention = block() if class_attention else None
        self.class_example_attention = block() if example_class_attention else None
        self.example_attention = block() if example_attention else None
        self.not_a_mask_embed = nn.Embedding(1, embed_dim)
        #: image-token rows (sequences x tokens) processed per pass of the fusion transformer (bounds the workspace)
        self.max_rows_per_pass = 6 * 1024 * 1024

    # ------------------------------------------------------------------ packed weights
    def _mask_host_weights(self) -> Dict[str, Any]:
        md = self.mask_downscaling

        def build():
            cpu = lambda t: t.detach().float().cpu().contiguous()  # noqa: E731
            return {"w0": cpu(md[0].weight).view(-1), "b0": cpu(md[0].bias), "g1": cpu(md[1].weight),
                    "be1": cpu(md[1].bias), "eps1": md[1].eps, "w3": cpu(md[3].weight).view(-1), "b3": cpu(md[3].bias),
                    "g2": cpu(md[4].weight), "be2": cpu(md[4].bias), "eps2": md[4].eps}

        return self.packed("mask_host", build, md[0].weight, md[0].bias, md[1].weight, md[1].bias, md[3].weight,
                           md[3].bias, md[4].weight, md[4].bias)

    def _pe_table4(self) -> torch.Tensor:
        ws = [e.weight for e in self.point_embeddings]
        return self.packed("pe_table4", lambda: torch.cat([w.detach().float() for w in ws]).contiguous(), *ws)

    def _class_codes(self, C: int, device) -> Optional[torch.Tensor]:
        ce = self.class_encoder
        if isinstance(ce, RandomMatrixEncoder):
            return ce.class_codes(C, device)
        if isinstance(ce, nn.Module):
            raise NotImplementedError(f"class encoder {type(ce).__name__} has no native path")
        return None  # the reference's identity lambda (build_lam.py:143)

    # ------------------------------------------------------------------ native forward
    def encode(self, feat: torch.Tensor, B: int, M: int, points, boxes, masks, flag_examples: torch.Tensor,
               feat_lead: int = 0) -> Dict[str, torch.Tensor]:
        """feat: token-major fp32 [B*(M+feat_lead)*T, D] features (support images of episode b start at image
        b*(M+feat_lead)+feat_lead).  points = (coords [B,M,C,P,2], labels [B,M,C,P]) | None, boxes = (xyxy
        [B,M,C,Bx,4], flags [B,M,C,Bx]) | None, masks = (masks [B,M,C,Hm,Wm], flags [B,M,C]) | None."""
        ops._require_cuda(feat)
        any_prompt = points[0] if points is not None else boxes[0] if boxes is not None else \
            masks[0] if masks is not None else None
        if any_prompt is None:
            raise ValueError("No prompts provided")  # prompt_encoder.py:562
        assert tuple(any_prompt.shape[:2]) == (B, M)
        C = any_prompt.shape[2]
        D = self.embed_dim
        h, w = self.image_embedding_size
        T = h * w
        S = B * M * C
        dev = feat.device
        assert feat.dtype == torch.float32 and feat.shape == (B * (M + feat_lead) * T, D), (feat.shape, B, M, T, D)

        # ---- sparse tokens (prompt_encoder.py:596-629)
        if points is not None or boxes is not None:
            pts = lab = bx = bfl = None
            if points is not None:
                pts = points[0].reshape(S, -1, 2).float().contiguous()
                lab = points[1].reshape(S, -1).float().contiguous()
            if boxes is not None:
                bx = boxes[0].reshape(S, -1, 4).float().contiguous()
                bfl = boxes[1].reshape(S, -1).float().contiguous()
            gauss = f32(self, "gauss", self.pe_layer.positional_encoding_gaussian_matrix)
            sparse = ops.embed_sparse(pts, lab, bx, bfl, gauss, f32(self, "nap", self.not_a_point_embed.weight).view(-1),
                                      self._pe_table4(), S, D, self.input_image_size[1], self.input_image_size[0])
            n = sparse.shape[1]
            sparse = sparse.view(S * n, D)
        else:
            n = 1
            sparse = torch.empty((S, D), dtype=torch.float32, device=dev)
            ops.add_layernorm(f32(self, "nse", self.no_sparse_embedding.weight), None, None, None, 0.0, rows=S, d=D,
                              y_out=sparse, x_mod=1)
        sparse = run_attention_mlp_block(self.sparse_embedding_attention, sparse, B * M, C * n)

        # ---- class code (RandomMatrixEncoder.forward_with_rows, prompt_encoder.py:250-264)
        code = self._class_codes(C, dev)
        if code is not None:
            sparse = ops.add_bcast(sparse, code, n, C)

        # ---- dense mask embedding, channels 0..15 (mask_downscaling[0..5]; [6] is folded into build_src)
        m16 = mflags = None
        if masks is not None:
            mk, mf = masks
            Hm, Wm = mk.shape[-2:]
            m16 = ops.mask_downscale(mk.reshape(S, Hm, Wm).float().contiguous(), self._mask_host_weights())
            if (Hm // 4, Wm // 4) != (h, w):     # prompt_encoder.py:787-793 (bilinear commutes with the 1x1 conv)
                m16 = ops.resize_bilinear(m16, h, w)
            mflags = (mf.reshape(S) != 0).to(torch.uint8).contiguous()
        md6 = self.mask_downscaling[6]
        w6 = f32(self, "md6.w", md6.weight).view(D, 16)
        b6 = f32(self, "md6.b", md6.bias)
        nam = f32(self, "nam", self.not_a_mask_embed.weight).view(-1)
        nom = f32(self, "nom", self.no_mask_embed.weight).view(-1)
        pe = self.dense_pe_tokens()

        # ---- fusion transformer, whole episodes per pass
        per_episode = M * C
        ep_per_pass = max(1, self.max_rows_per_pass // (per_episode * T))
        pooled = []
        for e0 in range(0, B, ep_per_pass):
            ne = min(ep_per_pass, B - e0)
            s0, ns = e0 * per_episode, ne * per_episode
            src = ops.build_src(feat, None if m16 is None else m16[s0:s0 + ns],
                                None if mflags is None else mflags[s0:s0 + ns], w6, b6, nam, nom, code, ns, T, D, C, M,
                                feat_lead=feat_lead, seq_offset=s0)
            _, _, p = run_two_way(self.transformer, src, None, pe, sparse[s0 * n:(s0 + ns) * n], ns, T, n,
                                  want_queries=False, pool=True, pe_cached=True)
            pooled.append(p)
            del src
        emb = pooled[0] if len(pooled) == 1 else torch.cat(pooled)       # [S, D] = [B, M, C, D]

        # ---- prompt_class_information_merge (prompt_encoder.py:696-717; the key masks passed there are no-ops)
        if self.class_attention is not None:
            emb = run_attention_mlp_block(self.class_attention, emb, B * M, C)
        if self.example_attention is not None:
            e = ops.permute_rows(emb, B, M, C)                           # b m c d -> b c m d
            e = run_attention_mlp_block(self.example_attention, e, B * C, M)
            emb = ops.permute_rows(e, B, C, M)
        if self.class_example_attention is not None:
            emb = run_attention_mlp_block(self.class_example_attention, emb, B, M * C)

        # ---- average over examples, ignoring padding (prompt_encoder.py:738-745)
        fe8 = (flag_examples != 0).to(torch.uint8).contiguous()
        emb = emb.view(B, M, C, D)
        class_emb = ops.masked_mean(emb, fe8)
        return {BatchKeys.FLAG_EXAMPLES: flag_examples, ResultDict.CLASS_EMBS: class_emb,
                ResultDict.EXAMPLES_CLASS_EMBS: emb}

    def forward(self, image_embeddings: torch.Tensor, points, boxes, masks, flag_examples, chunk_size=None):
        """Reference signature (prompt_encoder.py:752-827): image_embeddings [B, M, D, h, w] fp32.  `chunk_size` is
        accepted and ignored: results are only defined for chunk_size=None in the reference.
        The `class_examples_src` entry (S x D x h x w, consumed only by the out-of-scope AffinityDecoder) is not
        produced — the fused tokens are pooled inside the last LayerNorm kernel."""
        B, M, D, h, w = image_embeddings.shape
        assert (h, w) == tuple(self.image_embedding_size)
        feat, _ = ops.nchw_to_tokens(image_embeddings.float().contiguous().view(B * M, D, h, w))
        return self.encode(feat, B, M, points, boxes, masks, flag_examples, feat_lead=0)
