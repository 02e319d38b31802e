"""Two-way transformer (tokens <-> image) and the token-set attention block, on the native kernels.

Mirror of label_anything/models/transformer.py (`TwoWayTransformer` :157-252, `TwoWayAttentionBlock` :255-329)
and of `AttentionMLPBlock` / `Attention` / `MLPBlock` in label_anything/models/common.py:19-37,57-184.  The
classes keep the reference's attribute names and state-dict keys and own the parameters; the arithmetic is the
launch sequences `run_two_way` and `run_attention_mlp_block` below.

Data layout: everything is token-major.  `tokens` are the S*n sparse / class tokens ([S*n, D], fp32 residual
stream + bf16 copies feeding the GEMMs); `keys` are the S*T image tokens.  Per layer the image side costs ONE
projection GEMM over [S*T, D] with the k/v(/q) weights of both cross attentions concatenated, the positional
encoding enters as a projected table (pe @ W^T, constant per model) added inside the attention kernels, and
the attention matrices are never materialised.

Exact simplifications used (no approximation):
  * the reference's key_mask / query_mask are no-ops (common.py:117-139) -> no masks;
  * with a single token per sequence (n == 1: mask-only prompts) the image->token softmax is over one key, i.e.
    exactly 1, so that attention's output is out_proj(v_proj(token)) for every image position: a per-sequence
    vector folded into the following add+LayerNorm (no image-side q projection, no [S*T, D] branch tensor);
  * the prompt encoder only consumes the image-token output (prompt_encoder.py:681-685), so the final
    token->image attention (transformer.py:245-250) is skipped there (`want_queries=False`);
  * with a single token per sequence the token->image attention needs no projection of the image tokens: by
    associativity the k / v projections move to the query side (`_pooled_token_to_image`, csrc/la_poolattn.cu) and the
    S*T image tokens are read once instead of projected (2.6 TFLOP, 10 GB per layer), written and read again.
"""
from __future__ import annotations

from typing import Optional, Tuple, Type

import torch
import torch.nn as nn

from . import ops
from .common import Attention, AttentionMLPBlock, MLPBlock, NativeModule, bf16_weight, f32

__all__ = ["TwoWayTransformer", "TwoWayAttentionBlock", "run_two_way", "run_attention_mlp_block"]


class TwoWayAttentionBlock(NativeModule):
    def __init__(self, embedding_dim: int, num_heads: int, mlp_dim: int = 2048,
                 activation: Type[nn.Module] = nn.ReLU, attention_downsample_rate: int = 2,
                 skip_first_layer_pe: bool = False, dropout: float = 0.0) -> None:
        super().__init__()
        self.self_attn = Attention(embedding_dim, num_heads, dropout=dropout)
        self.norm1 = nn.LayerNorm(embedding_dim)
        self.cross_attn_token_to_image = Attention(embedding_dim, num_heads,
                                                   downsample_rate=attention_downsample_rate, dropout=dropout)
        self.norm2 = nn.LayerNorm(embedding_dim)
        self.mlp = MLPBlock(embedding_dim, mlp_dim, activation, dropout=dropout)
        self.norm3 = nn.LayerNorm(embedding_dim)
        self.norm4 = nn.LayerNorm(embedding_dim)
        self.cross_attn_image_to_token = Attention(embedding_dim, num_heads,
                                                   downsample_rate=attention_downsample_rate, dropout=dropout)
        self.skip_first_layer_pe = skip_first_layer_pe


class TwoWayTransformer(NativeModule):
    def __init__(self, depth: int, embedding_dim: int, num_heads: int, mlp_dim: int,
                 activation: Type[nn.Module] = nn.ReLU, attention_downsample_rate: int = 2,
                 dropout: float = 0.0) -> None:
        super().__init__()
        self.depth = depth
        self.embedding_dim = embedding_dim
        self.num_heads = num_heads
        self.mlp_dim = mlp_dim
        self.layers = nn.ModuleList()
        self.attention_downsample_rate = attention_downsample_rate
        for i in range(depth):
            self.layers.append(TwoWayAttentionBlock(embedding_dim=embedding_dim, num_heads=num_heads,
                                                    mlp_dim=mlp_dim, activation=activation,
                                                    attention_downsample_rate=attention_downsample_rate,
                                                    dropout=dropout, skip_first_layer_pe=(i == 0)))
        self.final_attn_token_to_image = Attention(embedding_dim, num_heads,
                                                   downsample_rate=attention_downsample_rate, dropout=dropout)
        self.norm_final_attn = nn.LayerNorm(embedding_dim)

    def forward(self, image_embedding: torch.Tensor, image_pe: torch.Tensor, point_embedding: torch.Tensor,
                query_mask: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, torch.Tensor]:
        """Reference signature (transformer.py:206-252): image_embedding [S, D, h, w], image_pe [1 or S, D, h, w]
        (identical rows), point_embedding [S, n, D] -> (tokens [S, n, D], image tokens [S, h*w, D]); fp32."""
        S, D, h, w = image_embedding.shape
        n = point_embedding.shape[1]
        k32, k16 = ops.nchw_to_tokens(image_embedding.float().contiguous(), want_f32=True, want_bf16=True)
        pe, _ = ops.nchw_to_tokens(image_pe[:1].float().contiguous())
        tok = point_embedding.float().contiguous().view(S * n, D)
        q, keys, _ = run_two_way(self, k16, k32, pe, tok, S, h * w, n, want_queries=True, want_keys_f32=True)
        return q.view(S, n, D), keys.view(S, h * w, D)


# ------------------------------------------------------------------------------------------------------------
# weight packing
# ------------------------------------------------------------------------------------------------------------
def _cat_w(att: Attention, names: Tuple[str, ...]) -> torch.Tensor:
    mods = [getattr(att, n + "_proj") for n in names]
    return att.packed("wcat:" + "".join(names),
                      lambda: torch.cat([m.weight.detach() for m in mods]).to(torch.bfloat16).contiguous(),
                      *[m.weight for m in mods])


def _cat_b(att: Attention, names: Tuple[str, ...]) -> torch.Tensor:
    mods = [getattr(att, n + "_proj") for n in names]
    return att.packed("bcat:" + "".join(names),
                      lambda: torch.cat([m.bias.detach() for m in mods]).float().contiguous(),
                      *[m.bias for m in mods])


def _w(att: Attention, name: str) -> torch.Tensor:
    return bf16_weight(att, name, getattr(att, name + "_proj").weight)


def _b(att: Attention, name: str) -> torch.Tensor:
    return f32(att, name + ".b", getattr(att, name + "_proj").bias)


def _pe_table(att: Attention, name: str, pe: torch.Tensor, cached: bool) -> torch.Tensor:
    """pe [T, D] fp32 -> pe @ W_name^T  [T, internal_dim] fp32 (bias stays in the projection GEMM).  Constant per
    (model, resolution) on the native `Lam` path, whose `pe` is the prompt encoder's cached `dense_pe_tokens()` tensor
    (`cached=True`): computed once at weight-packing time.  A transient `pe` (the reference-signature `forward`
    methods build one per call) is NOT cached: the allocator may hand the same address to a different positional
    encoding, which a (data_ptr, _version) key could not tell apart."""
    w = getattr(att, name + "_proj").weight
    make = lambda: (pe.double() @ w.detach().double().t()).float().contiguous()   # noqa: E731
    if not cached:
        with torch.no_grad():
            return make()
    return att.packed(f"pe:{name}:{pe.shape[0]}", make, w, pe)


def _pooled_token_to_image(att: Attention, tq: torch.Tensor, keys16: torch.Tensor, pe: torch.Tensor, pe_cached: bool,
                           S: int, T: int, H: int, Dc: int) -> torch.Tensor:
    """Attention(q = tokens + pe_q, k = keys + pe, v = keys) for ONE query token per sequence (transformer.py:311-318,
    common.py:97-148) without projecting the S*T image tokens:
        q_h . k_{h,t} = u_h . x_t + u_h . pe_t + q_h . b_k[h],   u_h = W_k[h]^T q_h       (the last term is constant
        o_h = W_v[h] (sum_t p_{h,t} x_t) + b_v[h]                                          over t: gone in the softmax)
    tq bf16 [S, Dc] (projected query incl. bias) -> bf16 [S, Dc] (input of out_proj).  Every product is a la_gemm_bf16
    over S*H rows; the image tokens are touched only by la_attention_pooled_bf16."""
    dh = Dc // H
    wk_t = att.packed("wk_t", lambda: att.k_proj.weight.detach().t().to(torch.bfloat16).contiguous(), att.k_proj.weight)
    def make_pe16():   # rows padded to a multiple of 8 (the GEMM's N granularity; 900 tokens at 480 px): the extra
        t = torch.zeros(((pe.shape[0] + 7) // 8 * 8, pe.shape[1]), dtype=torch.bfloat16, device=pe.device)   # columns
        t[:pe.shape[0]] = pe.to(torch.bfloat16)                                                      # of e are never read
        return t

    pe16 = att.packed(f"pe16:{pe.shape[0]}", make_pe16, pe) if pe_cached else make_pe16()
    qb = ops.head_rows(tq, S, H, dh, expand=True)                          # [S*H, Dc], row (s, h) = q_h in its columns
    u = ops.gemm(qb, wk_t, None)                                           # [S*H, D]   u_h = W_k[h]^T q_h
    e = ops.gemm(u, pe16, None, out_dtype=torch.float32)                   # [S*H, T]   u_h . pe_t
    y = ops.attention_pooled(keys16, u, e, dh ** -0.5, S, T, H)            # [S*H, D]   sum_t p x_t
    ov = ops.gemm(y, _w(att, "v"), _b(att, "v"))                           # [S*H, Dc]  W_v y_h + b_v (all head blocks)
    return ops.head_rows(ov, S, H, dh, expand=False)                       # [S, Dc]    head h's block of row (s, h)


def _ln(mod: NativeModule, name: str, ln: nn.LayerNorm):
    return f32(mod, name + ".w", ln.weight), f32(mod, name + ".b", ln.bias), ln.eps


def _act_code(mlp: MLPBlock) -> int:
    if isinstance(mlp.act, nn.GELU):
        assert getattr(mlp.act, "approximate", "none") == "none", "native GELU is the exact erf form"
        return ops.ACT_GELU
    if isinstance(mlp.act, nn.ReLU):
        return ops.ACT_RELU
    raise NotImplementedError(f"activation {type(mlp.act).__name__} has no native epilogue (GELU / ReLU only)")


def _empty(rows: int, d: int, dtype, dev) -> torch.Tensor:
    return torch.empty((rows, d), dtype=dtype, device=dev)


def _cast_bf16(x: torch.Tensor) -> torch.Tensor:
    rows, d = x.shape
    y = _empty(rows, d, torch.bfloat16, x.device)
    ops.add_layernorm(x, None, None, None, 0.0, rows=rows, d=d, y_out=y)
    return y


# ------------------------------------------------------------------------------------------------------------
# AttentionMLPBlock  (common.py:151-184): a = LN(attn(x, x, x) + x); out = LN(mlp(a) + a) -- the SAME LayerNorm
# ------------------------------------------------------------------------------------------------------------
def run_attention_mlp_block(blk: AttentionMLPBlock, x: torch.Tensor, n_seq: int, seq_len: int) -> torch.Tensor:
    """x fp32 [n_seq*seq_len, D] -> fp32 [n_seq*seq_len, D]."""
    rows, D = x.shape
    assert rows == n_seq * seq_len
    att = blk.attn
    Di, H = att.internal_dim, att.num_heads
    g, b, eps = _ln(blk, "norm", blk.norm)
    dev = x.device
    xb = _cast_bf16(x)
    qkv = ops.gemm(xb, _cat_w(att, ("q", "k", "v")), _cat_b(att, ("q", "k", "v")))
    o = ops.attention_tokens(qkv[:, :Di], qkv[:, Di:2 * Di], qkv[:, 2 * Di:], n_seq, seq_len, seq_len, H, Di // H)
    o = ops.gemm(o, _w(att, "out"), _b(att, "out"))
    a32, a16 = _empty(rows, D, torch.float32, dev), _empty(rows, D, torch.bfloat16, dev)
    ops.add_layernorm(x, o, g, b, eps, rows=rows, d=D, y_out=a16, y2_out=a32)
    mlp = blk.mlp
    h = ops.gemm(a16, bf16_weight(mlp, "lin1", mlp.lin1.weight), f32(mlp, "lin1.b", mlp.lin1.bias), act=_act_code(mlp))
    m = ops.gemm(h, bf16_weight(mlp, "lin2", mlp.lin2.weight), f32(mlp, "lin2.b", mlp.lin2.bias))
    out = _empty(rows, D, torch.float32, dev)
    ops.add_layernorm(a32, m, g, b, eps, rows=rows, d=D, y_out=out)
    return out


# ------------------------------------------------------------------------------------------------------------
# TwoWayTransformer  (transformer.py:206-252, 298-329)
# ------------------------------------------------------------------------------------------------------------
def run_two_way(tw: TwoWayTransformer, keys16: torch.Tensor, keys32: Optional[torch.Tensor], pe: torch.Tensor,
                tokens: torch.Tensor, S: int, T: int, n: int, *, want_queries: bool, pool: bool = False,
                want_keys_f32: bool = False, pe_cached: bool = False):
    """keys16 bf16 [S*T, D] (the image tokens; keys32 = the same in fp32 when available), pe fp32 [T, D] (dense
    positional encoding, shared by all sequences; pe_cached: it is the model's cached dense_pe_tokens() tensor, so the
    tables projected from it may be cached too), tokens fp32 [S*n, D] (also the tokens' positional encoding).

    Returns (queries fp32 [S*n, D] | None, keys, pooled):
      pool=True  -> keys is None and pooled = mean over the T image tokens of the last layer's output [S, D] fp32
                    (the normalised tokens are never written: prompt_encoder.py:733-735 fused into norm4);
      pool=False -> keys = last layer's image tokens, bf16 [S*T, D] (fp32 when want_keys_f32), pooled = None.
    """
    D, H = tw.embedding_dim, tw.num_heads
    dev = keys16.device
    R, RT = S * n, S * T
    assert keys16.shape == (RT, D) and tokens.shape == (R, D) and pe.shape == (T, D)
    assert tokens.dtype == torch.float32 and tokens.is_contiguous()
    qpe = tokens                     # query_pe = the original token embeddings (transformer.py:240)
    tok32 = tokens
    tok16: Optional[torch.Tensor] = None     # bf16(queries)
    tokpe16: Optional[torch.Tensor] = None   # bf16(queries + query_pe)
    pooled = None
    depth = len(tw.layers)
    for li, layer in enumerate(tw.layers):
        last = li == depth - 1
        # ---- (1) token self-attention --------------------------------------------------------------------
        sa = layer.self_attn
        Ds = sa.internal_dim
        g1, b1, e1 = _ln(layer, "norm1", layer.norm1)
        if layer.skip_first_layer_pe:
            xb = tok16 if tok16 is not None else _cast_bf16(tok32)
            qkv = ops.gemm(xb, _cat_w(sa, ("q", "k", "v")), _cat_b(sa, ("q", "k", "v")))
            q_, k_, v_ = qkv[:, :Ds], qkv[:, Ds:2 * Ds], qkv[:, 2 * Ds:]
            resid = None             # queries are REPLACED by the attention output (transformer.py:302-303)
        else:
            if tokpe16 is None:      # first layer of a transformer built without skip_first_layer_pe
                tokpe16, tok16 = _empty(R, D, torch.bfloat16, dev), _cast_bf16(tok32)
                ops.add_layernorm(tok32, None, None, None, 0.0, rows=R, d=D, pe=qpe, ype_out=tokpe16)
            qk = ops.gemm(tokpe16, _cat_w(sa, ("q", "k")), _cat_b(sa, ("q", "k")))
            q_, k_ = qk[:, :Ds], qk[:, Ds:]
            v_ = ops.gemm(tok16, _w(sa, "v"), _b(sa, "v"))
            resid = tok32
        o = ops.attention_tokens(q_, k_, v_, S, n, n, H, Ds // H)
        o = ops.gemm(o, _w(sa, "out"), _b(sa, "out"))
        tok32n, tok16, tokpe16 = (_empty(R, D, torch.float32, dev), _empty(R, D, torch.bfloat16, dev),
                                  _empty(R, D, torch.bfloat16, dev))
        ops.add_layernorm(resid, o, g1, b1, e1, rows=R, d=D, y_out=tok16, y2_out=tok32n, pe=qpe, ype_out=tokpe16)
        tok32 = tok32n

        # ---- image-side projections: k, v of token->image (+ q of image->token when n > 1) ----------------
        t2i, i2t = layer.cross_attn_token_to_image, layer.cross_attn_image_to_token
        Dc = t2i.internal_dim
        assert i2t.internal_dim == Dc
        need_q = n > 1
        # One query token per sequence (mask-only prompts): no image-side projection at all -- the k / v projections
        # are moved to the query side by associativity and the image tokens are read once, as they are
        # (csrc/la_poolattn.cu).
        pooled = (not need_q) and ops.attention_pooled_supported(n, H, D, T)
        if pooled:
            proj = None
        elif need_q:
            w_img = layer.packed("w_img3", lambda: torch.cat([t2i.k_proj.weight, t2i.v_proj.weight, i2t.q_proj.weight])
                                 .detach().to(torch.bfloat16).contiguous(),
                                 t2i.k_proj.weight, t2i.v_proj.weight, i2t.q_proj.weight)
            b_img = layer.packed("b_img3", lambda: torch.cat([t2i.k_proj.bias, t2i.v_proj.bias, i2t.q_proj.bias])
                                 .detach().float().contiguous(), t2i.k_proj.bias, t2i.v_proj.bias, i2t.q_proj.bias)
        else:
            w_img, b_img = _cat_w(t2i, ("k", "v")), _cat_b(t2i, ("k", "v"))
        if not pooled:
            proj = ops.gemm(keys16, w_img, b_img)                  # [S*T, 2Dc | 3Dc] bf16

        # ---- (2) tokens attend to the image ----------------------------------------------------------------
        tq = ops.gemm(tokpe16, _w(t2i, "q"), _b(t2i, "q"))
        if pooled:
            o = _pooled_token_to_image(t2i, tq, keys16, pe, pe_cached, S, T, H, Dc)
        else:
            o = ops.attention_tokens(tq, proj[:, :Dc], proj[:, Dc:2 * Dc], S, n, T, H, Dc // H,
                                     k_add=_pe_table(t2i, "k", pe, pe_cached))
        o = ops.gemm(o, _w(t2i, "out"), _b(t2i, "out"))
        g2, b2, e2 = _ln(layer, "norm2", layer.norm2)
        tok32n, tok16 = _empty(R, D, torch.float32, dev), _empty(R, D, torch.bfloat16, dev)
        ops.add_layernorm(tok32, o, g2, b2, e2, rows=R, d=D, y_out=tok16, y2_out=tok32n)
        tok32 = tok32n

        # ---- (3) token MLP ---------------------------------------------------------------------------------
        mlp = layer.mlp
        hmid = ops.gemm(tok16, bf16_weight(mlp, "lin1", mlp.lin1.weight), f32(mlp, "lin1.b", mlp.lin1.bias),
                        act=_act_code(mlp))
        m = ops.gemm(hmid, bf16_weight(mlp, "lin2", mlp.lin2.weight), f32(mlp, "lin2.b", mlp.lin2.bias))
        g3, b3, e3 = _ln(layer, "norm3", layer.norm3)
        tok32n, tok16, tokpe16 = (_empty(R, D, torch.float32, dev), _empty(R, D, torch.bfloat16, dev),
                                  _empty(R, D, torch.bfloat16, dev))
        ops.add_layernorm(tok32, m, g3, b3, e3, rows=R, d=D, y_out=tok16, y2_out=tok32n, pe=qpe, ype_out=tokpe16)
        tok32 = tok32n

        # ---- (4) image attends to the tokens ---------------------------------------------------------------
        tv = ops.gemm(tok16, _w(i2t, "v"), _b(i2t, "v"))
        delta = seq_add = None
        if need_q:
            tk = ops.gemm(tokpe16, _w(i2t, "k"), _b(i2t, "k"))
            o = ops.attention_tokens(proj[:, 2 * Dc:], tk, tv, S, T, n, H, Dc // H, q_add=_pe_table(i2t, "q", pe, pe_cached))
            delta = ops.gemm(o, _w(i2t, "out"), _b(i2t, "out"))    # [S*T, D] bf16
            del o
        else:
            seq_add = ops.gemm(tv, _w(i2t, "out"), _b(i2t, "out"), out_dtype=torch.float32)   # [S, D]
        proj = None
        g4, b4, e4 = _ln(layer, "norm4", layer.norm4)
        x_in, d1, d2 = (keys32, delta, None) if keys32 is not None else (None, keys16, delta)
        if last and pool:
            pooled = ops.add_layernorm_meanpool(x_in, d1, g4, b4, e4, S, T, D, delta2=d2, seq_add=seq_add)
            keys16 = keys32 = None
        else:
            new16 = _empty(RT, D, torch.bfloat16, dev)
            # fp32 copy of the image tokens: residual of the next layer's norm4.  When the transformer output is only
            # mean-pooled (prompt encoder) the bf16 copy doubles as the residual: its rounding error (2^-9 relative)
            # averages out over the T pooled tokens, and it saves an 8-byte/element HBM round trip.
            need32 = ((not last) and not pool) or want_keys_f32
            new32 = _empty(RT, D, torch.float32, dev) if need32 else None
            ops.add_layernorm(x_in, d1, g4, b4, e4, rows=RT, d=D, y_out=new16, y2_out=new32, delta2=d2,
                              seq_add=seq_add, seq_rows=T)
            keys16, keys32 = new16, new32
        del delta

    queries = None
    if want_queries:
        assert keys16 is not None, "pooling and the final token->image attention are mutually exclusive"
        fa = tw.final_attn_token_to_image
        Df = fa.internal_dim
        kv = ops.gemm(keys16, _cat_w(fa, ("k", "v")), _cat_b(fa, ("k", "v")))
        tq = ops.gemm(tokpe16, _w(fa, "q"), _b(fa, "q"))
        o = ops.attention_tokens(tq, kv[:, :Df], kv[:, Df:], S, n, T, H, Df // H, k_add=_pe_table(fa, "k", pe, pe_cached))
        o = ops.gemm(o, _w(fa, "out"), _b(fa, "out"))
        gf, bf, ef = _ln(tw, "norm_final", tw.norm_final_attn)
        queries = _empty(R, D, torch.float32, dev)
        ops.add_layernorm(tok32, o, gf, bf, ef, rows=R, d=D, y_out=queries)
    keys = keys32 if (want_keys_f32 and keys32 is not None) else keys16
    return queries, keys, pooled
