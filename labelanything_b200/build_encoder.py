"""Encoder factories with the reference's names (label_anything/models/build_encoder.py).

`build_vit_{b,l,h}` build the SAM ViT (`ImageEncoderViT`); `ViTModelWrapper` subclasses
`transformers.ViTModel` exactly like the reference (:83-100) so HF / MAE checkpoints load through
`from_pretrained`, but its forward runs on the native kernels (CLS kept through all layers, bicubic-resized
position table, final LayerNorm, CLS dropped, `b (h w) c -> b c h w`).
"""
from __future__ import annotations

from functools import partial

import torch
from transformers import ViTModel

from . import ops
from .common import NativeModule
from .image_encoder import ImageEncoderViT
from .vit_engine import BlockWeights, VitSpec, run_vit, tokens_to_nchw

vit_configs = dict(  # build_encoder.py:9-28
    vit_h=dict(encoder_embed_dim=1280, encoder_depth=32, encoder_num_heads=16,
               encoder_global_attn_indexes=[7, 15, 23, 31]),
    vit_l=dict(encoder_embed_dim=1024, encoder_depth=24, encoder_num_heads=16,
               encoder_global_attn_indexes=[5, 11, 17, 23]),
    vit_b=dict(encoder_embed_dim=768, encoder_depth=12, encoder_num_heads=12,
               encoder_global_attn_indexes=[2, 5, 8, 11]),
)


def _build_vit(encoder_embed_dim, encoder_depth, encoder_num_heads, encoder_global_attn_indexes, checkpoint=None,
               use_sam_checkpoint=False, project_last_hidden=True):
    """build_encoder.py:43-80"""
    if encoder_embed_dim % encoder_num_heads or encoder_embed_dim // encoder_num_heads != 64:
        # ViT-H: 1280 / 16 = head_dim 80.  The tcgen05 attention kernels stage 64-element (128-byte, one swizzle atom)
        # head slices; there is no head_dim-80 path, and none is faked: fail here, before a multi-GB checkpoint load,
        # like every other out-of-scope option of the reference.
        raise NotImplementedError(
            f"labelanything_b200: SAM ViT with head_dim {encoder_embed_dim // max(encoder_num_heads, 1)} "
            f"(embed_dim {encoder_embed_dim}, {encoder_num_heads} heads) is not supported; the native attention kernels "
            "are built for head_dim 64 (vit_b, vit_l)")
    vit = ImageEncoderViT(depth=encoder_depth, embed_dim=encoder_embed_dim, img_size=1024, mlp_ratio=4,
                          norm_layer=partial(torch.nn.LayerNorm, eps=1e-6), num_heads=encoder_num_heads,
                          patch_size=16, qkv_bias=True, use_rel_pos=True,
                          global_attn_indexes=encoder_global_attn_indexes, project_last_hidden=project_last_hidden,
                          window_size=14, out_chans=256)
    if checkpoint is not None:
        weights = torch.load(checkpoint, map_location="cpu")
        if use_sam_checkpoint:
            weights = {k[len("image_encoder."):]: v for k, v in weights.items() if k.startswith("image_encoder")}
        vit.load_state_dict(weights)
    return vit


def build_vit_h(**kwargs):
    return _build_vit(**vit_configs["vit_h"], **kwargs)


def build_vit_l(**kwargs):
    return _build_vit(**vit_configs["vit_l"], **kwargs)


def build_vit_b(**kwargs):
    return _build_vit(**vit_configs["vit_b"], **kwargs)


class ViTModelWrapper(ViTModel):
    """HF ViT whose forward returns the spatial last hidden state [B, C, H/16, W/16] (build_encoder.py:83-100)."""

    max_images_per_chunk = 64

    # -- packed-weight cache (same contract as common.NativeModule, which we cannot inherit from) ----------
    _cache = NativeModule._cache
    packed = NativeModule.packed
    _packed_eager = NativeModule._packed_eager
    _packed_checked = NativeModule._packed_checked

    def __getstate__(self):
        state = self.__dict__.copy()
        state.pop("_la_cache", None)
        return state

    def _apply(self, fn, *a, **k):
        self.__dict__.pop("_la_cache", None)
        return super()._apply(fn, *a, **k)

    def _w(self, name, t):
        return self.packed("w:" + name, lambda: t.detach().reshape(t.shape[0], -1).to(torch.bfloat16).contiguous(), t)

    def _f(self, name, t):
        return self.packed("f:" + name, lambda: t.detach().float().contiguous(), t)

    def _spec(self, gh: int, gw: int) -> VitSpec:
        cfg = self.config
        d, heads = cfg.hidden_size, cfg.num_attention_heads
        assert gh == gw, "native HF-ViT path expects square inputs"
        if getattr(cfg, "hidden_act", "gelu") != "gelu":
            raise NotImplementedError(f"labelanything_b200: HF ViT hidden_act={cfg.hidden_act!r}; the native MLP epilogue "
                                      "is the exact-erf GELU ('gelu') of the ViT / ViT-MAE checkpoints")
        if d != heads * 64:
            raise NotImplementedError(f"labelanything_b200: HF ViT head_dim {d // heads}; native attention is head_dim 64")
        blocks = []
        for i, layer in enumerate(self.encoder.layer):
            att = layer.attention.attention
            wkv = self.packed(f"l{i}.wkv", lambda att=att: torch.cat([att.key.weight, att.value.weight]).detach()
                              .to(torch.bfloat16).contiguous(), att.key.weight, att.value.weight)
            bkv = None
            if att.key.bias is not None:
                bkv = self.packed(f"l{i}.bkv", lambda att=att: torch.cat([att.key.bias, att.value.bias]).detach()
                                  .float().contiguous(), att.key.bias, att.value.bias)
            blocks.append(BlockWeights(
                self._f(f"l{i}.n1w", layer.layernorm_before.weight), self._f(f"l{i}.n1b", layer.layernorm_before.bias),
                self._w(f"l{i}.q", att.query.weight), None if att.query.bias is None else self._f(f"l{i}.qb", att.query.bias),
                wkv, bkv,
                self._w(f"l{i}.o", layer.attention.output.dense.weight), self._f(f"l{i}.ob", layer.attention.output.dense.bias),
                self._f(f"l{i}.n2w", layer.layernorm_after.weight), self._f(f"l{i}.n2b", layer.layernorm_after.bias),
                self._w(f"l{i}.fc1", layer.intermediate.dense.weight), self._f(f"l{i}.fc1b", layer.intermediate.dense.bias),
                self._w(f"l{i}.fc2", layer.output.dense.weight), self._f(f"l{i}.fc2b", layer.output.dense.bias)))
        return VitSpec(d=d, heads=heads, eps=cfg.layer_norm_eps, blocks=blocks, grid=gh, n_cls=1,
                       final_ln_w=self._f("lnw", self.layernorm.weight), final_ln_b=self._f("lnb", self.layernorm.bias))

    def _pos_table(self, gh: int, gw: int) -> torch.Tensor:
        """[1 + gh*gw, d] fp32 position table, bicubic-resized when the grid differs from the pre-training one
        (transformers modeling_vit.py:61-98, interpolate_pos_encoding).  Constant per resolution -> cached."""
        pos = self.embeddings.position_embeddings

        def build():
            p = pos.detach().float()
            n_pos = p.shape[1] - 1
            if n_pos == gh * gw:
                return p[0].contiguous()
            side = int(n_pos ** 0.5)
            grid = p[:, 1:].reshape(1, side, side, -1).permute(0, 3, 1, 2)
            grid = torch.nn.functional.interpolate(grid, size=(gh, gw), mode="bicubic", align_corners=False)
            return torch.cat([p[0, :1], grid.permute(0, 2, 3, 1).reshape(gh * gw, -1)]).contiguous()

        return self.packed(f"pos:{gh}x{gw}", build, pos)

    def encode_tokens(self, pixel_values: torch.Tensor, out_dtype: torch.dtype = torch.float32):
        ops._require_cuda(pixel_values)
        I, C, H, W = pixel_values.shape
        assert self.config.patch_size == 16 and H % 16 == 0 and W % 16 == 0
        gh, gw = H // 16, W // 16
        spec = self._spec(gh, gw)
        d = spec.d
        proj = self.embeddings.patch_embeddings.projection
        w_pe, b_pe = self._w("patch", proj.weight), self._f("patch.b", proj.bias)
        cls = self._f("cls", self.embeddings.cls_token).view(-1)
        pos = self._pos_table(gh, gw)
        pixel_values = pixel_values.float().contiguous()
        outs = []
        # balanced chunks (208 images -> 4 x 52, not 3 x 64 + 16): every launch of a kernel then has the same size, and
        # the persistent kernels' last-wave loss is paid on fewer, larger launches
        n_chunks = -(-I // self.max_images_per_chunk)
        per_chunk = -(-I // n_chunks)
        for s in range(0, I, per_chunk):
            n = min(per_chunk, I - s)
            cols = ops.im2col_patch16(pixel_values[s:s + n])
            patch = ops.gemm(cols, w_pe, b_pe)
            del cols
            x = torch.empty((n * (gh * gw + 1), d), dtype=torch.float32, device=pixel_values.device)
            ops.embed_tokens(patch, cls, pos, x, n, gh * gw + 1, 1, d)
            del patch
            outs.append(run_vit(spec, x, n, out_dtype))
        return (outs[0] if len(outs) == 1 else torch.cat(outs)), gh

    def forward(self, pixel_values):
        feats, g = self.encode_tokens(pixel_values, torch.float32)
        return tokens_to_nchw(feats, pixel_values.shape[0], g)


def build_vit_b_mae(project_last_hidden=False):
    return ViTModelWrapper.from_pretrained("facebook/vit-mae-base")


def build_vit_b_imagenet_i21k(project_last_hidden=False):
    return ViTModelWrapper.from_pretrained("google/vit-base-patch16-224-in21k")


def build_vit_from_config(hidden_size=768, num_hidden_layers=12, num_attention_heads=12, intermediate_size=3072,
                          image_size=224, project_last_hidden=False):
    """Randomly initialised HF ViT of the given geometry (offline stand-in for `from_pretrained`)."""
    from transformers import ViTConfig

    return ViTModelWrapper(ViTConfig(hidden_size=hidden_size, num_hidden_layers=num_hidden_layers,
                                     num_attention_heads=num_attention_heads, intermediate_size=intermediate_size,
                                     image_size=image_size, patch_size=16))


def build_encoder(name, **kwargs):
    if name in ENCODERS:
        return ENCODERS[name](**kwargs)
    raise ValueError(f"unknown encoder {name!r}; labelanything_b200 provides {sorted(ENCODERS)}")


ENCODERS = {  # build_encoder.py:144-152 (pyramid / DINO-8 backbones are out of the hot-path scope)
    "vit_h": build_vit_h,
    "vit_l": build_vit_l,
    "vit_b": build_vit_b,
    "vit_b_mae": build_vit_b_mae,
    "vit_b_imagenet_i21k": build_vit_b_imagenet_i21k,
}
