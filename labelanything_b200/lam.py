"""`Lam` — the LabelAnything model — with the reference's surface, running on the native kernels.

Mirror of label_anything/models/lam.py:24-453: `forward` / `_forward` / `prepare_query_example_embeddings` /
`prepare_prompts` / `generate_class_embeddings` / `predict` / `postprocess_masks` / `init_pretrained_weights` /
`get_learnable_params`, sub-modules `image_encoder`, `neck`, `prompt_encoder`, `mask_decoder` (same state-dict keys).

Inside, features stay token-major ([images * h*w, D]) from the encoder's last kernel to the decoder's first one;
NCHW tensors exist only at the public boundary (`embeddings` input, `image_encoder(x)` output).
There is no CPU / eager path: inputs must live on a CUDA (sm_100a) device.
"""
from __future__ import annotations

from typing import Any, Dict, Optional, Tuple

import torch
import torch.nn as nn

from . import ops
from .common import SAM_EMBED_DIM, NativeModule
from .transformer import TwoWayTransformer
from .utils import BatchKeys, ResultDict, get_preprocess_shape
from .vit_engine import pack_neck, run_neck


class Lam(NativeModule):
    mask_threshold: float = 0.0
    image_format: str = "RGB"

    def __init__(self, image_encoder: Optional[nn.Module], prompt_encoder: nn.Module, mask_decoder: nn.Module,
                 neck: Optional[nn.Module], image_size: int = 1024, custom_preprocess: bool = True) -> None:
        super().__init__()
        self.image_size = image_size
        self.image_encoder = image_encoder
        self.prompt_encoder = prompt_encoder
        self.mask_decoder = mask_decoder
        self.class_embeddings = None
        self.neck = neck
        self.custom_preprocess = custom_preprocess

    # ------------------------------------------------------------------ features
    def _features(self, batched_input: Dict[str, Any]) -> Tuple[torch.Tensor, int, int, int]:
        """-> (token-major fp32 features [B*N*T, D], B, N, grid) for either input mode (lam.py:138-170)."""
        if "embeddings" in batched_input:
            emb = batched_input["embeddings"]
            if isinstance(emb, dict):
                raise NotImplementedError("feature pyramids (PyramidNeck) are outside the native hot path")
            ops._require_cuda(emb)
            B, N, C, H, W = emb.shape
            assert H == W, "native kernels expect square feature maps"
            flat = emb.reshape(B * N, C, H, W).float().contiguous()
            if self.neck is not None:
                _, t16 = ops.nchw_to_tokens(flat, want_f32=False, want_bf16=True)
                feats = run_neck(pack_neck(self, self.neck), t16, B * N, H, torch.float32)
            else:
                feats, _ = ops.nchw_to_tokens(flat)
            return feats, B, N, H
        if "images" in batched_input:
            images = batched_input["images"]
            ops._require_cuda(images)
            B, N = images.shape[:2]
            flat = images.reshape(B * N, *images.shape[2:])
            if self.image_encoder is None:
                raise ValueError("this model was built without an image encoder (lam_no_vit): pass 'embeddings'")
            if self.neck is not None:
                t16, g = self.image_encoder.encode_tokens(flat, torch.bfloat16)
                feats = run_neck(pack_neck(self, self.neck), t16, B * N, g, torch.float32)
            else:
                feats, g = self.image_encoder.encode_tokens(flat, torch.float32)
            return feats, B, N, g
        raise ValueError("Either 'images' or 'embeddings' must be provided.")  # lam.py:165

    def prepare_query_example_embeddings(self, batched_input):
        """Reference-shaped result (lam.py:138-170): (query [B, D, h, w], support [B, M, D, h, w]) fp32."""
        emb = self.prepare_embeddings_example(batched_input)
        return emb[:, 0], emb[:, 1:]

    def prepare_embeddings_example(self, batched_input):
        """[B, N, D, h, w] fp32 (lam.py:173-190)."""
        feats, B, N, g = self._features(batched_input)
        return ops.tokens_to_nchw(feats, B * N, g, g).view(B, N, -1, g, g)

    def prepare_embeddings(self, batched_input, chunk_size=None):
        """lam.py:192-212 (chunking changes nothing numerically; the native encoder chunks internally)."""
        return self.prepare_embeddings_example(batched_input)

    def prepare_prompts(self, batched_input):
        """Drop prompt types whose flags are all zero (lam.py:214-239); one host sync instead of three."""
        present = [(k, f) for k, f in ((BatchKeys.PROMPT_POINTS, BatchKeys.FLAG_POINTS),
                                       (BatchKeys.PROMPT_BBOXES, BatchKeys.FLAG_BBOXES),
                                       (BatchKeys.PROMPT_MASKS, BatchKeys.FLAG_MASKS)) if k in batched_input]
        out = {BatchKeys.PROMPT_POINTS: None, BatchKeys.PROMPT_BBOXES: None, BatchKeys.PROMPT_MASKS: None}
        if present:
            any_set = torch.stack([(batched_input[f] != 0).any() for _, f in present]).tolist()
            for (k, f), on in zip(present, any_set):
                if on:
                    out[k] = (batched_input[k], batched_input[f])
        return (out[BatchKeys.PROMPT_POINTS], out[BatchKeys.PROMPT_BBOXES], out[BatchKeys.PROMPT_MASKS],
                batched_input[BatchKeys.FLAG_EXAMPLES])

    def get_dense_pe(self):
        return self.prompt_encoder.get_dense_pe()

    # ------------------------------------------------------------------ forward paths
    def _forward(self, batched_input) -> Tuple[torch.Tensor, Dict[str, torch.Tensor]]:
        feats, B, N, g = self._features(batched_input)
        M, T = N - 1, g * g
        points, boxes, masks, flag_examples = self.prepare_prompts(batched_input)
        pe_result = self.prompt_encoder.encode(feats, B, M, points, boxes, masks, flag_examples, feat_lead=1)
        q32, q16 = ops.copy_slabs(feats, B, T, N * T, 0, want_f32=True, want_bf16=True)
        seg = self.mask_decoder.decode(q32, q16, self.prompt_encoder.dense_pe_tokens(),
                                       pe_result[ResultDict.CLASS_EMBS], B, g, g, pe_cached=True)
        return seg, pe_result

    def forward(self, batched_input: Dict[str, Any]) -> Dict[str, torch.Tensor]:
        """batched_input: the reference's dict (lam.py:57-100; data/utils.py:43-58) -> {"logits" [B, C, Hmax, Wmax],
        "class_examples_embeddings" [B, M, C, D]}."""
        seg, pe_result = self._forward(batched_input)
        seg = self.postprocess_masks(seg, batched_input["dims"], flag_gts=batched_input.get("flag_gts"))
        return {ResultDict.LOGITS: seg, ResultDict.EXAMPLES_CLASS_EMBS: pe_result[ResultDict.EXAMPLES_CLASS_EMBS]}

    def generate_class_embeddings(self, example_dict, chunk_size=None):
        """lam.py:349-360: every image of `example_dict` is a support image."""
        feats, B, N, g = self._features(example_dict)
        points, boxes, masks, flag_examples = self.prepare_prompts(example_dict)
        return self.prompt_encoder.encode(feats, B, N, points, boxes, masks, flag_examples, feat_lead=0)

    def predict(self, batched_input, class_embeddings=None):
        """lam.py:362-381: decode the query image(s) against cached class embeddings."""
        if class_embeddings is None and self.class_embeddings is None:
            return self.forward(batched_input)
        if class_embeddings is None:
            class_embeddings = self.class_embeddings
        feats, B, N, g = self._features(batched_input)
        T = g * g
        q32, q16 = ops.copy_slabs(feats, B, T, N * T, 0, want_f32=True, want_bf16=True)
        ce = class_embeddings[ResultDict.CLASS_EMBS]
        seg = self.mask_decoder.decode(q32, q16, self.prompt_encoder.dense_pe_tokens(), ce, B, g, g, pe_cached=True)
        return self.postprocess_masks(seg, batched_input["dims"].unsqueeze(1))

    def postprocess_masks(self, masks: torch.Tensor, original_sizes: torch.Tensor,
                          flag_gts: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Bilinear to image_size, crop the un-padded region, bilinear to each query's original size, pad to the
        batch maximum with -inf (background: 0); classes absent from `flag_gts` -> -inf.  One kernel
        (lam.py:383-453, 92-93).  original_sizes [B, M+1, 2] (H, W), row 0 = query."""
        ops._require_cuda(masks)
        sizes_host = original_sizes.detach().to("cpu", torch.int64)
        max_h, max_w = (int(v) for v in sizes_host.view(-1, 2).max(dim=0).values)
        rows = []
        for oh, ow in sizes_host[:, 0, :].tolist():
            ih, iw = get_preprocess_shape(oh, ow, self.image_size) if self.custom_preprocess else \
                (self.image_size, self.image_size)
            rows.append((oh, ow, ih, iw))
        sizes = torch.tensor(rows, dtype=torch.int32).to(masks.device, non_blocking=True)
        fg = None
        if flag_gts is not None:
            fg = (flag_gts != 0).to(device=masks.device, dtype=torch.uint8).contiguous()
        return ops.postprocess_masks(masks.float().contiguous(), sizes, fg, self.image_size, max_h, max_w)

    # ------------------------------------------------------------------ weights / optimiser plumbing
    def init_pretrained_weights(self, weights):
        """Initialise from a SAM checkpoint (lam.py:241-319)."""
        def sub(prefix):
            return {k[len(prefix):]: v for k, v in weights.items() if k.startswith(prefix)}

        if self.image_encoder is not None:
            self.image_encoder.load_state_dict(sub("image_encoder."))
        pe = self.prompt_encoder
        if pe.pe_layer.positional_encoding_gaussian_matrix.shape[1] == 2 * SAM_EMBED_DIM:
            pe.pe_layer.load_state_dict(sub("prompt_encoder.pe_layer."))
            pe.point_embeddings.load_state_dict(sub("prompt_encoder.point_embeddings."))
            pe.not_a_point_embed.load_state_dict(sub("prompt_encoder.not_a_point_embed."))
            pe.mask_downscaling.load_state_dict(sub("prompt_encoder.mask_downscaling."))
            pe.no_mask_embed.load_state_dict(sub("prompt_encoder.no_mask_embed."))
            tw = sub("mask_decoder.transformer.")
            if pe.transformer.attention_downsample_rate == 2:
                pe.transformer.load_state_dict(tw)
            if (isinstance(self.mask_decoder.transformer, TwoWayTransformer)
                    and self.mask_decoder.transformer.attention_downsample_rate == 2):
                self.mask_decoder.transformer.load_state_dict(dict(tw))
            self.mask_decoder.output_upscaling.load_state_dict(sub("mask_decoder.output_upscaling."))

    def get_learnable_params(self, training_params: dict):
        """lam.py:321-347"""
        def not_encoder(x):
            return "image_encoder" not in x[0]

        freeze = training_params.get("freeze_backbone", False)
        if freeze and "backbone_lr" in training_params:
            raise ValueError("Cannot freeze the backbone and set a learning rate for it at the same time.")
        if freeze:
            for p in self.image_encoder.parameters():
                p.requires_grad = False
            return [x[1] for x in filter(not_encoder, list(self.named_parameters()))]
        if "backbone_lr" in training_params:
            return [{"params": self.image_encoder.parameters(), "lr": training_params["backbone_lr"]},
                    {"params": [x[1] for x in filter(not_encoder, list(self.named_parameters()))]}]
        return self.parameters()
