"""Training step of the frozen / pre-computed-encoder configuration (BASELINE config 4).

What the reference does per step (label_anything/experiment/run.py:425-590, experiment/utils.py:266-288):
    result = model(input_dict); loss = LabelAnythingLoss(result, gt); accelerator.backward(loss); optimizer.step()
with the model = `lam_no_vit` on pre-computed embeddings (parameters/trainval/coco/mael.yaml), DDP over the GPUs
(`find_unused_parameters=True`: prompt types absent from a batch leave their embeddings without gradient) and AdamW.

Here:
  * `train_forward(lam, batch)` is the differentiable launch sequence of `Lam.forward(embeddings)` -- neck, prompt
    encoder, mask decoder, postprocess_masks -- written against the SAME parameter-holding modules as the inference
    path (same state-dict keys), op for op as the reference computes it (no inference-only algebra: every projection
    of the two-way transformers is materialised, because its weight needs a gradient).  Every op is a
    `train_ops` autograd Function whose forward and backward are native launches;
  * `FlatAdamW` keeps parameters, gradients and the two moments in flat fp32 buckets (the parameters / .grad of the
    modules are views), all-reduces the gradient bucket with ONE NCCL call (the reference: DDP's bucketed all-reduce)
    and updates everything with one `la_adamw_f32` launch; parameters that received no gradient on any rank are
    skipped, like torch.optim.AdamW skips `grad is None`;
  * `train_step(lam, loss_fn, opt, batch, gt)` = forward + loss + backward + all-reduce + update.
"""
from __future__ import annotations

from typing import Any, Dict, List, Optional, Tuple

import torch
import torch.nn as nn

from . import ops
from . import train_ops as T
from .common import Attention, AttentionMLPBlock, MLPBlock
from .lam import Lam
from .prompt_encoder import PromptImageEncoder, RandomMatrixEncoder
from .transformer import TwoWayTransformer
from .utils import BatchKeys, ResultDict, get_preprocess_shape

__all__ = ["train_forward", "make_plan", "FlatAdamW", "train_step", "GraphedTrainStep", "ConstantWithWarmup"]


# ----------------------------------------------------------------------------------------------------------------
# building blocks (reference: label_anything/models/common.py, transformer.py)
# ----------------------------------------------------------------------------------------------------------------
def _lin(x: torch.Tensor, m: nn.Linear, act: int = ops.ACT_NONE) -> torch.Tensor:
    return T.linear(x, m.weight, m.bias, act)


def _ln(x: torch.Tensor, m, act: int = ops.ACT_NONE) -> torch.Tensor:
    return T.layernorm(x, m.weight, m.bias, m.eps, act)


def _attn(att: Attention, q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, n_seq: int, nq: int, nk: int) -> torch.Tensor:
    """common.py:97-148: projections, per-head softmax(QK^T / sqrt(dh)) V, output projection (masks are no-ops)."""
    o = T.attention(_lin(q, att.q_proj), _lin(k, att.k_proj), _lin(v, att.v_proj), n_seq, nq, nk, att.num_heads)
    return _lin(o, att.out_proj)


def _mlp(m: MLPBlock, x: torch.Tensor) -> torch.Tensor:
    """common.py:19-37"""
    if isinstance(m.act, nn.ReLU):
        return _lin(_lin(x, m.lin1, ops.ACT_RELU), m.lin2)
    if isinstance(m.act, nn.GELU) and getattr(m.act, "approximate", "none") == "none":
        return _lin(T.gelu(_lin(x, m.lin1)), m.lin2)
    raise NotImplementedError(f"activation {type(m.act).__name__} has no native training path (GELU / ReLU only)")


def _attention_mlp_block(blk: AttentionMLPBlock, x: torch.Tensor, n_seq: int, L: int) -> torch.Tensor:
    """common.py:151-184: a = norm(attn(x, x, x) + x); out = norm(mlp(a) + a) -- the same LayerNorm twice."""
    a = _ln(T.add(_attn(blk.attn, x, x, x, n_seq, L, L), x), blk.norm)
    return _ln(T.add(_mlp(blk.mlp, a), a), blk.norm)


def _two_way(tw: TwoWayTransformer, keys: torch.Tensor, pe: torch.Tensor, tokens: torch.Tensor, S: int, Tn: int, n: int,
             want_queries: bool) -> Tuple[Optional[torch.Tensor], torch.Tensor]:
    """transformer.py:206-252,298-329.  keys fp32 [S*Tn, D] image tokens, pe fp32 [Tn, D] (constant), tokens fp32
    [S*n, D] (also their positional encoding) -> (queries [S*n, D] | None, keys [S*Tn, D])."""
    queries, qpe = tokens, tokens
    for layer in tw.layers:
        if layer.skip_first_layer_pe:
            queries = _attn(layer.self_attn, queries, queries, queries, S, n, n)
        else:
            q = T.add(queries, qpe)
            queries = T.add(queries, _attn(layer.self_attn, q, q, queries, S, n, n))
        queries = _ln(queries, layer.norm1)
        q = T.add(queries, qpe)
        k = T.add_bcast(keys, pe, 1, Tn)
        queries = _ln(T.add(queries, _attn(layer.cross_attn_token_to_image, q, k, keys, S, n, Tn)), layer.norm2)
        queries = _ln(T.add(queries, _mlp(layer.mlp, queries)), layer.norm3)
        q = T.add(queries, qpe)
        keys = _ln(T.add(keys, _attn(layer.cross_attn_image_to_token, k, q, queries, S, Tn, n)), layer.norm4)
    if not want_queries:
        return None, keys
    q = T.add(queries, qpe)
    k = T.add_bcast(keys, pe, 1, Tn)
    queries = _ln(T.add(queries, _attn(tw.final_attn_token_to_image, q, k, keys, S, n, Tn)), tw.norm_final_attn)
    return queries, keys


# ----------------------------------------------------------------------------------------------------------------
# neck, prompt encoder, mask decoder
# ----------------------------------------------------------------------------------------------------------------
def _neck(neck: nn.Sequential, x: torch.Tensor, n_img: int, g: int) -> torch.Tensor:
    """build_lam.py:150-171: Conv2d 1x1 (no bias) -> LayerNorm2d -> Conv2d 3x3 (no bias) -> LayerNorm2d."""
    c1, n1, c3, n2 = neck[0], neck[1], neck[2], neck[3]
    y = _ln(T.linear(x, c1.weight, c1.bias), n1)
    return _ln(T.conv3x3(y, c3.weight, c3.bias, n_img, g, g), n2)


def _prompt_encoder(pe_mod: PromptImageEncoder, feat: torch.Tensor, B: int, M: int, points, boxes, masks,
                    flag_examples: torch.Tensor, class_rows: Optional[torch.Tensor] = None) -> Dict[str, torch.Tensor]:
    """prompt_encoder.py:752-827 (+ 564-669, 696-750).  feat fp32 [B*M*T, D] support features."""
    any_prompt = points[0] if points is not None else boxes[0] if boxes is not None else \
        masks[0] if masks is not None else None
    if any_prompt is None:
        raise ValueError("No prompts provided")  # prompt_encoder.py:562
    C = any_prompt.shape[2]
    D = pe_mod.embed_dim
    h, w = pe_mod.image_embedding_size
    Tn, S = h * w, B * M * C
    dev = feat.device

    # ---- sparse tokens
    if points is not None or boxes is not None:
        pts = lab = bx = bfl = None
        if points is not None:
            pts = points[0].reshape(S, -1, 2).float().contiguous()
            lab = points[1].reshape(S, -1).float().contiguous()
        if boxes is not None:
            bx = boxes[0].reshape(S, -1, 4).float().contiguous()
            bfl = boxes[1].reshape(S, -1).float().contiguous()
        tab4 = torch.cat([e.weight for e in pe_mod.point_embeddings])
        gauss = pe_mod.pe_layer.positional_encoding_gaussian_matrix.detach().float().contiguous()
        sparse = T.embed_sparse(tab4, pe_mod.not_a_point_embed.weight, pts, lab, bx, bfl, gauss, S, D,
                                pe_mod.input_image_size[1], pe_mod.input_image_size[0])
        n = sparse.shape[1]
        sparse = sparse.view(S * n, D)
    else:
        n = 1
        zero = torch.zeros((S, D), dtype=torch.float32, device=dev)
        sparse = T.add_bcast(zero, pe_mod.no_sparse_embedding.weight, 1, 1)
    sparse = _attention_mlp_block(pe_mod.sparse_embedding_attention, sparse, B * M, C * n)

    # ---- class code
    code = None
    ce = pe_mod.class_encoder
    if isinstance(ce, RandomMatrixEncoder):
        rows = class_rows if class_rows is not None else ce.sample_rows(C, dev)
        code = ce.pos_embedding[0, 0].index_select(0, rows)              # [C, D], differentiable gather of the bank
        sparse = T.add_bcast(sparse, code, n, C)
    elif isinstance(ce, nn.Module):
        raise NotImplementedError(f"class encoder {type(ce).__name__} has no native path")

    # ---- dense mask embedding (mask_downscaling; the bilinear resize commutes with the last 1x1 convolution)
    dense = mflags = None
    if masks is not None:
        mk, mf = masks
        Hm, Wm = mk.shape[-2:]
        m16 = T.mask_downscale(mk.reshape(S, Hm, Wm).float().contiguous(), pe_mod.mask_downscaling)
        if (Hm // 4, Wm // 4) != (h, w):
            m16 = T.resize_bilinear(m16, h, w)
        md6 = pe_mod.mask_downscaling[6]
        dense = T.linear(m16.view(S * Tn, 16), md6.weight, md6.bias)
        mflags = (mf.reshape(S) != 0).to(torch.uint8).contiguous()
        alt = pe_mod.not_a_mask_embed.weight
    else:
        alt = pe_mod.no_mask_embed.weight
    src = T.src_combine(feat, dense, alt, mflags, S, Tn, D, C)
    if code is not None:
        src = T.add_bcast(src, code, Tn, C)

    pe = pe_mod.dense_pe_tokens()
    _, fused = _two_way(pe_mod.transformer, src, pe, sparse, S, Tn, n, want_queries=False)
    emb = T.segment_mean(fused, S, Tn)                                    # [S, D] = [B, M, C, D]

    if pe_mod.class_attention is not None:
        emb = _attention_mlp_block(pe_mod.class_attention, emb, B * M, C)
    if pe_mod.example_attention is not None:
        e = T.permute_rows(emb, B, M, C)
        e = _attention_mlp_block(pe_mod.example_attention, e, B * C, M)
        emb = T.permute_rows(e, B, C, M)
    if pe_mod.class_example_attention is not None:
        emb = _attention_mlp_block(pe_mod.class_example_attention, emb, B, M * C)

    fe8 = (flag_examples != 0).to(torch.uint8).contiguous()
    emb4 = emb.view(B, M, C, D)
    return {ResultDict.CLASS_EMBS: T.masked_mean(emb4, fe8), ResultDict.EXAMPLES_CLASS_EMBS: emb4}


def _mask_decoder(md, query: torch.Tensor, pe: torch.Tensor, class_emb: torch.Tensor, B: int, h: int, w: int) -> torch.Tensor:
    """mask_decoder.py:316-363: query fp32 [B*h*w, D], class_emb [B, C, D] -> low-resolution logits [B, C, 4h, 4w]."""
    D = md.attention_dim
    Tn = h * w
    C = class_emb.shape[1]
    queries, keys = _two_way(md.transformer, query, pe, class_emb.reshape(B * C, D), B, Tn, C, want_queries=True)
    c = queries
    for i, layer in enumerate(md.class_mlp.layers):
        c = _lin(c, layer, ops.ACT_NONE if i == md.class_mlp.num_layers - 1 else ops.ACT_RELU)

    up = md.output_upscaling
    c1, c2 = up[0].weight.shape[1], up[3].weight.shape[1]
    # ConvTranspose2d(k = s = 2): per-pixel GEMM to (ky, kx, co) columns, then the pixel shuffle as a row permutation
    u = T.linear(keys, up[0].weight.permute(2, 3, 1, 0).reshape(4 * c1, D), up[0].bias.repeat(4))
    u = T.permute_rows(u.view(B * Tn * 2, 2 * c1), B * h, w, 2).view(B * Tn * 4, c1)
    x = _ln(u, up[1], ops.ACT_GELU)
    u = T.linear(x, up[3].weight.permute(2, 3, 1, 0).reshape(4 * c2, c1), up[3].bias.repeat(4))
    x = T.permute_rows(u.view(B * Tn * 8, 2 * c2), B * 2 * h, 2 * w, 2).view(B * Tn * 16, c2)
    H4, W4 = 4 * h, 4 * w
    if md.spatial_convs is not None:
        mods = list(md.spatial_convs)
        i = 0
        while i < len(mods):
            x = T.conv3x3(x, mods[i].weight, mods[i].bias, B, H4, W4)
            if i + 1 < len(mods):
                x = _ln(x, mods[i + 1], ops.ACT_GELU)
                i += 3
            else:
                i += 1
    return T.classify(x, c.view(B, C, c2), B, H4 * W4).view(B, C, H4, W4)


# ----------------------------------------------------------------------------------------------------------------
# Lam.forward(embeddings) with gradients
# ----------------------------------------------------------------------------------------------------------------
def make_plan(lam: Lam, batched_input: Dict[str, Any]) -> Dict[str, Any]:
    """Everything `train_forward` needs from the HOST side of a batch, computed once: which prompt types are present
    (lam.py:214-239 reads the flags back), the per-episode sizes of `postprocess_masks` (lam.py:383-453 reads `dims`
    back) and the class-code rows.  A plan can be reused for every batch of the same geometry -- which is what a captured
    CUDA graph of the step requires (`GraphedTrainStep`)."""
    points, boxes, masks, _ = lam.prepare_prompts(batched_input)
    dims = batched_input["dims"]
    sizes_host = dims.detach().to("cpu", torch.int64)
    max_h, max_w = (int(v) for v in sizes_host.view(-1, 2).max(dim=0).values)
    rows = []
    for oh, ow in sizes_host[:, 0, :].tolist():
        ih, iw = get_preprocess_shape(oh, ow, lam.image_size) if lam.custom_preprocess else (lam.image_size, lam.image_size)
        rows.append((oh, ow, ih, iw))
    dev = (batched_input["embeddings"] if "embeddings" in batched_input else batched_input["images"]).device
    plan = {"points": points is not None, "boxes": boxes is not None, "masks": masks is not None,
            "sizes": torch.tensor(rows, dtype=torch.int32).to(dev), "out_hw": (max_h, max_w), "class_rows": None}
    ce = lam.prompt_encoder.class_encoder
    if isinstance(ce, RandomMatrixEncoder):
        C = batched_input[BatchKeys.FLAG_EXAMPLES].shape[2]
        plan["class_rows"] = ce.sample_rows(C, dev)
    return plan


def train_forward(lam: Lam, batched_input: Dict[str, Any], plan: Optional[Dict[str, Any]] = None) -> Dict[str, torch.Tensor]:
    """Differentiable `Lam.forward` on pre-computed embeddings (lam.py:57-170 with the `embeddings` key) or on images
    through a FROZEN image encoder: returns
    {"logits" [B, C, Hmax, Wmax], "class_examples_embeddings" [B, M, C, D]} attached to the autograd graph of the
    parameters of lam.neck / lam.prompt_encoder / lam.mask_decoder.  With a `plan` (make_plan) the call performs no
    host synchronisation at all."""
    if plan is None:
        plan = make_plan(lam, batched_input)
    if "embeddings" in batched_input:
        emb = batched_input["embeddings"]
        ops._require_cuda(emb)
        B, N, Ce, H, W = emb.shape
        assert H == W, "native kernels expect square feature maps"
        g, Tn, M = H, H * W, N - 1
        feats, _ = ops.nchw_to_tokens(emb.reshape(B * N, Ce, H, W).float().contiguous())   # [B*N*T, Ce] (input: no grad)
    elif "images" in batched_input and lam.image_encoder is not None:
        # FROZEN encoder (train_params.freeze_backbone, lam.py:321-347): the ViT runs on its inference kernels without an
        # autograd graph -- it has no backward kernels -- and hands its token-major features to the differentiable part
        images = batched_input["images"]
        ops._require_cuda(images)
        B, N = images.shape[:2]
        if any(p.requires_grad for p in lam.image_encoder.parameters()):
            raise NotImplementedError("the native training step trains neck + prompt encoder + mask decoder; freeze the "
                                      "image encoder (get_learnable_params({'freeze_backbone': True})) or pass "
                                      "pre-computed 'embeddings' (the ViT has no backward kernels)")
        with torch.no_grad():
            feats, g = lam.image_encoder.encode_tokens(images.reshape(B * N, *images.shape[2:]), torch.float32)
        Tn, M = g * g, N - 1
    else:
        raise ValueError("Either 'images' or 'embeddings' must be provided.")  # lam.py:165
    if lam.neck is not None:
        feats = _neck(lam.neck, feats, B * N, g)
    D = feats.shape[1]
    per_ep = feats.view(B, N * Tn * D)
    query = per_ep[:, :Tn * D].reshape(B * Tn, D)                    # image 0 of every episode
    support = per_ep[:, Tn * D:].reshape(B * M * Tn, D)
    bi = batched_input
    points = (bi[BatchKeys.PROMPT_POINTS], bi[BatchKeys.FLAG_POINTS]) if plan["points"] else None
    boxes = (bi[BatchKeys.PROMPT_BBOXES], bi[BatchKeys.FLAG_BBOXES]) if plan["boxes"] else None
    masks = (bi[BatchKeys.PROMPT_MASKS], bi[BatchKeys.FLAG_MASKS]) if plan["masks"] else None
    pe_result = _prompt_encoder(lam.prompt_encoder, support, B, M, points, boxes, masks, bi[BatchKeys.FLAG_EXAMPLES],
                                class_rows=plan["class_rows"])
    seg = _mask_decoder(lam.mask_decoder, query, lam.prompt_encoder.dense_pe_tokens(), pe_result[ResultDict.CLASS_EMBS],
                        B, g, g)
    fg = batched_input.get("flag_gts")
    if fg is not None:
        fg = (fg != 0).to(device=seg.device, dtype=torch.uint8).contiguous()
    logits = T.postprocess_masks(seg, plan["sizes"], fg, lam.image_size, *plan["out_hw"])    # lam.py:383-453
    return {ResultDict.LOGITS: logits, ResultDict.EXAMPLES_CLASS_EMBS: pe_result[ResultDict.EXAMPLES_CLASS_EMBS]}


# ----------------------------------------------------------------------------------------------------------------
# optimiser over flat buckets + the one gradient all-reduce
# ----------------------------------------------------------------------------------------------------------------
class FlatAdamW:
    """torch.optim.AdamW semantics over flat fp32 buckets (experiment/run.py:172-200 builds AdamW over
    `get_learnable_params`; DDP all-reduces the gradients, run.py:122-124).

    The parameters become views of one flat buffer, their `.grad` views of a second one whose tail holds one "used"
    counter per parameter; after backward ONE all-reduce (SUM) of that buffer over the process group delivers summed
    gradients and which parameters received a gradient on any rank; the update divides by the world size and runs
    `la_adamw_f32` over maximal runs of used parameters with equal step counts (normally one launch).  Parameters
    unused on every rank keep their value, moments and step count, like `grad is None` in torch.optim.AdamW."""

    def __init__(self, params, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 1e-2,
                 process_group=None) -> None:
        self.params: List[nn.Parameter] = [p for p in params if p.requires_grad]
        assert self.params, "no trainable parameters"
        dev = self.params[0].device
        assert all(p.device == dev and p.dtype == torch.float32 for p in self.params), \
            "FlatAdamW: fp32 parameters on one device"
        self.lr, self.betas, self.eps, self.weight_decay = lr, betas, eps, weight_decay
        self.group = process_group
        self.offsets, n = [], 0
        for p in self.params:
            self.offsets.append(n)
            n += (p.numel() + 3) // 4 * 4                     # 16-byte aligned views
        self.numel = n
        self.flat_p = torch.zeros(n, dtype=torch.float32, device=dev)
        self.flat_g = torch.zeros(n + len(self.params), dtype=torch.float32, device=dev)   # gradients | used counters
        self.exp_avg = torch.zeros(n, dtype=torch.float32, device=dev)
        self.exp_avg_sq = torch.zeros(n, dtype=torch.float32, device=dev)
        self.steps = [0] * len(self.params)
        self._used = [False] * len(self.params)
        with torch.no_grad():
            for i, (p, o) in enumerate(zip(self.params, self.offsets)):
                self.flat_p[o:o + p.numel()].copy_(p.detach().reshape(-1))
                p.data = self.flat_p[o:o + p.numel()].view(p.shape)
                p.grad = self.flat_g[o:o + p.numel()].view(p.shape)
                p.register_post_accumulate_grad_hook(self._mark(i))
        self.last_allreduce_ms: Optional[float] = None

    def attach(self, module: nn.Module) -> None:
        """Keep `module.state_dict()` checkpoint-friendly: the parameters are views of ONE flat storage now, which
        `safetensors.torch.save_model` (huggingface_hub's `save_pretrained`, accelerate's `save_state`) refuses ("none is
        covering the entire storage").  A state-dict hook hands out copies of such views instead.  Idempotent; called by
        `train_step` / `GraphedTrainStep`."""
        if module.__dict__.get("_la_flat_state_dict_hook"):
            return

        def hook(mod, state_dict, prefix, local_metadata):
            for k, v in list(state_dict.items()):
                if torch.is_tensor(v) and v.untyped_storage().nbytes() > v.numel() * v.element_size():
                    state_dict[k] = v.detach().clone()

        module.register_state_dict_post_hook(hook)
        module.__dict__["_la_flat_state_dict_hook"] = True

    def _mark(self, i: int):
        def hook(_p):
            self._used[i] = True
        return hook

    def _check_views(self) -> None:
        for i in (0, len(self.params) - 1):
            if self.params[i].data_ptr() != self.flat_p.data_ptr() + 4 * self.offsets[i]:
                raise RuntimeError("FlatAdamW: the parameters no longer live in the optimiser's flat bucket (the model was "
                                   "moved / re-materialised, e.g. model.to(...) or load_state_dict(assign=True), after the "
                                   "optimiser was built); build a new FlatAdamW")

    def zero_grad(self) -> None:
        self._check_views()
        self.flat_g.zero_()
        self._used = [False] * len(self.params)
        for p, o in zip(self.params, self.offsets):     # autograd may have replaced .grad (it does not when one is set)
            if p.grad is None or p.grad.data_ptr() != self.flat_g.data_ptr() + 4 * o:
                p.grad = self.flat_g[o:o + p.numel()].view(p.shape)

    def used_runs(self, used: List[bool]) -> List[Tuple[int, int, int]]:
        """Maximal runs [(first offset, one-past-last offset, step count)] of consecutive used parameters that share a
        step count."""
        return self.used_runs_for(self.steps, used)

    def used_runs_for(self, steps: List[int], used: List[bool]) -> List[Tuple[int, int, int]]:
        runs, i, n = [], 0, len(self.params)
        while i < n:
            if not used[i]:
                i += 1
                continue
            j = i
            while j + 1 < n and used[j + 1] and steps[j + 1] == steps[i]:
                j += 1
            end = self.offsets[j + 1] if j + 1 < n else self.numel
            runs.append((self.offsets[i], end, steps[i]))
            i = j + 1
        return runs

    def reduce_gradients(self, timed: bool = False) -> Tuple[List[bool], int]:
        """The step's only collective: one all-reduce (SUM) of [gradients | used counters] over the process group
        (NCCL over NVLink on the GPUs; device-agnostic, so the host logic is testable with gloo).  Returns (which
        parameters received a gradient on any rank, world size)."""
        import torch.distributed as dist

        tail = self.flat_g[self.numel:]
        tail.copy_(torch.tensor([1.0 if u else 0.0 for u in self._used], dtype=torch.float32), non_blocking=True)
        if not (dist.is_available() and dist.is_initialized() and dist.get_world_size(self.group) > 1):
            return list(self._used), 1
        timed = timed and self.flat_g.is_cuda
        if timed:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
        dist.all_reduce(self.flat_g, group=self.group)
        if timed:
            e1.record()
            e1.synchronize()
            self.last_allreduce_ms = e0.elapsed_time(e1)
        return (tail > 0).tolist(), dist.get_world_size(self.group)

    def step(self, timed: bool = False) -> None:
        used, world = self.reduce_gradients(timed)
        self.apply(used, world)

    def apply(self, used: List[bool], world: int) -> None:
        """AdamW update of the parameters in `used` from the (already reduced) gradient bucket."""
        for i, u in enumerate(used):
            if u:
                self.steps[i] += 1
        for lo, hi, step in self.used_runs(used):        # runs are formed on the incremented step counts
            T.adamw_step(self.flat_p[lo:hi], self.flat_g[lo:hi], self.exp_avg[lo:hi], self.exp_avg_sq[lo:hi], self.lr,
                         self.betas[0], self.betas[1], self.eps, self.weight_decay, step, 1.0 / world)
        # the kernel wrote through raw pointers: bump the version counters so that the packed-weight caches of the
        # inference path (keyed on data_ptr / _version, common.NativeModule.packed) see the update
        torch.autograd.graph.increment_version([p for p, u in zip(self.params, used) if u])


def train_step(lam: Lam, loss_fn, opt: FlatAdamW, batched_input: Dict[str, Any], gt: torch.Tensor,
               timed: bool = False, loss_normalizer: float = 1.0) -> Dict[str, Any]:
    """One optimisation step: forward, LabelAnythingLoss, backward, gradient all-reduce, AdamW (run.py:425-590).
    `loss_normalizer` divides the loss before backward, as `Run._backward` does (run.py:359-361)."""
    opt.attach(lam)
    opt.zero_grad()
    result = train_forward(lam, batched_input)
    loss = loss_fn(result, gt)
    value = loss["value"] if isinstance(loss, dict) else loss
    (value if loss_normalizer == 1.0 else value / loss_normalizer).backward()
    opt.step(timed=timed)
    return {"loss": loss, **result}


class ConstantWithWarmup:
    """The reference's default schedule (`scheduler: constant_with_warmup`, parameters/trainval/coco/mael.yaml:34-37,
    transformers.get_constant_schedule_with_warmup): lr = base * min(1, step / num_warmup_steps).  `step()` sets `opt.lr`,
    which eager steps pass by value and `GraphedTrainStep` uploads before every replay."""

    def __init__(self, opt: FlatAdamW, num_warmup_steps: int) -> None:
        self.opt, self.base_lr, self.num_warmup_steps, self.last_step = opt, opt.lr, int(num_warmup_steps), 0
        self._set()

    def _set(self) -> None:
        w = self.num_warmup_steps
        self.opt.lr = self.base_lr * (min(1.0, self.last_step / max(1.0, w)) if w > 0 else 1.0)

    def step(self) -> None:
        self.last_step += 1
        self._set()

    def get_last_lr(self) -> List[float]:
        return [self.opt.lr]


def _loss_value(loss_fn, logits: torch.Tensor, gt: torch.Tensor) -> torch.Tensor:
    """`LabelAnythingLoss.logits_loss` (loss/__init__.py:67-92) without the `.item()` of its logging dictionary: the
    summed value only, as a device tensor (the component weight applied twice, like the reference)."""
    from .loss import FROM_CLASS_WEIGHTS

    wm = cw = None
    if loss_fn.class_weighting:
        cw, _ = ops.label_class_weights(gt.contiguous(), logits.shape[1])
        wm = FROM_CLASS_WEIGHTS
    total = None
    for k, comp in loss_fn.components.items():
        v = loss_fn.weights[k] * (loss_fn.weights[k] * comp(logits, gt, weight_matrix=wm, class_weights=cw))
        total = v if total is None else total + v
    return total


class GraphedTrainStep:
    """The whole optimisation step -- forward, loss, backward, gradient all-reduce, AdamW -- captured ONCE in a CUDA
    graph and replayed per batch.  The eager step issues ~960 native launches plus autograd bookkeeping and is bound by
    the host (27.6 ms against ~12 ms of GPU work for BASELINE configs[3]); a replay is one launch.

    What a graph fixes, and how each piece is handled:
      * geometry (shapes, prompt types present, original sizes): the `plan` of the example batch; `__call__` copies a new
        batch of the same geometry into the static input tensors;
      * which parameters receive a gradient: taken from eager warm-up steps, constant for a geometry;
      * the optimiser's step count and learning rate: bias corrections and `opt.lr` live in device memory
        (`la_adamw_f32_dev`) and are refreshed before every replay, so a scheduler that sets `opt.lr` keeps working;
      * RandomMatrixEncoder rows: drawn before every replay into a static tensor.
    The loss value and the outputs are static device tensors (no `.item()` inside the step)."""

    def __init__(self, lam: Lam, loss_fn, opt: FlatAdamW, example_input: Dict[str, Any], example_gt: torch.Tensor,
                 warmup: int = 3, collective_in_graph: bool = False) -> None:
        import torch.distributed as dist

        self.lam, self.loss_fn, self.opt = lam, loss_fn, opt
        opt.attach(lam)
        self.static = {k: (v.clone() if torch.is_tensor(v) else v) for k, v in example_input.items()}
        self.gt = example_gt.clone()
        self.plan = make_plan(lam, self.static)
        self.world = dist.get_world_size(opt.group) if (dist.is_available() and dist.is_initialized()) else 1
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):                     # eager warm-up steps on a side stream (torch.cuda.graph protocol)
            for _ in range(max(1, warmup)):
                opt.zero_grad()
                out = train_forward(lam, self.static, self.plan)
                _loss_value(loss_fn, out[ResultDict.LOGITS], self.gt).backward()
                used, _ = opt.reduce_gradients()
                opt.apply(used, self.world)
        torch.cuda.current_stream().wait_stream(side)
        self.used = used
        dev = opt.flat_p.device
        self.runs = opt.used_runs_for([s + 1 if u else s for s, u in zip(opt.steps, used)], used)
        n_runs = max(1, len(self.runs))
        self.bc = torch.ones((n_runs, 3), dtype=torch.float32, device=dev)
        # ring of pinned staging rows for the bias corrections: a slot is rewritten only after its last upload finished
        self._bc_host = [torch.ones((n_runs, 3), dtype=torch.float32).pin_memory() for _ in range(4)]
        self._bc_done = [torch.cuda.Event() for _ in range(4)]
        self._bc_slot = 0
        self._run_steps = [st for _, _, st in self.runs]
        torch.cuda.synchronize()
        # Inside the graph the gradients are taken with torch.autograd.grad w.r.t. fresh leaf ALIASES of the parameters
        # (same storage) and copied into the bucket.  The parameters themselves cannot be differentiated under capture
        # once an eager step has used them: their AccumulateGrad nodes stay bound to the stream of their first use (the
        # legacy stream), and autograd buffers incoming gradients on that stream ("operation would make the legacy stream
        # depend on a capturing blocking stream").
        views = [opt.flat_g[o:o + p.numel()].view(p.shape) for p, o, u in zip(opt.params, opt.offsets, used) if u]
        alias = {id(p): p.detach().requires_grad_(True) for p, u in zip(opt.params, used) if u}
        swapped = []
        for mod in lam.modules():
            for name, p in list(mod._parameters.items()):
                if p is not None and id(p) in alias:
                    mod._parameters[name] = alias[id(p)]
                    swapped.append((mod, name, p))
        T._DERIVED.clear()                    # no operand cached by the eager steps may stand in for a captured launch
        opt.flat_g.zero_()
        torch.cuda.synchronize()
        def update():
            for i, (lo, hi, _) in enumerate(self.runs):
                T.adamw_step_dev(opt.flat_p[lo:hi], opt.flat_g[lo:hi], opt.exp_avg[lo:hi], opt.exp_avg_sq[lo:hi],
                                 opt.betas[0], opt.betas[1], opt.eps, opt.weight_decay, self.bc[i], 1.0 / self.world)

        # world > 1: by default the all-reduce is launched BETWEEN two graphs (forward + backward | update) -- one more
        # launch per step; `collective_in_graph=True` makes it a node of a single graph (NCCL supports capture; verified
        # on 2 GPUs here)
        self.graph = torch.cuda.CUDAGraph()
        self.graph_update = None
        try:
            with torch.cuda.graph(self.graph):
                out = train_forward(lam, self.static, self.plan)
                self.loss = _loss_value(loss_fn, out[ResultDict.LOGITS], self.gt)
                grads = torch.autograd.grad(self.loss, [alias[id(p)] for p, u in zip(opt.params, used) if u])
                torch._foreach_copy_(views, list(grads))
                del grads
                if self.world > 1 and collective_in_graph:
                    dist.all_reduce(opt.flat_g[:opt.numel], group=opt.group)
                if self.world == 1 or collective_in_graph:
                    update()
            if self.world > 1 and not collective_in_graph:
                self.graph_update = torch.cuda.CUDAGraph()
                with torch.cuda.graph(self.graph_update, pool=self.graph.pool()):
                    update()
        finally:
            for mod, name, p in swapped:
                mod._parameters[name] = p
        self.loss = self.loss.detach()
        out = {k: (v.detach() if torch.is_tensor(v) else v) for k, v in out.items()}
        self.out = out

    def _fill_bc(self) -> None:
        b1, b2 = self.opt.betas
        slot = self._bc_slot
        self._bc_slot = (slot + 1) % len(self._bc_host)
        self._bc_done[slot].synchronize()
        host = self._bc_host[slot]
        for i, st in enumerate(self._run_steps):
            host[i, 0] = 1.0 - b1 ** st
            host[i, 1] = (1.0 - b2 ** st) ** 0.5
            host[i, 2] = self.opt.lr                      # read per replay: learning-rate schedules keep working
        self.bc.copy_(host, non_blocking=True)
        self._bc_done[slot].record()

    def __call__(self, batched_input: Optional[Dict[str, Any]] = None, gt: Optional[torch.Tensor] = None) -> Dict[str, Any]:
        """One step on `batched_input` / `gt` (same geometry as the example; None = the tensors already in place)."""
        if batched_input is not None:
            for k, v in batched_input.items():
                if torch.is_tensor(v):
                    dst = self.static[k]
                    if v.shape != dst.shape:
                        raise ValueError(f"GraphedTrainStep: '{k}' has shape {tuple(v.shape)}, the captured step "
                                         f"{tuple(dst.shape)}; capture a new step for a new geometry")
                    dst.copy_(v, non_blocking=True)
        if gt is not None:
            self.gt.copy_(gt, non_blocking=True)
        if self.plan["class_rows"] is not None:
            ce = self.lam.prompt_encoder.class_encoder
            self.plan["class_rows"].copy_(ce.sample_rows(self.plan["class_rows"].numel(), self.plan["class_rows"].device))
        self._fill_bc()
        self.graph.replay()
        if self.graph_update is not None:
            import torch.distributed as dist

            dist.all_reduce(self.opt.flat_g[:self.opt.numel], group=self.opt.group)   # the step's one collective
            self.graph_update.replay()
        opt = self.opt
        for i, u in enumerate(self.used):
            if u:
                opt.steps[i] += 1
        self._run_steps = [st + 1 for st in self._run_steps]
        torch.autograd.graph.increment_version([p for p, u in zip(opt.params, self.used) if u])
        return {"loss": {"value": self.loss}, **self.out}
