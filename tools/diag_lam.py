"""Stage-by-stage error of the native Lam (embeddings path) against the CPU oracle (diagnostic, GPU box)."""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "oracle"))
import lam_oracle as O
from labelanything_b200 import ops
from labelanything_b200.build_lam import build_lam_no_vit
from labelanything_b200.synthetic import load_synth_weights
from labelanything_b200.vit_engine import pack_neck, run_neck


def rep(name, a, b):
    a, b = a.float().cpu(), b.float().cpu()
    fin = torch.isfinite(b)
    e = (a[fin] - b[fin]).abs()
    std = b[fin].std().item()
    print(f"{name:40s} max={e.max().item():.5f} mean={e.mean().item():.6f} std={std:.4f} max/std={e.max().item()/std:.4f} mean/std={e.mean().item()/std:.5f}")


B, M, C, S, D, Ce = 2, 3, 3, 256, 256, 384
g = torch.Generator().manual_seed(7)
h = S // 16
ep = {"embeddings": torch.randn(B, M + 1, Ce, h, h, generator=g),
      "flag_examples": torch.ones(B, M, C, dtype=torch.uint8)}
ep["prompt_masks"] = (torch.rand(B, M, C, 256, 256, generator=g) > 0.5).float()
ep["flag_masks"] = torch.ones(B, M, C, dtype=torch.uint8)
ep["dims"] = torch.full((B, M + 1, 2), S, dtype=torch.int64)
lam = build_lam_no_vit(image_embed_dim=Ce, embed_dim=D, image_size=S, spatial_convs=3, custom_preprocess=False)
load_synth_weights(lam, seed=3)
sd = {k: v.clone() for k, v in lam.state_dict().items()}
cfg = {"image_size": S, "image_embedding_size": (h, h), "has_neck": True, "spatial_convs": 3, "class_attention": False,
       "example_attention": False, "example_class_attention": True, "custom_preprocess": False}
with torch.no_grad():
    ref = O.lam_forward(sd, cfg, dict(ep), return_intermediates=True)
    lam = lam.cuda()
    epc = {k: v.cuda() for k, v in ep.items()}
    feats, _, N, gg = lam._features(epc)
    rep("neck features", ops.tokens_to_nchw(feats, B * N, gg, gg).view(B, N, D, gg, gg), ref["features"])
    seg, pe = lam._forward(epc)
    rep("class_examples_embeddings", pe["class_examples_embeddings"], ref["class_examples_embeddings"])
    rep("class_embeddings", pe["class_embeddings"], ref["class_embs"])
    rep("low_res_logits", seg, ref["low_res_logits"])
    # decoder alone on the ORACLE's features and class embeddings
    f_or = ref["features"][:, 0].contiguous().cuda()
    q32, q16 = ops.nchw_to_tokens(f_or, want_f32=True, want_bf16=True)
    seg2 = lam.mask_decoder.decode(q32, q16, lam.prompt_encoder.dense_pe_tokens(), ref["class_embs"].cuda(), B, gg, gg)
    rep("decoder on oracle inputs", seg2, ref["low_res_logits"])
    # prompt encoder alone on the oracle's features
    feat_or, _ = ops.nchw_to_tokens(ref["features"].flatten(0, 1).contiguous().cuda())
    pts, bxs, msk, fe = lam.prepare_prompts(epc)
    pe2 = lam.prompt_encoder.encode(feat_or, B, M, pts, bxs, msk, fe, feat_lead=1)
    rep("prompt encoder on oracle features", pe2["class_examples_embeddings"], ref["class_examples_embeddings"])
    # pieces of the decoder
    dec_sd = sd
    qe = ref["features"][:, 0]
    gauss = sd["prompt_encoder.pe_layer.positional_encoding_gaussian_matrix"]
    cls_o, keys_o = O.two_way_transformer(sd, "mask_decoder.transformer", qe, O.dense_pe(gauss, h, h), ref["class_embs"])
    from labelanything_b200.transformer import run_two_way
    q_n, k_n, _ = run_two_way(lam.mask_decoder.transformer, q16, q32, lam.prompt_encoder.dense_pe_tokens(),
                              ref["class_embs"].cuda().view(B * C, D).contiguous(), B, h * h, C, want_queries=True,
                              want_keys_f32=True)
    rep("decoder transformer queries", q_n.view(B, C, D), cls_o)
    rep("decoder transformer keys", k_n.view(B, h * h, D), keys_o)
