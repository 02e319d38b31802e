"""Debug: which rows fail in the fused attention, with / without bias and with forced O rescales."""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tools"))
import torch

from labelanything_b200 import ops
import diag_attention as D


def run(name, L, heads, qscale, bias_scale, gsz):
    g = torch.Generator(device="cuda").manual_seed(2)
    qkv = torch.randn(L, 3 * heads * 64, device="cuda", generator=g)
    qkv[:, : heads * 64] *= qscale
    qkv = qkv.to(torch.bfloat16)
    rel_h = rel_w = bh = bw = None
    if gsz:
        rel_h = torch.randn(2 * gsz - 1, 64, device="cuda", generator=g) * bias_scale
        rel_w = torch.randn(2 * gsz - 1, 64, device="cuda", generator=g) * bias_scale
        qh = qkv[:, : heads * 64].reshape(L, heads, 64).permute(1, 0, 2).contiguous()
        bh = D.rev_table_bias(qh, rel_h, 128)
        bw = D.rev_table_bias(qh, rel_w, 128)
    out = torch.zeros(L, heads * 64, device="cuda", dtype=torch.bfloat16)
    ops.attention(qkv, qkv, 1, L, heads, 0.125, out, 0, heads * 64, 2 * heads * 64, bh, bw, grid_hw=gsz)
    torch.cuda.synchronize()
    ref = D.ref_attention(qkv, 1, L, heads, 0.125, rel_h, rel_w, gsz)
    err = (out.float() - ref).abs().view(L, heads, 64)
    bad = (~(err < 2e-2 * (1 + ref.abs().view(L, heads, 64)))).any(-1)   # [L, heads], NaN counts as bad
    nan = torch.isnan(out.float()).view(L, heads, 64).any(-1)
    print(f"{name}: bad row-heads {int(bad.sum())} / {bad.numel()}  nan row-heads {int(nan.sum())}")
    if bad.any():
        idx = bad.nonzero()
        rows = idx[:, 0]
        print("   by 128-row tile:", torch.bincount(rows // 128, minlength=L // 128).tolist())
        print("   by warp quarter:", torch.bincount((rows % 128) // 32, minlength=4).tolist())
        print("   by head:", torch.bincount(idx[:, 1], minlength=heads).tolist())
        r0, h0 = idx[0].tolist()
        print("   first bad", r0, h0, "out", out[r0, h0 * 64:h0 * 64 + 4].tolist(), "ref", ref[r0, h0 * 64:h0 * 64 + 4].tolist())


run("plain 4096 qscale 1", 4096, 4, 1.0, 0, 0)
run("plain 4096 qscale 4 (forced rescales)", 4096, 4, 4.0, 0, 0)
run("global bias 0", 4096, 4, 1.0, 0.0, 64)
run("global bias 0.1", 4096, 4, 1.0, 0.1, 64)
run("global bias 0.1 qscale 4", 4096, 4, 4.0, 0.1, 64)
