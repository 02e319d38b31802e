import sys, torch
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "tests"))
import test_kernels_gpu as T
from labelanything_b200 import ops

def run(mode, n_seq, L, gsz, qscale, heads):
    g = T._gen(L + n_seq)
    qkv = torch.randn(n_seq * L, 3 * heads * 64, device="cuda", generator=g)
    qkv[:, : heads * 64] *= qscale
    qkv = qkv.to(torch.bfloat16)
    rel_h = rel_w = bh = bw = None
    if gsz:
        rel_h = torch.randn(2 * gsz - 1, 64, device="cuda", generator=g) * 0.1
        rel_w = torch.randn(2 * gsz - 1, 64, device="cuda", generator=g) * 0.1
        qh = qkv[:, : heads * 64].reshape(n_seq * L, heads, 64).permute(1, 0, 2).contiguous()
        pad = 128 if gsz == 64 else 32
        bh, bw = T._rev_bias(ops, qh, rel_h, pad), T._rev_bias(ops, qh, rel_w, pad)
    ref = T._ref_attention(qkv, n_seq, L, heads, 0.125, rel_h, rel_w, gsz)
    # host emulation of the lazy-rescale decisions (64-key tiles, log2 units, threshold 8)
    q, k, v = qkv.float().view(n_seq, L, 3, heads, 64).permute(2, 0, 3, 1, 4)
    att = (q * 0.125) @ k.transpose(-1, -2)
    if rel_h is not None:
        idx = torch.arange(gsz, device=qkv.device)[:, None] - torch.arange(gsz, device=qkv.device)[None, :] + gsz - 1
        Rh, Rw = rel_h.to(torch.bfloat16).float()[idx], rel_w.to(torch.bfloat16).float()[idx]
        q5 = q.reshape(n_seq, heads, gsz, gsz, 64)
        b_h = torch.einsum("bnhwc,hkc->bnhwk", q5, Rh)
        b_w = torch.einsum("bnhwc,wkc->bnhwk", q5, Rw)
        att = (att.view(n_seq, heads, gsz, gsz, gsz, gsz) + b_h[..., :, None] + b_w[..., None, :]).view(n_seq, heads, L, L)
    tile = 64 if L != 196 else 112
    nt = (L + tile - 1) // tile
    att2 = att * 1.4426950408889634
    pad = nt * tile - L
    if pad:
        att2 = torch.nn.functional.pad(att2, (0, pad), value=float("-inf"))
    tmax = att2.view(n_seq, heads, L, nt, tile).amax(-1)            # [n_seq, heads, L, nt]
    m = tmax[..., 0].clone()
    resc = torch.zeros_like(m, dtype=torch.int32)
    for j in range(1, nt):
        need = tmax[..., j] > m + 8.0
        m = torch.where(need, tmax[..., j], m)
        resc += need.int()
    resc_rows = resc.permute(0, 2, 1).reshape(n_seq * L, heads)      # [rows, heads]
    print("rows needing >=1 rescale:", int((resc_rows > 0).sum()), "of", resc_rows.numel())
    outs = []
    for rep in range(3):
        out = torch.zeros(n_seq * L, heads * 64, device="cuda", dtype=torch.bfloat16)
        ops.attention(qkv, qkv, n_seq, L, heads, 0.125, out, 0, heads * 64, 2 * heads * 64, bh, bw, grid_hw=gsz)
        torch.cuda.synchronize()
        outs.append(out)
        err = (out.float() - ref).abs()
        bad = err > 2e-2 + 2e-2 * ref.abs()
        rows = bad.any(1).nonzero().flatten().tolist()
        badrh = bad.view(n_seq * L, heads, 64).any(-1)
        print("   bad (row,head) pairs:", int(badrh.sum()), " of which needed rescale:", int((badrh & (resc_rows > 0)).sum()),
              " warps with a rescale row:", int((resc_rows.view(-1, 32, heads) > 0).any(1).sum()))
        print(mode, n_seq, L, gsz, qscale, "rep", rep, "bad", int(bad.sum()), "max", err.max().item(), "rows", rows[:10],
              "cols", bad.any(0).nonzero().flatten().tolist()[:6], "...", flush=True)
        for r in rows[:3]:
            h = bad[r].nonzero().flatten()[0].item() // 64
            print("   row", r, "head", h, "out", out[r, h*64:h*64+4].float().tolist(), "ref", ref[r, h*64:h*64+4].tolist(),
                  "ratio", (out[r, h*64:h*64+8].float() / ref[r, h*64:h*64+8]).tolist())
    print("   deterministic:", all(torch.equal(outs[0], o) for o in outs))

run("global", 1, 4096, 64, 1.0, 4)
run("global", 2, 4096, 64, 4.0, 4)
run("plain", 1, 4096, 0, 4.0, 12)
