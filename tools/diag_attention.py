"""GPU diagnostic: fused attention (3 modes) and row kernels vs torch references. Writes gpurun_out/diag_attention.log"""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch
import torch.nn.functional as F

from labelanything_b200 import ops

out_dir = ROOT / "gpurun_out"
out_dir.mkdir(exist_ok=True)
log = open(out_dir / "diag_attention.log", "w")


def P(*a):
    s = " ".join(str(x) for x in a)
    print(s, flush=True)
    log.write(s + "\n")
    log.flush()


def report(name, y, r, tol):
    err = (y.float() - r.float()).abs()
    bad = err > tol * (1 + r.float().abs())
    P(f"{name}: max_abs_err={err.max().item():.3e} mean={err.mean().item():.3e} bad={int(bad.sum())}/{bad.numel()} nan={int(torch.isnan(y.float()).sum())}")
    if bad.any():
        idx = bad.nonzero()
        P("   first bad", idx[:6].tolist())
        rows = torch.unique(idx[:, 0])
        P(f"   bad rows n={len(rows)} min={rows.min().item()} max={rows.max().item()}")
    return not bad.any()


def rev_table_bias(q_heads, rel, pad_to):
    # q_heads [heads, rows, 64] bf16 ; rel [2g-1, 64] -> fp32 [heads, rows, pad_to] = q @ reversed(rel)^T
    trev = torch.flip(rel, dims=[0]).to(torch.bfloat16)
    w = torch.zeros(pad_to, 64, device=rel.device, dtype=torch.bfloat16)
    w[: trev.shape[0]] = trev
    outs = []
    for h in range(q_heads.shape[0]):
        outs.append(ops.gemm(q_heads[h], w, None, out_dtype=torch.float32))
    return torch.stack(outs).permute(1, 0, 2).contiguous()  # [rows, heads, pad]


def rel_operand(rel_h, rel_w, pad):
    w = torch.zeros(2 * pad, 64, device=rel_h.device, dtype=torch.bfloat16)
    w[: rel_h.shape[0]] = torch.flip(rel_h, dims=[0]).to(torch.bfloat16)
    w[pad: pad + rel_w.shape[0]] = torch.flip(rel_w, dims=[0]).to(torch.bfloat16)
    return w


def ref_attention(qkv, n_seq, L, heads, scale, rel_h=None, rel_w=None, g=0):
    q, k, v = qkv.float().view(n_seq, L, 3, heads, 64).permute(2, 0, 3, 1, 4)
    att = (q * scale) @ k.transpose(-1, -2)
    if rel_h is not None:
        idx = torch.arange(g, device=qkv.device)[:, None] - torch.arange(g, device=qkv.device)[None, :] + g - 1
        Rh = rel_h.to(torch.bfloat16).float()[idx]
        Rw = rel_w.to(torch.bfloat16).float()[idx]
        q5 = q.reshape(n_seq, heads, g, g, 64)
        bh = torch.einsum("bnhwc,hkc->bnhwk", q5, Rh)
        bw = torch.einsum("bnhwc,wkc->bnhwk", q5, Rw)
        att = (att.view(n_seq, heads, g, g, g, g) + bh[..., :, None] + bw[..., None, :]).view(n_seq, heads, L, L)
    att = att.softmax(-1)
    return (att @ v).transpose(1, 2).reshape(n_seq * L, heads * 64)


def test_plain(n_seq=2, L=901, heads=12):
    g = torch.Generator(device="cuda").manual_seed(1)
    qkv = torch.randn(n_seq * L, 3 * heads * 64, device="cuda", generator=g).to(torch.bfloat16)
    out = torch.zeros(n_seq * L, heads * 64, device="cuda", dtype=torch.bfloat16)
    ops.attention(qkv, qkv, n_seq, L, heads, 0.125, out, 0, heads * 64, 2 * heads * 64)
    torch.cuda.synchronize()
    return report(f"attention plain L={L}", out, ref_attention(qkv, n_seq, L, heads, 0.125), 2e-2)


def test_global(n_seq=1, heads=12):
    L, gsz = 4096, 64
    g = torch.Generator(device="cuda").manual_seed(2)
    qkv = torch.randn(n_seq * L, 3 * heads * 64, device="cuda", generator=g).to(torch.bfloat16)
    rel_h = torch.randn(127, 64, device="cuda", generator=g) * 0.1
    rel_w = torch.randn(127, 64, device="cuda", generator=g) * 0.1
    qh = qkv[:, : heads * 64].reshape(n_seq * L, heads, 64).permute(1, 0, 2).contiguous()
    bh = rev_table_bias(qh, rel_h, 128)
    bw = rev_table_bias(qh, rel_w, 128)
    out = torch.zeros(n_seq * L, heads * 64, device="cuda", dtype=torch.bfloat16)
    ops.attention(qkv, qkv, n_seq, L, heads, 0.125, out, 0, heads * 64, 2 * heads * 64, bh, bw, grid_hw=64)
    torch.cuda.synchronize()
    ref = ref_attention(qkv, n_seq, L, heads, 0.125, rel_h, rel_w, gsz)
    return report("attention global64", out, ref, 2e-2)


def test_window(n_img=2, heads=12):
    L, gsz, nwin, hw = 196, 14, 5, 64
    n_seq = n_img * nwin * nwin
    g = torch.Generator(device="cuda").manual_seed(3)
    qkv = torch.randn(n_seq * L, 3 * heads * 64, device="cuda", generator=g).to(torch.bfloat16)
    rel_h = torch.randn(27, 64, device="cuda", generator=g) * 0.1
    rel_w = torch.randn(27, 64, device="cuda", generator=g) * 0.1
    op = rel_operand(rel_h, rel_w, 32)
    ok = True
    out = torch.zeros(n_seq * L, heads * 64, device="cuda", dtype=torch.bfloat16)
    ops.attention_window(qkv, qkv, n_seq, heads, 0.125, out, 0, heads * 64, 2 * heads * 64, op, 32)
    torch.cuda.synchronize()
    ref = ref_attention(qkv, n_seq, L, heads, 0.125, rel_h, rel_w, gsz)
    ok &= report("attention window14 (identity rows)", out, ref, 2e-2)
    out2 = torch.zeros(n_img * hw * hw, heads * 64, device="cuda", dtype=torch.bfloat16)
    ops.attention_window(qkv, qkv, n_seq, heads, 0.125, out2, 0, heads * 64, 2 * heads * 64, op, 32, out_mode=1,
                         nwin=nwin, img_hw=hw)
    torch.cuda.synchronize()
    r = ref.view(n_img, nwin, nwin, 14, 14, -1).permute(0, 1, 3, 2, 4, 5).reshape(n_img, 70, 70, -1)[:, :hw, :hw]
    ok &= report("attention window14 (unpartition)", out2, r.reshape(n_img * hw * hw, -1), 2e-2)
    return ok


def test_rowops():
    ok = True
    g = torch.Generator(device="cuda").manual_seed(4)
    I, hw, d = 2, 64, 768
    rows = I * hw * hw
    x = torch.randn(rows, d, device="cuda", generator=g)
    delta = torch.randn(rows, d, device="cuda", generator=g).to(torch.bfloat16)
    gamma = torch.randn(d, device="cuda", generator=g)
    beta = torch.randn(d, device="cuda", generator=g)
    # identity map with in-place residual
    x1 = x.clone()
    y = torch.empty(rows, d, device="cuda", dtype=torch.bfloat16)
    ops.add_layernorm(x1, delta, gamma, beta, 1e-6, rows=rows, d=d, x_out=x1, y_out=y)
    xr = x + delta.float()
    ok &= report("add_ln x_out", x1, xr, 1e-6)
    ok &= report("add_ln y", y, F.layer_norm(xr, (d,), gamma, beta, 1e-6), 1e-2)
    # window partition
    nwin, win = 5, 14
    orow = I * nwin * nwin * win * win
    y2 = torch.full((orow, d), 7.0, device="cuda", dtype=torch.bfloat16)
    x2 = x.clone()
    ops.add_layernorm(x2, delta, gamma, beta, 1e-6, rows=orow, d=d, x_out=x2, y_out=y2, map_mode=1, win=win,
                      nwin=nwin, hw=hw)
    ln = F.layer_norm(xr, (d,), gamma, beta, 1e-6).view(I, hw, hw, d)
    ln = F.pad(ln, (0, 0, 0, 6, 0, 6)).view(I, nwin, win, nwin, win, d).permute(0, 1, 3, 2, 4, 5).reshape(orow, d)
    ok &= report("add_ln window-partition y", y2, ln, 1e-2)
    ok &= report("add_ln window-partition x_out", x2, xr, 1e-6)
    # drop cls, fp32 out, no delta
    L = 901
    xs = torch.randn(3 * L, d, device="cuda", generator=g)
    y3 = torch.empty(3 * (L - 1), d, device="cuda")
    ops.add_layernorm(xs, None, gamma, beta, 1e-12, rows=3 * L, d=d, y_out=y3, map_mode=2, seq_len=L)
    ok &= report("add_ln drop-cls fp32", y3, F.layer_norm(xs, (d,), gamma, beta, 1e-12).view(3, L, d)[:, 1:].reshape(-1, d), 1e-5)
    # cast only with pos broadcast
    pos = torch.randn(hw * hw, d, device="cuda", generator=g)
    x4 = torch.empty(rows, d, device="cuda")
    ops.add_layernorm(pos, delta, None, None, 0.0, rows=rows, d=d, x_out=x4, x_mod=hw * hw)
    ok &= report("add_ln pos broadcast", x4, pos.repeat(I, 1) + delta.float(), 1e-6)
    # embed tokens
    patch = torch.randn(I * 900, d, device="cuda", generator=g).to(torch.bfloat16)
    cls = torch.randn(d, device="cuda", generator=g)
    pos2 = torch.randn(901, d, device="cuda", generator=g)
    x5 = torch.empty(I * 901, d, device="cuda")
    ops.embed_tokens(patch, cls, pos2, x5, I, 901, 1, d)
    r5 = torch.cat([cls.view(1, 1, d).expand(I, 1, d), patch.float().view(I, 900, d)], 1) + pos2
    ok &= report("embed_tokens", x5, r5.reshape(-1, d), 1e-6)
    # im2col patch
    img = torch.randn(2, 3, 480, 480, device="cuda", generator=g)
    pm = ops.im2col_patch16(img)
    rp = F.unfold(img, 16, stride=16).transpose(1, 2).reshape(-1, 768)
    ok &= report("im2col_patch16", pm, rp.to(torch.bfloat16), 1e-6)
    # im2col 3x3
    f = torch.randn(2 * 30 * 30, 64, device="cuda", generator=g).to(torch.bfloat16)
    c3 = ops.im2col_3x3(f, 2, 30, 30, 64)
    fr = f.float().view(2, 30, 30, 64).permute(0, 3, 1, 2)
    r3 = F.unfold(fr, 3, padding=1).view(2, 64, 9, 900).permute(0, 3, 2, 1).reshape(-1, 9 * 64)
    ok &= report("im2col_3x3", c3, r3, 1e-6)
    return ok


def bench_attention():
    heads = 12
    for name, n_seq, L, gsz in [("global64", 8, 4096, 64), ("window14", 8 * 25, 196, 14), ("plain901", 32, 901, 0)]:
        qkv = torch.randn(n_seq * L, 3 * heads * 64, device="cuda").to(torch.bfloat16)
        out = torch.zeros(n_seq * L, heads * 64, device="cuda", dtype=torch.bfloat16)
        bh = bw = None
        if gsz == 64:
            bh = torch.randn(n_seq * L, heads, 128, device="cuda") * 0.1
            bw = torch.randn(n_seq * L, heads, 128, device="cuda") * 0.1
        if gsz == 14:
            op = rel_operand(torch.randn(27, 64, device="cuda") * 0.1, torch.randn(27, 64, device="cuda") * 0.1, 32)
            f = lambda: ops.attention_window(qkv, qkv, n_seq, heads, 0.125, out, 0, heads * 64, 2 * heads * 64, op, 32)
        else:
            f = lambda: ops.attention(qkv, qkv, n_seq, L, heads, 0.125, out, 0, heads * 64, 2 * heads * 64, bh, bw, grid_hw=gsz)
        for _ in range(3):
            f()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        for _ in range(10):
            f()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        fl = 4.0 * n_seq * heads * L * L * 64
        q, k, v = qkv.view(n_seq, L, 3, heads, 64).permute(2, 0, 3, 1, 4)
        for _ in range(3):
            F.scaled_dot_product_attention(q, k, v)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(10):
            F.scaled_dot_product_attention(q, k, v)
        e1.record()
        torch.cuda.synchronize()
        ms_t = e0.elapsed_time(e1) / 10
        P(f"bench attention {name}: ours {ms:.3f} ms = {fl/ms/1e9:.1f} TFLOP/s ; torch sdpa (no bias) {ms_t:.3f} ms = {fl/ms_t/1e9:.1f} TFLOP/s")


if __name__ == "__main__":
    P(torch.cuda.get_device_name(0))
    ok = True
    for fn in (test_rowops, test_plain, test_window, test_global):
        try:
            ok &= fn()
        except Exception as e:  # keep going: we want as much signal per GPU call as possible
            P(f"{fn.__name__} EXCEPTION: {e}")
            ok = False
    P("ALL_OK" if ok else "SOME_FAILED")
    try:
        bench_attention()
    except Exception as e:
        P("bench EXCEPTION", e)
    sys.path.insert(0, str(ROOT / "tools"))
    import diag_gemm
    diag_gemm.bench(32768, 3072, 768, act=1)
    diag_gemm.bench(32768, 3072, 768, act=0)
