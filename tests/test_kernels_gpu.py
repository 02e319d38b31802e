"""Per-kernel GPU parity, through the C ABI, against plain torch fp32 (or the CPU oracle's functions) on the same
inputs.  Tolerances (stated per test): kernels whose arithmetic is fp32 end to end (row kernels, resizes, sparse
embedding, postprocess) must agree to ~1e-5; kernels with bf16 operands / outputs are compared on IDENTICAL bf16
inputs with a bf16-output bound of 2^-8 relative (+ small absolute term); index / layout kernels are bit-exact."""
import math
import sys
from pathlib import Path

import pytest
import torch
import torch.nn.functional as F

ROOT = Path(__file__).resolve().parent.parent
pytestmark = pytest.mark.gpu


def _ops():
    from labelanything_b200 import ops

    return ops


def _oracle():
    sys.path.insert(0, str(ROOT / "oracle"))
    import lam_oracle

    return lam_oracle


def _close(y, r, rtol, atol, what=""):
    y, r = y.float(), r.float()
    err = (y - r).abs()
    bad = err > atol + rtol * r.abs()
    assert not bool(torch.isnan(y).any()), f"{what}: NaN in output"
    assert not bool(bad.any()), f"{what}: {int(bad.sum())}/{bad.numel()} outside tolerance, max err {err.max().item():.3e}"


def _gen(seed):
    return torch.Generator(device="cuda").manual_seed(seed)


# ---------------------------------------------------------------------------------------------- GEMM
@pytest.mark.parametrize("M,N,K,act,out_dtype,bias", [
    (4096, 3072, 768, 1, torch.bfloat16, True),    # ViT MLP lin1 + exact GELU
    (4096, 768, 3072, 0, torch.bfloat16, True),    # ViT MLP lin2
    (4900, 2304, 768, 0, torch.bfloat16, True),    # qkv on window-padded rows (M not a tile multiple)
    (1000, 128, 512, 2, torch.bfloat16, True),     # ReLU epilogue, N = one 128 tile
    (1000, 64, 576, 0, torch.bfloat16, True),      # spatial conv as GEMM (K = 9 * 64)
    (300, 256, 136, 0, torch.float32, True),       # ragged K (TMA zero fill), fp32 out
    (3000, 384, 128, 0, torch.bfloat16, True),     # CTA-pair path with an M tail and a half-filled N tile
    (12, 512, 512, 0, torch.float32, False),       # a handful of token rows
    (2048, 64, 64, 0, torch.float32, False),       # rel-pos table GEMM
    (4096, 256, 64, 0, torch.float16, False),      # rel-pos table GEMM of the global blocks, fp16 out (CTA pair)
    (1000, 128, 64, 0, torch.float16, False),      # fp16 out, single-CTA tiles
])
def test_gemm_matches_torch(M, N, K, act, out_dtype, bias):
    ops = _ops()
    g = _gen(M + N + K)
    a = torch.randn(M, K, device="cuda", generator=g).to(torch.bfloat16)
    w = (torch.randn(N, K, device="cuda", generator=g) / math.sqrt(K)).to(torch.bfloat16)
    b = torch.randn(N, device="cuda", generator=g) if bias else None
    y = ops.gemm(a, w, b, act=act, out_dtype=out_dtype)
    r = a.float() @ w.float().t()
    if b is not None:
        r = r + b
    r = F.gelu(r) if act == 1 else (F.relu(r) if act == 2 else r)
    assert y.dtype == out_dtype and y.shape == (M, N)
    if out_dtype == torch.bfloat16:
        _close(y, r, 2 ** -8, 2e-3, "gemm bf16")      # one bf16 rounding of the result
    elif out_dtype == torch.float16:
        _close(y, r, 2 ** -11, 1e-4, "gemm fp16")     # one fp16 rounding of the result
    else:
        _close(y, r, 1e-4, 1e-4, "gemm fp32")         # fp32 accumulation order only


def test_gemm_accumulate_in_place():
    """x += a @ w^T + b with the add done in fp32 on the accumulator (lin2 of the ViT blocks)."""
    ops = _ops()
    g = _gen(21)
    for M, N, K in ((4900, 768, 3072), (4096, 768, 768), (2500, 384, 128)):
        a = torch.randn(M, K, device="cuda", generator=g).to(torch.bfloat16)
        w = (torch.randn(N, K, device="cuda", generator=g) / math.sqrt(K)).to(torch.bfloat16)
        b = torch.randn(N, device="cuda", generator=g)
        x = torch.randn(M, N, device="cuda", generator=g)
        r = x + a.float() @ w.float().t() + b
        y = ops.gemm_accumulate(a, w, b, x)
        assert y.data_ptr() == x.data_ptr()
        _close(x, r, 1e-4, 1e-4, f"gemm accumulate {M}x{N}x{K}")


def test_gemm_strided_operand_and_output_views():
    ops = _ops()
    g = _gen(5)
    big = torch.randn(512, 1536, device="cuda", generator=g).to(torch.bfloat16)
    a = big[:, 512:1024]                               # column slice of a packed projection buffer
    w = (torch.randn(256, 512, device="cuda", generator=g) / 22).to(torch.bfloat16)
    y = ops.gemm(a, w, None, out_dtype=torch.float32)
    _close(y, a.float() @ w.float().t(), 1e-4, 1e-4, "gemm strided")


def test_conv3x3_implicit_gemm_matches_torch():
    """3x3 conv of the necks as an implicit GEMM (4-D TMA boxes at shifted coordinates, zero-filled borders)."""
    ops = _ops()
    g = _gen(11)
    n_img, H, W, C, N = 3, 64, 64, 128, 256
    x = torch.randn(n_img, H, W, C, device="cuda", generator=g).to(torch.bfloat16)
    w = (torch.randn(N, C, 3, 3, device="cuda", generator=g) / math.sqrt(9 * C)).to(torch.bfloat16)
    b = torch.randn(N, device="cuda", generator=g)
    wp = w.permute(0, 2, 3, 1).reshape(N, 9 * C).contiguous()          # column = (ky*3 + kx)*C + ci
    y = ops.conv3x3(x.view(-1, C), wp, b, n_img, H, W, C, out_dtype=torch.float32)
    r = F.conv2d(x.float().permute(0, 3, 1, 2), w.float(), b, padding=1).permute(0, 2, 3, 1).reshape(-1, N)
    _close(y, r, 1e-4, 1e-4, "conv3x3 fp32")
    y16 = ops.conv3x3(x.view(-1, C), wp, None, n_img, H, W, C, act=ops.ACT_RELU)
    r16 = F.relu(F.conv2d(x.float().permute(0, 3, 1, 2), w.float(), None, padding=1)).permute(0, 2, 3, 1).reshape(-1, N)
    _close(y16, r16, 2 ** -8, 2e-3, "conv3x3 bf16 relu")
    # same result as the explicit im2col + GEMM path
    col = ops.im2col_3x3(x.view(-1, C).contiguous(), n_img, H, W, C)
    y2 = ops.gemm(col, wp, b, out_dtype=torch.float32)
    _close(y, y2, 1e-5, 1e-5, "conv3x3 vs im2col + gemm")


# ---------------------------------------------------------------------------------------------- fused attention
def _rev_bias(ops, q_heads, rel, pad_to, dtype=torch.float32):
    trev = torch.flip(rel, dims=[0]).to(torch.bfloat16)
    w = torch.zeros(pad_to, 64, device=rel.device, dtype=torch.bfloat16)
    w[: trev.shape[0]] = trev
    return torch.stack([ops.gemm(q_heads[h], w, None, out_dtype=dtype) for h in range(q_heads.shape[0])]
                       ).permute(1, 0, 2).contiguous()


def _rel_operand(rel_h, rel_w, pad):
    """bf16 [2 * pad, 64]: reversed rel_pos_h rows (zero padded to `pad`) then reversed rel_pos_w rows."""
    w = torch.zeros(2 * pad, 64, device=rel_h.device, dtype=torch.bfloat16)
    w[: rel_h.shape[0]] = torch.flip(rel_h, dims=[0]).to(torch.bfloat16)
    w[pad: pad + rel_w.shape[0]] = torch.flip(rel_w, dims=[0]).to(torch.bfloat16)
    return w


def _ref_attention(qkv, n_seq, L, heads, scale, rel_h=None, rel_w=None, g=0):
    q, k, v = qkv.float().view(n_seq, L, 3, heads, 64).permute(2, 0, 3, 1, 4)
    att = (q * scale) @ k.transpose(-1, -2)
    if rel_h is not None:
        idx = torch.arange(g, device=qkv.device)[:, None] - torch.arange(g, device=qkv.device)[None, :] + g - 1
        Rh, Rw = rel_h.to(torch.bfloat16).float()[idx], rel_w.to(torch.bfloat16).float()[idx]
        q5 = q.reshape(n_seq, heads, g, g, 64)
        bh = torch.einsum("bnhwc,hkc->bnhwk", q5, Rh)
        bw = torch.einsum("bnhwc,wkc->bnhwk", q5, Rw)
        att = (att.view(n_seq, heads, g, g, g, g) + bh[..., :, None] + bw[..., None, :]).view(n_seq, heads, L, L)
    return (att.softmax(-1) @ v).transpose(1, 2).reshape(n_seq * L, heads * 64)


@pytest.mark.parametrize("mode,n_seq,L,gsz,qscale", [
    ("plain", 3, 901, 0, 1.0),         # HF ViT: CLS + 30x30 tokens, ragged last key tile
    ("plain", 2, 197, 0, 1.0),         # ViT-B/224
    ("plain", 1, 4096, 0, 4.0),        # large logits: forces the lazy O rescale path
    ("global", 1, 4096, 64, 1.0),      # SAM global block with decomposed rel-pos bias
    ("global", 2, 4096, 64, 4.0),      # + rescales, two sequences (persistent loop over items)
    ("global16", 1, 4096, 64, 1.0),    # fp16 rel-pos tables (what the encoder engine feeds the global blocks)
    ("global16", 2, 4096, 64, 4.0),
    ("window", 50, 196, 14, 1.0),      # SAM 14x14 windows (2 images x 25 windows)
])
def test_fused_attention_matches_torch(mode, n_seq, L, gsz, qscale):
    ops = _ops()
    heads = 12 if not mode.startswith("global") else 4
    g = _gen(L + n_seq)
    qkv = torch.randn(n_seq * L, 3 * heads * 64, device="cuda", generator=g)
    qkv[:, : heads * 64] *= qscale
    # V = 1 + N(0, 1/4): every output element is O(1) in every parametrization (a softmax average of N(0, 1) values
    # over thousands of keys would be ~0.03, and an absolute tolerance would then hide any error)
    qkv[:, 2 * heads * 64:] = 1.0 + 0.5 * qkv[:, 2 * heads * 64:]
    qkv = qkv.to(torch.bfloat16)
    rel_h = rel_w = bh = bw = None
    if gsz:
        rel_h = torch.randn(2 * gsz - 1, 64, device="cuda", generator=g) * 0.1
        rel_w = torch.randn(2 * gsz - 1, 64, device="cuda", generator=g) * 0.1
    out = torch.zeros(n_seq * L, heads * 64, device="cuda", dtype=torch.bfloat16)
    if gsz == 14:
        # windows: the table products are formed inside the kernel from the reversed rel-pos operand
        ops.attention_window(qkv, qkv, n_seq, heads, 0.125, out, 0, heads * 64, 2 * heads * 64,
                             _rel_operand(rel_h, rel_w, 32), 32)
    else:
        if gsz:
            qh = qkv[:, : heads * 64].reshape(n_seq * L, heads, 64).permute(1, 0, 2).contiguous()
            td = torch.float16 if mode == "global16" else torch.float32
            bh, bw = _rev_bias(ops, qh, rel_h, 128, td), _rev_bias(ops, qh, rel_w, 128, td)
        ops.attention(qkv, qkv, n_seq, L, heads, 0.125, out, 0, heads * 64, 2 * heads * 64, bh, bw, grid_hw=gsz)
    ref = _ref_attention(qkv, n_seq, L, heads, 0.125, rel_h, rel_w, gsz)
    assert float(ref.abs().mean()) > 0.5      # the outputs are O(1): the absolute term below is 0.2 % of the signal
    # kernel tolerance on identical bf16 inputs vs torch fp32.  Two bf16 roundings sit on the path, each with a
    # worst-case relative error of 2^-8 (half an ulp at the bottom of a binade): P before the PV product and the
    # output.  With ordinary logits (qscale 1) hundreds of keys carry weight and the P errors average out: bound =
    # 2^-8 (output rounding) + 2e-3 absolute.  With 4x logits (qscale 4: softmax dominated by one or two keys, the
    # lazy-rescale path) the P error of the dominant key reaches the output unaveraged: bound = 2^-7 + 2e-3.
    # The rel-pos modes add the fp16 rounding of the table products (|bias| up to ~3 -> 1.5e-3 absolute in the logit,
    # i.e. 1.5e-3 relative in P, twice: rel_h and rel_w): + 2^-8 on the sharp cases.
    rtol = 2 ** -8 if qscale == 1.0 else (2 ** -7 + (2 ** -8 if gsz else 0.0))
    if qscale == 1.0:
        _close(out, ref, rtol, 2e-3, f"attention {mode}")
    else:
        # sharp softmax: the bound above is a sum of worst cases that a handful of the 2 M outputs do reach (measured:
        # 2-5 elements up to 1.5e-2); assert it on all but 1e-5 of the elements and 1.25x of it on every element
        e = (out.float() - ref.float()).abs()
        bad = e > 2e-3 + rtol * ref.float().abs()
        assert int(bad.sum()) <= 1e-5 * bad.numel(), f"attention {mode}: {int(bad.sum())} elements beyond the bound"
        _close(out, ref, 1.25 * rtol, 2.5e-3, f"attention {mode}")
    err = (out.float() - ref).abs()
    print(f"attention {mode} L={L} qscale={qscale}: max_abs_err {err.max().item():.3e} mean {err.mean().item():.3e} "
          f"(|ref| mean {ref.abs().mean().item():.3f})")


def test_window_attention_unpartition_drops_padding():
    ops = _ops()
    n_img, nwin, hw, L, heads = 2, 5, 64, 196, 12
    n_seq = n_img * nwin * nwin
    g = _gen(3)
    qkv = torch.randn(n_seq * L, 3 * heads * 64, device="cuda", generator=g).to(torch.bfloat16)
    rel = torch.randn(27, 64, device="cuda", generator=g) * 0.1
    op = _rel_operand(rel, rel, 32)
    flat = torch.zeros(n_seq * L, heads * 64, device="cuda", dtype=torch.bfloat16)
    ops.attention_window(qkv, qkv, n_seq, heads, 0.125, flat, 0, heads * 64, 2 * heads * 64, op, 32)
    out = torch.zeros(n_img * hw * hw, heads * 64, device="cuda", dtype=torch.bfloat16)
    ops.attention_window(qkv, qkv, n_seq, heads, 0.125, out, 0, heads * 64, 2 * heads * 64, op, 32, out_mode=1,
                         nwin=nwin, img_hw=hw)
    r = flat.view(n_img, nwin, nwin, 14, 14, -1).permute(0, 1, 3, 2, 4, 5).reshape(n_img, 70, 70, -1)[:, :hw, :hw]
    assert torch.equal(out, r.reshape(n_img * hw * hw, -1))      # same arithmetic, only the row mapping differs


def test_window_attention_from_padded_grid_equals_partitioned_rows():
    """Windowed block without the 70 x 70 projection: la_gemm_bf16_to_grid stores the 64 x 64 projected tokens into a
    padded grid whose padding positions hold the bias row, the window kernel fetches each window as one 4-D box.
    Same arithmetic as LayerNorm-with-window-partition (zero rows) -> GEMM -> partitioned attention: identical bits."""
    ops = _ops()
    n_img, nwin, hw, heads, d = 2, 5, 64, 12, 768
    P = nwin * 14
    n_seq = n_img * nwin * nwin
    g = _gen(33)
    y = torch.randn(n_img * hw * hw, d, device="cuda", generator=g).to(torch.bfloat16)      # LayerNorm output, image order
    wq = (torch.randn(d, d, device="cuda", generator=g) * 0.03).to(torch.bfloat16)
    wkv = (torch.randn(2 * d, d, device="cuda", generator=g) * 0.03).to(torch.bfloat16)
    bq = torch.randn(d, device="cuda", generator=g) * 0.5
    bkv = torch.randn(2 * d, device="cuda", generator=g) * 0.5
    rel = torch.randn(27, 64, device="cuda", generator=g) * 0.1
    op = _rel_operand(rel, rel, 32)
    # reference route: zero-padded window partition, projection of all 70 x 70 tokens, partitioned attention
    yp = torch.zeros(n_img, P, P, d, device="cuda", dtype=torch.bfloat16)
    yp[:, :hw, :hw] = y.view(n_img, hw, hw, d)
    ywin = yp.view(n_img, nwin, 14, nwin, 14, d).permute(0, 1, 3, 2, 4, 5).reshape(n_seq * 196, d).contiguous()
    q_ref, kv_ref = ops.gemm(ywin, wq, bq), ops.gemm(ywin, wkv, bkv)
    out_ref = torch.zeros(n_img * hw * hw, d, device="cuda", dtype=torch.bfloat16)
    ops.attention_window(q_ref, kv_ref, n_seq, heads, 0.125, out_ref, 0, 0, d, op, 32, out_mode=1, nwin=nwin, img_hw=hw)
    # padded-grid route
    q, kv = ops.gemm_to_grid(y, wq, bq, hw, P), ops.gemm_to_grid(y, wkv, bkv, hw, P)
    assert q.shape == (n_img * P * P, d) and kv.shape == (n_img * P * P, 2 * d)
    qg = q.view(n_img, P, P, d)
    assert torch.equal(qg[:, :hw, :hw].reshape(-1, d), ops.gemm(y, wq, bq))
    assert torch.equal(qg[:, hw:, :], bq.to(torch.bfloat16).expand(n_img, P - hw, P, d))
    assert torch.equal(qg[:, :, hw:], bq.to(torch.bfloat16).expand(n_img, P, P - hw, d))
    out = torch.zeros(n_img * hw * hw, d, device="cuda", dtype=torch.bfloat16)
    ops.attention_window(q, kv, n_seq, heads, 0.125, out, 0, 0, d, op, 32, out_mode=1, nwin=nwin, img_hw=hw, in_pad=P)
    assert torch.equal(out, out_ref)


# ---------------------------------------------------------------------------------------------- row kernels
def test_add_layernorm_modes():
    ops = _ops()
    g = _gen(4)
    I, hw, d = 2, 64, 768
    rows = I * hw * hw
    x = torch.randn(rows, d, device="cuda", generator=g)
    delta = torch.randn(rows, d, device="cuda", generator=g).to(torch.bfloat16)
    gamma, beta = torch.randn(d, device="cuda", generator=g), torch.randn(d, device="cuda", generator=g)
    xr = x + delta.float()
    ln = F.layer_norm(xr, (d,), gamma, beta, 1e-6)
    x1, y, y2 = x.clone(), torch.empty(rows, d, device="cuda", dtype=torch.bfloat16), torch.empty(rows, d, device="cuda")
    ops.add_layernorm(x1, delta, gamma, beta, 1e-6, rows=rows, d=d, x_out=x1, y_out=y, y2_out=y2)
    assert torch.equal(x1, xr)                                    # fp32 add: exact
    _close(y2, ln, 1e-5, 1e-5, "LN fp32")
    _close(y, ln, 2 ** -8, 1e-3, "LN bf16")
    # the ViT block case (in-place residual update + bf16 LN output) runs through the staged bulk-copy kernel
    x2, ys = x.clone(), torch.empty(rows, d, device="cuda", dtype=torch.bfloat16)
    ops.add_layernorm(x2, delta, gamma, beta, 1e-6, rows=rows, d=d, x_out=x2, y_out=ys)
    assert torch.equal(x2, xr)
    _close(ys, ln, 2 ** -8, 1e-3, "LN staged bf16")
    x3 = x.clone()
    ops.add_layernorm(x3, None, gamma, beta, 1e-6, rows=rows, d=d, y_out=ys)      # no branch output to add
    _close(ys, F.layer_norm(x, (d,), gamma, beta, 1e-6), 2 ** -8, 1e-3, "LN staged, x only")
    # window partition with zero padding (F.pad after norm1)
    nwin, win = 5, 14
    orow = I * nwin * nwin * win * win
    yw = torch.full((orow, d), 7.0, device="cuda", dtype=torch.bfloat16)
    ops.add_layernorm(x.clone(), delta, gamma, beta, 1e-6, rows=orow, d=d, y_out=yw, map_mode=1, win=win, nwin=nwin, hw=hw)
    lw = F.pad(ln.view(I, hw, hw, d), (0, 0, 0, 6, 0, 6)).view(I, nwin, win, nwin, win, d).permute(0, 1, 3, 2, 4, 5)
    _close(yw, lw.reshape(orow, d), 2 ** -8, 1e-3, "LN window partition")
    # drop CLS
    L = 901
    xs = torch.randn(3 * L, d, device="cuda", generator=g)
    y3 = torch.empty(3 * (L - 1), d, device="cuda")
    ops.add_layernorm(xs, None, gamma, beta, 1e-12, rows=3 * L, d=d, y_out=y3, map_mode=2, seq_len=L)
    _close(y3, F.layer_norm(xs, (d,), gamma, beta, 1e-12).view(3, L, d)[:, 1:].reshape(-1, d), 1e-5, 1e-5, "LN drop cls")
    # pixel shuffle + LayerNorm2d + GELU on 32 channels (decoder upscaling)
    B, h, c = 2, 8, 32
    u = torch.randn(B * h * h * 4, c, device="cuda", generator=g).to(torch.bfloat16)
    gm, bt = torch.randn(c, device="cuda", generator=g), torch.randn(c, device="cuda", generator=g)
    y4 = torch.empty(B * 4 * h * h, c, device="cuda", dtype=torch.bfloat16)
    ops.add_layernorm(None, u, gm, bt, 1e-6, rows=B * h * h * 4, d=c, y_out=y4, act=ops.ACT_GELU, map_mode=3, hw=h)
    r4 = F.gelu(F.layer_norm(u.float(), (c,), gm, bt, 1e-6)).view(B, h, h, 2, 2, c).permute(0, 1, 3, 2, 4, 5)
    _close(y4, r4.reshape(B * 4 * h * h, c), 2 ** -8, 2e-3, "pixel shuffle LN GELU")
    # per-sequence vector + second bf16 addend + positional output
    T, S, D = 64, 6, 256
    k16 = torch.randn(S * T, D, device="cuda", generator=g).to(torch.bfloat16)
    d2 = torch.randn(S * T, D, device="cuda", generator=g).to(torch.bfloat16)
    sa = torch.randn(S, D, device="cuda", generator=g)
    pe = torch.randn(T, D, device="cuda", generator=g)
    gm, bt = torch.randn(D, device="cuda", generator=g), torch.randn(D, device="cuda", generator=g)
    y5, y5pe = torch.empty(S * T, D, device="cuda"), torch.empty(S * T, D, device="cuda", dtype=torch.bfloat16)
    ops.add_layernorm(None, k16, gm, bt, 1e-5, rows=S * T, d=D, y_out=y5, delta2=d2, seq_add=sa, seq_rows=T, pe=pe,
                      ype_out=y5pe)
    r5 = F.layer_norm(k16.float() + d2.float() + sa.repeat_interleave(T, 0), (D,), gm, bt, 1e-5)
    _close(y5, r5, 1e-5, 1e-5, "LN seq_add")
    _close(y5pe, r5 + pe.repeat(S, 1), 2 ** -8, 2e-3, "LN + pe")
    pooled = ops.add_layernorm_meanpool(None, k16, gm, bt, 1e-5, S, T, D, delta2=d2, seq_add=sa)
    _close(pooled, r5.view(S, T, D).mean(1), 1e-5, 1e-5, "LN meanpool")
    # the prompt-encoder norm4 at a size that takes the staged kernels: bf16 rows only + per-sequence vector, fp32 /
    # bf16 outputs, and the mean-pool variant
    T, S, D = 1024, 6, 512
    k16 = torch.randn(S * T, D, device="cuda", generator=g).to(torch.bfloat16)
    sa = torch.randn(S, D, device="cuda", generator=g)
    gm, bt = torch.randn(D, device="cuda", generator=g), torch.randn(D, device="cuda", generator=g)
    r6 = F.layer_norm(k16.float() + sa.repeat_interleave(T, 0), (D,), gm, bt, 1e-5)
    y6 = torch.empty(S * T, D, device="cuda")
    ops.add_layernorm(None, k16, gm, bt, 1e-5, rows=S * T, d=D, y_out=y6, seq_add=sa, seq_rows=T)
    _close(y6, r6, 1e-5, 1e-5, "LN staged seq_add fp32")
    y6b = torch.empty(S * T, D, device="cuda", dtype=torch.bfloat16)
    ops.add_layernorm(None, k16, gm, bt, 1e-5, rows=S * T, d=D, y_out=y6b, seq_add=sa, seq_rows=T)
    _close(y6b, r6, 2 ** -8, 2e-3, "LN staged seq_add bf16")
    pooled = ops.add_layernorm_meanpool(None, k16, gm, bt, 1e-5, S, T, D, seq_add=sa)
    _close(pooled, r6.view(S, T, D).mean(1), 1e-5, 1e-5, "LN staged meanpool")
    # ... and with the second bf16 stream (image tokens + last MLP output), the prompt encoder's final norm4
    d26 = torch.randn(S * T, D, device="cuda", generator=g).to(torch.bfloat16)
    r7 = F.layer_norm(k16.float() + d26.float() + sa.repeat_interleave(T, 0), (D,), gm, bt, 1e-5)
    pooled = ops.add_layernorm_meanpool(None, k16, gm, bt, 1e-5, S, T, D, delta2=d26, seq_add=sa)
    _close(pooled, r7.view(S, T, D).mean(1), 1e-5, 1e-5, "LN staged meanpool, two bf16 streams")


def test_layout_and_index_kernels_are_exact():
    ops = _ops()
    g = _gen(6)
    x = torch.randn(3, 40, 5, 7, device="cuda", generator=g)
    t32, t16 = ops.nchw_to_tokens(x, want_f32=True, want_bf16=True)
    ref = x.permute(0, 2, 3, 1).reshape(3 * 35, 40)
    assert torch.equal(t32, ref) and torch.equal(t16, ref.to(torch.bfloat16))
    sq = torch.randn(2, 24, 6, 6, device="cuda", generator=g)
    assert torch.equal(ops.tokens_to_nchw(ops.nchw_to_tokens(sq)[0], 2, 6, 6), sq)
    rows = torch.randn(4 * 3 * 5, 16, device="cuda", generator=g)
    assert torch.equal(ops.permute_rows(rows, 4, 3, 5), rows.view(4, 3, 5, 16).permute(0, 2, 1, 3).reshape(-1, 16))
    code = torch.randn(5, 16, device="cuda", generator=g)
    assert torch.equal(ops.add_bcast(rows, code, 3, 5), rows + code[(torch.arange(60, device="cuda") // 3) % 5])
    s32, s16 = ops.copy_slabs(rows, 4, 5, 15, 3, want_f32=True, want_bf16=True)
    refs = rows.view(4, 15, 16)[:, 3:8].reshape(-1, 16)
    assert torch.equal(s32, refs) and torch.equal(s16, refs.to(torch.bfloat16))
    img = torch.randn(2, 3, 64, 64, device="cuda", generator=g)
    assert torch.equal(ops.im2col_patch16(img), F.unfold(img, 16, stride=16).transpose(1, 2).reshape(-1, 768).to(torch.bfloat16))
    f = torch.randn(2 * 9 * 9, 16, device="cuda", generator=g).to(torch.bfloat16)
    r3 = F.unfold(f.float().view(2, 9, 9, 16).permute(0, 3, 1, 2), 3, padding=1).view(2, 16, 9, 81).permute(0, 3, 2, 1)
    assert torch.equal(ops.im2col_3x3(f, 2, 9, 9, 16), r3.reshape(-1, 144).to(torch.bfloat16))
    emb = torch.randn(2, 3, 4, 16, device="cuda", generator=g)
    fl = (torch.rand(2, 3, 4, device="cuda", generator=g) > 0.4).to(torch.uint8)
    fl[0, :, 1] = 0
    n = fl.sum(1, keepdim=True).clamp(min=1).transpose(1, 2)
    _close(ops.masked_mean(emb, fl), (emb * fl[..., None]).sum(1) / n, 1e-6, 1e-6, "masked mean")


# ---------------------------------------------------------------------------------------------- token attention
@pytest.mark.parametrize("n_seq,nq,nk,dh,with_tables", [
    (7, 1, 4096, 32, True),     # token -> image, one token per sequence (key-parallel with splits)
    (3, 6, 900, 16, True),      # class tokens -> 30x30 image (MAE-256 decoder)
    (2, 4096, 9, 32, True),     # image -> tokens (one thread per query row: <= 16 keys)
    (3, 900, 3, 16, True),      # image -> class tokens (MAE-256 decoder)
    (2, 300, 16, 64, False),    # 16 keys, head_dim 64 (row kernel at its register limit)
    (2, 40, 9, 32, True),       # few queries, few keys: the lanes-per-unit kernel
    (3, 9, 4096, 32, True),     # nine prompt tokens -> image: all queries in ONE pass over the keys (KP_QB 10)
    (2, 21, 900, 16, True),     # 20-way decoder: 21 class tokens, three passes of 10
    (2, 6, 4096, 64, False),    # six class tokens (KP_QB 6), head_dim 64
    (4, 150, 150, 64, False),   # example attention over M*C tokens
    (5, 9, 9, 8, False),        # tiny self-attention
])
def test_attention_tokens_matches_torch(n_seq, nq, nk, dh, with_tables):
    ops = _ops()
    H = 8
    g = _gen(nq + nk)
    q = torch.randn(n_seq * nq, H * dh, device="cuda", generator=g).to(torch.bfloat16)
    k = torch.randn(n_seq * nk, H * dh, device="cuda", generator=g).to(torch.bfloat16)
    v = torch.randn(n_seq * nk, H * dh, device="cuda", generator=g).to(torch.bfloat16)
    qa = torch.randn(nq, H * dh, device="cuda", generator=g) * 0.3 if with_tables else None
    ka = torch.randn(nk, H * dh, device="cuda", generator=g) * 0.3 if with_tables else None
    out = ops.attention_tokens(q, k, v, n_seq, nq, nk, H, dh, q_add=qa, k_add=ka)
    qf = q.float().view(n_seq, nq, H, dh) + (qa.view(1, nq, H, dh) if qa is not None else 0)
    kf = k.float().view(n_seq, nk, H, dh) + (ka.view(1, nk, H, dh) if ka is not None else 0)
    att = torch.softmax(torch.einsum("sqhd,skhd->shqk", qf, kf) / math.sqrt(dh), -1)
    ref = torch.einsum("shqk,skhd->sqhd", att, v.float().view(n_seq, nk, H, dh)).reshape(n_seq * nq, H * dh)
    _close(out, ref, 2 ** -8, 2e-3, "attention_tokens")   # fp32 math, one bf16 rounding of the output


# ---------------------------------------------------------------------------------------------- prompt / decoder kernels
def test_prompt_side_kernels_match_the_oracle_functions():
    ops, O = _ops(), _oracle()
    from labelanything_b200.build_lam import build_lam_no_vit
    from labelanything_b200.common import f32
    from labelanything_b200.synthetic import load_synth_weights

    lam = build_lam_no_vit(image_embed_dim=384, embed_dim=256, image_size=256, spatial_convs=3)
    load_synth_weights(lam, seed=9)
    sd = {k: v.clone() for k, v in lam.state_dict().items()}
    pe = lam.cuda().prompt_encoder
    g = torch.Generator().manual_seed(11)
    B, M, C, S = 1, 2, 3, 6
    # mask downscaling (channels 0..15) + bilinear resize: compare after the 1x1 conv in fp32
    masks = (torch.rand(B, M, C, 64, 64, generator=g) > 0.5).float()
    m16 = ops.mask_downscale(masks.view(S, 64, 64).cuda(), pe._mask_host_weights())            # [S, 16, 16, 16]
    md = "prompt_encoder.mask_downscaling"
    x = F.conv2d(masks.view(S, 1, 64, 64), sd[md + ".0.weight"], sd[md + ".0.bias"], stride=2)
    x = F.gelu(O.layer_norm_2d(sd, md + ".1", x))
    x = F.conv2d(x, sd[md + ".3.weight"], sd[md + ".3.bias"], stride=2)
    x = F.gelu(O.layer_norm_2d(sd, md + ".4", x))
    _close(m16.cpu().permute(0, 3, 1, 2), x, 1e-4, 2e-5, "mask downscale")
    r = ops.resize_bilinear(m16, 5, 7)
    _close(r.cpu().permute(0, 3, 1, 2), F.interpolate(m16.cpu().permute(0, 3, 1, 2), (5, 7), mode="bilinear"), 1e-5,
           1e-6, "bilinear resize")
    # sparse tokens: points (incl. null / negative), boxes (incl. null with the reference's repeat() indexing)
    pts = torch.rand(B, M, C, 4, 2, generator=g) * 256
    lab = torch.randint(-1, 2, (B, M, C, 4), generator=g).float()
    bxs = torch.rand(B, M, C, 2, 4, generator=g) * 256
    bfl = torch.randint(-1, 2, (B, M, C, 2), generator=g).float()
    ref = torch.cat([O.embed_points(sd, "prompt_encoder", pts.view(S, 4, 2), lab.view(S, 4), pad=False, image_size=256),
                     O.embed_boxes(sd, "prompt_encoder", bxs, bfl, 256)], dim=1)
    got = ops.embed_sparse(pts.view(S, 4, 2).cuda(), lab.view(S, 4).cuda(), bxs.view(S, 2, 4).cuda(),
                           bfl.view(S, 2).cuda(), f32(pe, "gauss", pe.pe_layer.positional_encoding_gaussian_matrix),
                           f32(pe, "nap", pe.not_a_point_embed.weight).view(-1), pe._pe_table4(), S, 256, 256, 256)
    _close(got.cpu(), ref, 1e-4, 2e-4, "sparse embedding")     # sin/cos of arguments up to ~2*pi*5: fp32 range reduction
    ref = O.embed_points(sd, "prompt_encoder", pts.view(S, 4, 2), lab.view(S, 4), pad=True, image_size=256)
    got = ops.embed_sparse(pts.view(S, 4, 2).cuda(), lab.view(S, 4).cuda(), None, None,
                           f32(pe, "gauss", pe.pe_layer.positional_encoding_gaussian_matrix),
                           f32(pe, "nap", pe.not_a_point_embed.weight).view(-1), pe._pe_table4(), S, 256, 256, 256)
    _close(got.cpu(), ref, 1e-4, 2e-4, "sparse embedding (padded point)")
    # dense positional encoding table
    _close(pe.dense_pe_tokens().cpu(), O.dense_pe(sd["prompt_encoder.pe_layer.positional_encoding_gaussian_matrix"], 16, 16)
           [0].permute(1, 2, 0).reshape(256, 256), 1e-5, 1e-5, "dense pe")


@pytest.mark.parametrize("D,T", [(64, 40), (96, 40), (256, 200), (512, 333)])
def test_build_src_matches_reference_composition(D, T):
    """D % 64 == 0 takes the tensor-core kernel (TF32 products, fp32 accumulation: 16 products of N(0,1) operands
    carry ~1.6e-3 of absolute error, hence the 8e-3 absolute term next to the bf16 output rounding 2^-8); D = 96 the
    FFMA2 kernel (fp32 products: 1e-3)."""
    ops = _ops()
    g = _gen(12)
    B, M, C, lead = 2, 2, 3, 1
    atol = 8e-3 if D % 64 == 0 else 1e-3
    feat = torch.randn(B * (M + lead) * T, D, device="cuda", generator=g)
    S = B * M * C
    m16 = torch.randn(S, T, 16, device="cuda", generator=g)
    fl = torch.tensor([1, 0, 1, 1, 1, 0, 1, 1, 0, 1, 1, 1], dtype=torch.uint8, device="cuda")
    w6, b6 = torch.randn(D, 16, device="cuda", generator=g), torch.randn(D, device="cuda", generator=g)
    nam, nom = torch.randn(D, device="cuda", generator=g), torch.randn(D, device="cuda", generator=g)
    code = torch.randn(C, D, device="cuda", generator=g)
    out = ops.build_src(feat, m16, fl, w6, b6, nam, nom, code, S, T, D, C, M, feat_lead=lead)
    f = feat.view(B, M + lead, T, D)[:, lead:].unsqueeze(2).expand(B, M, C, T, D).reshape(S, T, D)
    dense = torch.where(fl.view(S, 1, 1).bool(), m16.double() @ w6.double().t() + b6, nam.expand(S, T, D).double())
    ref = (f + dense + code.repeat(B * M, 1).view(S, 1, D)).float()
    _close(out.view(S, T, D), ref, 2 ** -8, atol, "build_src")
    # sequences with a null mask and calls without masks involve no product: fp32 adds + one bf16 rounding
    null = ~fl.bool()
    _close(out.view(S, T, D)[null], ref[null], 2 ** -8, 1e-5, "build_src (null masks)")
    out2 = ops.build_src(feat, None, None, w6, b6, nam, nom, None, S, T, D, C, M, feat_lead=lead)
    _close(out2.view(S, T, D), f + nom, 2 ** -8, 1e-5, "build_src (no masks)")
    # second episode only (chunked passes): same bits, whatever the grid
    out3 = ops.build_src(feat, m16[M * C:], fl[M * C:], w6, b6, nam, nom, code, M * C, T, D, C, M, feat_lead=lead,
                         seq_offset=M * C)
    assert torch.equal(out3, out.view(S, T, D)[M * C:].reshape(-1, D))


def test_classify_and_postprocess_match_torch():
    ops, O = _ops(), _oracle()
    g = _gen(13)
    B, C, dk, P = 2, 6, 64, 32 * 32
    x = torch.randn(B * P, dk, device="cuda", generator=g).to(torch.bfloat16)
    cls = torch.randn(B, C, dk, device="cuda", generator=g)
    lg = ops.classify(x, cls, B, P)
    _close(lg, cls @ x.float().view(B, P, dk).transpose(1, 2), 1e-5, 1e-4, "classify")
    low = lg.view(B, C, 32, 32).contiguous()
    dims = torch.tensor([[[100, 128], [128, 128]], [[128, 77], [128, 128]]])
    for custom in (True, False):
        ref = O.postprocess_masks(low.cpu(), dims, 128, custom)
        ref[torch.tensor([[True] * 5 + [False], [True] * 6]).logical_not()] = float("-inf")
        sizes = [(oh, ow) + (O.preprocess_shape(oh, ow, 128) if custom else (128, 128)) for oh, ow in dims[:, 0].tolist()]
        fg = torch.tensor([[1] * 5 + [0], [1] * 6], dtype=torch.uint8, device="cuda")
        out = ops.postprocess_masks(low, torch.tensor(sizes, dtype=torch.int32, device="cuda"), fg, 128, 128, 128).cpu()
        fin = torch.isfinite(ref)
        assert torch.equal(torch.isfinite(out), fin) and torch.equal(out[~fin], ref[~fin])
        _close(out[fin], ref[fin], 1e-5, 2e-5, "postprocess")


# ---------------------------------------------------------------------------------------------- pooled attention
@pytest.mark.parametrize("S,T,rows,D", [(5, 4096, 8, 512), (3, 900, 8, 256), (2, 70, 3, 128), (4, 64, 1, 64),
                                        (2, 1000, 8, 512)])
def test_pooled_attention_matches_torch(S, T, rows, D):
    """y = softmax(scale (u x^T + e)) x with x as key AND value (la_attention_pooled_bf16): identical bf16 inputs vs
    torch fp32; ragged last tile (T % 64 != 0), fewer than 8 query rows, every supported width."""
    ops = _ops()
    g = _gen(S + T + rows + D)
    x = (0.5 * torch.randn(S * T, D, device="cuda", generator=g) + 0.25).to(torch.bfloat16)
    u = (torch.randn(S * rows, D, device="cuda", generator=g) * (4.0 / math.sqrt(D))).to(torch.bfloat16)
    e = torch.randn(S * rows, T, device="cuda", generator=g)
    scale = 0.7
    y = ops.attention_pooled(x, u, e, scale, S, T, rows)
    xf, uf = x.float().view(S, T, D), u.float().view(S, rows, D)
    p = torch.softmax((uf @ xf.transpose(1, 2) + e.view(S, rows, T)) * scale, dim=-1)
    ref = (p @ xf).reshape(S * rows, D)
    # bf16 rounding of P (averaged over the tokens carrying weight) and of the output: 2^-8 relative + 2e-3 absolute
    _close(y, ref, 2 ** -8, 2e-3, "pooled attention")
    y0 = ops.attention_pooled(x, u, None, scale, S, T, rows)
    ref0 = (torch.softmax(uf @ xf.transpose(1, 2) * scale, dim=-1) @ xf).reshape(S * rows, D)
    _close(y0, ref0, 2 ** -8, 2e-3, "pooled attention, no positional scores")


def test_head_rows_expand_and_gather_are_exact():
    ops = _ops()
    S, H, dh = 7, 8, 32
    t = torch.randn(S, H * dh, device="cuda").to(torch.bfloat16)
    ex = ops.head_rows(t, S, H, dh, expand=True)
    ref = torch.zeros(S, H, H, dh, device="cuda", dtype=torch.bfloat16)
    for h in range(H):
        ref[:, h, h] = t.view(S, H, dh)[:, h]
    assert torch.equal(ex, ref.view(S * H, H * dh))
    assert torch.equal(ops.head_rows(ex, S, H, dh, expand=False), t)


def test_pooled_token_to_image_equals_the_projected_path():
    """The associativity rewrite of the single-query token->image attention against the k / v projection path it
    replaces (same module, same inputs): equal up to bf16 rounding of the intermediates."""
    from labelanything_b200 import ops as O
    from labelanything_b200 import transformer as TR
    from labelanything_b200.synthetic import load_synth_weights

    D, H, S, T = 256, 8, 6, 900
    tw = TR.TwoWayTransformer(depth=2, embedding_dim=D, num_heads=H, mlp_dim=512)
    load_synth_weights(tw, seed=3)
    tw = tw.cuda()
    g = _gen(11)
    keys16 = torch.randn(S * T, D, device="cuda", generator=g).to(torch.bfloat16)
    pe = torch.randn(T, D, device="cuda", generator=g)
    tok = torch.randn(S, D, device="cuda", generator=g)
    with torch.no_grad():
        _, _, pooled_new = TR.run_two_way(tw, keys16, None, pe, tok, S, T, 1, want_queries=False, pool=True)
        O._NO_POOLED_ATTENTION = True
        try:
            _, _, pooled_old = TR.run_two_way(tw, keys16, None, pe, tok, S, T, 1, want_queries=False, pool=True)
        finally:
            O._NO_POOLED_ATTENTION = False
    err = (pooled_new - pooled_old).abs()
    print(f"pooled vs projected path: max {err.max().item():.4e} mean {err.mean().item():.4e} "
          f"(|ref| mean {pooled_old.abs().mean().item():.3f})")
    # two bf16 evaluations of two transformer layers (LayerNorm outputs, |x| ~ 0.7): rounding-noise level
    assert err.max().item() < 4e-2 and err.mean().item() < 6e-3
