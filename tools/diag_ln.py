"""GPU timing of the add+LayerNorm kernels (staged vs register-load): python tools/diag_ln.py"""
import os
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch

from labelanything_b200 import ops


def timeit(f, iters=10):
    for _ in range(3):
        f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(iters):
        f()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    # A/B against the register-load kernels: build a variant with -DLA_LN_UNSTAGED=1 (sources=("la_rowops.cu",)) and
    # point LA_B200_LIB at it
    tag = os.environ.get("LA_B200_LIB", "product").rsplit("_", 1)[-1]
    g = torch.Generator(device="cuda").manual_seed(0)
    # (a) ViT block: x += delta; y = LN(x) bf16
    rows, d = 131072, 768
    x = torch.randn(rows, d, device="cuda", generator=g)
    delta = torch.randn(rows, d, device="cuda", generator=g).to(torch.bfloat16)
    gm, bt = torch.randn(d, device="cuda"), torch.randn(d, device="cuda")
    y = torch.empty(rows, d, device="cuda", dtype=torch.bfloat16)
    ms = timeit(lambda: ops.add_layernorm(x, delta, gm, bt, 1e-6, rows=rows, d=d, x_out=x, y_out=y))
    print(f"[{tag}] vit block d768 rows {rows}: {ms:.3f} ms = {rows * d * 12 / ms / 1e6:.0f} GB/s")
    ms = timeit(lambda: ops.add_layernorm(x, None, gm, bt, 1e-6, rows=rows, d=d, y_out=y))
    print(f"[{tag}] vit LN1 (x only) d768 rows {rows}: {ms:.3f} ms = {rows * d * 6 / ms / 1e6:.0f} GB/s")
    n_img, nwin, win, hw = 32, 5, 14, 64
    orow = n_img * nwin * nwin * win * win
    yw = torch.empty(orow, d, device="cuda", dtype=torch.bfloat16)
    ms = timeit(lambda: ops.add_layernorm(x, None, gm, bt, 1e-6, rows=orow, d=d, y_out=yw, map_mode=1, win=win, nwin=nwin, hw=hw))
    print(f"[{tag}] vit LN1 window partition (x only) rows {orow}: {ms:.3f} ms = {(rows * d * 4 + orow * d * 2) / ms / 1e6:.0f} GB/s")
    ms = timeit(lambda: ops.add_layernorm(x, delta, gm, bt, 1e-6, rows=orow, d=d, x_out=x, y_out=yw, map_mode=1, win=win, nwin=nwin, hw=hw))
    print(f"[{tag}] vit LN1 window partition (x += delta) rows {orow}: {ms:.3f} ms = {(rows * d * 10 + orow * d * 2) / ms / 1e6:.0f} GB/s")
    # (b) prompt-encoder norm4: bf16 rows + per-sequence vector -> bf16
    S, T, d = 300, 4096, 512
    k16 = torch.randn(S * T, d, device="cuda", generator=g).to(torch.bfloat16)
    sa = torch.randn(S, d, device="cuda", generator=g)
    gm, bt = torch.randn(d, device="cuda"), torch.randn(d, device="cuda")
    y = torch.empty(S * T, d, device="cuda", dtype=torch.bfloat16)
    ms = timeit(lambda: ops.add_layernorm(None, k16, gm, bt, 1e-5, rows=S * T, d=d, y_out=y, seq_add=sa, seq_rows=T))
    print(f"[{tag}] norm4 d512 rows {S * T}: {ms:.3f} ms = {S * T * d * 4 / ms / 1e6:.0f} GB/s")
    ms = timeit(lambda: ops.add_layernorm_meanpool(None, k16, gm, bt, 1e-5, S, T, d, seq_add=sa))
    print(f"[{tag}] meanpool d512 rows {S * T}: {ms:.3f} ms = {S * T * d * 2 / ms / 1e6:.0f} GB/s")
    # (c) neck: fp32 rows -> bf16
    rows, d = 131072, 512
    x = torch.randn(rows, d, device="cuda", generator=g)
    y = torch.empty(rows, d, device="cuda", dtype=torch.bfloat16)
    ms = timeit(lambda: ops.add_layernorm(x, None, gm, bt, 1e-6, rows=rows, d=d, y_out=y))
    print(f"[{tag}] neck d512 rows {rows}: {ms:.3f} ms = {rows * d * 6 / ms / 1e6:.0f} GB/s")


if __name__ == "__main__":
    main()
