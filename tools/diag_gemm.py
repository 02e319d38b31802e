"""GPU diagnostic for the tcgen05 GEMM: several shapes vs torch.matmul, with error-structure dumps.
Run under gpurun; writes gpurun_out/diag_gemm.log."""
import json
import os
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch

from labelanything_b200 import ops

out_dir = ROOT / "gpurun_out"
out_dir.mkdir(exist_ok=True)
log = open(out_dir / "diag_gemm.log", "w")


def P(*a):
    s = " ".join(str(x) for x in a)
    print(s, flush=True)
    log.write(s + "\n")
    log.flush()


def ref_gemm(a, w, bias, act):
    y = a.float() @ w.float().t()
    if bias is not None:
        y = y + bias
    if act == 1:
        y = torch.nn.functional.gelu(y)
    elif act == 2:
        y = torch.relu(y)
    return y


def run(M, N, K, act=0, out_dtype=torch.bfloat16, use_bias=True, seed=0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    a = torch.randn(M, K, device="cuda", generator=g).to(torch.bfloat16)
    w = (torch.randn(N, K, device="cuda", generator=g) / K ** 0.5).to(torch.bfloat16)
    bias = torch.randn(N, device="cuda", generator=g) if use_bias else None
    y = ops.gemm(a, w, bias, act=act, out_dtype=out_dtype)
    torch.cuda.synchronize()
    r = ref_gemm(a, w, bias, act)
    err = (y.float() - r).abs()
    tol = 2e-2 if out_dtype == torch.bfloat16 else 2e-3
    bad = err > tol * (1 + r.abs())
    nbad = int(bad.sum())
    P(f"M={M} N={N} K={K} act={act} out={out_dtype} bias={use_bias}: max_abs_err={err.max().item():.3e} bad={nbad}/{M*N}")
    if nbad:
        idx = bad.nonzero()
        P("  first bad:", idx[:8].tolist())
        rows = torch.unique(idx[:, 0])
        cols = torch.unique(idx[:, 1])
        P(f"  bad rows: n={len(rows)} min={rows.min().item()} max={rows.max().item()} ; bad cols: n={len(cols)} min={cols.min().item()} max={cols.max().item()}")
        P("  y[0,:8] =", y[0, :8].float().tolist())
        P("  r[0,:8] =", r[0, :8].tolist())
        # structure probes: is y a permutation of r within a row?
        ys = torch.sort(y[0].float()).values
        rs = torch.sort(r[0]).values
        P("  row0 sorted diff:", (ys - rs).abs().max().item())
    return nbad == 0


def bench(M, N, K, act=0, iters=20):
    a = torch.randn(M, K, device="cuda").to(torch.bfloat16)
    w = torch.randn(N, K, device="cuda").to(torch.bfloat16)
    bias = torch.randn(N, device="cuda")
    out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    for _ in range(3):
        ops.gemm(a, w, bias, act=act, out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(iters):
        ops.gemm(a, w, bias, act=act, out=out)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    tf = 2.0 * M * N * K / ms / 1e9
    # torch reference
    for _ in range(3):
        torch.nn.functional.linear(a, w)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(iters):
        torch.nn.functional.linear(a, w)
    e1.record()
    torch.cuda.synchronize()
    ms_t = e0.elapsed_time(e1) / iters
    P(f"bench M={M} N={N} K={K} act={act}: ours {ms:.3f} ms = {tf:.1f} TFLOP/s ; torch(cuBLAS) {ms_t:.3f} ms = {2.0*M*N*K/ms_t/1e9:.1f} TFLOP/s")


if __name__ == "__main__":
    P(torch.cuda.get_device_name(0))
    ok = True
    ok &= run(128, 256, 64, use_bias=False)
    ok &= run(128, 256, 128, use_bias=False)
    ok &= run(128, 256, 768)
    ok &= run(256, 512, 768)
    ok &= run(4096, 2304, 768)
    ok &= run(4900, 768, 768)
    ok &= run(1000, 128, 512)
    ok &= run(1000, 64, 576)
    ok &= run(4096, 3072, 768, act=1)
    ok &= run(4096, 768, 3072, act=2)
    ok &= run(4096, 768, 3072, out_dtype=torch.float32)
    ok &= run(300, 256, 136, out_dtype=torch.float32)
    P("ALL_OK" if ok else "SOME_FAILED")
    if ok:
        bench(131072, 768, 768)
        bench(131072, 1536, 768)
        bench(131072, 3072, 768, act=1)
        bench(131072, 768, 3072)
        bench(8192, 8192, 8192)
