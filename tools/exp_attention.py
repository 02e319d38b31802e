"""A/B of attention-kernel builds on one GPU: python tools/exp_attention.py [tag ...]
Each tag is a library built by `python -m labelanything_b200.build --variant <tag> <defines>` ("product" = the in-tree
product library).  Every build runs in its own process (LA_B200_LIB), checks the three modes against torch fp32 on
identical bf16 inputs and times them; torch sdpa (no bias) is timed beside them.  Appends to gpurun_out/exp_attention.log"""
import os
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def child():
    import torch
    import torch.nn.functional as F

    from labelanything_b200 import ops
    sys.path.insert(0, str(ROOT / "tools"))
    from diag_attention import ref_attention, rel_operand, rev_table_bias

    heads = 12
    tag = os.environ.get("EXP_TAG", "?")

    def timeit(f, n=10):
        for _ in range(3):
            f()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        for _ in range(n):
            f()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n

    res = [tag]
    # ---- correctness (small) ----
    g = torch.Generator(device="cuda").manual_seed(2)
    L = 4096
    qkv = torch.randn(L, 3 * heads * 64, device="cuda", generator=g)
    qkv[:, 2 * heads * 64:] = 1.0 + 0.5 * qkv[:, 2 * heads * 64:]
    qkv = qkv.to(torch.bfloat16)
    rel_h = torch.randn(127, 64, device="cuda", generator=g) * 0.1
    rel_w = torch.randn(127, 64, device="cuda", generator=g) * 0.1
    qh = qkv[:, : heads * 64].reshape(L, heads, 64).permute(1, 0, 2).contiguous()
    bh = rev_table_bias(qh, rel_h, 128).half()
    bw = rev_table_bias(qh, rel_w, 128).half()
    out = torch.zeros(L, heads * 64, device="cuda", dtype=torch.bfloat16)
    ops.attention(qkv, qkv, 1, L, heads, 0.125, out, 0, heads * 64, 2 * heads * 64, bh, bw, grid_hw=64)
    ref = ref_attention(qkv, 1, L, heads, 0.125, rel_h, rel_w, 64)
    e = (out.float() - ref).abs()
    res.append(f"global16 err max {e.max().item():.2e} mean {e.mean().item():.2e}")
    for Lp in (901, 197):
        q2 = torch.randn(3 * Lp, 3 * heads * 64, device="cuda", generator=g)
        q2[:, 2 * heads * 64:] = 1.0 + 0.5 * q2[:, 2 * heads * 64:]
        q2 = q2.to(torch.bfloat16)
        o2 = torch.zeros(3 * Lp, heads * 64, device="cuda", dtype=torch.bfloat16)
        ops.attention(q2, q2, 3, Lp, heads, 0.125, o2, 0, heads * 64, 2 * heads * 64)
        e = (o2.float() - ref_attention(q2, 3, Lp, heads, 0.125)).abs()
        res.append(f"plain{Lp} err max {e.max().item():.2e}")
    n_seq = 50
    q3 = torch.randn(n_seq * 196, 3 * heads * 64, device="cuda", generator=g)
    q3[:, 2 * heads * 64:] = 1.0 + 0.5 * q3[:, 2 * heads * 64:]
    q3 = q3.to(torch.bfloat16)
    rh, rw = torch.randn(27, 64, device="cuda", generator=g) * 0.1, torch.randn(27, 64, device="cuda", generator=g) * 0.1
    op = rel_operand(rh, rw, 32)
    o3 = torch.zeros(n_seq * 196, heads * 64, device="cuda", dtype=torch.bfloat16)
    ops.attention_window(q3, q3, n_seq, heads, 0.125, o3, 0, heads * 64, 2 * heads * 64, op, 32)
    e = (o3.float() - ref_attention(q3, n_seq, 196, heads, 0.125, rh, rw, 14)).abs()
    res.append(f"window err max {e.max().item():.2e} mean {e.mean().item():.2e}")
    # ---- timing at in-step sizes: 37 images = 48 items per SM (global), 52 images of windows, 64 HF images ----
    for name, n_seq, L, gsz in [("global64", 37, 4096, 64), ("window14", 52 * 25, 196, 14), ("plain901", 64, 901, 0)]:
        qkv = torch.randn(n_seq * L, 3 * heads * 64, device="cuda").to(torch.bfloat16)
        out = torch.zeros(n_seq * L, heads * 64, device="cuda", dtype=torch.bfloat16)
        bh = bw = None
        if gsz == 64:
            bh = (torch.randn(n_seq * L, heads, 128, device="cuda") * 0.1).half()
            bw = (torch.randn(n_seq * L, heads, 128, device="cuda") * 0.1).half()
        if gsz == 14:
            f = lambda: ops.attention_window(qkv, qkv, n_seq, heads, 0.125, out, 0, heads * 64, 2 * heads * 64, op, 32)
        else:
            f = lambda: ops.attention(qkv, qkv, n_seq, L, heads, 0.125, out, 0, heads * 64, 2 * heads * 64, bh, bw,
                                      grid_hw=gsz)
        ms = timeit(f)
        fl = 4.0 * n_seq * heads * L * L * 64
        s = f"{name} {ms:.3f} ms {fl / ms / 1e9:.0f} TF"
        if tag == "product":
            q, k, v = qkv.view(n_seq, L, 3, heads, 64).permute(2, 0, 3, 1, 4)
            ms_t = timeit(lambda: F.scaled_dot_product_attention(q, k, v))
            s += f" (sdpa no-bias {ms_t:.3f} ms {fl / ms_t / 1e9:.0f} TF)"
        res.append(s)
        del qkv, out, bh, bw
    print(" | ".join(res), flush=True)


if __name__ == "__main__":
    if os.environ.get("EXP_CHILD"):
        child()
        sys.exit(0)
    tags = sys.argv[1:] or ["product"]
    out_dir = ROOT / "gpurun_out"
    out_dir.mkdir(exist_ok=True)
    with open(out_dir / "exp_attention.log", "a") as log:
        for tag in tags:
            env = dict(os.environ, EXP_CHILD="1", EXP_TAG=tag)
            if tag != "product":
                env["LA_B200_LIB"] = str(ROOT / "labelanything_b200" / "_variants" / f"liblabelanything_b200_{tag}.so")
            r = subprocess.run([sys.executable, __file__], env=env, capture_output=True, text=True, timeout=200)
            line = r.stdout.strip().splitlines()[-1] if r.stdout.strip() else f"{tag}: FAILED rc={r.returncode} {r.stderr[-800:]}"
            print(line, flush=True)
            log.write(line + "\n")
            log.flush()
