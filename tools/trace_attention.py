"""Timeline of one CTA of the fused attention kernel (clock64 stamps): python tools/trace_attention.py [global|plain|window] [f16]"""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch

from labelanything_b200 import _native, ops

mode = sys.argv[1] if len(sys.argv) > 1 else "global"
tab_dtype = torch.float16 if len(sys.argv) > 2 and sys.argv[2] == "f16" else torch.float32
heads = 12
n_seq, L, gsz = {"global": (8, 4096, 64), "plain": (8, 4096, 0), "window": (200, 196, 14)}[mode]
qkv = torch.randn(n_seq * L, 3 * heads * 64, device="cuda").to(torch.bfloat16)
out = torch.zeros(n_seq * L, heads * 64, device="cuda", dtype=torch.bfloat16)
bh = bw = None
if gsz == 64:
    bh = (torch.randn(n_seq * L, heads, 128, device="cuda") * 0.1).to(tab_dtype)
    bw = (torch.randn(n_seq * L, heads, 128, device="cuda") * 0.1).to(tab_dtype)
tr = torch.zeros(5, 192, 4, dtype=torch.int64, device="cuda")
op = torch.zeros(64, 64, device="cuda", dtype=torch.bfloat16)
op[:27] = (torch.randn(27, 64, device="cuda") * 0.1).to(torch.bfloat16)
op[32:59] = (torch.randn(27, 64, device="cuda") * 0.1).to(torch.bfloat16)


def run():
    if gsz == 14:
        ops.attention_window(qkv, qkv, n_seq, heads, 0.125, out, 0, heads * 64, 2 * heads * 64, op, 32)
    else:
        ops.attention(qkv, qkv, n_seq, L, heads, 0.125, out, 0, heads * 64, 2 * heads * 64, bh, bw, grid_hw=gsz)


for _ in range(3):
    run()
e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
e0.record()
for _ in range(10):
    run()
e1.record()
torch.cuda.synchronize()
print(f"{mode} tables={tab_dtype}: {e0.elapsed_time(e1) / 10:.3f} ms per launch")
_native.lib().la_attention_set_trace(tr.data_ptr())
run()
torch.cuda.synchronize()
_native.lib().la_attention_set_trace(None)
t = tr.cpu()
t0 = int(t[t > 0].min())
t = (t - t0).clamp(min=-1)
print("tile | MMA: P_A seen, PV_A issued, P_B seen, PV_B issued | softA: waitS, gotS, max done, P arrived | softB: same")
for j in (list(range(0, 4)) + list(range(58, 72)) + list(range(124, 134))) if mode != "window" else list(range(0, 24)):
    print(j, t[0, j].tolist(), t[1, j].tolist(), t[2, j].tolist(), t[3, j].tolist(), t[4, j].tolist())
d = t[1, 8:60]
print("item boundaries (softmax A): O ready, epilogue done, T ready, prologue done")
for i in range(0, 8):
    print(i, t[3, i].tolist(), " B:", t[4, i].tolist())
print("softmax A steady state: mean wait-for-S", float((d[:, 1] - d[:, 0]).float().mean()), "pass1", float((d[:, 2] - d[:, 1]).float().mean()),
      "pass2+store", float((d[:, 3] - d[:, 2]).float().mean()), "period", float((d[1:, 3] - d[:-1, 3]).float().mean()))
d = t[2, 8:60]
print("softmax B steady state: mean wait-for-S", float((d[:, 1] - d[:, 0]).float().mean()), "pass1", float((d[:, 2] - d[:, 1]).float().mean()),
      "pass2+store", float((d[:, 3] - d[:, 2]).float().mean()), "period", float((d[1:, 3] - d[:-1, 3]).float().mean()))
m = t[0, 8:60]
print("MMA: P_A seen -> PV_A issued", float((m[:, 1] - m[:, 0]).float().mean()), "PV_A issued -> P_B seen", float((m[:, 2] - m[:, 1]).float().mean()),
      "P_B seen -> PV_B issued", float((m[:, 3] - m[:, 2]).float().mean()))
