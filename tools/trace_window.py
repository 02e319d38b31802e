"""Timeline of one CTA of the second-generation window-attention kernel (clock64 stamps, -DLA_ATT_TRACE build):
LA_B200_LIB=.../liblabelanything_b200_<tag>.so python tools/trace_window.py"""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch

from labelanything_b200 import _native, ops

heads, n_seq = 12, 52 * 25
qkv = torch.randn(n_seq * 196, 3 * heads * 64, device="cuda").to(torch.bfloat16)
out = torch.zeros(n_seq * 196, heads * 64, device="cuda", dtype=torch.bfloat16)
op = torch.zeros(64, 64, device="cuda", dtype=torch.bfloat16)
op[:27] = (torch.randn(27, 64, device="cuda") * 0.1).to(torch.bfloat16)
op[32:59] = (torch.randn(27, 64, device="cuda") * 0.1).to(torch.bfloat16)
ITEMS, EVENTS = 64, 6
tr = torch.zeros(4, ITEMS, EVENTS, dtype=torch.int64, device="cuda")


def run():
    ops.attention_window(qkv, qkv, n_seq, heads, 0.125, out, 0, heads * 64, 2 * heads * 64, op, 32)


for _ in range(3):
    run()
e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
e0.record()
for _ in range(10):
    run()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print(f"window: {ms:.3f} ms per launch, {4 * n_seq * heads * 196 * 196 * 64 / ms / 1e9:.0f} TF, "
      f"{n_seq * heads / 148:.1f} items per CTA")
_native.lib().la_attention_set_trace(tr.data_ptr())
run()
torch.cuda.synchronize()
_native.lib().la_attention_set_trace(None)
t = tr.cpu()
t0 = int(t[t > 0].min())
t = (t - t0).clamp(min=-1)
print("item | issuer A: inputs+TMEM free, T/S issued, P seen, PV issued | issuer B | softmax A: T+S ready, prologue done, "
      "pass 1 done, P delivered, O ready, stores issued | softmax B")
for i in list(range(0, 4)) + list(range(20, 30)):
    print(i, t[0, i, :4].tolist(), t[1, i, :4].tolist(), t[2, i].tolist(), t[3, i].tolist())
a = t[2, 8:60].float()
b = t[3, 8:60].float()
for name, d in (("A", a), ("B", b)):
    print(f"softmax {name}: period {float((d[1:, 0] - d[:-1, 0]).mean()):.0f}  prologue {float((d[:, 1] - d[:, 0]).mean()):.0f}  "
          f"pass1 {float((d[:, 2] - d[:, 1]).mean()):.0f}  pass2+store {float((d[:, 3] - d[:, 2]).mean()):.0f}  "
          f"wait O {float((d[:, 4] - d[:, 3]).mean()):.0f}  epilogue {float((d[:, 5] - d[:, 4]).mean()):.0f}  "
          f"next T+S wait {float((d[1:, 0] - d[:-1, 5]).mean()):.0f}")
ia = t[0, 8:60].float()
print(f"issuer A: wait inputs {float((ia[1:, 0] - ia[:-1, 3]).mean()):.0f}  T/S issue {float((ia[:, 1] - ia[:, 0]).mean()):.0f}  "
      f"wait P {float((ia[:, 2] - ia[:, 1]).mean()):.0f}  PV issue {float((ia[:, 3] - ia[:, 2]).mean()):.0f}")
