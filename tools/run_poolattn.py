"""Run la_attention_pooled_bf16 at the bench shape (for ncu captures / timing): python tools/run_poolattn.py [n_seq] [reps]"""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch

from labelanything_b200 import ops

S = int(sys.argv[1]) if len(sys.argv) > 1 else 1200
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
T, D, R = 4096, 512, 8
x = torch.randn(S * T, D, device="cuda").to(torch.bfloat16)
u = (torch.randn(S * R, D, device="cuda") * 0.2).to(torch.bfloat16)
e = torch.randn(S * R, T, device="cuda")
for _ in range(2):
    ops.attention_pooled(x, u, e, 0.17, S, T, R)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
e0.record()
for _ in range(reps):
    ops.attention_pooled(x, u, e, 0.17, S, T, R)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
print(f"pooled attention S={S} T={T} D={D} rows={R}: {ms:.3f} ms = {2.0 * S * T * D / ms / 1e6:.0f} GB/s of x")
