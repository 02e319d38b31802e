"""Run the mean-pool LayerNorm of the prompt encoder at the BASELINE size (for ncu captures / timing):
python tools/run_meanpool.py [n_launches]"""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch

from labelanything_b200 import ops

S, T, D = 1200, 4096, 512
k16 = torch.randn(S * T, D, device="cuda", dtype=torch.bfloat16)
sa = torch.randn(S, D, device="cuda")
gm, bt = torch.randn(D, device="cuda"), torch.randn(D, device="cuda")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 3
e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
for i in range(n):
    if i == n - 1:
        e0.record()
    out = ops.add_layernorm_meanpool(None, k16, gm, bt, 1e-5, S, T, D, seq_add=sa)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
print(f"meanpool {S}x{T}x{D}: {ms:.3f} ms = {S * T * D * 2 / ms / 1e6:.0f} GB/s")
