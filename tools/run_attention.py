"""Run the fused attention kernel a few times (for ncu captures):
python tools/run_attention.py [global|global16|window|plain] [repeats] [sequences]
global16 = the product's configuration of the 64x64 blocks: ONE fp16 table [rows, heads, 256] holding rel_h | rel_w."""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch

from labelanything_b200 import ops

mode = sys.argv[1] if len(sys.argv) > 1 else "global"
heads = 12
n_seq, L, gsz = {"global": (8, 4096, 64), "global16": (32, 4096, 64), "window": (200, 196, 14), "plain": (32, 901, 0)}[mode]
if len(sys.argv) > 3:
    n_seq = int(sys.argv[3])
qkv = torch.randn(n_seq * L, 3 * heads * 64, device="cuda").to(torch.bfloat16)
out = torch.zeros(n_seq * L, heads * 64, device="cuda", dtype=torch.bfloat16)
bh = bw = op = None
if gsz == 64 and mode == "global16":
    tab = (torch.randn(n_seq * L, heads, 256, device="cuda") * 0.1).to(torch.float16)
    bh, bw = tab[:, :, :128], tab[:, :, 128:]
elif gsz == 64:
    bh = torch.randn(n_seq * L, heads, 128, device="cuda") * 0.1
    bw = torch.randn(n_seq * L, heads, 128, device="cuda") * 0.1
if gsz == 14:
    op = torch.zeros(64, 64, device="cuda", dtype=torch.bfloat16)
    op[:27] = (torch.randn(27, 64, device="cuda") * 0.1).to(torch.bfloat16)
    op[32:59] = (torch.randn(27, 64, device="cuda") * 0.1).to(torch.bfloat16)
for _ in range(int(sys.argv[2]) if len(sys.argv) > 2 else 3):
    if gsz == 14:
        ops.attention_window(qkv, qkv, n_seq, heads, 0.125, out, 0, heads * 64, 2 * heads * 64, op, 32)
    else:
        ops.attention(qkv, qkv, n_seq, L, heads, 0.125, out, 0, heads * 64, 2 * heads * 64, bh, bw, grid_hw=gsz)
torch.cuda.synchronize()
print("done", mode)
