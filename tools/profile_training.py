"""Per-kernel-family time of one native training step (BASELINE configs[3] shape, see bench.training_measurement)."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))

import torch

from labelanything_b200 import ops
from labelanything_b200.build_lam import build_lam_no_vit
from labelanything_b200.loss import LabelAnythingLoss
from labelanything_b200.synthetic import load_synth_weights, make_episode
from labelanything_b200.training import FlatAdamW, train_step

B, N, K, S = 2, 2, 5, 480
prompts = sys.argv[1] if len(sys.argv) > 1 else "mixed"
lam = build_lam_no_vit(image_embed_dim=1024, embed_dim=256, image_size=S, spatial_convs=3, class_attention=False,
                       example_attention=False, example_class_attention=True, custom_preprocess=False)
load_synth_weights(lam, seed=4)
lam = lam.cuda().train()
ep = {k: v.cuda() for k, v in make_episode(B, N, K, S, seed=400, prompts=prompts, embeddings=(1024, 30)).items()}
gt = torch.randint(0, N + 1, (B, S // 16, S // 16)).repeat_interleave(16, 1).repeat_interleave(16, 2).cuda()
loss_fn = LabelAnythingLoss({"focal": {"weight": 1.0, "gamma": 2.0}}, class_weighting=True)
opt = FlatAdamW(lam.parameters(), lr=5e-5)
for _ in range(3):
    train_step(lam, loss_fn, opt, ep, gt)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    train_step(lam, loss_fn, opt, ep, gt)
e1.record()
torch.cuda.synchronize()
print(f"{prompts}: {e0.elapsed_time(e1) / 5:.2f} ms per step")
with ops.profile() as prof:
    train_step(lam, loss_fn, opt, ep, gt)
    torch.cuda.synchronize()
fam = {}
for name, fl, by, a, b in prof.records:
    f = fam.setdefault(name.split(".")[0] if name.startswith("gemm") else name, [0, 0.0])
    f[0] += 1
    f[1] += a.elapsed_time(b)
tot = sum(v[1] for v in fam.values())
print(f"{prof.launches} launches, {tot:.2f} ms in native kernels")
for k, v in sorted(fam.items(), key=lambda kv: -kv[1][1]):
    print(f"  {k:28s} {v[0]:5d} launches {v[1]:8.3f} ms")
