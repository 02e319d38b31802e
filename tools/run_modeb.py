"""One `Lam.forward(embeddings)` of the SAM-512 model (neck, prompt encoder, mask decoder, postprocess) after a warm-up,
for `ncu -k ...` captures of the kernels outside the ViT: python tools/run_modeb.py [episodes] [mask|mixed]"""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch

import bench
from labelanything_b200.synthetic import make_episode

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
prompts = sys.argv[2] if len(sys.argv) > 2 else "mask"
lam = bench._build_model()
lam.prompt_encoder.class_encoder.fixed_rows = torch.arange(bench.N_WAYS + 1)
lam = lam.cuda()
ep = {k: v.cuda() for k, v in make_episode(B, bench.N_WAYS, bench.K_SHOTS, bench.IMAGE_SIZE, seed=7, prompts=prompts,
                                           embeddings=(768, 64)).items()}
with torch.no_grad():
    for _ in range(2):
        out = lam(ep)["logits"]
torch.cuda.synchronize()
print("done", tuple(out.shape))
