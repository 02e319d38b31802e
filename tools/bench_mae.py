"""Secondary measurement (BASELINE.json configs[1]): MAE-256 (HF ViT-B, 480 px), 1-way 1-shot, batch 32 on one B200.

    python tools/bench_mae.py [--batch 32] [--steps 10]

Model per parameters/trainval/coco20i/mae_noembs.yaml:42-55 of the reference, random-init synthetic
weights, synthetic episodes, inputs resident in HBM, CUDA events.  Not the bench.py metric (that is configs[2])."""
import argparse
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch

from labelanything_b200.build_encoder import build_vit_from_config
from labelanything_b200.build_lam import build_lam
from labelanything_b200.synthetic import load_synth_weights, make_episode


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--steps", type=int, default=10)
    args = ap.parse_args()
    # random-init HF ViT-B (no checkpoint download: there is no network)
    lam = build_lam(build_vit=lambda project_last_hidden: build_vit_from_config(), image_embed_dim=768, embed_dim=256, image_size=480, spatial_convs=3, class_attention=False,
                    example_attention=False, example_class_attention=True,
                    class_encoder={"name": "RandomMatrixEncoder", "bank_size": 100, "embed_dim": 256},
                    custom_preprocess=False)
    load_synth_weights(lam, seed=0)
    lam.prompt_encoder.class_encoder.fixed_rows = torch.arange(2)
    lam = lam.cuda()
    ep = {k: v.cuda() for k, v in make_episode(args.batch, 1, 1, 480, seed=3).items()}
    with torch.no_grad():
        for _ in range(3):
            lam(ep)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            out = lam(ep)["logits"]
        e1.record()
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    gflop = 373.8 * args.batch   # 373.8 GFLOP per MAE-256 1-way 1-shot episode
    print(json.dumps({"config": "MAE-256 (HF ViT-B 480 px) 1-way 1-shot", "batch": args.batch, "ms_per_step": ms,
                      "episodes_per_s": args.batch / ms * 1e3, "tflops": gflop / ms,
                      "logits": list(out.shape)}))


if __name__ == "__main__":
    main()
