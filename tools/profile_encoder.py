"""GPU helper: run the native SAM ViT-B (1024 px) encoder on a few synthetic images.

  python tools/profile_encoder.py time [n_images]   -> CUDA-event timing, writes gpurun_out/encoder_time.log
  python tools/profile_encoder.py once [n_images]   -> one warm-up + one pass (to be wrapped in ncu)
"""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch

from labelanything_b200.build_encoder import build_vit_b
from labelanything_b200.synthetic import synth_tensor


def main():
    mode = sys.argv[1] if len(sys.argv) > 1 else "time"
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 8
    vit = build_vit_b(project_last_hidden=False)
    sd = vit.state_dict()
    vit.load_state_dict({k: synth_tensor("image_encoder." + k, tuple(v.shape), 0) for k, v in sd.items()})
    vit = vit.cuda()
    g = torch.Generator(device="cuda").manual_seed(0)
    img = torch.randn(n, 3, 1024, 1024, device="cuda", generator=g)
    with torch.no_grad():
        vit.encode_tokens(img)  # warm-up (packs weights)
        torch.cuda.synchronize()
        if mode == "once":
            vit.encode_tokens(img)
            torch.cuda.synchronize()
            return
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        times = []
        for _ in range(5):
            e0.record()
            vit.encode_tokens(img)
            e1.record()
            torch.cuda.synchronize()
            times.append(e0.elapsed_time(e1))
    ms = min(times)
    flops = 965.6e9 * n
    line = (f"SAM ViT-B 1024 encoder, {n} images: best {ms:.2f} ms ({ms / n:.3f} ms/img), "
            f"{flops / ms / 1e9:.1f} TFLOP/s algorithmic; all runs {['%.2f' % t for t in times]}")
    print(line)
    out = ROOT / "gpurun_out"
    out.mkdir(exist_ok=True)
    (out / "encoder_time.log").write_text(line + "\n")


if __name__ == "__main__":
    main()
