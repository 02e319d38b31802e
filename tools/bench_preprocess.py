"""Row f3 throughput: GPU preprocessing of COCO-sized uint8 images vs the reference's host pipeline (PIL resize + ToTensor
+ normalise + pad, one image at a time like the data loader workers).  python tools/bench_preprocess.py"""
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import numpy as np
import torch

from labelanything_b200.transforms import ImagePreprocessor

rng = np.random.default_rng(0)
imgs = [rng.integers(0, 256, (480, 640, 3), dtype=np.uint8) for _ in range(26)]      # one 5-way 5-shot episode
pre = ImagePreprocessor(1024)
dev_imgs = [torch.from_numpy(i).cuda() for i in imgs]
for _ in range(3):
    pre(dev_imgs)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
e0.record()
for _ in range(10):
    out = pre(dev_imgs)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
pinned = [torch.from_numpy(i).pin_memory() for i in imgs]
t0 = time.perf_counter()
for _ in range(10):
    out = pre(pinned)
torch.cuda.synchronize()
ms_h2d = (time.perf_counter() - t0) * 100
algo = sum(i.size for i in imgs) + out.numel() * 4
print(f"GPU: 26 images 480x640 -> 3x1024x1024 fp32: {ms:.3f} ms per episode resident ({algo / ms / 1e6:.0f} GB/s algorithmic), "
      f"{ms_h2d:.3f} ms incl. H2D from pinned memory = {1e3 / ms_h2d:.0f} episodes/s")
try:
    from PIL import Image
    from torchvision.transforms import Compose, ToTensor
    from torchvision.transforms.functional import resize

    mean, std = torch.tensor([0.485, 0.456, 0.406]).view(3, 1, 1), torch.tensor([0.229, 0.224, 0.225]).view(3, 1, 1)
    pil = [Image.fromarray(i) for i in imgs]
    t0 = time.perf_counter()
    for im in pil:
        x = ToTensor()(resize(im, (768, 1024)))
        x = torch.nn.functional.pad((x - mean) / std, (0, 0, 0, 256))
    cpu_ms = (time.perf_counter() - t0) * 1e3
    print(f"reference host pipeline (PIL + torch, 1 thread): {cpu_ms:.1f} ms per episode = {1e3 / cpu_ms:.1f} episodes/s")
except ImportError as e:
    print("PIL / torchvision not importable:", e)
