"""HBM roofline of the post-logits kernel at the BASELINE size, next to the torch ops the
reference runs for the same step: python tools/bench_metrics.py"""
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch

from labelanything_b200 import ops
from labelanything_b200.metrics import chain_label_map

B, C, H, W, G = 8, 6, 1024, 1024, 81
logits = torch.randn(B, C, H, W, device="cuda")
gt = torch.randint(0, C, (B, H, W), device="cuda")
gt[:, :, :16] = -100
classes = [[[3 + 5 * b, 10 + b, 20 + b, 40 + b, 60 + b]] for b in range(B)]
categories = {k: {} for k in range(1, G)}
table_host = chain_label_map(classes, categories, map_len=C)
table = table_host.cuda()
values = table_host.tolist()
conf = torch.zeros(G, G, dtype=torch.int64, device="cuda")
bad = torch.zeros(1, dtype=torch.int64, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")     # > L2


def timed(fn, iters=20):
    for _ in range(3):
        fn()
    ms = 0.0
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ms += e0.elapsed_time(e1)
    return ms / iters


def reference_ops():
    preds = logits.argmax(dim=1)
    outs = [preds.clone(), gt.clone()]
    for i in range(B):                                    # to_global_multiclass, data/utils.py:583-589
        for j in range(C - 1):
            for t in outs:
                t[i] = torch.where(t[i] == j + 1, values[i][j + 1], t[i])
    keep = outs[1] != -100
    return torch.bincount(outs[1][keep] * G + outs[0][keep], minlength=G * G)


px = B * H * W
res = {}
# segmentation-like labels: 32x32-pixel blocks of one class, predictions agreeing with them on ~90 % of the blocks
blocks = torch.randint(0, C, (B, H // 32, W // 32), device="cuda")
gt_seg = blocks.repeat_interleave(32, 1).repeat_interleave(32, 2).contiguous()
pred_blocks = torch.where(torch.rand(blocks.shape, device="cuda") < 0.9, blocks, torch.randint_like(blocks, C))
logits_seg = (torch.randn(B, C, H, W, device="cuda") * 0.1
              + 4.0 * torch.nn.functional.one_hot(pred_blocks.repeat_interleave(32, 1).repeat_interleave(32, 2), C).permute(0, 3, 1, 2)).contiguous()
ms = timed(lambda: ops.label_confusion(logits_seg, None, gt_seg, table, conf, bad, want_preds=True, want_gt=True))
res["segmentation-like labels, confmat + global label maps"] = {"ms": ms, "GB/s": px * (4 * C + 24) / ms / 1e6,
                                                                 "algorithmic_bytes": px * (4 * C + 24)}
for name, wp, wg in (("random labels, confmat only", False, False), ("random labels, confmat + global label maps", True, True)):
    ms = timed(lambda: ops.label_confusion(logits, None, gt, table, conf, bad, want_preds=wp, want_gt=wg))
    nbytes = px * (4 * C + 8 + (16 if wp else 0))
    res[name] = {"ms": ms, "GB/s": nbytes / ms / 1e6, "algorithmic_bytes": nbytes}
ms_ref = timed(reference_ops, iters=5)
peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text()) if (ROOT / "MEASURED_PEAKS.json").exists() else {}
print(json.dumps({"workload": f"logits {B}x{C}x{H}x{W} fp32, gt int64, {G} global classes", "kernel": res,
                  "torch_ops_of_the_reference_ms": ms_ref, "hbm_peak_GB/s": peaks.get("hbm_gbs")}, indent=1))
