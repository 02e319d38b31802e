"""HBM roofline of the fused focal loss at the BASELINE size, next to the torch ops of the
reference's LabelAnythingLoss for the same step: python tools/bench_loss.py"""
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch
import torch.nn.functional as F

from labelanything_b200 import ops

B, C, H, W = 8, 6, 1024, 1024
x = torch.randn(B, C, H, W, device="cuda") * 2
t = torch.randint(0, C, (B, H, W), device="cuda")
t[:, :, :16] = -100
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")     # > L2
one = torch.ones((), device="cuda")


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


def ref_weights(labels, num_classes, ignore_index=-100):      # loss/utils.py:17-42 with torch ops
    wl = labels.clone()
    wl += 1
    wl[wl == ignore_index + 1] = 0
    weights = torch.ones(num_classes + 1, device=labels.device)
    classes, counts = wl.unique(return_counts=True)
    weights[classes] = 1 / torch.log(1.1 + counts / counts.sum())
    weights[0] = 0
    return weights[wl]


def ref_forward(xx):
    wm = ref_weights(t, C)
    ce = F.cross_entropy(xx, t, reduction="none")
    pt = torch.exp(-ce)
    return torch.mean(torch.pow(1 - pt, 2.0) * wm * ce)


def ref_fwd_bwd():
    xx = x.detach().requires_grad_(True)
    ref_forward(xx).backward()


px = B * H * W
cw, _ = ops.label_class_weights(t, C)
res = {
    "class weights (label histogram)": (timed(lambda: ops.label_class_weights(t, C)), px * 8),
    "loss value": (timed(lambda: ops.focal_loss(x, t, cw, 2.0)), px * (4 * C + 8)),
    "gradient w.r.t. logits": (timed(lambda: ops.focal_loss(x, t, cw, 2.0, want_loss=False, want_grad=True, grad_scale=one)),
                               px * (8 * C + 8)),
}
out = {k: {"ms": ms, "GB/s": nb / ms / 1e6, "algorithmic_bytes": nb} for k, (ms, nb) in res.items()}
peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text()) if (ROOT / "MEASURED_PEAKS.json").exists() else {}
print(json.dumps({"workload": f"logits {B}x{C}x{H}x{W} fp32, target int64, gamma 2, class weighting", "kernels": out,
                  "torch_ops_of_the_reference_forward_ms": timed(lambda: ref_forward(x), 5),
                  "torch_ops_of_the_reference_forward_backward_ms": timed(ref_fwd_bwd, 5),
                  "hbm_peak_GB/s": peaks.get("hbm_gbs")}, indent=1))
