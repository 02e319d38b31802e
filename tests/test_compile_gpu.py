"""The native ops as torch custom ops: opcheck of schema / fake registration on real launches, `torch.compile(model)`
(the reference compiles its model on request, label_anything/experiment/run.py:167-169) must give the eager result bit
for bit, the launches must follow the TENSOR's device, and the model must construct and forward under
DistributedDataParallel(find_unused_parameters=True) (run.py:122-131, experiment/utils.py:266-288)."""
import os
import subprocess
import sys
from pathlib import Path

import pytest
import torch

ROOT = Path(__file__).resolve().parent.parent
pytestmark = pytest.mark.gpu


def _small_lam():
    from labelanything_b200.build_lam import build_lam_no_vit
    from labelanything_b200.synthetic import load_synth_weights, make_episode

    lam = build_lam_no_vit(image_embed_dim=384, embed_dim=256, image_size=256, spatial_convs=3, custom_preprocess=False)
    load_synth_weights(lam, seed=1)
    ep = make_episode(2, 2, 1, 256, seed=2, prompts="mixed", embeddings=(384, 16))
    return lam, ep


def test_opcheck_schema_and_fake_registration():
    from labelanything_b200 import ops

    g = torch.Generator(device="cuda").manual_seed(0)
    a = torch.randn(300, 128, device="cuda", generator=g).to(torch.bfloat16)
    w = torch.randn(256, 128, device="cuda", generator=g).to(torch.bfloat16)
    b = torch.randn(256, device="cuda", generator=g)
    out = torch.empty(300, 256, device="cuda")
    torch.library.opcheck(torch.ops.labelanything_b200.la_gemm_bf16.default,
                          (a, a.stride(0), w, w.stride(0), b, out, out.stride(0), ops.DT_F32, 300, 256, 128, ops.ACT_NONE),
                          test_utils=("test_schema", "test_faketensor"))
    x = torch.randn(64, 256, device="cuda", generator=g)
    y = torch.empty(64, 256, device="cuda", dtype=torch.bfloat16)
    gm = torch.randn(256, device="cuda", generator=g)
    torch.library.opcheck(torch.ops.labelanything_b200.la_add_layernorm.default,
                          (x, 0, None, None, None, 0, None, gm, gm, 1e-6, 0, y, ops.DT_BF16, None, None, 0, None, 64, 256,
                           0, 0, 0, 0, 0), test_utils=("test_schema", "test_faketensor"))
    # the op called through the dispatcher == the wrapper's direct launch (opcheck itself works on clones)
    torch.ops.labelanything_b200.la_gemm_bf16(a, a.stride(0), w, w.stride(0), b, out, out.stride(0), ops.DT_F32, 300, 256,
                                               128, ops.ACT_NONE)
    ref = ops.gemm(a, w, b, out_dtype=torch.float32)
    assert torch.equal(ref, out)


@pytest.mark.parametrize("backend", ["aot_eager", "inductor"])
def test_torch_compile_equals_eager_bit_for_bit(backend):
    lam, ep = _small_lam()
    lam = lam.cuda()
    ep = {k: v.cuda() for k, v in ep.items()}
    with torch.no_grad():
        eager = lam(ep)
        compiled = torch.compile(lam, backend=backend)
        out = compiled(ep)
        out2 = compiled(ep)      # second call: cached graphs
    for k in ("logits", "class_examples_embeddings"):
        assert torch.equal(eager[k], out[k]), f"{backend}: {k} differs from eager"
        assert torch.equal(eager[k], out2[k])


def test_launches_follow_the_tensor_device_not_the_current_one():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    lam, ep = _small_lam()
    with torch.no_grad():
        torch.cuda.set_device(0)
        ref = lam.cuda(0)({k: v.cuda(0) for k, v in ep.items()})["logits"]
        out = lam.to("cuda:1")({k: v.to("cuda:1") for k, v in ep.items()})["logits"]   # current device stays 0
        assert out.device.index == 1 and torch.cuda.current_device() == 0
    assert torch.equal(ref.cpu(), out.cpu())


_DDP_CHILD = r"""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, os.environ["LA_ROOT"])
from labelanything_b200.build_lam import build_lam_no_vit
from labelanything_b200.loss import LabelAnythingLoss
from labelanything_b200.synthetic import load_synth_weights, make_episode

class WrapperModule(torch.nn.Module):          # shape of label_anything/experiment/utils.py:266-288
    def __init__(self, model, loss):
        super().__init__()
        self.model, self.loss = model, loss
    def forward(self, input_dict, gt):
        result_dict = self.model(input_dict)
        loss = self.loss(result_dict["logits"], gt)
        return {"loss": loss, **result_dict}

rank = int(os.environ["RANK"])
dist.init_process_group("gloo")               # two ranks on ONE GPU: NCCL refuses duplicate devices, gloo stages via host
torch.cuda.set_device(0)
lam = build_lam_no_vit(image_embed_dim=384, embed_dim=256, image_size=256, spatial_convs=3, custom_preprocess=False)
load_synth_weights(lam, seed=1 + rank)        # different weights per rank: DDP must broadcast rank 0's
wm = WrapperModule(lam, LabelAnythingLoss({"focal": {"weight": 1.0, "gamma": 2.0}}, class_weighting=True)).cuda()
ddp = torch.nn.parallel.DistributedDataParallel(wm, device_ids=[0], find_unused_parameters=True)
ep = {k: v.cuda() for k, v in make_episode(1, 2, 1, 256, seed=5, prompts="mixed", embeddings=(384, 16)).items()}
gt = torch.randint(0, 3, (1, 256, 256), generator=torch.Generator().manual_seed(3)).cuda()
with torch.no_grad():
    out = ddp(ep, gt)
vals = [torch.zeros(2) for _ in range(2)]
dist.all_gather(vals, torch.stack([out["loss"]["value"].float().cpu(), out["logits"].float().cpu().abs().mean()]))
assert torch.equal(vals[0], vals[1]), vals     # same weights (broadcast), same inputs -> same result on both ranks
assert torch.isfinite(vals[0]).all()
if rank == 0:
    print("DDP_OK", vals[0].tolist())
dist.destroy_process_group()
"""


def test_ddp_wrapper_constructs_and_forwards_two_ranks(tmp_path):
    script = tmp_path / "ddp_child.py"
    script.write_text(_DDP_CHILD)
    env = dict(os.environ, LA_ROOT=str(ROOT), MASTER_ADDR="127.0.0.1", MASTER_PORT="29731", WORLD_SIZE="2")
    procs = [subprocess.Popen([sys.executable, str(script)], env=dict(env, RANK=str(r)), stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=300)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), "\n".join(outs)
    assert "DDP_OK" in outs[0]
