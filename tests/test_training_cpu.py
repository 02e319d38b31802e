"""Row f1 (training step) without a GPU: the gradient oracle is pinned to the autograd gradients of the UNMODIFIED
reference (tests/golden/train_f1.pt, oracle/make_golden.py train), and the host logic of the flat-bucket optimiser
(run formation, one all-reduce of [gradients | used counters] over two gloo ranks) is exercised on CPU tensors."""
import os
import sys
from pathlib import Path

import pytest
import torch
import torch.nn.functional as F

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "oracle"))

import lam_oracle as O  # noqa: E402  (checker only)
import loss_oracle as LO  # noqa: E402

from labelanything_b200.build_lam import build_lam_no_vit  # noqa: E402
from labelanything_b200.synthetic import load_synth_weights  # noqa: E402
from labelanything_b200.training import FlatAdamW  # noqa: E402

GOLD = ROOT / "tests" / "golden" / "train_f1.pt"


def sample_index(numel, n=4096):          # oracle/make_golden.py::sample_index
    return torch.arange(numel) if numel <= n else torch.linspace(0, numel - 1, n).long()


def scale_matrices(module, gain):          # oracle/make_golden.py::scale_matrices
    if gain != 1.0:
        with torch.no_grad():
            for p in module.parameters():
                if p.dim() >= 2 and p.shape[0] > 1:
                    p.mul_(gain)


def reference_loss(logits, gt):
    """LabelAnythingLoss({"focal": {"weight": 1, "gamma": 2}}, class_weighting=True) restated with torch ops
    (loss/__init__.py:67-92, loss/focal.py:17-25); the weight map comes from the pinned numpy oracle."""
    wm, _ = LO.get_weight_matrix_from_labels(gt.numpy().copy(), logits.shape[1])
    ce = F.cross_entropy(logits, gt, reduction="none")
    pt = torch.exp(-ce)
    return (torch.pow(1 - pt, 2.0) * torch.from_numpy(wm) * ce).mean()


def build_case(case):
    lam = build_lam_no_vit(**case["build"])
    load_synth_weights(lam, seed=case["weights_seed"])
    scale_matrices(lam, case.get("weight_gain", 1.0))
    if case["class_rows"] is not None:
        lam.prompt_encoder.class_encoder.fixed_rows = case["class_rows"]
    return lam


@pytest.mark.parametrize("name", ["mixed", "masks_only", "mixed_scaled"])
def test_oracle_gradients_match_the_reference_autograd(name):
    case = torch.load(GOLD, weights_only=False)["cases"][name]
    lam = build_case(case)
    sd = {k: v.detach().clone().requires_grad_(v.is_floating_point()) for k, v in lam.state_dict().items()}
    out = O.lam_forward(sd, case["cfg"], dict(case["episode"]), case["class_rows"])
    ref = case["logits"]
    fin = torch.isfinite(ref)
    assert torch.equal(torch.isfinite(out["logits"]), fin)
    assert (out["logits"][fin] - ref[fin]).abs().max() < 5e-5 * max(1.0, float(ref[fin].abs().max()))
    loss = reference_loss(out["logits"], case["gt"])
    assert abs(float(loss) - float(case["loss"])) < 1e-5 * max(1.0, abs(float(case["loss"])))
    loss.backward()
    names = dict(lam.named_parameters()).keys()
    checked = 0
    for k in names:
        want = case["grads"][k]
        got = sd[k].grad
        if want is None:
            assert got is None or float(got.abs().max()) == 0.0, k
            continue
        g = got.reshape(-1)
        assert abs(float(g.double().norm()) - float(want["norm"])) <= 2e-3 * float(want["norm"]) + 1e-7, k
        v = g[sample_index(g.numel())]
        assert (v - want["values"]).abs().max() <= 2e-3 * float(want["values"].abs().max()) + 1e-7, k
        checked += 1
    assert checked > 200


def test_used_runs_group_consecutive_parameters_with_equal_step_counts():
    opt = object.__new__(FlatAdamW)
    opt.params = [None] * 6
    opt.offsets = [0, 8, 12, 40, 44, 100]
    opt.numel = 120
    opt.steps = [3, 3, 3, 1, 3, 3]
    assert opt.used_runs([True] * 6) == [(0, 40, 3), (40, 44, 1), (44, 120, 3)]
    assert opt.used_runs([True, False, True, True, False, True]) == [(0, 8, 3), (12, 40, 3), (40, 44, 1), (100, 120, 3)]
    assert opt.used_runs([False] * 6) == []


def _rank(rank, world, port, q):
    import torch.distributed as dist

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    params = [torch.nn.Parameter(torch.randn(5)), torch.nn.Parameter(torch.randn(3, 2)), torch.nn.Parameter(torch.randn(7))]
    opt = FlatAdamW(params)
    opt.zero_grad()
    # rank 0 uses parameters 0 and 1, rank 1 only parameter 1; parameter 2 is unused everywhere
    loss = (params[1] * (rank + 1)).sum() + (params[0].sum() * 2 if rank == 0 else 0)
    loss.backward()
    used, w = opt.reduce_gradients()
    q.put((rank, used, w, params[0].grad.clone(), params[1].grad.clone(), params[2].grad.clone(),
           params[0].grad.data_ptr() == opt.flat_g.data_ptr()))
    dist.destroy_process_group()


def test_two_rank_gloo_gradient_bucket_allreduce():
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_rank, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    for rank, used, w, g0, g1, g2, in_bucket in res:
        assert w == 2 and used == [True, True, False]          # used on ANY rank
        assert torch.allclose(g0, torch.full((5,), 2.0))       # summed over ranks (the update divides by the world size)
        assert torch.allclose(g1, torch.full((3, 2), 3.0))
        assert float(g2.abs().max()) == 0.0
        assert in_bucket                                       # autograd accumulated straight into the flat bucket


def test_checkpoints_can_be_written_after_the_parameters_moved_into_the_flat_bucket(tmp_path):
    """huggingface_hub's save_pretrained / accelerate's save_state go through safetensors.torch.save_model, which refuses
    views of a larger storage: FlatAdamW.attach makes state_dict() hand out copies."""
    from safetensors.torch import load_file, save_model

    lam = build_lam_no_vit(image_embed_dim=64, embed_dim=128, image_size=128, spatial_convs=3)
    opt = FlatAdamW(lam.parameters())
    with pytest.raises(RuntimeError):
        save_model(lam, str(tmp_path / "views.safetensors"))
    opt.attach(lam)
    opt.attach(lam)                                   # idempotent
    save_model(lam, str(tmp_path / "m.safetensors"))
    back = load_file(str(tmp_path / "m.safetensors"))
    sd = lam.state_dict()
    assert set(back) == set(sd) and all(torch.equal(back[k], v) for k, v in sd.items())
    lam.load_state_dict(back)                         # loads in place: the parameters stay views of the bucket
    p0 = opt.params[0]
    assert p0.data_ptr() == opt.flat_p.data_ptr() + 4 * opt.offsets[0]


def test_derived_operand_cache_follows_tensor_identity_version_and_lifetime():
    """train_ops._derived: operands derived from a tensor (bf16 splits, transposes) are shared between the ops that
    consume the same tensor, recomputed after an in-place update, and dropped when the source dies (so an address the
    allocator hands out again cannot hit a stale entry)."""
    import gc

    from labelanything_b200 import train_ops as T

    T._DERIVED.clear()
    calls = []

    def build(tag):
        calls.append(tag)
        return object()

    x = torch.zeros(4, 8)
    a = T._derived("t", x, lambda: build("a"))
    assert T._derived("t", x, lambda: build("a2")) is a and calls == ["a"]          # shared
    assert T._derived("other", x, lambda: build("b")) is not a                       # another derivation of the same tensor
    x.add_(1)                                                                        # in-place update: new version
    assert T._derived("t", x, lambda: build("c")) is not a and calls == ["a", "b", "c"]
    n = len(T._DERIVED)
    del x
    gc.collect()
    assert len(T._DERIVED) < n and all(ref() is not None for ref, _ in T._DERIVED.values())
    with T.precision("bf16x3"):
        assert T._PRECISION == "bf16x3"
        with T.precision("bf16x6"):
            assert T._PRECISION == "bf16x6"
        assert T._PRECISION == "bf16x3"
    assert T._PRECISION == "bf16"


def test_flat_adamw_notices_parameters_that_left_the_bucket():
    lin = torch.nn.Linear(4, 3)
    opt = FlatAdamW(lin.parameters())
    opt.zero_grad()
    lin.weight.data = lin.weight.data.clone()            # what model.to(other_device) does to a parameter
    with pytest.raises(RuntimeError, match="flat bucket"):
        opt.zero_grad()


def test_constant_with_warmup_matches_the_transformers_schedule():
    """lr_lambda of transformers.get_constant_schedule_with_warmup: step / max(1, warmup) below the warm-up, then 1."""
    from transformers import get_constant_schedule_with_warmup

    from labelanything_b200.training import ConstantWithWarmup

    lin = torch.nn.Linear(4, 3)
    ref_opt = torch.optim.AdamW(lin.parameters(), lr=5e-5)
    ref = get_constant_schedule_with_warmup(ref_opt, num_warmup_steps=4)
    opt = FlatAdamW(torch.nn.Linear(4, 3).parameters(), lr=5e-5)
    sched = ConstantWithWarmup(opt, 4)
    for _ in range(7):
        assert abs(sched.get_last_lr()[0] - ref.get_last_lr()[0]) < 1e-12
        ref_opt.step()
        ref.step()
        sched.step()
