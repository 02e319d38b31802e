#!/usr/bin/env python
"""Benchmark of the LabelAnything hot path on B200 (BASELINE.json metric: episodes/s, query+support forward,
SAM ViT-B 1024 px, 5-way 5-shot).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl native|reference] [--batch B]

One step = `Lam.forward` over a batch of B synthetic episodes per GPU (each: 1 query + 25 support images at
1024 px, 150 mask prompts) through the `images` key, SAM-512 model (parameters/trainval/other/COCO_vit.yaml:47-63 of
the reference), random-init synthetic weights.  For N > 1 launch with torchrun (one rank per GPU); episodes are
sharded across ranks with no data-path collective (weak scaling), timing is the max over ranks.

Prints ONE JSON line (rank 0).  `value`: inputs resident in HBM.  `e2e`: the same step with the batch in pinned host
memory, H2D copies and the D2H read of the logits inside the timed region.  `roofline`: the kernel family with the
largest share of the step, timed live with CUDA events around every launch.  `cpu_baseline`: the CPU oracle (a
port of the reference's PyTorch forward) timed on this box's host cores on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

N_WAYS, K_SHOTS, IMAGE_SIZE, EMBED_DIM = 5, 5, 1024, 512
METRIC = "episodes/sec (query+support fwd) ViT-B 1024 5-way 5-shot"
SAM512 = dict(image_embed_dim=768, embed_dim=EMBED_DIM, image_size=IMAGE_SIZE, use_vit_sam_neck=False, spatial_convs=3,
              class_attention=False, example_attention=True, example_class_attention=False,
              class_encoder={"name": "RandomMatrixEncoder", "bank_size": 100, "embed_dim": EMBED_DIM},
              custom_preprocess=True)
# algorithmic FLOPs of one episode (26 images x (965.64 + 22.55) + 150 x 10.855 + 28.7 GFLOP)
EPISODE_GFLOP = 26 * (965.64 + 22.55) + 150 * 10.855 + 28.7


# DRAM traffic per launch (MB) of the kernels that can dominate the step, from the committed `ncu --set full` captures
# (profiles/r01_ncu_{attention,gemm,layernorm}_v3.txt) taken at the bench's launch size (one 32-image encoder chunk)
NCU_TRAFFIC_MB = {"attention.L4096": 1600.9, "attention.L196": 904.8, "gemm.n3072.k768": 964.0, "gemm.n768.k3072": 1007.2,
                  "gemm.n768.k768": 361.7, "gemm.n1536.k768": 554.9, "add_layernorm.d768.map0": 1151.0}


def _peaks() -> dict:
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        d["_source"] = "measured"
        return d
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "_source": "fallback"}


class ClockSampler:
    """nvidia-smi clock / throttle-reason sampling during the timed region (B200_PROFILING.md recipe)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int) -> None:
        self.lines: list[str] = []
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _pump(self) -> None:
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
                power.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "power_w_max": max(power),
                "samples": len(sm), "reasons": sorted(reasons)}


def _dist_env():
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    return rank, local, world


def _build_model():
    from labelanything_b200.build_lam import build_lam_vit_b
    from labelanything_b200.synthetic import load_synth_weights

    lam = build_lam_vit_b(**SAM512)
    load_synth_weights(lam, seed=0)
    return lam


# ----------------------------------------------------------------------------------------------------------
# CPU baseline: the oracle (port of the reference forward) on a bounded sample of the workload
# ----------------------------------------------------------------------------------------------------------
def cpu_reference_sample(sd, seed: int = 0, threads: int | None = None) -> dict:
    """Time 1 image through the SAM ViT-B encoder + neck, the prompt encoder on the 6 prompt sequences of one
    support image, and the mask decoder + postprocess of one query — all with the CPU oracle — and extrapolate
    to one 5-way 5-shot episode: 26 images, 150 sequences, 1 decode."""
    import torch

    sys.path.insert(0, str(ROOT / "oracle"))
    import lam_oracle as O  # the timed CPU baseline (bench.py's cpu_baseline / --impl reference legs only)

    from labelanything_b200.synthetic import make_episode

    cores = threads or os.cpu_count() or 1
    torch.set_num_threads(cores)
    cfg = {"image_size": IMAGE_SIZE, "image_embedding_size": (64, 64), "has_neck": True, "spatial_convs": 3,
           "class_attention": False, "example_attention": True, "example_class_attention": False,
           "custom_preprocess": True,
           "encoder": {"kind": "sam", "num_heads": 12, "depth": 12, "global_attn": [2, 5, 8, 11], "window": 14}}
    ep = make_episode(1, N_WAYS, 1, IMAGE_SIZE, seed=seed)   # M = 5 support images generated, 1 used
    C = N_WAYS + 1
    with torch.no_grad():
        t0 = time.perf_counter()
        enc = O.encode_images(sd, cfg, ep["images"][0, :1], chunk=1)
        feat = O.neck(sd, "neck", enc)
        t_img = time.perf_counter() - t0
        support = feat.unsqueeze(0)                           # [1, 1, D, 64, 64]
        masks = (ep["prompt_masks"][:, :1], ep["flag_masks"][:, :1])
        t0 = time.perf_counter()
        pe = O.prompt_encoder(sd, "prompt_encoder", cfg, support, None, None, masks, ep["flag_examples"][:, :1],
                              class_rows=torch.arange(C))
        t_seq = (time.perf_counter() - t0) / C
        gauss = sd["prompt_encoder.pe_layer.positional_encoding_gaussian_matrix"]
        t0 = time.perf_counter()
        low = O.mask_decoder(sd, "mask_decoder", cfg, feat, O.dense_pe(gauss, 64, 64), pe["class_embeddings"])
        O.postprocess_masks(low, ep["dims"][:, :2], IMAGE_SIZE, True)
        t_dec = time.perf_counter() - t0
    M = N_WAYS * K_SHOTS
    t_episode = (M + 1) * t_img + M * C * t_seq + t_dec
    return {"value": 1.0 / t_episode, "unit": "episodes/s", "cores": cores, "kind": "port",
            "sample": (f"CPU oracle (port of the reference PyTorch forward, fp32): 1 image encoder+neck {t_img:.2f}s, "
                       f"{C} prompt sequences {t_seq * C:.2f}s, 1 decode+postprocess {t_dec:.2f}s; extrapolated to "
                       f"26 images + 150 sequences + 1 decode = {t_episode:.1f}s/episode"),
            "seconds_per_episode": t_episode}


def run_reference(args) -> None:
    rank, _, world = _dist_env()
    if rank != 0:
        return
    lam = _build_model()
    sd = {k: v.clone() for k, v in lam.state_dict().items()}
    del lam
    vals = []
    for i in range(args.warmup + args.steps):
        r = cpu_reference_sample(sd, seed=i)
        if i >= args.warmup:
            vals.append(r)
    t_ep = statistics.mean(v["seconds_per_episode"] for v in vals)
    last = vals[-1]
    value = 1.0 / t_ep
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "episodes/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 * t_ep, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "SAM ViT-B 1024px 5-way 5-shot (SAM-512 model), CPU forward, bounded sample "
                                   "extrapolated per episode", "sample": last["sample"]},
            "cpu_baseline": {"value": value, "unit": "episodes/s", "cores": last["cores"], "kind": "port",
                             "sample": last["sample"]},
            "e2e": {"value": value, "unit": "episodes/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------------------
# native arm
# ----------------------------------------------------------------------------------------------------------
def run_native(args) -> None:
    import torch
    import torch.distributed as dist

    from labelanything_b200 import _native, ops
    from labelanything_b200.synthetic import make_episode

    rank, local, world = _dist_env()
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py (native arm) needs a CUDA device; there is no CPU fallback")
    torch.cuda.set_device(local)
    _native.check(_native.lib().la_device_check(), "device_check")
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    B = args.batch
    lam = _build_model()
    sd_cpu = {k: v.clone() for k, v in lam.state_dict().items()} if (rank == 0 and world == 1 and not args.no_cpu) else None
    lam.prompt_encoder.class_encoder.fixed_rows = torch.arange(N_WAYS + 1)
    lam = lam.cuda()
    if args.chunk > 0:
        lam.image_encoder.max_images_per_chunk = args.chunk

    host = make_episode(B, N_WAYS, K_SHOTS, IMAGE_SIZE, seed=100 + rank)
    host = {k: v.pin_memory() for k, v in host.items()}
    dev = {k: v.cuda(non_blocking=True) for k, v in host.items()}
    h2d_bytes = sum(v.numel() * v.element_size() for v in host.values())
    out_host = None

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_resident():
        with torch.no_grad():
            return lam(dev)["logits"]

    # End to end: every step copies its inputs from pinned host memory and reads the logits back.  Uploads run on
    # a copy stream into one of two device input buffers, so the H2D copy of step k+1 overlaps the kernels of
    # step k (the usual double-buffered input pipeline); the D2H read of step k is queued behind its kernels.
    copy_stream = torch.cuda.Stream()
    dev2 = [{k: torch.empty_like(v, device="cuda") for k, v in host.items()} for _ in range(2)]
    uploaded = [torch.cuda.Event() for _ in range(2)]
    consumed = [torch.cuda.Event() for _ in range(2)]

    def upload(slot):
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(consumed[slot])      # the step that last used this buffer has finished reading it
            for k, v in host.items():
                dev2[slot][k].copy_(v, non_blocking=True)
            uploaded[slot].record(copy_stream)

    d2h_stream = torch.cuda.Stream()
    out_ready = torch.cuda.Event()

    def run_e2e(n_steps):
        nonlocal out_host
        cur = torch.cuda.current_stream()
        for sl in range(2):
            consumed[sl].record(cur)
        upload(0)
        for i in range(n_steps):
            slot = i & 1
            if i + 1 < n_steps:
                upload(slot ^ 1)
            cur.wait_event(uploaded[slot])
            with torch.no_grad():
                logits = lam(dev2[slot])["logits"]
            consumed[slot].record(cur)
            if out_host is None:
                out_host = torch.empty(logits.shape, dtype=logits.dtype, pin_memory=True)
            # the logits go back on their own stream, so the read-back of step k overlaps the kernels of step k+1
            out_ready.record(cur)
            with torch.cuda.stream(d2h_stream):
                d2h_stream.wait_event(out_ready)
                out_host.copy_(logits, non_blocking=True)
                logits.record_stream(d2h_stream)
        cur.wait_stream(d2h_stream)   # the timed region ends when the last result is in host memory

    for _ in range(args.warmup):
        step_resident()
    barrier()

    # ---- timed region 1: inputs resident in HBM, every launch bracketed by CUDA events ----------------------
    sampler = ClockSampler(local) if rank == 0 else None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ops.profile() as prof:
        barrier()
        e0.record()
        for _ in range(args.steps):
            step_resident()
        e1.record()
        barrier()
    ms = e0.elapsed_time(e1)
    detail = prof.summary()
    launches = prof.launches
    fam: dict = {}   # kernel families (gemm.n2304.k768 -> gemm)
    for k, v in detail.items():
        f = fam.setdefault(k.split(".")[0], {"launches": 0, "ms": 0.0, "flops": 0.0, "bytes": 0.0})
        for kk in f:
            f[kk] += v[kk]
    clocks = sampler.stop() if sampler else None

    # ---- timed region 2: end to end from pinned host memory ---------------------------------------------------
    run_e2e(2)
    barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    run_e2e(args.steps)
    f1.record()
    barrier()
    ms_e2e = f0.elapsed_time(f1)
    d2h_bytes = out_host.numel() * out_host.element_size()

    t = torch.tensor([ms, ms_e2e], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, ms_e2e = t.tolist()
    episodes = B * world * args.steps
    value = episodes / (ms / 1000.0)
    e2e = episodes / (ms_e2e / 1000.0)

    if rank == 0:
        peaks = _peaks()
        kernel_ms = sum(f["ms"] for f in fam.values())
        # dominant kernel = the launch shape with the largest share of the step (e.g. attention.L4096: the 64x64
        # global-attention instantiation of la_attention_bf16 on one 32-image chunk)
        top = max(detail, key=lambda k: detail[k]["ms"])
        f = detail[top]
        tensor_bound = f["flops"] > 0 and top.split(".")[0] in ("gemm", "gemm_acc", "attention", "conv3x3")
        if tensor_bound:
            achieved, peak, unit = f["flops"] / f["ms"] / 1e9, peaks["bf16_tflops_sustained"], "TFLOP/s"
        else:
            achieved, peak, unit = f["bytes"] / f["ms"] / 1e6, peaks["hbm_gbs"], "GB/s"
        # the captures were taken on 32-image launches; encoder launches scale linearly with the images per chunk
        n_img = B * (N_WAYS * K_SHOTS + 1)
        cap = lam.image_encoder.max_images_per_chunk
        per_chunk = -(-n_img // -(-n_img // cap))
        traffic = NCU_TRAFFIC_MB.get(top)
        if traffic is not None:
            traffic *= per_chunk / 32.0
        roofline = {"kernel": top, "bound": "tensor" if tensor_bound else "hbm", "achieved": achieved, "peak": peak,
                    "unit": unit, "frac": achieved / peak,
                    "traffic": None if traffic is None else traffic * 1e6, "traffic_unit": "bytes per launch",
                    "traffic_source": "ncu --set full dram__bytes_read.sum + dram__bytes_write.sum on a 32-image launch "
                                      "(profiles/r01_ncu_attention_v4.txt, r01_ncu_{gemm,layernorm}_v3.txt), scaled to this run's images per chunk",
                    "algorithmic_per_launch": {"flops": f["flops"] / f["launches"], "bytes": f["bytes"] / f["launches"]},
                    "peak_source": peaks["_source"],
                    "avg_launch_ms": f["ms"] / f["launches"], "share_of_kernel_time": f["ms"] / kernel_ms,
                    "whole_step": {"achieved": EPISODE_GFLOP * episodes / ms, "unit": "TFLOP/s",
                                   "frac": EPISODE_GFLOP * episodes / ms / peaks["bf16_tflops_sustained"]},
                    "families": {k: {"launches": v["launches"] // args.steps, "ms_per_step": v["ms"] / args.steps,
                                     "tflops": (v["flops"] / v["ms"] / 1e9) if v["flops"] else None,
                                     "gbs": (v["bytes"] / v["ms"] / 1e6) if v["bytes"] else None}
                                 for k, v in sorted(fam.items(), key=lambda kv: -kv[1]["ms"])}}
        cpu = None
        if sd_cpu is not None:
            cpu = cpu_reference_sample(sd_cpu)
            cpu.pop("seconds_per_episode", None)
        if args.detail:
            for k, v in sorted(detail.items(), key=lambda kv: -kv[1]["ms"])[:24]:
                print(f"# {k:34s} {v['launches'] // args.steps:5d} launches/step {v['ms'] / args.steps:8.2f} ms/step "
                      f"{(v['flops'] / v['ms'] / 1e9) if v['flops'] else 0:8.1f} TFLOP/s {v['bytes'] / v['ms'] / 1e6:8.0f} GB/s",
                      file=sys.stderr)
        line = {"metric": METRIC, "value": value, "unit": "episodes/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
                "config": {"workload": f"SAM ViT-B 1024px 5-way 5-shot inference, batch={B} episodes per GPU "
                                       f"(SAM-512: embed_dim 512, Lam.forward(images): 26 images + 150 mask-prompt "
                                       f"sequences per episode)", "mode": "A (images)", "episodes_per_step_per_gpu": B,
                           "l2": "inputs larger than L2 (images 12.6 MB each, > 2 GB per step)",
                           "parallelism": f"episode-sharded x{world}, no data-path collective"},
                "e2e": {"value": e2e, "unit": "episodes/s", "h2d_bytes_per_step": h2d_bytes,
                        "d2h_bytes_per_step": d2h_bytes, "ms_per_step": ms_e2e / args.steps,
                        "pipeline": "double-buffered inputs: H2D of step k+1 on a copy stream overlaps step k; D2H of the logits on a third stream overlaps step k+1"},
                "gpu_launches": launches, "roofline": roofline, "cpu_baseline": cpu, "clocks": clocks}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--batch", type=int, default=8, help="episodes per GPU per step")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--chunk", type=int, default=0, help="images per encoder chunk (0 = the model's default)")
    ap.add_argument("--detail", action="store_true", help="print the per-kernel-shape breakdown to stderr")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_native(args)


if __name__ == "__main__":
    main()
