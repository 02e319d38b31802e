#!/usr/bin/env python
"""Benchmark of the LabelAnything hot path on B200 (BASELINE.json metric: episodes/s, query+support forward,
SAM ViT-B 1024 px, 5-way 5-shot).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl native|reference] [--batch B]

One step = `Lam.forward` over a batch of B synthetic episodes per GPU (each: 1 query + 25 support images at
1024 px, 150 mask prompts) through the `images` key, SAM-512 model (parameters/trainval/other/COCO_vit.yaml:47-63 of
the reference), random-init synthetic weights.  For N > 1 launch with torchrun (one rank per GPU); episodes are
sharded across ranks with no data-path collective (weak scaling), timing is the max over ranks.

Prints ONE JSON line (rank 0).  `value`: inputs resident in HBM, K steps timed with ONE pair of CUDA events (no
per-launch instrumentation).  `e2e`: the same step with the batch in pinned host memory, H2D copies and the D2H read
of the logits inside the timed region.  `roofline`: a separate, instrumented pass (every native launch bracketed by
CUDA events on the launching stream): the launch shape with the largest share of the step.  `secondary` (N = 1): the
other entry modes / BASELINE.json configs, each labelled (mode B `embeddings`, mode C `generate_class_embeddings` +
`predict`, config 2 MAE-256 batch 32, config 5 20-way 5-shot).  `parity`: drift of the logits against the committed
golden tensors of the unmodified reference.  `cpu_baseline` / `--impl reference`: the UNMODIFIED reference installed
under baseline/_ref (baseline/reference_arm.py), timed on this box's host cores on a bounded sample of the same
workload and scaled to episodes/s; the CPU oracle (kind "port") only if baseline/_ref is absent.
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
NCU_TRAFFIC_SOURCE = ("ncu --set full dram__bytes_read.sum + dram__bytes_write.sum on a 32-image launch "
                      "(profiles/r02_ncu_attention_global_v2.txt: q/k/v/out 805 MB + the fp16 rel-pos table 805 MB; "
                      "r01_ncu_attention_window_v2.txt, r01_ncu_{gemm,layernorm}_v3.txt), scaled to this run's images per chunk")
NCU_TRAFFIC_MB = {"attention.L4096": 1608.5, "attention.L196": 904.8, "gemm.n3072.k768": 964.0, "gemm.n768.k3072": 1007.2,
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
# CPU baseline / reference arm
# ----------------------------------------------------------------------------------------------------------
def _config(B: int, world: int) -> dict:
    """The workload description shared by both arms (the driver compares them)."""
    return {"workload": f"SAM ViT-B 1024px 5-way 5-shot inference, batch={B} episodes per GPU "
                        f"(SAM-512: embed_dim 512, Lam.forward(images): 26 images + 150 mask-prompt "
                        f"sequences per episode)", "mode": "A (images)", "episodes_per_step_per_gpu": B,
            "l2": "inputs larger than L2 (images 12.6 MB each, > 2 GB per step)",
            "parallelism": f"episode-sharded x{world}, no data-path collective"}


class ReferenceCpu:
    """The reference's own CPU forward on a bounded sample of the workload (baseline/reference_arm.py); falls back
    to the CPU oracle (a port, kind "port") only when baseline/_ref is not there."""

    def __init__(self) -> None:
        from baseline import reference_arm as R

        self.kind = "reference" if R.available() else "port"
        self.cores = os.cpu_count() or 1
        if self.kind == "reference":
            lam = R.build_reference_lam("build_lam_vit_b", SAM512, seed=0, n_classes=N_WAYS + 1)
            # a step = encoder on 1 image + the rest of a 5-way 1-SHOT sub-episode (6 images, 30 prompt sequences)
            self.sampler = R.SamEpisodeSampler(lam, N_WAYS, 1, IMAGE_SIZE, threads=self.cores)
            self.full = R.SamEpisodeSampler(lam, N_WAYS, K_SHOTS, IMAGE_SIZE, threads=self.cores)
        else:
            self.sd = {k: v.clone() for k, v in _build_model().state_dict().items()}

    def step(self, seed: int = 0) -> dict:
        """-> {"t_step": seconds actually spent, "t_episode": seconds per whole episode it scales to, "sample": str}"""
        if self.kind == "port":
            r = cpu_port_sample(self.sd, seed=seed, threads=self.cores)
            return {"t_step": r["t_step"], "t_episode": r["seconds_per_episode"], "sample": r["sample"]}
        r = self.sampler.step(seed, n_img=1)
        M, C = N_WAYS * K_SHOTS, N_WAYS + 1
        # encoder: per image; forward(embeddings): per prompt sequence (99 % of it is the per-sequence two-way
        # transformer) -- the sub-episode has M/K_SHOTS support images, i.e. 1/K_SHOTS of the sequences
        t_episode = (M + 1) * r["t_enc"] + K_SHOTS * r["t_rest"]
        sample = (f"unmodified reference (baseline/_ref), fp32, {self.cores} threads: image_encoder on 1 image "
                  f"{r['t_enc']:.2f}s + Lam.forward(embeddings) on a {N_WAYS}-way 1-shot sub-episode ({N_WAYS * C} of the "
                  f"{M * C} prompt sequences, neck, decode, postprocess) {r['t_rest']:.2f}s; scaled to {M + 1} images + "
                  f"{K_SHOTS} x the sub-episode = {t_episode:.1f}s/episode")
        return {"t_step": r["t_step"], "t_episode": t_episode, "sample": sample}

    def whole_episode(self, seed: int = 0) -> dict:
        r = self.full.full_episode(seed)
        return {"t_step": r["t_step"], "t_episode": r["t_episode"], "sample": self.full.describe(r)}


def cpu_port_sample(sd, seed: int = 0, threads: int | None = None) -> dict:
    """Fallback when baseline/_ref is absent: 1 image through the SAM ViT-B encoder + neck, the prompt encoder on the 6
    prompt sequences of one support image, and the mask decoder + postprocess of one query -- all with the CPU
    oracle -- scaled to one 5-way 5-shot episode: 26 images, 150 sequences, 1 decode."""
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
    return {"t_step": t_img + t_seq * C + t_dec, "seconds_per_episode": t_episode,
            "sample": (f"CPU oracle (port of the reference PyTorch forward, fp32, {cores} threads; baseline/_ref absent): "
                       f"1 image encoder+neck {t_img:.2f}s, {C} prompt sequences {t_seq * C:.2f}s, 1 decode+postprocess "
                       f"{t_dec:.2f}s; scaled to 26 images + 150 sequences + 1 decode = {t_episode:.1f}s/episode")}


def run_reference(args) -> None:
    """`--impl reference`: K timed steps (after W warm-ups) of the reference's own CPU implementation, each a bounded
    sample of the 5-way 5-shot episode; `ms_per_step` is the time actually spent per step, `value` the episodes/s it
    scales to (`episodes_per_step` = the fraction of an episode one step amounts to).  `--full-episode`: one whole
    episode, nothing scaled."""
    rank, _, world = _dist_env()
    if rank != 0:
        return
    ref = ReferenceCpu()
    rs = []
    if args.full_episode:
        ref.step(0)                                   # warm-up: one image + sub-episode
        rs = [ref.whole_episode(seed=1)]
        steps, warmup = 1, 1
    else:
        for i in range(args.warmup + args.steps):
            r = ref.step(seed=i)
            if i >= args.warmup:
                rs.append(r)
        steps, warmup = args.steps, args.warmup
    t_step = statistics.mean(r["t_step"] for r in rs)
    t_ep = statistics.mean(r["t_episode"] for r in rs)
    value = 1.0 / t_ep
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "episodes/s", "n_gpus": args.gpus,
            "steps": steps, "warmup": warmup, "ms_per_step": 1000.0 * t_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": _config(args.batch, args.gpus), "episodes_per_step": t_step / t_ep,
            "cpu_baseline": {"value": value, "unit": "episodes/s", "cores": ref.cores, "kind": ref.kind,
                             "sample": rs[-1]["sample"]},
            "e2e": {"value": value, "unit": "episodes/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------------------
# native arm
# ----------------------------------------------------------------------------------------------------------
def _timed(fn, steps: int, warmup: int = 2) -> float:
    """ms per call of fn() on the current stream (CUDA events, after warm-up)."""
    import torch

    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def secondary_measurements(lam, steps: int = 3) -> dict:
    """The other entry modes and BASELINE.json configs, one GPU, inputs resident in HBM, each
    labelled with its own workload and algorithmic FLOPs.  `lam` is the SAM-512 model of the headline."""
    import torch

    from labelanything_b200.build_encoder import build_vit_from_config
    from labelanything_b200.build_lam import build_lam
    from labelanything_b200.synthetic import load_synth_weights, make_episode

    out = {}
    cuda = lambda ep: {k: v.cuda() for k, v in ep.items()}   # noqa: E731
    with torch.no_grad():
        # ---- mode B: precomputed encoder embeddings through the `embeddings` key (lam.py:139-146) ----
        B = 8
        ep = cuda(make_episode(B, N_WAYS, K_SHOTS, IMAGE_SIZE, seed=7, embeddings=(768, 64)))
        ms = _timed(lambda: lam(ep)["logits"], steps)
        gf = 26 * 22.55 + 150 * 10.855 + 28.7
        out["mode_B_embeddings"] = {
            "workload": f"SAM-512 5-way 5-shot, Lam.forward(embeddings): neck + prompt encoder + decoder, batch {B}",
            "ms_per_step": ms, "episodes_per_s": B / ms * 1e3, "gflop_per_episode": gf, "tflops": gf * B / ms}
        # the same with point + box + mask prompts (9 sparse tokens per sequence instead of 1): the token <-> image
        # attentions run on la_attention_tokens (CUDA cores) instead of the pooled single-token path
        epm = cuda(make_episode(B, N_WAYS, K_SHOTS, IMAGE_SIZE, seed=7, prompts="mixed", embeddings=(768, 64)))
        ms = _timed(lambda: lam(epm)["logits"], steps)
        out["mode_B_mixed_prompts"] = {
            "workload": f"SAM-512 5-way 5-shot, Lam.forward(embeddings) with 5 points + 2 boxes + 1 mask per (example, "
                        f"class): 9 sparse tokens per sequence, batch {B}",
            "ms_per_step": ms, "episodes_per_s": B / ms * 1e3}
        del ep, epm
        # ---- mode C: class embeddings once, then predict per query (lam.py:349-381) ----
        sup = make_episode(1, N_WAYS, K_SHOTS, IMAGE_SIZE, seed=8)
        sup_in = cuda({k: (v[:, 1:] if k in ("images", "dims") else v) for k, v in sup.items()})
        t0 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0[0].record()
        ce = lam.generate_class_embeddings(sup_in)
        t0[1].record()
        torch.cuda.synchronize()
        from labelanything_b200.utils import ResultDict

        # the same cached class embeddings serve every query of the step (lam.py:362-381 decodes B queries against
        # class_embeddings[B, C, D])
        ce = dict(ce)
        ce[ResultDict.CLASS_EMBS] = ce[ResultDict.CLASS_EMBS].expand(B, -1, -1).contiguous()
        q = {"images": torch.randn(B, 1, 3, IMAGE_SIZE, IMAGE_SIZE, device="cuda"),
             "dims": torch.full((B, 2), IMAGE_SIZE, dtype=torch.int64, device="cuda")}
        ms = _timed(lambda: lam.predict(q, ce), steps)
        gf = 965.64 + 22.55 + 28.7
        out["mode_C_predict"] = {
            "workload": f"SAM-512 5-way 5-shot, generate_class_embeddings once ({t0[0].elapsed_time(t0[1]):.1f} ms, "
                        f"untimed) + predict on {B} query images per step",
            "ms_per_step": ms, "queries_per_s": B / ms * 1e3, "gflop_per_query": gf, "tflops": gf * B / ms}
        del q, sup_in, ce
        # ---- config 5: 20-way 5-shot, one episode per GPU (101 images, 2100 prompt sequences) ----
        lam.prompt_encoder.class_encoder.fixed_rows = torch.arange(21)
        ep = cuda(make_episode(1, 20, 5, IMAGE_SIZE, seed=9))
        ms = _timed(lambda: lam(ep)["logits"], max(2, steps - 1), warmup=1)
        gf = 101 * (965.64 + 22.55) + 2100 * 10.855 + 28.9
        out["config5_20way_5shot"] = {
            "workload": "SAM-512 20-way 5-shot, Lam.forward(images), 1 episode per GPU (101 images, 2100 sequences)",
            "ms_per_step": ms, "episodes_per_s": 1e3 / ms, "gflop_per_episode": gf, "tflops": gf / ms}
        lam.prompt_encoder.class_encoder.fixed_rows = torch.arange(N_WAYS + 1)
        del ep
        torch.cuda.empty_cache()
        # ---- config 2: MAE-256 (HF ViT-B, 480 px), 1-way 1-shot, batch 32 ----
        mae = build_lam(build_vit=lambda project_last_hidden: build_vit_from_config(), image_embed_dim=768,
                        embed_dim=256, image_size=480, spatial_convs=3, class_attention=False,
                        example_attention=False, example_class_attention=True,
                        class_encoder={"name": "RandomMatrixEncoder", "bank_size": 100, "embed_dim": 256},
                        custom_preprocess=False)
        load_synth_weights(mae, seed=0)
        mae.prompt_encoder.class_encoder.fixed_rows = torch.arange(2)
        mae = mae.cuda()
        ep = cuda(make_episode(32, 1, 1, 480, seed=3))
        ms = _timed(lambda: mae(ep)["logits"], max(steps, 5), warmup=3)
        out["config2_mae256_b32"] = {
            "workload": "MAE-256 (HF ViT-B 480px, embed 256) 1-way 1-shot, Lam.forward(images), batch 32",
            "ms_per_step": ms, "episodes_per_s": 32 / ms * 1e3, "gflop_per_episode": 373.8, "tflops": 373.8 * 32 / ms}
        del mae, ep
        torch.cuda.empty_cache()
    return out


def parity_drift(lam) -> dict | None:
    """Drift of the native logits (bf16 operands) against the golden tensors of the UNMODIFIED fp32 reference for one
    SAM-512 5-way 5-shot episode (tests/golden/sam512_5w5s.pt, oracle/make_golden.py) -- reported, not asserted here
    (the assertions live in tests/)."""
    import torch

    from labelanything_b200.synthetic import make_episode

    p = ROOT / "tests" / "golden" / "sam512_5w5s.pt"
    if not p.exists():
        return None
    g = torch.load(p, weights_only=False)
    if g.get("weights_seed", 0) != 0:
        return None
    rows = lam.prompt_encoder.class_encoder.fixed_rows
    lam.prompt_encoder.class_encoder.fixed_rows = g["class_rows"]
    ep = {k: v.cuda() for k, v in make_episode(**g["episode_args"]).items()}
    with torch.no_grad():
        out = lam(ep)["logits"][..., ::8, ::8].float().cpu()
    lam.prompt_encoder.class_encoder.fixed_rows = rows
    ref = g["logits_sub8"].float()
    fin = torch.isfinite(ref)
    err = (out[fin] - ref[fin]).abs()
    return {"parity_max_abs": err.max().item(), "parity_mean_abs": err.mean().item(), "logit_std": ref[fin].std().item(),
            "inf_pattern_equal": bool(torch.equal(torch.isfinite(out), fin)),
            "against": "tests/golden/sam512_5w5s.pt: logits of the unmodified fp32 reference, same weights and episode"}


def training_measurement(rank: int, world: int, steps: int = 5, warmup: int = 3) -> dict:
    """BASELINE.json configs[3]: one optimisation step of `lam_no_vit` on pre-computed ViT-MAE-L
    embeddings (parameters/trainval/coco/mael.yaml: image_embed_dim 1024, embed_dim 256, 480 px, focal loss + class
    weighting, AdamW), 2 episodes of 2-way 5-shot per GPU with point + box + mask prompts.  Runs on EVERY rank: forward,
    loss, backward, ONE all-reduce of the flat gradient bucket over NCCL, one AdamW launch.  Timed like the headline
    (barrier + synchronize, CUDA events, max over ranks); the all-reduce is timed in separate steps (its event pair
    costs a host sync)."""
    import torch
    import torch.distributed as dist

    from labelanything_b200 import ops
    from labelanything_b200.build_lam import build_lam_no_vit
    from labelanything_b200.loss import LabelAnythingLoss
    from labelanything_b200.synthetic import load_synth_weights, make_episode
    from labelanything_b200.training import FlatAdamW, train_step

    B, N, K, S = 2, 2, 5, 480
    lam = build_lam_no_vit(image_embed_dim=1024, embed_dim=256, image_size=S, spatial_convs=3, class_attention=False,
                           example_attention=False, example_class_attention=True, custom_preprocess=False)
    load_synth_weights(lam, seed=4)
    lam = lam.cuda().train()
    ep = {k: v.cuda() for k, v in make_episode(B, N, K, S, seed=400 + rank, prompts="mixed", embeddings=(1024, 30)).items()}
    g = torch.Generator().manual_seed(500 + rank)
    gt = torch.randint(0, N + 1, (B, S // 16, S // 16), generator=g).repeat_interleave(16, 1).repeat_interleave(16, 2).cuda()
    loss_fn = LabelAnythingLoss({"focal": {"weight": 1.0, "gamma": 2.0}}, class_weighting=True)
    opt = FlatAdamW(lam.parameters(), lr=5e-5)

    def sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    losses = []
    for _ in range(warmup):
        losses.append(float(train_step(lam, loss_fn, opt, ep, gt)["loss"]["value"]))
    sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        r = train_step(lam, loss_fn, opt, ep, gt)
    e1.record()
    sync()
    losses.append(float(r["loss"]["value"]))
    ms = e0.elapsed_time(e1) / steps
    with ops.profile() as prof:
        train_step(lam, loss_fn, opt, ep, gt)
        torch.cuda.synchronize()
    kernel_ms = sum(a.elapsed_time(b) for _, _, _, a, b in prof.records)
    gemm_ms = sum(a.elapsed_time(b) for n, _, _, a, b in prof.records if n.startswith("gemm"))
    ar = []
    for _ in range(3):
        train_step(lam, loss_fn, opt, ep, gt, timed=True)
        if opt.last_allreduce_ms is not None:
            ar.append(opt.last_allreduce_ms)
    # the same step captured once in a CUDA graph (forward, loss, backward, all-reduce, AdamW) and replayed
    from labelanything_b200.training import GraphedTrainStep

    gstep = GraphedTrainStep(lam, loss_fn, opt, ep, gt, warmup=1)
    for _ in range(warmup):
        gstep()
    sync()
    g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    g0.record()
    for _ in range(steps):
        r = gstep()
    g1.record()
    sync()
    losses.append(float(r["loss"]["value"]))
    ms_graph = g0.elapsed_time(g1) / steps
    # the same configuration from IMAGES through a frozen ViT-MAE-L (BASELINE configs[3] as worded: the encoder runs on its
    # inference kernels inside the captured step, only neck + prompt encoder + decoder are trained)
    del gstep
    from labelanything_b200.build_encoder import build_vit_from_config
    from labelanything_b200.build_lam import build_lam

    lam_l = build_lam(build_vit=lambda project_last_hidden: build_vit_from_config(
        hidden_size=1024, num_hidden_layers=24, num_attention_heads=16, intermediate_size=4096, image_size=224),
        image_embed_dim=1024, embed_dim=256, image_size=S, spatial_convs=3, class_attention=False, example_attention=False,
        example_class_attention=True, custom_preprocess=False)
    load_synth_weights(lam_l, seed=4)
    lam_l = lam_l.cuda().train()
    ep_i = {k: v.cuda() for k, v in make_episode(B, N, K, S, seed=400 + rank, prompts="mixed").items()}
    opt_l = FlatAdamW(lam_l.get_learnable_params({"freeze_backbone": True}), lr=5e-5)
    gstep_l = GraphedTrainStep(lam_l, loss_fn, opt_l, ep_i, gt, warmup=2)
    for _ in range(warmup):
        gstep_l()
    sync()
    i0, i1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    i0.record()
    for _ in range(steps):
        gstep_l()
    i1.record()
    sync()
    ms_images = i0.elapsed_time(i1) / steps
    del gstep_l, lam_l, opt_l
    t = torch.tensor([ms, max(ar) if ar else 0.0, ms_graph, ms_images], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, ar_ms, ms_graph, ms_images = float(t[0]), float(t[1]), float(t[2]), float(t[3])
    return {"workload": f"BASELINE configs[3]: lam_no_vit on pre-computed ViT-MAE-L embeddings (1024 x 30 x 30), embed 256, "
                        f"480 px, {N}-way {K}-shot, {B} episodes per GPU, point + box + mask prompts; forward + focal loss "
                        f"+ backward + gradient all-reduce + AdamW, bf16 GEMM operands / fp32 accumulation and state",
            "ms_per_step": ms_graph, "episodes_per_s": B * world / ms_graph * 1e3, "n_gpus": world,
            "step": ("one CUDA-graph replay per step (GraphedTrainStep)" if world == 1 else
                     "two CUDA-graph replays per step (forward + backward | AdamW) with the NCCL all-reduce of the "
                     "gradient bucket between them (GraphedTrainStep)") + "; eager launch sequence: see eager_ms_per_step",
            "eager_ms_per_step": ms, "eager_episodes_per_s": B * world / ms * 1e3,
            "from_images_frozen_vit_l": {"workload": f"the same step from images [{B}, {N * K + 1}, 3, {S}, {S}] through a "
                                                     "frozen ViT-MAE-L (1024 / 24 / 16) inside the captured graph",
                                         "ms_per_step": ms_images, "episodes_per_s": B * world / ms_images * 1e3},
            "allreduce_ms": ar_ms if world > 1 else None, "grad_bucket_bytes": int(opt.flat_g.numel() * 4),
            "trainable_parameters": int(sum(p.numel() for p in opt.params)),
            "native_launches_per_step": prof.launches, "kernel_ms_per_step": kernel_ms, "gemm_ms_per_step": gemm_ms,
            "loss_first_last": [losses[0], losses[-1]]}


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

    # The image tensor (2.7 of the 2.9 GB) goes up in slices of one encoder chunk each, every slice with its own event:
    # the encoder's chunk c waits for slice c only (ImageEncoderViT.chunk_ready), so a step whose upload could not be
    # hidden behind the previous step (the first one) starts computing after a quarter of the transfer.
    n_img = B * (N_WAYS * K_SHOTS + 1)
    cap = lam.image_encoder.max_images_per_chunk
    per_chunk = -(-n_img // -(-n_img // cap))
    slices = [(s0, min(per_chunk, n_img - s0)) for s0 in range(0, n_img, per_chunk)]
    slice_up = [[torch.cuda.Event() for _ in slices] for _ in range(2)]
    host_flat = host["images"].view(n_img, *host["images"].shape[2:])
    active = {"slot": None}

    def upload(slot):
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(consumed[slot])      # the step that last used this buffer has finished reading it
            for k, v in host.items():
                if k != "images":
                    dev2[slot][k].copy_(v, non_blocking=True)
            uploaded[slot].record(copy_stream)          # everything but the images
            dst = dev2[slot]["images"].view(n_img, *host["images"].shape[2:])
            for i, (s0, n) in enumerate(slices):
                dst[s0:s0 + n].copy_(host_flat[s0:s0 + n], non_blocking=True)
                slice_up[slot][i].record(copy_stream)

    def chunk_ready(first, n):
        slot = active["slot"]
        if slot is not None:
            for i, (s0, m) in enumerate(slices):
                if s0 < first + n and first < s0 + m:
                    torch.cuda.current_stream().wait_event(slice_up[slot][i])

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
            active["slot"] = slot
            lam.image_encoder.chunk_ready = chunk_ready
            with torch.no_grad():
                logits = lam(dev2[slot])["logits"]
            lam.image_encoder.chunk_ready = None
            active["slot"] = None
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

    for _ in range(max(args.warmup, 3)):
        step_resident()
    barrier()

    # ---- timed region 1 (the headline): inputs resident in HBM, one event pair around K steps -------------------
    sampler = ClockSampler(local) if rank == 0 else None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        step_resident()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
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

    # ---- instrumented pass (NOT the headline): every native launch bracketed by CUDA events ----------------------
    prof_steps = min(args.steps, 3)
    with ops.profile() as prof:
        barrier()
        p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        p0.record()
        for _ in range(prof_steps):
            step_resident()
        p1.record()
        barrier()
    ms_prof = p0.elapsed_time(p1)
    detail = prof.summary()
    launches_per_step = prof.launches // prof_steps
    fam: dict = {}   # kernel families (gemm.n2304.k768 -> gemm)
    for k, v in detail.items():
        f = fam.setdefault(k.split(".")[0], {"launches": 0, "ms": 0.0, "flops": 0.0, "bytes": 0.0})
        for kk in f:
            f[kk] += v[kk]

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
        # global-attention instantiation of la_attention_bf16 on one encoder chunk)
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
        # whole-step figure PER GPU: the aggregate algorithmic rate divided by the number of GPUs that produced it
        step_tflops = EPISODE_GFLOP * episodes / ms / world
        roofline = {"kernel": top, "bound": "tensor" if tensor_bound else "hbm", "achieved": achieved, "peak": peak,
                    "unit": unit, "frac": achieved / peak,
                    "traffic": None if traffic is None else traffic * 1e6, "traffic_unit": "bytes per launch",
                    "traffic_source": NCU_TRAFFIC_SOURCE,
                    "algorithmic_per_launch": {"flops": f["flops"] / f["launches"], "bytes": f["bytes"] / f["launches"]},
                    "peak_source": peaks["_source"],
                    "avg_launch_ms": f["ms"] / f["launches"], "share_of_kernel_time": f["ms"] / kernel_ms,
                    "timed_in": f"separate instrumented pass of {prof_steps} steps ({ms_prof / prof_steps:.1f} ms/step with "
                                f"two event records per launch; the headline pass has none)",
                    "whole_step": {"achieved": step_tflops, "unit": "TFLOP/s per GPU",
                                   "frac": step_tflops / peaks["bf16_tflops_sustained"]},
                    "families": {k: {"launches": v["launches"] // prof_steps, "ms_per_step": v["ms"] / prof_steps,
                                     "tflops": (v["flops"] / v["ms"] / 1e9) if v["flops"] else None,
                                     "gbs": (v["bytes"] / v["ms"] / 1e6) if v["bytes"] else None}
                                 for k, v in sorted(fam.items(), key=lambda kv: -kv[1]["ms"])}}
        if args.detail:
            for k, v in sorted(detail.items(), key=lambda kv: -kv[1]["ms"])[:28]:
                print(f"# {k:34s} {v['launches'] // prof_steps:5d} launches/step {v['ms'] / prof_steps:8.2f} ms/step "
                      f"{(v['flops'] / v['ms'] / 1e9) if v['flops'] else 0:8.1f} TFLOP/s {v['bytes'] / v['ms'] / 1e6:8.0f} GB/s",
                      file=sys.stderr)
    training = None
    if not args.no_training:
        del dev
        torch.cuda.empty_cache()
        training = training_measurement(rank, world)
    if rank == 0:
        secondary = parity = cpu = None
        if world == 1:
            del dev2
            torch.cuda.empty_cache()
            parity = parity_drift(lam)
            if not args.no_secondary:
                secondary = secondary_measurements(lam)
            if not args.no_cpu:
                ref = ReferenceCpu()
                ref.step(0)                       # warm-up (first-touch allocations, thread pool)
                r = ref.step(1)
                cpu = {"value": 1.0 / r["t_episode"], "unit": "episodes/s", "cores": ref.cores, "kind": ref.kind,
                       "sample": r["sample"]}
        line = {"metric": METRIC, "value": value, "unit": "episodes/s", "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
                "config": _config(B, world),
                "e2e": {"value": e2e, "unit": "episodes/s", "h2d_bytes_per_step": h2d_bytes,
                        "d2h_bytes_per_step": d2h_bytes, "ms_per_step": ms_e2e / args.steps,
                        "pipeline": "double-buffered inputs: H2D of step k+1 on a copy stream overlaps step k, the images in slices of one encoder chunk (chunk c waits for slice c only); D2H of the logits on a third stream overlaps step k+1"},
                "gpu_launches": launches_per_step * args.steps, "roofline": roofline, "parity": parity,
                "secondary": secondary, "training": training, "cpu_baseline": cpu, "clocks": clocks}
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
    ap.add_argument("--no-training", action="store_true", help="skip the training-step measurement (BASELINE configs[3])")
    ap.add_argument("--no-secondary", action="store_true", help="skip the labelled secondary measurements (N = 1)")
    ap.add_argument("--full-episode", action="store_true",
                    help="--impl reference: time ONE whole episode through the reference (minutes) instead of K samples")
    ap.add_argument("--chunk", type=int, default=0, help="images per encoder chunk (0 = the model's default)")
    ap.add_argument("--detail", action="store_true", help="print the per-kernel-shape breakdown to stderr")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_native(args)


if __name__ == "__main__":
    main()
