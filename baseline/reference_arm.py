"""The reference arm of bench.py: the UNMODIFIED reference (`label_anything`, installed with
`pip install --no-deps --target baseline/_ref`) timed on the host CPU cores.

BASELINE INFRASTRUCTURE ONLY -- never imported by labelanything_b200.  Nothing here touches the native library:
the models are built by the reference's own builders (`label_anything.models.build_lam_vit_b`, `build_lam_vit_mae_b`)
and driven through the reference's own `Lam.forward`.

The SAM 1024-px episode is fed the way SURVEY.md §8d / BASELINE.md §4 prescribe: the reference encoder on <= 2 images
per call (one un-chunked 26-image call materialises 26 x 12 x 4096^2 fp32 attention matrices, > 62 GB), the result handed
to `Lam.forward` through the `embeddings` key -- the same arithmetic as the `images` key
(label_anything/models/lam.py:139-146 vs 158-163).
"""
from __future__ import annotations

import os
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def available() -> bool:
    return (ROOT / "baseline" / "_ref" / "label_anything" / "models" / "lam.py").exists()


def build_reference_lam(builder: str, kwargs: dict, seed: int = 0, n_classes: int | None = None):
    """The reference's own model (its builder, its modules), synthetic weights of labelanything_b200.synthetic
    (a pure function of parameter name / shape / seed, so both arms hold identical weights)."""
    import torch

    from baseline import ref_shim
    from labelanything_b200.synthetic import load_synth_weights

    models = ref_shim.import_reference()
    lam = getattr(models, builder)(**kwargs)
    load_synth_weights(lam, seed=seed)
    lam.eval()
    ce = getattr(lam.prompt_encoder, "class_encoder", None)
    if ce is not None and hasattr(ce, "sample_rows") and n_classes is not None:
        rows = torch.arange(n_classes)
        ce.sample_rows = lambda C, device=None: rows[:C]     # pins torch.randperm, as in the native arm
    return lam


class SamEpisodeSampler:
    """Bounded samples of one SAM ViT-B 1024-px N-way K-shot episode through the unmodified reference.

    `step(n_img)`: the reference image encoder on `n_img` (<= 2) images of the episode in one call, then the rest of
    the episode -- neck on all M+1 feature maps, prompt encoder on all M*C sequences, mask decoder, postprocess --
    through `Lam.forward` with the `embeddings` key; encoder outputs not computed in this step are copies of the
    computed ones (the arithmetic downstream does not depend on their values).  Returns the two times and the
    episode time they scale to: (M+1)/n_img encoder calls + one `forward(embeddings)`.
    `full_episode()`: every image encoded (2 per call), nothing scaled."""

    def __init__(self, lam, n_ways: int, k_shots: int, image_size: int, threads: int | None = None):
        import torch

        self.torch = torch
        self.lam = lam
        self.n_ways, self.k_shots, self.image_size = n_ways, k_shots, image_size
        self.cores = threads or os.cpu_count() or 1
        torch.set_num_threads(self.cores)

    def _episode(self, seed: int):
        from labelanything_b200.synthetic import make_episode

        return make_episode(1, self.n_ways, self.k_shots, self.image_size, seed=seed)

    def _forward_embeddings(self, ep, emb):
        batch = {k: v for k, v in ep.items() if k != "images"}
        batch["embeddings"] = emb
        return self.lam(batch)

    def step(self, seed: int = 0, n_img: int = 2) -> dict:
        torch = self.torch
        ep = self._episode(seed)
        n_total = ep["images"].shape[1]
        with torch.no_grad():
            t0 = time.perf_counter()
            feats = self.lam.image_encoder(ep["images"][0, :n_img])
            t_enc = time.perf_counter() - t0
            reps = -(-n_total // n_img)
            emb = feats.repeat(reps, 1, 1, 1)[:n_total].unsqueeze(0).contiguous()
            t0 = time.perf_counter()
            out = self._forward_embeddings(ep, emb)
            t_rest = time.perf_counter() - t0
        assert out["logits"].shape[1] == self.n_ways + 1
        t_episode = (n_total / n_img) * t_enc + t_rest
        return {"t_step": t_enc + t_rest, "t_enc": t_enc, "t_rest": t_rest, "t_episode": t_episode, "n_img": n_img,
                "n_total": n_total}

    def full_episode(self, seed: int = 0) -> dict:
        torch = self.torch
        ep = self._episode(seed)
        n_total = ep["images"].shape[1]
        with torch.no_grad():
            t0 = time.perf_counter()
            feats = [self.lam.image_encoder(ep["images"][0, i:i + 2]) for i in range(0, n_total, 2)]
            emb = torch.cat(feats).unsqueeze(0)
            t_enc = time.perf_counter() - t0
            out = self._forward_embeddings(ep, emb)
            t_all = time.perf_counter() - t0
        return {"t_step": t_all, "t_enc": t_enc, "t_rest": t_all - t_enc, "t_episode": t_all, "n_img": n_total,
                "n_total": n_total, "logits": out["logits"]}

    def describe(self, r: dict) -> str:
        if r["n_img"] == r["n_total"]:
            return (f"unmodified reference (baseline/_ref), fp32, {self.cores} threads: one whole {self.n_ways}-way "
                    f"{self.k_shots}-shot episode, encoder 2 images per call {r['t_enc']:.1f}s + forward(embeddings) "
                    f"{r['t_rest']:.1f}s")
        return (f"unmodified reference (baseline/_ref), fp32, {self.cores} threads: image_encoder on {r['n_img']} of the "
                f"{r['n_total']} images {r['t_enc']:.2f}s + Lam.forward(embeddings) of the whole episode "
                f"({r['n_total'] - 1} x {self.n_ways + 1} prompt sequences, decode, postprocess) {r['t_rest']:.2f}s; "
                f"scaled to {r['n_total']}/{r['n_img']} encoder calls + 1 forward = {r['t_episode']:.1f}s/episode")
