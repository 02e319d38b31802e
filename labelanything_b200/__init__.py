"""labelanything_b200 — B200-native (sm_100a) implementation of the LabelAnything hot path.

Image encoder (SAM ViT / HF ViT) -> visual prompt encoder -> two-way-transformer mask decoder, behind the
reference's `label_anything.models.LabelAnything` / `Lam` module surface.  All arithmetic runs in hand-written
CUDA kernels reached through the C ABI in include/labelanything_b200.h; there is no CPU or eager fallback.
"""
__version__ = "0.1.0"
