"""Embedding store of the reference's recommended training mode: one
`<image id, 12 digits>.safetensors` per image with key "embedding" (C x h x w fp32) and optionally
"<dataset>_gt" — written by label_anything/preprocess.py:53-75,163-172,202-206, read back by
label_anything/data/coco.py:251-275,490-505 (`load_file` per image, then `torch.stack`).

Same files, byte for byte (the safetensors container: u64 header length, compact JSON header padded with spaces to
8 bytes, tensors ordered by descending dtype rank then name, raw little-endian data), but batched:
  * load: headers parsed once, every file's payload is read by a thread pool STRAIGHT into its slot of one pinned
    [n, C, h, w] staging buffer (no per-file tensors, no torch.stack copy), then one asynchronous host-to-device copy;
  * save: one device-to-host copy of the whole batch into pinned memory, files written by the pool.
Host-side I/O only — the token layout conversion of the `embeddings` input path is `la_nchw_to_tokens` (lam.py).
"""
from __future__ import annotations

import json
import os
import struct
from concurrent.futures import ThreadPoolExecutor
from typing import Optional, Sequence

import numpy as np
import torch

__all__ = ["EmbeddingStore", "read_safetensors_header", "write_safetensors"]

_DTYPES = {"F64": torch.float64, "F32": torch.float32, "F16": torch.float16, "BF16": torch.bfloat16,
           "I64": torch.int64, "I32": torch.int32, "I16": torch.int16, "I8": torch.int8, "U8": torch.uint8,
           "BOOL": torch.bool}
_NAMES = {v: k for k, v in _DTYPES.items()}
# file order of the safetensors writer: descending position in its dtype enumeration (BOOL < U8 < I8 < I16 < F16 < BF16
# < I32 < F32 < F64 < I64), then by name
_RANK = {n: i for i, n in enumerate(["BOOL", "U8", "I8", "I16", "F16", "BF16", "I32", "F32", "F64", "I64"])}


def _header_bytes(entries: list[tuple[str, torch.dtype, tuple, int]], metadata: Optional[dict] = None) -> bytes:
    """entries: (name, dtype, shape, nbytes) in file order -> the padded JSON header."""
    parts, off = [], 0
    if metadata:
        parts.append('"__metadata__":' + json.dumps(metadata, separators=(",", ":")))
    for name, dtype, shape, nbytes in entries:
        parts.append(f'{json.dumps(name)}:{{"dtype":"{_NAMES[dtype]}","shape":[{",".join(str(int(s)) for s in shape)}],'
                     f'"data_offsets":[{off},{off + nbytes}]}}')
        off += nbytes
    h = ("{" + ",".join(parts) + "}").encode()
    return h + b" " * (-len(h) % 8)


def write_safetensors(path: str, tensors: dict, metadata: Optional[dict] = None) -> None:
    """Byte-identical to safetensors.torch.save_file(tensors, path, metadata) for contiguous CPU tensors."""
    items = sorted(tensors.items(), key=lambda kv: (-_RANK[_NAMES[kv[1].dtype]], kv[0]))
    entries = [(k, v.dtype, tuple(v.shape), v.numel() * v.element_size()) for k, v in items]
    head = _header_bytes(entries, metadata)
    with open(path, "wb") as f:
        f.write(struct.pack("<Q", len(head)))
        f.write(head)
        for _, v in items:
            if not (v.device.type == "cpu" and v.is_contiguous()):
                raise ValueError("write_safetensors: tensors must be contiguous CPU tensors")
            f.write(v.reshape(-1).view(torch.uint8).numpy().tobytes() if v.dtype == torch.bfloat16 else v.numpy().tobytes())


def read_safetensors_header(path: str) -> tuple[dict, int]:
    """-> ({name: (dtype, shape, begin, end)}, offset of the data section)."""
    with open(path, "rb") as f:
        (n,) = struct.unpack("<Q", f.read(8))
        if n > 100 << 20:
            raise ValueError(f"{path}: implausible safetensors header length {n}")
        head = json.loads(f.read(n))
    out = {}
    for k, v in head.items():
        if k == "__metadata__":
            continue
        out[k] = (_DTYPES[v["dtype"]], tuple(v["shape"]), int(v["data_offsets"][0]), int(v["data_offsets"][1]))
    return out, 8 + n


def _read_into(path: str, offset: int, dst: np.ndarray) -> None:
    view = memoryview(dst.reshape(-1).view(np.uint8))
    fd = os.open(path, os.O_RDONLY)
    try:
        done = 0
        while done < len(view):
            got = os.preadv(fd, [view[done:]], offset + done)
            if got <= 0:
                raise IOError(f"{path}: truncated payload")
            done += got
    finally:
        os.close(fd)


class EmbeddingStore:
    """Directory of per-image embedding files (`emb_dir` of the reference's datasets, coco.py:251-275)."""

    def __init__(self, emb_dir: str, name: Optional[str] = None, load_gts: bool = False, workers: int = 8):
        self.emb_dir = emb_dir
        self.name = name
        self.load_gts = load_gts
        self.pool = ThreadPoolExecutor(max_workers=workers)
        self._staging: Optional[torch.Tensor] = None

    def path(self, image_id) -> str:
        return os.path.join(self.emb_dir, f"{str(image_id).zfill(12)}.safetensors")

    def _stage(self, shape: tuple, dtype: torch.dtype) -> torch.Tensor:
        need = int(np.prod(shape))
        s = self._staging
        if s is None or s.dtype != dtype or s.numel() < need:
            s = torch.empty(need, dtype=dtype, pin_memory=torch.cuda.is_available())
            self._staging = s
        return s[:need].view(shape)

    # ---- read ----------------------------------------------------------------------------------------------------
    def load(self, image_ids: Sequence, device: torch.device | str = "cuda"):
        """-> (embeddings [n, C, h, w] on `device`, list of ground-truth tensors or None) — what the reference builds
        with load_file + torch.stack (coco.py:490-505) followed by the trainer's `.to(device)`."""
        paths = [self.path(i) for i in image_ids]
        heads = list(self.pool.map(read_safetensors_header, paths))
        if not paths:
            raise ValueError("EmbeddingStore.load: no image ids")
        for p, (h, _) in zip(paths, heads):
            if "embedding" not in h:
                raise KeyError(f"{p}: no 'embedding' tensor (pyramid stores are not supported)")
        dtype, shape = heads[0][0]["embedding"][:2]
        for p, (h, _) in zip(paths, heads):
            if h["embedding"][:2] != (dtype, shape):
                raise ValueError(f"{p}: embedding {h['embedding'][:2]} differs from the batch's {(dtype, shape)}")
        stage = self._stage((len(paths),) + shape, dtype)
        stage_np = stage.view(torch.uint8).numpy() if dtype == torch.bfloat16 else stage.numpy()
        jobs = [self.pool.submit(_read_into, p, start + h["embedding"][2], stage_np[i])
                for i, (p, (h, start)) in enumerate(zip(paths, heads))]
        gts = None
        if self.load_gts:
            key = f"{self.name}_gt"
            gts = []
            for p, (h, start) in zip(paths, heads):
                gd, gs, b, e = h[key]
                g = torch.empty(gs, dtype=gd)
                _read_into(p, start + b, g.numpy())
                gts.append(g)
        for j in jobs:
            j.result()
        dev = torch.device(device)
        if dev.type == "cpu":
            return stage.clone(), gts
        out = stage.to(dev, non_blocking=True)
        torch.cuda.current_stream(dev).synchronize()     # the staging buffer is reused by the next call
        return out, gts

    # ---- write ---------------------------------------------------------------------------------------------------
    def save(self, image_ids: Sequence, embeddings: torch.Tensor, gts: Optional[Sequence[torch.Tensor]] = None) -> None:
        """embeddings [n, C, h, w] (device or host) -> one file per image, as preprocess.py:65-73 writes them."""
        if embeddings.shape[0] != len(image_ids):
            raise ValueError("EmbeddingStore.save: one image id per embedding")
        os.makedirs(self.emb_dir, exist_ok=True)
        if embeddings.is_cuda:
            host = self._stage(tuple(embeddings.shape), embeddings.dtype)
            host.copy_(embeddings, non_blocking=True)
            torch.cuda.current_stream(embeddings.device).synchronize()
        else:
            host = embeddings.contiguous()

        def one(i):
            t = {"embedding": host[i]}
            if gts is not None:
                t[f"{self.name}_gt"] = gts[i].cpu().contiguous()
            write_safetensors(self.path(image_ids[i]), t)

        list(self.pool.map(one, range(len(image_ids))))
