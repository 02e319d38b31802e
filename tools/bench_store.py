"""Embedding store: batched load / save of one BASELINE batch of SAM-512 embeddings (208 files of
512 x 64 x 64 fp32 = 1.74 GB, page cache hot) next to the reference's per-file loop: python tools/bench_store.py"""
import json
import shutil
import sys
import tempfile
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch
from safetensors.torch import load_file, save_file

from labelanything_b200.embedding_store import EmbeddingStore

n, shape = 208, (512, 64, 64)
tmp = tempfile.mkdtemp(prefix="la_store_")
try:
    store = EmbeddingStore(tmp, workers=16)
    embs = torch.randn((n,) + shape, device="cuda")
    ids = list(range(n))
    nbytes = embs.numel() * 4

    def wall(fn, iters=3):
        fn()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(iters):
            fn()
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) / iters

    def ref_save():
        out = embs.cpu()
        for i in ids:
            save_file({"embedding": out[i]}, store.path(i))

    def ref_load():
        return torch.stack([load_file(store.path(i))["embedding"] for i in ids]).cuda()

    t_save = wall(lambda: store.save(ids, embs))
    t_load = wall(lambda: store.load(ids))
    assert torch.equal(store.load(ids)[0], embs)
    t_ref_save = wall(ref_save, 1)
    t_ref_load = wall(ref_load, 1)
    print(json.dumps({"workload": f"{n} files x {shape} fp32 = {nbytes / 1e9:.2f} GB, page cache hot",
                      "batched_load_to_gpu": {"s": t_load, "GB/s": nbytes / t_load / 1e9},
                      "batched_save_from_gpu": {"s": t_save, "GB/s": nbytes / t_save / 1e9},
                      "reference_loop_load_file_stack_cuda": {"s": t_ref_load, "GB/s": nbytes / t_ref_load / 1e9},
                      "reference_loop_cpu_save_file": {"s": t_ref_save, "GB/s": nbytes / t_ref_save / 1e9}}, indent=1))
finally:
    shutil.rmtree(tmp, ignore_errors=True)
