"""Embedding store with device tensors: batch save from the GPU, batch load to the GPU,
bit-exact against the safetensors library's view of the same files."""
import pytest
import torch
from safetensors.torch import load_file

pytestmark = pytest.mark.gpu


def test_device_round_trip(tmp_path):
    from labelanything_b200.embedding_store import EmbeddingStore

    store = EmbeddingStore(str(tmp_path), name="coco", load_gts=True)
    ids = list(range(40, 52))
    embs = torch.randn(len(ids), 32, 30, 30, device="cuda")
    gts = [torch.randint(0, 3, (20 + i, 31)) for i in range(len(ids))]
    store.save(ids, embs, gts)
    for k in (0, 5, 11):
        f = load_file(store.path(ids[k]))
        assert torch.equal(f["embedding"], embs[k].cpu()) and torch.equal(f["coco_gt"], gts[k])
    got, got_gts = store.load(ids[::-1])
    assert got.is_cuda and torch.equal(got, embs.flip(0))
    assert all(torch.equal(a, b) for a, b in zip(got_gts, gts[::-1]))
    got2, _ = store.load(ids[:3])                     # smaller batch reuses the staging buffer
    assert torch.equal(got2, embs[:3]) and torch.equal(got, embs.flip(0))
