"""Embedding store: files byte-identical to the safetensors library the reference writes them
with (preprocess.py:69-73), batched loads equal to its per-file `load_file` + `torch.stack` (coco.py:251-275,490-505)."""
import os

import pytest
import torch
from safetensors.torch import load_file, save_file

from labelanything_b200.embedding_store import EmbeddingStore, read_safetensors_header, write_safetensors


def test_written_files_are_byte_identical_to_safetensors(tmp_path):
    g = torch.Generator().manual_seed(0)
    cases = [
        {"embedding": torch.randn(8, 6, 6, generator=g)},
        {"embedding": torch.randn(4, 3, 2, generator=g), "coco_gt": torch.randint(0, 80, (5, 7), generator=g),
         "x_u8": torch.arange(3, dtype=torch.uint8)},
        {"embedding": torch.randn(2, 2, generator=g).to(torch.bfloat16), "stage1": torch.randn(3, generator=g).half()},
    ]
    for i, t in enumerate(cases):
        a, b = tmp_path / f"a{i}.safetensors", tmp_path / f"b{i}.safetensors"
        save_file(t, str(a))
        write_safetensors(str(b), t)
        assert a.read_bytes() == b.read_bytes()
        head, start = read_safetensors_header(str(b))
        assert set(head) == set(t) and start % 8 == 0
    save_file(cases[0], str(tmp_path / "m_a.safetensors"), metadata={"k": "v"})
    write_safetensors(str(tmp_path / "m_b.safetensors"), cases[0], metadata={"k": "v"})
    assert (tmp_path / "m_a.safetensors").read_bytes() == (tmp_path / "m_b.safetensors").read_bytes()


def test_batched_load_equals_the_reference_loader(tmp_path):
    g = torch.Generator().manual_seed(1)
    ids = [17, 4, 123456789012, 900]
    embs = torch.randn(len(ids), 16, 6, 6, generator=g)
    gts = [torch.randint(0, 5, (10 + i, 12), generator=g) for i in range(len(ids))]
    for i, e, gt in zip(ids, embs, gts):       # written the reference's way
        save_file({"embedding": e, "coco_gt": gt}, str(tmp_path / f"{str(i).zfill(12)}.safetensors"))
    store = EmbeddingStore(str(tmp_path), name="coco", load_gts=True, workers=3)
    got, got_gts = store.load(ids, device="cpu")
    want = torch.stack([load_file(store.path(i))["embedding"] for i in ids])
    assert torch.equal(got, want) and got.dtype == torch.float32
    assert all(torch.equal(a, b) for a, b in zip(got_gts, gts))
    again, _ = store.load(ids[::-1], device="cpu")                 # staging buffer reuse
    assert torch.equal(again, want.flip(0)) and torch.equal(got, want)


def test_save_round_trip_and_errors(tmp_path):
    store = EmbeddingStore(str(tmp_path / "out"), name="coco")
    embs = torch.randn(3, 4, 2, 2)
    store.save([1, 2, 3], embs)
    assert sorted(os.listdir(tmp_path / "out")) == [f"{str(i).zfill(12)}.safetensors" for i in (1, 2, 3)]
    assert torch.equal(load_file(store.path(2))["embedding"], embs[1])
    assert torch.equal(store.load([3, 1], device="cpu")[0], embs[[2, 0]])
    save_file({"embedding": torch.zeros(4, 3, 3)}, store.path(9))
    with pytest.raises(ValueError, match="differs"):
        store.load([1, 9], device="cpu")
    save_file({"stage1": torch.zeros(2)}, store.path(10))
    with pytest.raises(KeyError, match="embedding"):
        store.load([10], device="cpu")
    with pytest.raises(FileNotFoundError):
        store.load([77], device="cpu")


def test_writer_matches_safetensors_on_random_tensor_sets(tmp_path):
    g = torch.Generator().manual_seed(5)
    dtypes = [torch.float32, torch.float16, torch.bfloat16, torch.int64, torch.int32, torch.uint8, torch.bool,
              torch.float64, torch.int16, torch.int8]
    for trial in range(25):
        tensors = {}
        for k in range(int(torch.randint(1, 5, (1,), generator=g))):
            dt = dtypes[int(torch.randint(0, len(dtypes), (1,), generator=g))]
            shape = tuple(int(v) for v in torch.randint(0, 5, (int(torch.randint(0, 4, (1,), generator=g)),), generator=g))
            base = torch.randint(0, 100, shape, generator=g)
            tensors[f"t{trial}_{k}" if k else "embedding"] = base.to(dt)
        a, b = tmp_path / "a.safetensors", tmp_path / "b.safetensors"
        save_file(tensors, str(a))
        write_safetensors(str(b), tensors)
        assert a.read_bytes() == b.read_bytes(), {k: (v.dtype, tuple(v.shape)) for k, v in tensors.items()}
