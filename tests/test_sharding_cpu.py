"""Episode sharding across ranks: world_size-2 gloo processes on CPU exercise the host-side logic the
N-GPU bench uses — contiguous episode shards, no data-path collective, max-over-ranks timing, gather on rank 0."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from labelanything_b200.parallel import gather_logits, max_over_ranks, shard_episodes, shard_range
from labelanything_b200.synthetic import make_episode


def test_shard_range_partitions_without_gaps():
    for n in (0, 1, 7, 8, 9, 64):
        for world in (1, 2, 3, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [e - b for b, e in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(4, 2, 2)


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        batch = make_episode(5, 1, 1, 32, seed=0, embeddings=(8, 2))
        mine = shard_episodes(batch, rank, world)
        b, e = shard_range(5, rank, world)
        assert all(torch.equal(mine[k], batch[k][b:e]) for k in batch)
        # stand-in for the per-rank forward: a deterministic function of the local episodes only
        local = mine["embeddings"].flatten(1).sum(1, keepdim=True)
        dist.barrier()
        t = max_over_ranks([10.0 + rank, 3.0 - rank])
        gathered = gather_logits(local)
        if rank == 0:
            full = torch.cat(gathered)
            q.put((t, torch.equal(full, batch["embeddings"].flatten(1).sum(1, keepdim=True))))
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_sharding_round_trip():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    t, same = q.get()
    assert t == [11.0, 3.0] and same
