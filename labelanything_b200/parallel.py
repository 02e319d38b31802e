"""Episode sharding across the GPUs of one box.

Episodes are independent in `Lam.forward` (every op is batched over the episode dimension), so inference scales by
giving each rank a contiguous slice of the episode batch: weights are replicated, there is NO data-path collective.
`torch.distributed` (NCCL on GPUs, gloo in the CPU tests) is used only for the bookkeeping around it: a barrier
before/after a timed region, the max-over-ranks of device-side timings, and gathering per-rank results.
The reference's only collective is DDP's gradient all-reduce in training (experiment/run.py:122-131,359-361).
"""
from __future__ import annotations

from typing import Dict, List, Tuple

import torch
import torch.distributed as dist


def shard_range(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """[begin, end) of the items owned by `rank`: contiguous, sizes differ by at most one, earlier ranks take the
    remainder (same convention as torch.tensor_split)."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world {rank}/{world}")
    base, rem = divmod(n_items, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def shard_episodes(batch: Dict[str, torch.Tensor], rank: int, world: int) -> Dict[str, torch.Tensor]:
    """Slice every tensor of a `batched_input` dict along the episode dimension (dim 0)."""
    n = next(v.shape[0] for v in batch.values() if torch.is_tensor(v))
    b, e = shard_range(n, rank, world)
    return {k: (v[b:e] if torch.is_tensor(v) and v.dim() > 0 and v.shape[0] == n else v) for k, v in batch.items()}


def max_over_ranks(values: List[float], device=None) -> List[float]:
    """Element-wise max of per-rank scalars (device-side timings) over all ranks; identity without a process group."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return list(values)
    t = torch.tensor(values, dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.tolist()


def gather_logits(local: torch.Tensor) -> List[torch.Tensor] | None:
    """Collect per-rank logits on rank 0 (shapes may differ in the episode dimension and in H x W); None elsewhere."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return [local]
    out: List = [None] * dist.get_world_size() if dist.get_rank() == 0 else None
    dist.gather_object(local.cpu(), out, dst=0)
    return out
