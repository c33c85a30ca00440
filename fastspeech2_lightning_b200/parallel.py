"""Host-side data-parallel plumbing (one process per GPU, `torch.distributed`; NCCL on the B200 box,
`gloo` in the CPU tests).  The path shards by utterance: synthesis has no collective, training
has exactly one exchange per step — the all-reduce of the flat gradient (done inside `FusedAdamW.step`)."""
from __future__ import annotations

from typing import Sequence

import torch
import torch.distributed as dist


def shard_utterances(lengths: Sequence[int], rank: int, world: int, batch_size: int) -> list[list[int]]:
    """Length-sorted batches of `batch_size` utterance indices, dealt round-robin to the ranks (config 4:
    256 utterances, sorted-by-length batches of 32).  Every utterance lands on exactly one rank."""
    order = sorted(range(len(lengths)), key=lambda i: (-lengths[i], i))
    batches = [order[i: i + batch_size] for i in range(0, len(order), batch_size)]
    return batches[rank::world]


def allreduce_mean_(flat: torch.Tensor, group=None) -> torch.Tensor:
    """In-place mean over the ranks of a flat gradient buffer (sum all-reduce, then 1/world)."""
    if dist.is_available() and dist.is_initialized():
        world = dist.get_world_size(group)
        if world > 1:
            dist.all_reduce(flat, group=group)
            flat.mul_(1.0 / world)
    return flat


def max_over_ranks(x: float, device="cpu", group=None) -> float:
    """Timing is reported as the maximum over the ranks (bench.py)."""
    t = torch.tensor([x], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t[0])


def sum_over_ranks(x: float, device="cpu", group=None) -> float:
    t = torch.tensor([x], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return float(t[0])
