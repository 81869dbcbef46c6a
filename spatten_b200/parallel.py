"""Head sharding across GPUs -- host-side plumbing only (torch.distributed; NCCL on GPUs, gloo in CPU tests).

Scores, top-k and the KV gather are per head in the reference (kv_cache_token_pruning.py:51,59-69), so a rank
owns a contiguous block of heads, their KV cache and their prune state; token pruning needs no communication.
The single exchange is on the attention output: every rank needs all heads' [b, H*D] before o_proj.
north_star words it as "one all-reduce on the attention output": each rank contributes its own heads in a
zero-padded [N, b, H/N, D] buffer, so the sum is bit-identical to the single-GPU output.  The same exchange
as an all-gather moves half the bytes; both are offered.
"""
from __future__ import annotations

from dataclasses import dataclass

import torch
import torch.distributed as dist


@dataclass(frozen=True)
class HeadShard:
    """Contiguous head block [first, first + count) of rank `rank` out of `world`."""
    rank: int
    world: int
    num_heads: int

    def __post_init__(self):
        if self.num_heads % self.world != 0:
            raise ValueError(f"num_heads {self.num_heads} is not divisible by world size {self.world}")

    @property
    def count(self) -> int:
        return self.num_heads // self.world

    @property
    def first(self) -> int:
        return self.rank * self.count

    def slice(self) -> slice:
        return slice(self.first, self.first + self.count)


def exchange_heads(slots: torch.Tensor, shard: HeadShard, mode: str = "allreduce", group=None) -> torch.Tensor:
    """slots: [N, b, H/N, D]; slot `shard.rank` holds this rank's attention output.

    Returns `slots` with every rank's slot filled.  mode "allreduce": the other slots are zeroed and the buffer
    is summed across ranks (the literal north_star collective); "allgather": all_gather_into_tensor.
    Use `gathered_to_heads` to view the result as [b, H, D].
    """
    if slots.shape[0] != shard.world:
        raise ValueError("slots must have one leading slot per rank")
    if shard.world == 1:
        return slots
    if mode == "allgather":
        dist.all_gather(list(slots.unbind(0)), slots[shard.rank].clone(), group=group)
    elif mode == "allreduce":
        for r in range(shard.world):
            if r != shard.rank:
                slots[r].zero_()
        dist.all_reduce(slots, group=group)
    else:
        raise ValueError(f"unknown exchange mode {mode!r}")
    return slots


def gathered_to_heads(slots: torch.Tensor) -> torch.Tensor:
    """[N, b, H/N, D] -> [b, H, D] with heads in global order (rank-major head blocks)."""
    n, b, hl, d = slots.shape
    return slots.permute(1, 0, 2, 3).reshape(b, n * hl, d)
