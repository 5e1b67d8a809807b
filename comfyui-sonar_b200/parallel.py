"""Batch sharding across GPUs (one process per GPU, torch.distributed / NCCL plumbing).

The hot path is per element or per (batch, channel) plane, so it shards by batch with no data-path
collective. Two things need care so that a sharded run equals the un-sharded
reference:

* scale_noise reduces over the WHOLE batch (py/utils.py:100-106): the two device-resident sums are
  all-reduced (2 doubles) between the moments pass and the apply pass;
* random draws: every rank reserves the FULL draw on its (replicated) torch CUDA generator and
  materialises only its slice, using the Philox element mapping (ops.philox_fill(begin=...)).

The only bulk collective is the final gather of the result to the caller's device.
"""

from __future__ import annotations

import contextlib
from dataclasses import dataclass, field
from typing import Iterator, Sequence

import torch
import torch.distributed as dist


@dataclass
class ShardContext:
    rank: int
    world_size: int
    batch_sizes: Sequence[int]  # items held by each rank
    group: object | None = None
    collectives: int = field(default=0)

    @property
    def local_batch(self) -> int:
        return self.batch_sizes[self.rank]

    @property
    def total_batch(self) -> int:
        return sum(self.batch_sizes)

    @property
    def batch_begin(self) -> int:
        return sum(self.batch_sizes[: self.rank])


_ACTIVE: ShardContext | None = None


def active() -> ShardContext | None:
    return _ACTIVE


def split_sizes(total_batch: int, world_size: int) -> list[int]:
    base, rem = divmod(total_batch, world_size)
    return [base + (1 if r < rem else 0) for r in range(world_size)]


@contextlib.contextmanager
def sharded(total_batch: int, *, rank: int | None = None, world_size: int | None = None, group=None) -> Iterator[ShardContext]:
    """Declares that tensors seen by the noise graph hold this rank's slice of a global batch."""
    global _ACTIVE  # noqa: PLW0603
    if rank is None:
        rank = dist.get_rank(group) if dist.is_initialized() else 0
    if world_size is None:
        world_size = dist.get_world_size(group) if dist.is_initialized() else 1
    ctx = ShardContext(rank=rank, world_size=world_size, batch_sizes=split_sizes(total_batch, world_size), group=group)
    prev, _ACTIVE = _ACTIVE, ctx
    try:
        yield ctx
    finally:
        _ACTIVE = prev


def shard(x: torch.Tensor, ctx: ShardContext | None = None) -> torch.Tensor:
    """This rank's contiguous batch slice of a replicated tensor."""
    ctx = ctx or _ACTIVE
    if ctx is None or ctx.world_size == 1:
        return x
    b0 = ctx.batch_begin
    return x[b0 : b0 + ctx.local_batch].contiguous()


def global_count(local_numel: int, sums: torch.Tensor) -> int:
    """All-reduces the device sums when sharded; returns the element count they now cover."""
    ctx = _ACTIVE
    if ctx is None or ctx.world_size == 1:
        return local_numel
    dist.all_reduce(sums, op=dist.ReduceOp.SUM, group=ctx.group)
    ctx.collectives += 1
    if ctx.local_batch == 0:
        return local_numel
    return (local_numel // ctx.local_batch) * ctx.total_batch


def global_draw_geometry(local_shape: Sequence[int]) -> tuple[int, int]:
    """(global numel, element offset of this rank's slice) for a batch-leading tensor shape."""
    ctx = _ACTIVE
    numel = 1
    for s in local_shape:
        numel *= int(s)
    if ctx is None or ctx.world_size == 1 or len(local_shape) == 0:
        return numel, 0
    if int(local_shape[0]) != ctx.local_batch:
        raise RuntimeError(
            f"sharded draw of shape {tuple(local_shape)} does not lead with this rank's batch ({ctx.local_batch})",
        )
    per_item = numel // max(1, ctx.local_batch)
    return per_item * ctx.total_batch, per_item * ctx.batch_begin


def gather(x: torch.Tensor, ctx: ShardContext | None = None, *, dst: int | None = None) -> torch.Tensor | None:
    """Concatenates the batch shards (all_gather, or gather to `dst`). NCCL over NVLink on GPUs."""
    ctx = ctx or _ACTIVE
    if ctx is None or ctx.world_size == 1:
        return x
    x = x.contiguous()
    widest = max(ctx.batch_sizes)
    if x.shape[0] < widest:  # ragged split: pad to the widest shard, trim after the collective
        pad = torch.zeros((widest - x.shape[0], *x.shape[1:]), device=x.device, dtype=x.dtype)
        x = torch.cat((x, pad), dim=0)
    parts = [torch.empty_like(x) for _ in range(ctx.world_size)]
    if dst is None:
        dist.all_gather(parts, x, group=ctx.group)
    else:
        dist.gather(x, parts if ctx.rank == dst else None, dst=dst, group=ctx.group)
    ctx.collectives += 1
    if dst is not None and ctx.rank != dst:
        return None
    return torch.cat([p[:n] for p, n in zip(parts, ctx.batch_sizes)], dim=0)
