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


class PeerExchange:
    """Mailboxes in every rank's HBM, mapped into all peers with CUDA IPC (csrc/peer.cu): the two
    partial sums of scale_noise travel as direct NVLink stores issued by a kernel, and the consumer
    kernel waits for them on the device. Replaces one NCCL all-reduce (~45 us of launch + protocol
    latency for 16 bytes) per normalisation with ~2 us, with no host involvement."""

    def __init__(self, rank: int, world_size: int, group=None):
        import ctypes

        from . import _native

        if world_size > _native.PEER_MAX_RANKS:
            raise ValueError(f"peer exchange supports up to {_native.PEER_MAX_RANKS} ranks")
        self.rank, self.world_size = rank, world_size
        self.lib = _native.load()
        self.epoch = 0
        local = ctypes.c_void_p()
        _native.check(self.lib.sonar_peer_alloc(ctypes.byref(local)), "sonar_peer_alloc")
        self.local = local.value
        handle = ctypes.create_string_buffer(64)
        _native.check(self.lib.sonar_peer_get_handle(local, handle), "sonar_peer_get_handle")
        handles: list = [None] * world_size
        dist.all_gather_object(handles, bytes(handle.raw), group=group)
        self.mapped = (ctypes.c_void_p * _native.PEER_MAX_RANKS)()
        self._opened = []
        for r, raw in enumerate(handles):
            if r == rank:
                self.mapped[r] = self.local
                continue
            ptr = ctypes.c_void_p()
            _native.check(self.lib.sonar_peer_open_handle(raw, ctypes.byref(ptr)), "sonar_peer_open_handle")
            self.mapped[r] = ptr.value
            self._opened.append(ptr.value)
        dist.barrier(group=group)

    def next_epoch(self) -> float:
        """Reserves the next exchange epoch (for callers whose C-ABI call publishes by itself)."""
        self.epoch += 1
        return float(self.epoch)

    def publish(self, sums: torch.Tensor) -> float:
        """Stores this rank's (sum, sum^2) into every rank's mailbox; returns the epoch to wait for."""
        import ctypes

        from . import _native, ops

        self.epoch += 1
        stream = ctypes.c_void_p(ops._raw_stream(sums.device.index))  # noqa: SLF001
        _native.check(
            self.lib.sonar_peer_publish_sums(self.mapped, self.rank, self.world_size, ctypes.c_void_p(sums.data_ptr()), float(self.epoch), stream),
            "sonar_peer_publish_sums",
        )
        ops.LAUNCH_COUNT += 1
        return float(self.epoch)

    def allreduce_table(self, table: torch.Tensor) -> bool:
        """In-place sum over ranks of a small float64 table through the mailboxes (one launch, no host
        round trip). Returns False when the table is too large for a mailbox slot."""
        import ctypes

        from . import _native, ops

        if table.dtype != torch.float64 or not table.is_contiguous() or table.numel() > _native.PEER_TABLE_MAX:
            return False
        self.epoch += 1
        stream = ctypes.c_void_p(ops._raw_stream(table.device.index))  # noqa: SLF001
        _native.check(
            self.lib.sonar_peer_allreduce_table(
                self.mapped, self.rank, self.world_size, ctypes.c_void_p(table.data_ptr()), table.numel(), float(self.epoch), stream,
            ),
            "sonar_peer_allreduce_table",
        )
        ops.LAUNCH_COUNT += 1
        return True

    def close(self) -> None:
        for ptr in self._opened:
            self.lib.sonar_peer_close_handle(ptr)
        self._opened = []
        if self.local:
            self.lib.sonar_peer_free(self.local)
            self.local = None


_PEERS: list = []  # [(process group object, world size, rank, PeerExchange | None)]


def _group_object(group):
    """The process-group OBJECT a `group` argument stands for (None = the default group). The cache below holds it
    strongly and compares by identity: a group that was destroyed and re-created is a different object (an id() key could
    be reused by the new one and hand out mailboxes mapped for ranks that no longer exist)."""
    if group is not None or not dist.is_initialized():
        return group
    try:
        return dist.distributed_c10d._get_default_group()  # noqa: SLF001
    except Exception:  # noqa: BLE001
        return None


def peer_exchange(rank: int, world_size: int, group=None) -> PeerExchange | None:
    """One PeerExchange per process group (created collectively on first use), or None when the
    ranks cannot share memory (no NCCL / CUDA, or SONAR_B200_NO_PEER set)."""
    import os

    pg = _group_object(group)
    world = dist.get_world_size(group) if dist.is_initialized() else world_size
    for entry in _PEERS:
        if entry[0] is pg and entry[1] == world and entry[2] == rank:
            return entry[3]
    ok = (
        world_size > 1
        and dist.is_initialized()
        and torch.cuda.is_available()
        and dist.get_backend(group) == "nccl"
        and os.environ.get("SONAR_B200_NO_PEER") is None
    )
    if len(_PEERS) > 8:  # groups that went away: release their mailboxes
        for old in _PEERS[:-8]:
            if old[3] is not None:
                old[3].close()
        del _PEERS[:-8]
    exchange = PeerExchange(rank, world_size, group) if ok else None
    _PEERS.append((pg, world, rank, exchange))
    return exchange


@dataclass
class ShardContext:
    rank: int
    world_size: int
    batch_sizes: Sequence[int]  # items held by each rank
    group: object | None = None
    collectives: int = field(default=0)
    peers: PeerExchange | None = None

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
    if total_batch < world_size:
        # a rank without items would skip the statistics exchange its peers wait for (and stop advancing its replicated
        # generator): refuse the split instead of hanging
        raise ValueError(f"cannot shard a batch of {total_batch} over {world_size} ranks: every rank needs at least one item")
    ctx = ShardContext(rank=rank, world_size=world_size, batch_sizes=split_sizes(total_batch, world_size), group=group)
    ctx.peers = peer_exchange(rank, world_size, group)
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


def allreduce_table(table: torch.Tensor) -> None:
    """Sums a small float64 table over the ranks of the active sharding context, in place: peer
    mailboxes over NVLink when available, one NCCL all-reduce otherwise."""
    ctx = _ACTIVE
    if ctx is None or ctx.world_size == 1:
        return
    if ctx.peers is not None and ctx.peers.allreduce_table(table):
        return
    dist.all_reduce(table, op=dist.ReduceOp.SUM, group=ctx.group)
    ctx.collectives += 1


_BARRIER_TABLE: dict = {}


def device_barrier() -> bool:
    """Device-side rendezvous of the ranks of the active sharding context: every GPU's stream stalls until
    all ranks have reached this point (one peer-mailbox exchange over NVLink, no host synchronisation, so
    each host keeps enqueueing behind it). Aligns the GPUs to NVLink latency rather than to the skew of the
    host threads leaving an NCCL barrier. Returns False (and does nothing) without peer memory."""
    ctx = _ACTIVE
    if ctx is None or ctx.world_size == 1 or ctx.peers is None:
        return False
    dev = torch.cuda.current_device()
    table = _BARRIER_TABLE.get(dev)
    if table is None:
        table = _BARRIER_TABLE[dev] = torch.zeros(1, device=torch.device("cuda", dev), dtype=torch.float64)
    return ctx.peers.allreduce_table(table)


def global_numel(local_numel: int) -> int:
    """Element count of the un-sharded tensor behind a local tensor of `local_numel` elements."""
    ctx = _ACTIVE
    if ctx is None or ctx.world_size == 1 or ctx.local_batch == 0:
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
