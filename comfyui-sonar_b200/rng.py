"""Random draws for the noise generators: device Philox, torch-stream compatible.

Every base draw of the in-scope generators goes through `normal` / `uniform` below. On the product
path they reserve the draw on torch's CUDA generator (so the global RNG advances exactly as if
torch.randn(device='cuda') had run) and fill it with our Philox kernel; under `parallel.sharded`
each rank fills only its batch slice of the global draw.

Tests pin parity against the reference by *injecting* the base tensors: `injected(list_of_tensors)`
makes the next draws return the supplied tensors (checked for shape) instead of Philox values, the
same way the golden fixtures were recorded from the reference (tests/golden/make_golden.py).
"""

from __future__ import annotations

import contextlib
import math
from typing import Iterator, Sequence

import torch

from . import ops, parallel

_INJECT: list[torch.Tensor] | None = None


@contextlib.contextmanager
def injected(draws: Sequence[torch.Tensor]) -> Iterator[list]:
    """Feed recorded base draws, in call order, to the generators (test / parity harness only)."""
    global _INJECT  # noqa: PLW0603
    prev, _INJECT = _INJECT, list(draws)
    try:
        yield _INJECT
    finally:
        _INJECT = prev


def _take_injected(shape: Sequence[int], device: torch.device, dtype: torch.dtype) -> torch.Tensor | None:
    if _INJECT is None:
        return None
    if not _INJECT:
        raise RuntimeError("rng.injected: ran out of recorded draws")
    t = _INJECT.pop(0)
    if tuple(t.shape) != tuple(shape):
        raise RuntimeError(f"rng.injected: next recorded draw has shape {tuple(t.shape)}, generator asked for {tuple(shape)}")
    return t.to(device=device, dtype=dtype).contiguous().clone()


_PENDING: list | None = None


@contextlib.contextmanager
def batched() -> Iterator[None]:
    """Draws requested inside the block are reserved on the generator immediately (order and offsets
    exactly as without batching) but materialised together, by ONE launch, when the block exits.
    The returned tensors must not be read by a kernel before that."""
    global _PENDING  # noqa: PLW0603
    if _PENDING is not None:  # nested: the outermost block flushes
        yield
        return
    _PENDING = []
    try:
        yield
    finally:
        pending, _PENDING = _PENDING, None
        if pending:
            ops.philox_fill_batch(pending)


@contextlib.contextmanager
def _unbatched() -> Iterator[None]:
    global _PENDING  # noqa: PLW0603
    saved, _PENDING = _PENDING, None
    try:
        yield
    finally:
        _PENDING = saved


def _fill(shape, device, dtype, generator, kind, p0, p1, batch_sharded) -> torch.Tensor:
    out = torch.empty(tuple(shape), device=device, dtype=dtype)
    if out.numel() == 0:
        return out
    floats_per_el = 2 if dtype == torch.complex64 else 1
    total, begin = parallel.global_draw_geometry(shape) if batch_sharded else (out.numel(), 0)
    draw = ops.reserve_draw(total * floats_per_el, out.device, generator)
    if _PENDING is not None:
        _PENDING.append((draw, out, kind, p0, p1, begin * floats_per_el))
        return out
    return ops.philox_fill(draw, out, kind=kind, p0=p0, p1=p1, begin=begin * floats_per_el)


def normal(
    shape: Sequence[int],
    *,
    device: torch.device,
    dtype: torch.dtype = torch.float32,
    generator: torch.Generator | None = None,
    std: float = 1.0,
    batch_sharded: bool = True,
) -> torch.Tensor:
    """torch.randn(shape, device=device, dtype=dtype) * std, from the torch CUDA Philox stream.

    batch_sharded=False marks a draw that is shared by the whole batch (replicated on every rank)."""
    inj = _take_injected(shape, device, dtype)
    if inj is not None:
        return inj
    if dtype not in {torch.float32, torch.complex64}:
        with _unbatched():  # the cast reads the draw right away
            return normal(
                shape, device=device, dtype=torch.float32, generator=generator, std=std, batch_sharded=batch_sharded,
            ).to(dtype)
    eff = std / math.sqrt(2.0) if dtype == torch.complex64 else std
    return _fill(shape, device, dtype, generator, "normal", 0.0, float(eff), batch_sharded)


def uniform(
    shape: Sequence[int],
    *,
    device: torch.device,
    dtype: torch.dtype = torch.float32,
    generator: torch.Generator | None = None,
    low: float = 0.0,
    high: float = 1.0,
    batch_sharded: bool = True,
) -> torch.Tensor:
    """torch.empty(shape).uniform_(low, high) / torch.rand on CUDA."""
    inj = _take_injected(shape, device, dtype)
    if inj is not None:
        return inj
    if dtype != torch.float32:
        with _unbatched():
            return uniform(
                shape, device=device, generator=generator, low=low, high=high, batch_sharded=batch_sharded,
            ).to(dtype)
    return _fill(shape, device, dtype, generator, "uniform", float(low), float(high), batch_sharded)


def host_rand(n: int, generator: torch.Generator | None = None) -> torch.Tensor:
    """torch.rand(n) on the CPU generator -- the host draws that size pyramid levels
    (reference py/noise_generation.py:544-552, :627-629). Injectable like the device draws."""
    inj = _take_injected((n,), torch.device("cpu"), torch.float32)
    if inj is not None:
        return inj
    return torch.rand(n, dtype=torch.float32, generator=generator)
