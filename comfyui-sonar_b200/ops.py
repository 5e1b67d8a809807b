"""Tensor-level wrappers over the C ABI (include/sonar_b200.h).

Every function here takes CUDA tensors, enqueues hand-written kernels on torch's current stream and
returns CUDA tensors. Nothing in this module computes with torch ops on the hot path and nothing
falls back to the CPU: a non-CUDA tensor is an error.
"""

from __future__ import annotations

import ctypes
import math
from typing import NamedTuple, Sequence

import torch

from . import _native
from ._native import SonarStepParams

# enums of include/sonar_b200.h
STEP_EULER, STEP_DPMPP = 0, 1
MODE_CLASSIC, MODE_NEW, MODE_DENOISED = 0, 1, 2
BLEND_IDS = {"lerp": 0, "inject": 1, "subtract_b": 2}
HIST_NONE, HIST_PRESENT, HIST_INIT = 0, 1, 2
NOISE_NONE, NOISE_TENSOR, NOISE_PHILOX, NOISE_PHILOX_NORMALIZED, NOISE_TENSOR_NORMALIZED = 0, 1, 2, 3, 4

LAUNCH_COUNT = 0  # number of kernels launched through this module (bench.py's gpu_launches)


TRACE: list | None = None  # when a list: (name, start_event, end_event) per launch (bench.py roofline)


def _launch(name: str, fn, *args, launches: int = 1) -> None:
    """Calls one C-ABI entry point (which enqueues `launches` kernels on torch's current stream)."""
    global LAUNCH_COUNT  # noqa: PLW0603
    if TRACE is None:
        code = fn(*args)
    else:
        start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        start.record()
        code = fn(*args)
        end.record()
        TRACE.append((name, start, end))
    _native.check(code, name)
    LAUNCH_COUNT += launches


_LAST_DEVICE: int | None = None
_GRID_LIMIT = 0


def set_grid_limit(ctas_per_sm: int) -> None:
    """Co-scheduling hint for the next launches of this thread (sonar_set_grid_limit): at most `ctas_per_sm` CTAs per SM
    for the grid-stride streaming kernels and the batched FFT, 0 = default grids."""
    global _GRID_LIMIT  # noqa: PLW0603
    if ctas_per_sm != _GRID_LIMIT:
        _native.load().sonar_set_grid_limit(int(ctas_per_sm))
        _GRID_LIMIT = ctas_per_sm



def _prepare(*tensors: torch.Tensor | None) -> tuple[ctypes.CDLL, ctypes.c_void_p]:
    """Validates devices, binds the library's runtime to the tensors' device, returns (lib, stream)."""
    global _LAST_DEVICE  # noqa: PLW0603
    dev = None
    for t in tensors:
        if t is None:
            continue
        if not t.is_cuda:
            raise RuntimeError("sonar_b200 kernels take CUDA tensors only (there is no CPU fallback)")
        if not t.is_contiguous():
            raise RuntimeError("sonar_b200 kernels take contiguous tensors")
        if dev is None:
            dev = t.device
        elif t.device != dev:
            raise RuntimeError(f"tensors on different devices: {dev} vs {t.device}")
    if dev is None:
        raise RuntimeError("no tensor given")
    lib = _native.load()
    idx = dev.index if dev.index is not None else torch.cuda.current_device()
    if idx != _LAST_DEVICE:
        _native.check(lib.sonar_set_device(idx), "sonar_set_device")
        _LAST_DEVICE = idx
    return lib, ctypes.c_void_p(_raw_stream(idx))


_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)  # noqa: SLF001  (fast path, ~0.3 us)
if _raw_stream is None:  # pragma: no cover - older torch

    def _raw_stream(idx: int) -> int:
        return torch.cuda.current_stream(idx).cuda_stream


def _ptr(t: torch.Tensor | None) -> ctypes.c_void_p:
    return ctypes.c_void_p(0 if t is None else t.data_ptr())


def _f32(t: torch.Tensor, name: str) -> None:
    if t.dtype != torch.float32:
        raise TypeError(f"{name} must be float32, got {t.dtype}")


# --------------------------------------------------------------------------------------------
# Philox draws
# --------------------------------------------------------------------------------------------
class PhiloxDraw(NamedTuple):
    """Identity of one torch-compatible CUDA draw: what torch.randn(numel) would have produced."""

    seed: int
    offset: int
    grid_blocks: int
    numel: int  # float elements of the draw (2x for complex64)
    counter_offset: int  # how far the draw advances the generator


_POLICY_CACHE: dict[tuple[int, int], tuple[int, int]] = {}


def philox_policy(numel: int) -> tuple[int, int]:
    """(grid_blocks, counter_offset) of ATen's calc_execution_policy for the current device."""
    lib = _native.load()
    grid = ctypes.c_uint32(0)
    inc = ctypes.c_uint64(0)
    _native.check(lib.sonar_philox_policy(int(numel), ctypes.byref(grid), ctypes.byref(inc)), "sonar_philox_policy")
    return grid.value, inc.value


def philox_policy_cached(device_index: int, numel: int) -> tuple[int, int]:
    policy = _POLICY_CACHE.get((device_index, numel))
    if policy is None:
        with torch.cuda.device(device_index):
            policy = philox_policy(numel)
        if len(_POLICY_CACHE) > 4096:
            _POLICY_CACHE.clear()
        _POLICY_CACHE[(device_index, int(numel))] = policy
    return policy


def reserve_draw(numel: int, device: torch.device, generator: torch.Generator | None = None) -> PhiloxDraw:
    """Consumes `numel` values from torch's CUDA generator exactly like an ATen distribution kernel
    would (same offset arithmetic), without launching anything."""
    if not isinstance(device, torch.device):
        device = torch.device(device)
    if device.type != "cuda":
        raise RuntimeError("reserve_draw needs a CUDA device")
    idx = device.index if device.index is not None else torch.cuda.current_device()
    gen = generator if generator is not None else torch.cuda.default_generators[idx]
    grid, inc = philox_policy_cached(idx, numel)
    offset = gen.get_offset()
    if numel > 0:
        gen.set_offset(offset + inc)
    return PhiloxDraw(gen.initial_seed(), offset, grid, int(numel), inc)


def peek_draws(numel: int, device: torch.device, count: int, generator: torch.Generator | None = None) -> list[PhiloxDraw]:
    """The next `count` draws of `numel` values each, as consecutive torch calls would make them, WITHOUT advancing
    the generator (speculative look-ahead: the caller advances it draw by draw as the values are consumed, and
    discards the rest if somebody else used the generator in between)."""
    if not isinstance(device, torch.device):
        device = torch.device(device)
    idx = device.index if device.index is not None else torch.cuda.current_device()
    gen = generator if generator is not None else torch.cuda.default_generators[idx]
    grid, inc = philox_policy_cached(idx, numel)
    offset, seed = gen.get_offset(), gen.initial_seed()
    return [PhiloxDraw(seed, offset + j * inc, grid, int(numel), inc) for j in range(count)]


def philox_fill(
    draw: PhiloxDraw,
    out: torch.Tensor,
    *,
    kind: str = "normal",
    p0: float = 0.0,
    p1: float = 1.0,
    begin: int = 0,
) -> torch.Tensor:
    """Materialises elements [begin, begin + out.numel()) of `draw` into `out` (float32 storage)."""
    lib, stream = _prepare(out)
    count = out.numel() * (2 if out.is_complex() else 1)
    fn = lib.sonar_philox_normal_f32 if kind == "normal" else lib.sonar_philox_uniform_f32
    _launch(
        f"sonar_philox_{kind}_f32", fn,
        _ptr(out), begin, count, draw.numel, draw.seed, draw.offset, draw.grid_blocks, p0, p1, stream,
    )  # fmt: skip
    return out


def philox_fill_batch(items: Sequence[tuple]) -> None:
    """Materialises several draws with one launch per 16 draws of the same seed and device.
    items: (draw, out, kind, p0, p1, begin) as for philox_fill."""
    groups: dict = {}
    for it in items:
        if it[1].numel():
            groups.setdefault((it[1].device, it[0].seed), []).append(it)
    for (_dev, seed), group in groups.items():
        for first in range(0, len(group), _native.FILL_BATCH_MAX):
            chunk = group[first : first + _native.FILL_BATCH_MAX]
            if len(chunk) == 1:
                draw, out, kind, p0, p1, begin = chunk[0]
                philox_fill(draw, out, kind=kind, p0=p0, p1=p1, begin=begin)
                continue
            b = _native.SonarFillBatch()
            b.n, b.seed = len(chunk), seed
            for d, (draw, out, kind, p0, p1, begin) in zip(b.draws, chunk):
                d.out, d.begin = out.data_ptr(), begin
                d.count = out.numel() * (2 if out.is_complex() else 1)
                d.numel_total, d.offset, d.grid_blocks = draw.numel, draw.offset, draw.grid_blocks
                d.kind, d.p0, d.p1 = (0 if kind == "normal" else 1), p0, p1
            lib, stream = _prepare(*(it[1] for it in chunk))
            _launch("sonar_philox_fill_batch", lib.sonar_philox_fill_batch, ctypes.byref(b), stream)


def randn(
    shape: Sequence[int],
    *,
    device: torch.device,
    dtype: torch.dtype = torch.float32,
    generator: torch.Generator | None = None,
) -> torch.Tensor:
    """torch.randn(shape, device='cuda', dtype=float32|complex64), bit-exact, from our kernel."""
    if dtype not in {torch.float32, torch.complex64}:
        raise TypeError(f"sonar_b200.randn supports float32 and complex64, got {dtype}")
    out = torch.empty(tuple(shape), device=device, dtype=dtype)
    if out.numel() == 0:
        return out
    is_c = dtype == torch.complex64
    draw = reserve_draw(out.numel() * (2 if is_c else 1), out.device, generator)
    # complex normal: real/imag each N(0, 1/2) (ATen normal_ on view_as_real with std/sqrt(2))
    std = float(1.0 / math.sqrt(2.0)) if is_c else 1.0
    return philox_fill(draw, out, kind="normal", p0=0.0, p1=std)


def rand(
    shape: Sequence[int],
    *,
    device: torch.device,
    generator: torch.Generator | None = None,
    low: float = 0.0,
    high: float = 1.0,
) -> torch.Tensor:
    """torch.rand / torch.empty(shape).uniform_(low, high) on CUDA, bit-exact."""
    out = torch.empty(tuple(shape), device=device, dtype=torch.float32)
    if out.numel() == 0:
        return out
    draw = reserve_draw(out.numel(), out.device, generator)
    return philox_fill(draw, out, kind="uniform", p0=float(low), p1=float(high))


# --------------------------------------------------------------------------------------------
# moments / scale_noise
# --------------------------------------------------------------------------------------------
def new_sums(device: torch.device) -> torch.Tensor:
    return torch.zeros(2, device=device, dtype=torch.float64)


# Producer kernels (spectral, pyramid, perlin, blend, axpby) reduce {sum, sum^2} of what they write in
# the same launch. The double[2] slots come from a per-device ring that needs neither an allocation nor
# a memset per launch: every slot starts zeroed, and the producer that fills slot i also clears slot
# i+1 (stream order makes that safe: the last reader of slot i+1 was enqueued a whole ring ago).
# The slot travels with the output tensor as the attribute `_sonar_sums = (slot, tensor._version)`:
# scale_noise uses it instead of a moments pass as long as nobody has modified the tensor since (torch
# bumps _version on in-place ops; our own in-place kernels call drop_sums).
_SUMS_RING: dict = {}
_SUMS_RING_SLOTS = 256


def _ring(device: torch.device) -> list:
    """The ring of the (device, current stream) pair: the "producer of slot i clears slot i+1" argument is a
    stream-order argument, so every stream gets its own ring. [slots, position, base pointer, launches so far]."""
    idx = device.index if device.index is not None else torch.cuda.current_device()
    key = (idx, _raw_stream(idx))
    ring = _SUMS_RING.get(key)
    if ring is None:
        base = torch.zeros((_SUMS_RING_SLOTS, 2), device=device, dtype=torch.float64)
        ring = _SUMS_RING[key] = [list(base.unbind(0)), 0, base.data_ptr(), 0]
    return ring


def sums_slot(device: torch.device) -> tuple[torch.Tensor, int, int]:
    """(slot tensor, its pointer, pointer of the slot to clear) -- call sums_advance(device) after the
    launch succeeded (a slot that was handed out but never written must stay zero)."""
    slots, i, base_ptr, _ = _ring(device)
    return slots[i], base_ptr + 16 * i, base_ptr + 16 * ((i + 1) % _SUMS_RING_SLOTS)


def sums_advance(device: torch.device) -> list:
    ring = _ring(device)
    ring[1] = (ring[1] + 1) % _SUMS_RING_SLOTS
    ring[3] += 1
    return ring


def _sums_written(out: torch.Tensor, slot: torch.Tensor) -> torch.Tensor:
    """After a producer launch: advance the ring and tag `out` (an empty tensor launches nothing, so
    its slot stays untouched and is handed out again)."""
    if out.numel() == 0:
        return out
    ring = sums_advance(out.device)
    return attach_sums(out, slot, ring)


def attach_sums(t: torch.Tensor, slot: torch.Tensor, ring: list | None = None) -> torch.Tensor:
    # (slot, tensor version, ring, ring launch count when the slot was filled): the slot is recycled -- cleared by the
    # producer of the previous slot, refilled by a later one -- once the ring has gone round, so the tag expires then
    t._sonar_sums = (slot, t._version, ring, 0 if ring is None else ring[3])  # noqa: SLF001
    return t


def attached_sums(t: torch.Tensor) -> torch.Tensor | None:
    tag = getattr(t, "_sonar_sums", None)
    if tag is None or tag[1] != t._version:  # noqa: SLF001
        return None
    ring = tag[2]
    if ring is not None and ring[3] - tag[3] >= _SUMS_RING_SLOTS - 1:
        return None  # stale: the ring wrapped since this tensor was produced
    return tag[0]


def reshape_keep_sums(t: torch.Tensor, shape) -> torch.Tensor:
    """t.reshape(shape); a view of the same elements keeps the producer's statistics."""
    out = t.reshape(shape)
    slot = attached_sums(t)
    if slot is not None and out is not t and out.data_ptr() == t.data_ptr():
        tag = t._sonar_sums  # noqa: SLF001
        out._sonar_sums = (slot, out._version, tag[2], tag[3])  # noqa: SLF001
    return out


def drop_sums(t: torch.Tensor) -> None:
    if getattr(t, "_sonar_sums", None) is not None:
        t._sonar_sums = None  # noqa: SLF001


def moments(x: torch.Tensor, sums: torch.Tensor | None = None) -> torch.Tensor:
    """Accumulates (sum, sum of squares) of x into the device double[2] `sums`."""
    _f32(x, "x")
    if sums is None:
        sums = new_sums(x.device)
    lib, stream = _prepare(x, sums)
    _launch("sonar_moments_f32", lib.sonar_moments_f32, _ptr(x), x.numel(), _ptr(sums), stream)
    return sums


def philox_normal_moments(draw: PhiloxDraw, *, begin: int, count: int, sums: torch.Tensor) -> torch.Tensor:
    lib, stream = _prepare(sums)
    _launch("sonar_philox_normal_moments", lib.sonar_philox_normal_moments, begin, count, draw.numel, draw.seed, draw.offset, draw.grid_blocks, _ptr(sums), stream)
    return sums


def philox_normal_fill_moments(draw: PhiloxDraw, out: torch.Tensor, sums: torch.Tensor, *, begin: int = 0) -> torch.Tensor:
    """Materialises the slice of `draw` into `out` and writes its (sum, sum^2) into `sums`: one pass."""
    lib, stream = _prepare(out, sums)
    _launch(
        "sonar_philox_normal_fill_moments_f32", lib.sonar_philox_normal_fill_moments_f32,
        _ptr(out), begin, out.numel(), draw.numel, draw.seed, draw.offset, draw.grid_blocks, _ptr(sums), stream,
    )  # fmt: skip
    return out


def philox_normal_moments_batch(
    draw: PhiloxDraw, offsets: Sequence[int], *, begin: int, count: int, device: torch.device,
) -> torch.Tensor:
    """(len(offsets), 2) float64 = (sum, sum^2) of elements [begin, begin+count) of the normal draws
    that share `draw`'s seed / geometry and start at the given generator offsets. One launch (per 64
    draws): the statistics of every remaining ancestral-noise draw of a sampler run."""
    sums = torch.empty((len(offsets), 2), device=device, dtype=torch.float64)
    lib, stream = _prepare(sums)
    arr = (ctypes.c_uint64 * len(offsets))(*offsets)
    _launch(
        "sonar_philox_normal_moments_batch", lib.sonar_philox_normal_moments_batch,
        arr, len(offsets), begin, count, draw.numel, draw.seed, draw.grid_blocks, _ptr(sums), stream,
        launches=(len(offsets) + 63) // 64,
    )  # fmt: skip
    return sums


def norm_decisions(sums: torch.Tensor, count: int, *, threshold_std_devs: float = 2.5) -> torch.Tensor:
    """(K, 4) float32 = (mean, std, subtract flag, divide flag) of scale_noise for each row of `sums` (K, 2)."""
    k = sums.shape[0]
    out = torch.empty((k, 4), device=sums.device, dtype=torch.float32)
    lib, stream = _prepare(sums, out)
    _launch("sonar_norm_decisions", lib.sonar_norm_decisions, _ptr(sums), k, int(count), float(threshold_std_devs), _ptr(out), stream)
    return out


def scale_noise_apply(
    x: torch.Tensor,
    sums: torch.Tensor,
    count: int,
    factor: float = 1.0,
    *,
    threshold_std_devs: float = 2.5,
    out: torch.Tensor | None = None,
) -> torch.Tensor:
    """The conditional centre / rescale / multiply of scale_noise, decided on the device."""
    _f32(x, "x")
    out = x if out is None else out
    lib, stream = _prepare(x, out, sums)
    drop_sums(out)
    _launch("sonar_scale_noise_f32", lib.sonar_scale_noise_f32, _ptr(x), _ptr(out), x.numel(), _ptr(sums), int(count), float(factor), float(threshold_std_devs), stream)
    return out


def scale_noise_apply_peers(x: torch.Tensor, peers, epoch: float, count: int, factor: float = 1.0, *, threshold_std_devs: float = 2.5):
    """scale_noise whose global statistics arrive through the peer mailboxes (parallel.PeerExchange)."""
    _f32(x, "x")
    lib, stream = _prepare(x)
    drop_sums(x)
    _launch(
        "sonar_scale_noise_peers_f32", lib.sonar_scale_noise_peers_f32,
        _ptr(x), _ptr(x), x.numel(), ctypes.c_void_p(peers.local), peers.world_size, float(epoch), int(count),
        float(factor), float(threshold_std_devs), stream,
    )  # fmt: skip
    return x


# --------------------------------------------------------------------------------------------
# fused sonar step
# --------------------------------------------------------------------------------------------
_STEP_FN = None


def launch_step(params_ref, device_index: int | None) -> None:
    """sonar_step_f32 on torch's current stream of `device_index`. `params_ref` is ctypes.byref() of
    a SonarStepParams whose pointers the caller has validated (float32, contiguous, that device).
    The leanest path into the library: the sampler calls this once per model evaluation."""
    global _STEP_FN, _LAST_DEVICE, LAUNCH_COUNT  # noqa: PLW0603
    if _STEP_FN is None:
        _STEP_FN = _native.load().sonar_step_f32
    if device_index is None:
        device_index = torch.cuda.current_device()
    if device_index != _LAST_DEVICE:
        _native.check(_native.load().sonar_set_device(device_index), "sonar_set_device")
        _LAST_DEVICE = device_index
    stream = ctypes.c_void_p(_raw_stream(device_index))
    if TRACE is None:
        code = _STEP_FN(params_ref, stream)
    else:
        start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        start.record()
        code = _STEP_FN(params_ref, stream)
        end.record()
        TRACE.append(("sonar_step_f32", start, end))
    if code:
        _native.check(code, "sonar_step_f32")
    LAUNCH_COUNT += 1


def sonar_step(params: SonarStepParams, *tensors: torch.Tensor | None) -> None:
    """Launches the fused step; `tensors` are the live tensors behind the pointers (validation)."""
    _prepare(*tensors)
    launch_step(ctypes.byref(params), next(t for t in tensors if t is not None).device.index)


# --------------------------------------------------------------------------------------------
# element-wise combinators
# --------------------------------------------------------------------------------------------
def scale(x: torch.Tensor, factor: float) -> torch.Tensor:
    """x *= factor (in place) through the axpby kernel."""
    _f32(x, "x")
    lib, stream = _prepare(x)
    drop_sums(x)
    _launch("sonar_axpby_f32", lib.sonar_axpby_f32, _ptr(x), float(factor), None, 0.0, _ptr(x), x.numel(), None, None, stream)
    return x


def axpby(a: torch.Tensor, alpha: float, b: torch.Tensor | None, beta: float = 1.0, out: torch.Tensor | None = None):
    """out = a * alpha + b * beta; the moments of `out` ride along (attached_sums)."""
    _f32(a, "a")
    out = torch.empty_like(a) if out is None else out
    lib, stream = _prepare(a, b, out)
    slot, slot_ptr, clear_ptr = sums_slot(out.device)
    _launch("sonar_axpby_f32", lib.sonar_axpby_f32, _ptr(a), float(alpha), _ptr(b), float(beta), _ptr(out), a.numel(), slot_ptr, clear_ptr, stream)
    return _sums_written(out, slot)


def affine(x: torch.Tensor, pre_add: float, mul: float, post_add: float, out: torch.Tensor | None = None):
    """out = ((x + pre_add) * mul) + post_add, rounding after each step like the eager op chain."""
    _f32(x, "x")
    out = x if out is None else out
    lib, stream = _prepare(x, out)
    drop_sums(out)
    _launch("sonar_affine_f32", lib.sonar_affine_f32, _ptr(x), _ptr(out), x.numel(), float(pre_add), float(mul), float(post_add), stream)
    return out


def divide_scalar(x: torch.Tensor, divisor: float) -> torch.Tensor:
    """x /= divisor (in place), IEEE division like Tensor.div_(scalar)."""
    _f32(x, "x")
    lib, stream = _prepare(x)
    drop_sums(x)
    _launch("sonar_div_scalar_f32", lib.sonar_div_scalar_f32, _ptr(x), _ptr(x), x.numel(), float(divisor), stream)
    return x


def scale_by_std(x: torch.Tensor, sums: torch.Tensor, count: int, scale: float) -> torch.Tensor:
    """x *= scale / std(x) (in place), std from device sums."""
    _f32(x, "x")
    lib, stream = _prepare(x, sums)
    drop_sums(x)
    _launch("sonar_scale_by_std_f32", lib.sonar_scale_by_std_f32, _ptr(x), _ptr(x), x.numel(), _ptr(sums), int(count), float(scale), stream)
    return x


def blend(a: torch.Tensor, b: torch.Tensor, t, *, mode: str = "lerp", out: torch.Tensor | None = None) -> torch.Tensor:
    """BLENDING_MODES[mode](a, b, t); t is a python scalar, a one-element tensor or a full tensor."""
    _f32(a, "a")
    _f32(b, "b")
    if a.shape != b.shape:
        raise ValueError(f"blend: shape mismatch {tuple(a.shape)} vs {tuple(b.shape)}")
    t_tensor = None
    t_scalar = 0.0
    if isinstance(t, torch.Tensor):
        if t.numel() == 1:
            t_scalar = float(t.item())
        else:
            t_tensor = t.expand_as(a).contiguous() if t.shape != a.shape else t
            _f32(t_tensor, "t")
    else:
        t_scalar = float(t)
    out = torch.empty_like(a) if out is None else out
    lib, stream = _prepare(a, b, t_tensor, out)
    slot, slot_ptr, clear_ptr = sums_slot(out.device)
    _launch("sonar_blend_f32", lib.sonar_blend_f32, _ptr(a), _ptr(b), _ptr(t_tensor), t_scalar, _ptr(out), a.numel(), BLEND_IDS[mode], slot_ptr, clear_ptr, stream)
    return _sums_written(out, slot)


def composite(dst: torch.Tensor, src: torch.Tensor, mask: torch.Tensor, out: torch.Tensor | None = None):
    """dst * (1 - mask) + src * mask, mask (B, 1, H, W) broadcast over channels."""
    _f32(dst, "dst")
    b = dst.shape[0]
    hw = dst.shape[-2] * dst.shape[-1]
    channels = dst.numel() // (b * hw)
    if mask.numel() != b * hw:
        raise ValueError("composite: mask must have shape (batch, 1, H, W)")
    out = torch.empty_like(dst) if out is None else out
    lib, stream = _prepare(dst, src, mask, out)
    drop_sums(out)
    _launch("sonar_composite_f32", lib.sonar_composite_f32, _ptr(dst), _ptr(src), _ptr(mask), _ptr(out), b, channels, hw, stream)
    return out


def powerlaw(x: torch.Tensor, alpha: float, *, use_sign: bool, out: torch.Tensor | None = None) -> torch.Tensor:
    _f32(x, "x")
    out = torch.empty_like(x) if out is None else out
    lib, stream = _prepare(x, out)
    drop_sums(out)
    _launch("sonar_powerlaw_f32", lib.sonar_powerlaw_f32, _ptr(x), _ptr(out), x.numel(), float(alpha), int(use_sign), stream)
    return out


def _range_scratch(items: int, device: torch.device) -> torch.Tensor:
    lib = _native.load()
    return torch.empty(max(1, lib.sonar_item_range_scratch_bytes(items)), device=device, dtype=torch.uint8)


def div_item_max(x: torch.Tensor, *, use_abs: bool = True) -> torch.Tensor:
    """x / amax(|x|) over everything but the leading dim, in place."""
    _f32(x, "x")
    items = x.shape[0]
    scratch = _range_scratch(items, x.device)
    lib, stream = _prepare(x, scratch)
    drop_sums(x)
    _launch("sonar_item_div_max_f32", lib.sonar_item_div_max_f32, _ptr(x), _ptr(x), items, x.numel() // items, int(use_abs), _ptr(scratch), stream, launches=2)
    return x


def minmax_rescale(x: torch.Tensor, target_min: float, target_max: float, *, eps: float = 1e-7) -> torch.Tensor:
    _f32(x, "x")
    items = x.shape[0]
    out = torch.empty_like(x)
    scratch = _range_scratch(items, x.device)
    lib, stream = _prepare(x, out, scratch)
    _launch("sonar_item_minmax_rescale_f32", lib.sonar_item_minmax_rescale_f32, _ptr(x), _ptr(out), items, x.numel() // items, float(target_min), float(target_max), float(eps),
            _ptr(scratch), stream, launches=2)
    return out


# --------------------------------------------------------------------------------------------
# pyramid / resample / perlin
# --------------------------------------------------------------------------------------------
RESAMPLE_IDS = {"bilinear": 0, "nearest-exact": 1, "area": 2}


def pyramid_accumulate(
    base: torch.Tensor | None,
    levels: Sequence[torch.Tensor],
    weights: Sequence[float],
    *,
    out_hw: tuple[int, int],
    mode: str = "bilinear",
    base_scale: float = 1.0,
    out: torch.Tensor | None = None,
) -> torch.Tensor:
    """out = base_scale*base + sum_i weights[i] * resample(levels[i] -> out_hw), one pass."""
    if mode not in RESAMPLE_IDS:
        raise NotImplementedError(f"resample mode {mode!r} has no kernel (supported: {', '.join(RESAMPLE_IDS)})")
    if len(levels) > _native.PYRAMID_MAX_LEVELS:
        raise ValueError("too many pyramid levels")
    H, W = out_hw
    ref = base if base is not None else levels[0]
    planes = ref.numel() // (ref.shape[-2] * ref.shape[-1])
    if out is None:
        out = torch.empty((*ref.shape[:-2], H, W), device=ref.device, dtype=torch.float32)
    p = _native.SonarPyramidParams()
    p.out, p.base = out.data_ptr(), (0 if base is None else base.data_ptr())
    for i, (lv, wgt) in enumerate(zip(levels, weights)):
        _f32(lv, "level")
        if lv.numel() // (lv.shape[-2] * lv.shape[-1]) != planes:
            raise ValueError("pyramid level plane count mismatch")
        p.levels[i] = lv.data_ptr()
        p.level_h[i], p.level_w[i] = lv.shape[-2], lv.shape[-1]
        p.weights[i] = float(wgt)
    p.planes, p.H, p.W = planes, H, W
    p.n_levels, p.mode, p.base_scale = len(levels), RESAMPLE_IDS[mode], float(base_scale)
    lib, stream = _prepare(out, base, *levels)
    slot, p.sums, p.sums_clear = sums_slot(out.device)
    _launch("sonar_pyramid_accum_f32", lib.sonar_pyramid_accum_f32, ctypes.byref(p), stream)
    return _sums_written(out, slot)


def resample(x: torch.Tensor, height: int, width: int, *, mode: str = "bilinear") -> torch.Tensor:
    """F.interpolate(x, size=(height, width), mode=mode) for bilinear / nearest-exact / area."""
    _f32(x, "x")
    return pyramid_accumulate(None, [x.contiguous()], [1.0], out_hw=(height, width), mode=mode)


def perlin_accumulate(
    base: torch.Tensor | None,
    angles: Sequence[torch.Tensor],
    *,
    shape: tuple[int, int, int, int],
    div_fac: float = 2.0,
    blend_mode: str = "lerp",
) -> torch.Tensor:
    B, C, H, W = shape
    if len(angles) > _native.PERLIN_MAX_ITERS:
        raise ValueError("too many perlin iterations")
    out = torch.empty(shape, device=angles[0].device if angles else base.device, dtype=torch.float32)
    p = _native.SonarPerlinParams()
    p.out, p.base = out.data_ptr(), (0 if base is None else base.data_ptr())
    for i, a in enumerate(angles):
        _f32(a, "angles")
        if tuple(a.shape) != (C, H + 1, W + 1):
            raise ValueError(f"perlin angle grid must be {(C, H + 1, W + 1)}, got {tuple(a.shape)}")
        p.angles[i] = a.data_ptr()
    p.B, p.C, p.H, p.W = B, C, H, W
    p.iterations, p.blend_mode, p.div_fac = len(angles), BLEND_IDS[blend_mode], float(div_fac)
    lib, stream = _prepare(out, base, *angles)
    slot, p.sums, p.sums_clear = sums_slot(out.device)
    _launch("sonar_perlin_accum_f32", lib.sonar_perlin_accum_f32, ctypes.byref(p), stream)
    return _sums_written(out, slot)


# --------------------------------------------------------------------------------------------
# RNG-fused pyramid / Perlin / blend (csrc/noise_mix.cu)
# --------------------------------------------------------------------------------------------
TERM_PYRAMID, TERM_PERLIN = 1, 2
MIX_MAX_TABLES, PYRAMID_MAX_LEVELS = _native.MIX_MAX_TABLES, _native.PYRAMID_MAX_LEVELS
MIX_MAX_ELEMENTS = (1 << 32) - 1  # the fused kernel indexes the global tensor with 32 bits


def perlin_tables(draws: Sequence[PhiloxDraw], channels: int, height: int, width: int, *, blend_mode: str, device) -> list[torch.Tensor]:
    """(C, H, W) Perlin stencil of every iteration; the corner angles are the uniform [0, 2 pi) draws `draws`
    (C (H+1) (W+1) values each), regenerated in registers. One launch."""
    n = len(draws)
    if n > _native.MIX_MAX_TABLES:
        raise ValueError("too many perlin iterations for the fused kernel")
    tables = torch.empty((n, channels, height, width), device=device, dtype=torch.float32)
    lib, stream = _prepare(tables)
    ptrs = (ctypes.c_void_p * n)(*(tables[i].data_ptr() for i in range(n)))
    offsets = (ctypes.c_uint64 * n)(*(d.offset for d in draws))
    grids = (ctypes.c_uint32 * n)(*(d.grid_blocks for d in draws))
    _launch(
        "sonar_perlin_tables_f32", lib.sonar_perlin_tables_f32,
        ptrs, offsets, grids, n, draws[0].seed, channels, height, width, BLEND_IDS[blend_mode], stream,
    )  # fmt: skip
    return list(tables.unbind(0))


def _fill_term(dst: "_native.SonarMixTerm", term: dict | None, keep: list) -> None:
    if term is None:
        dst.kind = 0
        return
    dst.kind = term["kind"]
    dst.base_offset = term["base"].offset
    dst.full_level = -1
    if term["kind"] == TERM_PYRAMID:
        levels = term["levels"]
        if len(levels) > _native.PYRAMID_MAX_LEVELS:
            raise ValueError("too many pyramid levels")
        dst.n_levels, dst.mode, dst.base_scale = len(levels), RESAMPLE_IDS[term["mode"]], float(term.get("base_scale", 1.0))
        for i, (lv, wgt) in enumerate(levels):
            dst.weights[i] = float(wgt)
            if isinstance(lv, PhiloxDraw):  # THE full-size level, regenerated from the stream
                if dst.full_level >= 0:
                    raise ValueError("at most one full-size pyramid level can be fused")
                dst.full_level, dst.level_offset[i], dst.levels[i] = i, lv.offset, 0
            else:
                _f32(lv, "level")
                keep.append(lv)
                dst.levels[i], dst.level_h[i], dst.level_w[i] = lv.data_ptr(), lv.shape[-2], lv.shape[-1]
    elif term["kind"] == TERM_PERLIN:
        tables = term["tables"]
        dst.iterations, dst.div_fac = len(tables), float(term["div_fac"])
        dst.uniform_from, dst.uniform_to = 0.0, 1.0
        for i, t in enumerate(tables):
            keep.append(t)
            dst.tables[i] = t.data_ptr()


def noise_mix(shape: Sequence[int], term_a: dict, term_b: dict | None = None, *, begin: int = 0, blend_mode: str = "lerp",
              blend_t: float = 0.5, device) -> torch.Tensor:  # fmt: skip
    """blend(mode, A, B, t) -- or A alone -- for a (B, C, H, W) tensor whose terms are pyramid / Perlin noise (or plain
    draws) regenerated from the Philox stream in registers. A term is a dict: kind, base (PhiloxDraw of the full-size
    base draw), and for TERM_PYRAMID levels = [(PhiloxDraw | coarse tensor, weight)], mode; for TERM_PERLIN tables
    (from perlin_tables), div_fac. `begin` = element offset of this rank's slice in the global draw."""
    batch, channels, height, width = shape
    out = torch.empty(tuple(shape), device=device, dtype=torch.float32)
    if out.numel() == 0:
        return out
    base = term_a["base"]
    p = _native.SonarNoiseMixParams()
    keep: list = []
    _fill_term(p.a, term_a, keep)
    _fill_term(p.b, term_b, keep)
    if term_b is not None and (term_b["base"].numel != base.numel or term_b["base"].grid_blocks != base.grid_blocks):
        raise ValueError("fused terms must draw tensors of the same size")
    p.out, p.n, p.begin, p.numel_total = out.data_ptr(), out.numel(), int(begin), base.numel
    p.C, p.H, p.W = channels, height, width
    p.blend_mode, p.blend_t = BLEND_IDS[blend_mode], float(blend_t)
    p.grid_blocks, p.seed = base.grid_blocks, base.seed
    lib, stream = _prepare(out, *keep)
    slot, p.sums, p.sums_clear = sums_slot(out.device)
    _launch("sonar_noise_mix_f32", lib.sonar_noise_mix_f32, ctypes.byref(p), stream)
    return _sums_written(out, slot)


# --------------------------------------------------------------------------------------------
# spectral shaping
# --------------------------------------------------------------------------------------------
_SPECTRAL_SCRATCH: dict[tuple, torch.Tensor] = {}


def spectral_filter(
    *,
    real: torch.Tensor | None = None,
    spectrum: torch.Tensor | None = None,
    philox: tuple | None = None,
    mask: torch.Tensor | None,
    hw: tuple[int, int],
    out_scale: float,
    out: torch.Tensor | None = None,
    sums_into: tuple | None = None,
    segment_planes: int | None = None,
    table: torch.Tensor | None = None,
):
    """[rfft2 ->] mask -> irfft2 per (H, W) plane. Exactly one of real / spectrum / philox is given.

    real: (..., H, W) float32; spectrum: (..., H, W//2+1) complex64; mask: (H, W//2+1) float32.
    philox = (draw, begin, std, lead_shape, device): the half spectrum is torch.randn(complex64) regenerated inside
    the kernel from `draw` (floats [begin, begin + 2 * planes * H * (W//2+1)) of it), never stored.
    out: optional preallocated result. sums_into = (slot pointer, clear pointer or 0): accumulate the moments into a
    slot the caller manages instead of taking one from the ring (the result is then not tagged).
    segment_planes: the planes form runs of this many planes (one noise sample each); returns (out, table) with
    table (n_segments, 2) float64 = {sum, sum of squares} of each run (`table`: a zeroed buffer of the caller's).
    """
    H, W = hw
    wh = W // 2 + 1
    if (real is not None) + (spectrum is not None) + (philox is not None) != 1:
        raise ValueError("give exactly one of real / spectrum / philox")
    if real is not None:
        _f32(real, "real")
        if tuple(real.shape[-2:]) != (H, W):
            raise ValueError("real input plane size mismatch")
        lead, device = real.shape[:-2], real.device
    elif spectrum is not None:
        if spectrum.dtype != torch.complex64 or tuple(spectrum.shape[-2:]) != (H, wh):
            raise ValueError(f"spectrum must be complex64 (..., {H}, {wh})")
        lead, device = spectrum.shape[:-2], spectrum.device
    else:
        lead, device = tuple(philox[3]), philox[4]
    if mask is not None:
        _f32(mask, "mask")
        if mask.numel() != H * wh:
            raise ValueError(f"mask must have {H}x{wh} elements")
    if out is None:
        out = torch.empty((*lead, H, W), device=device, dtype=torch.float32)
    elif out.dtype != torch.float32 or tuple(out.shape) != (*lead, H, W):
        raise ValueError(f"out must be float32 {(*lead, H, W)}")
    planes = out.numel() // (H * W)
    lib, stream = _prepare(real, spectrum, mask, out)
    scratch = None
    need = lib.sonar_spectral_scratch_bytes(H, W)
    if need > 0:
        key = (device, H, W)
        scratch = _SPECTRAL_SCRATCH.get(key)
        if scratch is None:
            scratch = torch.empty(need, device=device, dtype=torch.uint8)
            _SPECTRAL_SCRATCH[key] = scratch
    p = _native.SonarSpectralParams()
    p.out = out.data_ptr()
    p.in_real = 0 if real is None else real.data_ptr()
    p.in_spec = 0 if spectrum is None else spectrum.data_ptr()
    p.mask = 0 if mask is None else mask.data_ptr()
    p.scratch = 0 if scratch is None else scratch.data_ptr()
    p.planes, p.H, p.W, p.out_scale = planes, H, W, float(out_scale)
    if philox is not None:
        draw, begin, std = philox[:3]
        p.philox_seed, p.philox_offset, p.philox_grid_blocks = draw.seed, draw.offset, draw.grid_blocks
        p.philox_std, p.philox_begin, p.philox_numel_total = float(std), int(begin), draw.numel
    if segment_planes is not None:
        if planes % segment_planes:
            raise ValueError("segment_planes must divide the plane count")
        if table is None:
            table = torch.zeros((planes // segment_planes, 2), device=device, dtype=torch.float64)
        elif table.dtype != torch.float64 or tuple(table.shape) != (planes // segment_planes, 2) or not table.is_contiguous():
            raise ValueError("table must be a contiguous float64 (n_segments, 2) tensor")
        p.sums, p.sums_clear, p.sums_segment_planes = table.data_ptr(), 0, int(segment_planes)
        _launch("sonar_spectral_filter_f32", lib.sonar_spectral_filter_f32, ctypes.byref(p), stream)
        drop_sums(out)
        return out, table
    if sums_into is not None:
        p.sums, p.sums_clear = sums_into
        _launch("sonar_spectral_filter_f32", lib.sonar_spectral_filter_f32, ctypes.byref(p), stream)
        drop_sums(out)
        return out
    slot, p.sums, p.sums_clear = sums_slot(out.device)
    _launch("sonar_spectral_filter_f32", lib.sonar_spectral_filter_f32, ctypes.byref(p), stream)
    return _sums_written(out, slot)


FILL_BATCH_MAX = _native.FILL_BATCH_MAX
_PLAN_CACHE: dict = {}


def spectral_plan_batched(height: int, width: int, planes: int) -> bool:
    """True when sonar_spectral_filter_f32 runs (height, width) planes on the batched shared-memory kernel (even
    width, lengths factoring into radices 2..16) -- the form that can regenerate its input from the Philox stream."""
    key = (height, width, planes)
    hit = _PLAN_CACHE.get(key)
    if hit is None:
        info = _native.SonarSpectralPlanInfo()
        _native.check(_native.load().sonar_spectral_plan(height, width, planes, 0, ctypes.byref(info)), "sonar_spectral_plan")
        hit = _PLAN_CACHE[key] = bool(info.batched)
    return hit


def pack_mixer(mixer: torch.Tensor) -> torch.Tensor:
    """The (C, C) mixer pre-tiled for the bulk-async GEMM: [ceil(C/64)][ceil(C/16)][64][20] floats, zero padded (one
    contiguous 5 KB block per A tile). Setup work, once per sampler; returns a tensor on the mixer's device."""
    c = mixer.shape[0]
    tiles_m, kblocks = -(-c // 64), -(-c // 16)
    src = torch.zeros((tiles_m * 64, kblocks * 16), dtype=torch.float32)
    src[:c, :c] = mixer.detach().to("cpu", torch.float32)
    packed = torch.zeros((tiles_m, kblocks, 64, 20), dtype=torch.float32)
    packed[..., :16] = src.reshape(tiles_m, 64, kblocks, 16).permute(0, 2, 1, 3)
    assert packed.numel() == _native.load().sonar_channel_mix_packed_floats(c)
    return packed.to(mixer.device).contiguous()


def channel_mix(noise: torch.Tensor, mixer: torch.Tensor, mixer_host: torch.Tensor | None = None,
                mixer_packed: torch.Tensor | None = None) -> torch.Tensor:  # fmt: skip
    """out[b, c] = sum_k mixer[c, k] * noise[b, k] per pixel (ChannelMixer.apply); the moments of the result ride along.
    mixer_host: the matrix on the CPU (C <= 8: passed by value); mixer_packed: pack_mixer(mixer) (larger C: bulk-async GEMM)."""
    _f32(noise, "noise")
    _f32(mixer, "mixer")
    if noise.ndim != 4:
        raise ValueError("channel_mix expects a (B, C, H, W) tensor")
    batch, channels, height, width = noise.shape
    if tuple(mixer.shape) != (channels, channels):
        raise ValueError("Channel count mismatch")
    out = torch.empty_like(noise)
    lib, stream = _prepare(noise, mixer, out, mixer_packed)
    host_ptr = None
    if mixer_host is not None and channels <= _native.MIXER_SMALL_MAX:
        if mixer_host.dtype != torch.float32 or not mixer_host.is_contiguous() or mixer_host.is_cuda:
            raise ValueError("mixer_host must be a contiguous float32 CPU tensor")
        host_ptr = ctypes.c_void_p(mixer_host.data_ptr())
    if mixer_packed is not None and mixer_packed.numel() != lib.sonar_channel_mix_packed_floats(channels):
        raise ValueError("mixer_packed does not match the channel count (see pack_mixer)")
    slot, slot_ptr, clear_ptr = sums_slot(out.device)
    _launch(
        "sonar_channel_mix_f32", lib.sonar_channel_mix_f32,
        _ptr(noise), _ptr(out), _ptr(mixer), host_ptr, _ptr(mixer_packed), batch, channels, height * width, slot_ptr, clear_ptr, stream,
    )  # fmt: skip
    return _sums_written(out, slot)


# --------------------------------------------------------------------------------------------
# FreeU-Extreme epilogue (reference py/nodes/freeu_extreme.py:183-227)
# --------------------------------------------------------------------------------------------
def freeu_hidden_mean(h: torch.Tensor) -> tuple[torch.Tensor, torch.Tensor]:
    """Channel mean of a (B, C, H, W) activation plus the per-item (min, max) partials freeu_apply needs."""
    _f32(h, "h")
    if h.ndim != 4:
        raise ValueError("freeu_hidden_mean expects a (B, C, H, W) activation")
    batch, channels, height, width = h.shape
    hidden = torch.empty((batch, 1, height, width), device=h.device, dtype=torch.float32)
    lib, stream = _prepare(h, hidden)
    rng = torch.empty(max(int(lib.sonar_freeu_range_bytes(batch)), 8), device=h.device, dtype=torch.uint8)
    _launch("sonar_freeu_hidden_mean_f32", lib.sonar_freeu_hidden_mean_f32, _ptr(h), _ptr(hidden), _ptr(rng), batch, channels,
            height * width, stream)
    return hidden, rng


def freeu_apply(
    x: torch.Tensor,
    filtered: torch.Tensor | None,
    hidden: tuple[torch.Tensor, torch.Tensor] | None,
    *,
    slice_offset: int,
    slice_channels: int,
    scale: float,
    blend: float = 1.0,
    blend_mode: str | None = None,
) -> torch.Tensor:
    """x[:, off:off+n] = blend(x[:, off:off+n], src * scale_map, blend) in place; src = `filtered` (dense
    (B, n, H, W)) or the slice itself; scale_map = scale, or 1 + (scale - 1) * normalised hidden mean."""
    _f32(x, "x")
    if x.ndim != 4:
        raise ValueError("freeu_apply expects a (B, C, H, W) activation")
    batch, channels, height, width = x.shape
    if filtered is not None:
        _f32(filtered, "filtered")
        if tuple(filtered.shape) != (batch, slice_channels, height, width):
            raise ValueError(f"filtered slice has shape {tuple(filtered.shape)}, expected {(batch, slice_channels, height, width)}")
    p = _native.SonarFreeuParams()
    p.x = x.data_ptr()
    p.filtered = 0 if filtered is None else filtered.data_ptr()
    p.hidden = 0 if hidden is None else hidden[0].data_ptr()
    p.hidden_range = 0 if hidden is None else hidden[1].data_ptr()
    p.batch, p.channels, p.hw = batch, channels, height * width
    p.slice_offset, p.slice_channels = int(slice_offset), int(slice_channels)
    p.scale, p.scale_minus_one, p.blend = float(scale), float(scale - 1.0), float(blend)
    p.use_blend = int(blend != 1.0)
    p.blend_mode = BLEND_IDS[blend_mode] if p.use_blend else 0
    lib, stream = _prepare(x, filtered, *(hidden or ()))
    _launch("sonar_freeu_apply_f32", lib.sonar_freeu_apply_f32, ctypes.byref(p), stream)
    return x


# --------------------------------------------------------------------------------------------
# DWT levels
# --------------------------------------------------------------------------------------------
DWT_MODE_IDS = {"symmetric": 0, "zero": 1, "reflect": 2, "periodic": 3, "periodization": 4}
DWT_MODE_ALIASES = {"per": "periodization"}


def make_filters(dec_lo, dec_hi, rec_lo, rec_hi) -> _native.SonarWaveletFilters:
    f = _native.SonarWaveletFilters()
    n = len(dec_lo)
    if n > _native.DWT_MAX_TAPS or n % 2 or n < 2:
        raise ValueError(f"unsupported wavelet filter length {n}")
    f.length = n
    for i in range(n):
        f.dec_lo[i], f.dec_hi[i], f.rec_lo[i], f.rec_hi[i] = dec_lo[i], dec_hi[i], rec_lo[i], rec_hi[i]
    return f


def dwt_coeff_len(n: int, filter_len: int, mode: str = "symmetric") -> int:
    """Coefficients per axis: (n + L - 1) // 2 for the expansive modes, ceil(n / 2) for periodization."""
    return (n + 1) // 2 if mode == "periodization" else (n + filter_len - 1) // 2


def dwt2_analysis(
    a: torch.Tensor,
    b: torch.Tensor | None,
    filters: _native.SonarWaveletFilters,
    *,
    mode: str,
    coeff_dtype: torch.dtype,
    valid_hw: tuple[int, int] | None = None,
) -> tuple[torch.Tensor, torch.Tensor]:
    """One analysis level of (a - b). a: (planes, rows, W); valid_hw selects the top-left (H, W)
    region of a buffer whose rows exceed H. Returns ll (planes, h, w), hi (planes, 3, h, w)."""
    planes, rows, cols = a.shape
    H, W = valid_hw if valid_hw is not None else (rows, cols)
    if W != cols:
        raise ValueError("dwt2_analysis needs the row length of the buffer to equal W")
    L = filters.length
    h, w = dwt_coeff_len(H, L, mode), dwt_coeff_len(W, L, mode)
    ll = torch.empty((planes, h, w), device=a.device, dtype=coeff_dtype)
    hi = torch.empty((planes, 3, h, w), device=a.device, dtype=coeff_dtype)
    p = _native.SonarDwtAnalysisParams()
    p.in_a, p.in_b = a.data_ptr(), (0 if b is None else b.data_ptr())
    p.ll, p.hi = ll.data_ptr(), hi.data_ptr()
    p.planes, p.H, p.W, p.in_stride_h, p.h, p.w = planes, H, W, rows, h, w
    p.mode = DWT_MODE_IDS[mode]
    p.in_is_f32 = 0
    if a.dtype == torch.float32 and coeff_dtype == torch.float64:
        p.in_is_f32 = 1
        if rows != H:
            raise ValueError("fp32 level-1 input must be dense")
    elif a.dtype != coeff_dtype:
        raise TypeError(f"dwt2_analysis: input {a.dtype} vs coefficient {coeff_dtype}")
    p.use_f64 = int(coeff_dtype == torch.float64)
    p.filters = filters
    lib, stream = _prepare(a, b, ll, hi)
    _launch("sonar_dwt2_analysis", lib.sonar_dwt2_analysis, ctypes.byref(p), stream)
    return ll, hi


def dwt2_synthesis(
    sets: Sequence[tuple[torch.Tensor, torch.Tensor, Sequence[float]]],
    filters: _native.SonarWaveletFilters,
    *,
    final: dict | None = None,
) -> torch.Tensor:
    """One synthesis level. sets: [(ll, hi, (s_ll, s_h0, s_h1, s_h2)), ...] (1 or 2 entries);
    ll may be one row/col larger than hi ("unpad"). final = dict(crop=(h, w), addend=, addend_scale=,
    x=, x_scale=, recon_sign=) turns it into the fp32 epilogue level."""
    ll0, hi0, _ = sets[0]
    planes, _, h, w = hi0.shape
    L = filters.length
    out_h, out_w = 2 * h - L + 2, 2 * w - L + 2
    p = _native.SonarDwtSynthesisParams()
    live = []
    for i, (ll, hi, sc) in enumerate(sets):
        p.ll[i], p.hi[i] = ll.data_ptr(), hi.data_ptr()
        p.ll_rows[i], p.ll_cols[i] = ll.shape[-2], ll.shape[-1]
        for j in range(4):
            p.scales[i][j] = float(sc[j])
        live += [ll, hi]
    p.n_sets, p.planes, p.h, p.w = len(sets), planes, h, w
    p.use_f64 = int(hi0.dtype == torch.float64)
    p.filters = filters
    if final is None:
        out = torch.empty((planes, out_h, out_w), device=hi0.device, dtype=hi0.dtype)
        p.out, p.out_f32 = out.data_ptr(), 0
    else:
        ch, cw = final["crop"]
        out = torch.empty((planes, ch, cw), device=hi0.device, dtype=torch.float32)
        p.out, p.out_f32 = 0, out.data_ptr()
        p.crop_h, p.crop_w = ch, cw
        addend, x = final.get("addend"), final.get("x")
        p.addend = 0 if addend is None else addend.data_ptr()
        p.addend_scale = float(final.get("addend_scale", 1.0))
        p.x = 0 if x is None else x.data_ptr()
        p.x_scale = float(final.get("x_scale", 1.0))
        p.recon_sign = float(final.get("recon_sign", 1.0))
        live += [addend, x]
    lib, stream = _prepare(out, *live)
    _launch("sonar_dwt2_synthesis", lib.sonar_dwt2_synthesis, ctypes.byref(p), stream)
    return out


def dwt2_synthesis_per(ll: torch.Tensor, hi: torch.Tensor, scales: Sequence[float], filters: _native.SonarWaveletFilters) -> torch.Tensor:
    """One reconstruction level of the non-expansive "periodization" transform: ll (planes, >=h, >=w), hi
    (planes, 3, h, w) -> (planes, 2h, 2w) in the coefficient dtype, bands multiplied by `scales` on load."""
    planes, _, h, w = hi.shape
    if ll.dtype != hi.dtype or ll.dtype not in (torch.float32, torch.float64):
        raise TypeError(f"dwt2_synthesis_per: coefficient dtypes {ll.dtype} / {hi.dtype}")
    if ll.shape[0] != planes or ll.shape[1] < h or ll.shape[2] < w:
        raise ValueError(f"dwt2_synthesis_per: ll {tuple(ll.shape)} does not cover hi {tuple(hi.shape)}")
    out = torch.empty((planes, 2 * h, 2 * w), device=hi.device, dtype=hi.dtype)
    p = _native.SonarDwtSynthesisParams()
    p.ll[0], p.hi[0] = ll.data_ptr(), hi.data_ptr()
    p.ll_rows[0], p.ll_cols[0] = ll.shape[1], ll.shape[2]
    for i in range(4):
        p.scales[0][i] = float(scales[i])
    p.n_sets, p.planes, p.h, p.w = 1, planes, h, w
    p.out, p.out_f32 = out.data_ptr(), 0
    p.use_f64 = int(hi.dtype == torch.float64)
    p.filters = filters
    lib, stream = _prepare(ll, hi, out)
    _launch("sonar_dwt2_synthesis_per", lib.sonar_dwt2_synthesis_per, ctypes.byref(p), stream)
    return out


def wcfg_fused_fits(height: int, width: int, filter_len: int, levels: int, *, use_f64: bool) -> bool:
    """True when the whole coefficient pyramid of one (height, width) plane fits one SM's shared memory."""
    if levels > _native.WCFG_MAX_LEVELS:
        return False
    return _native.load().sonar_wcfg_fused_smem_bytes(height, width, filter_len, levels, int(use_f64)) > 0


def wcfg_fused(
    a: torch.Tensor,
    b: torch.Tensor | None,
    filters: _native.SonarWaveletFilters,
    *,
    levels: int,
    mode: str,
    use_f64: bool,
    scale_ll: float,
    scale_hi: Sequence[Sequence[float]],
    addend: torch.Tensor | None = None,
    addend_scale: float = 1.0,
    x: torch.Tensor | None = None,
    x_scale: float = 1.0,
    recon_sign: float = 1.0,
) -> torch.Tensor:
    """x_scale*x + float(recon_sign*(IDWT(S (.) DWT(a - b)) + addend_scale*addend)), cropped to the
    input size, in ONE launch. a, b, addend, x: (planes, H, W) float32. scale_hi: [level][3], fine ->
    coarse; scale_ll scales the coarsest approximation band."""
    _f32(a, "a")
    planes, height, width = a.shape
    out = torch.empty_like(a)
    p = _native.SonarWcfgFusedParams()
    p.in_a, p.in_b, p.out = a.data_ptr(), (0 if b is None else b.data_ptr()), out.data_ptr()
    p.addend, p.addend_scale = (0 if addend is None else addend.data_ptr()), float(addend_scale)
    p.x, p.x_scale, p.recon_sign = (0 if x is None else x.data_ptr()), float(x_scale), float(recon_sign)
    p.planes, p.H, p.W, p.levels = planes, height, width, levels
    p.mode, p.use_f64 = DWT_MODE_IDS[mode], int(use_f64)
    p.scale_ll = float(scale_ll)
    for j, row in enumerate(scale_hi):
        for o in range(3):
            p.scale_hi[j][o] = float(row[o])
    p.filters = filters
    lib, stream = _prepare(a, b, out, addend, x)
    _launch("sonar_wcfg_fused", lib.sonar_wcfg_fused, ctypes.byref(p), stream)
    return out


# --------------------------------------------------------------------------------------------
# reference-latent guidance
# --------------------------------------------------------------------------------------------
GUIDANCE_LINEAR, GUIDANCE_EULER = 0, 1


def item_moments(x: torch.Tensor) -> torch.Tensor:
    """(items, 2) float64: sum and sum of squares of every leading-dim item of x."""
    _f32(x, "x")
    items = x.shape[0]
    sums = torch.empty((items, 2), device=x.device, dtype=torch.float64)
    lib, stream = _prepare(x, sums)
    _launch("sonar_item_moments_f32", lib.sonar_item_moments_f32, _ptr(x), items, x.numel() // max(1, items), _ptr(sums), stream)
    return sums


def guidance(
    x: torch.Tensor,
    ref: torch.Tensor,
    item_sums: torch.Tensor | None,
    *,
    kind: int,
    blend_mode: str | int = "lerp",
    factor: float = 0.0,
    sigma: float = 1.0,
    dt: float = 0.0,
) -> torch.Tensor:
    """guidance_linear / guidance_euler of the reference (py/sonar.py:380-411) in one pass: the target
    is ref * std_i + mean_i with the per-item statistics taken from `item_sums` (item_moments of x for
    LINEAR, of the denoised prediction for EULER)."""
    _f32(x, "x")
    _f32(ref, "ref")
    items = x.shape[0]
    per_item = x.numel() // max(1, items)
    if ref.numel() not in {per_item, x.numel()}:
        raise ValueError(f"guidance: reference latent {tuple(ref.shape)} does not broadcast over {tuple(x.shape)}")
    out = torch.empty_like(x)
    p = _native.SonarGuidanceParams()
    p.x, p.ref, p.out = x.data_ptr(), ref.data_ptr(), out.data_ptr()
    p.item_sums = 0 if item_sums is None else item_sums.data_ptr()
    p.items, p.per_item, p.ref_items = items, per_item, (1 if ref.numel() == per_item and items != 1 else items)
    p.kind, p.blend_mode = kind, (blend_mode if isinstance(blend_mode, int) else BLEND_IDS[blend_mode])
    p.factor, p.sigma, p.dt = float(factor), float(sigma), float(dt)
    lib, stream = _prepare(x, ref, item_sums, out)
    _launch("sonar_guidance_f32", lib.sonar_guidance_f32, ctypes.byref(p), stream)
    return out
