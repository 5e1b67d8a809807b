"""`torch.library` custom-op layer: the C-ABI entry points as `torch.ops.sonar_b200.*`.

north_star / SURVEY.md 8b: the host code calls the hand-written kernels "through a thin C-ABI / torch.library
custom-op layer". The C ABI (include/sonar_b200.h, bound with ctypes in `_native.py`, wrapped per tensor in `ops.py`)
is the seam; this module registers the same entry points with the PyTorch dispatcher so that graph tooling, other
extensions and C++ callers reach them by name. Every op is registered for the CUDA dispatch key ONLY: calling one
with CPU tensors fails in the dispatcher ("no kernel for the CPU backend") -- there is no fallback to hide behind.

The samplers' per-step lane keeps calling `ops.launch_step` directly (one ctypes call, ~5 us of host time; a
dispatcher round trip costs about as much again, and the C2 step is host bound).

    torch.ops.sonar_b200.step(x, denoised, hist, noise, ...) -> (x', hist')       sonar_step_f32
    torch.ops.sonar_b200.philox_normal_(out, seed, offset, mean, std)              sonar_philox_normal_f32
    torch.ops.sonar_b200.philox_uniform_(out, seed, offset, low, high)             sonar_philox_uniform_f32
    torch.ops.sonar_b200.randn_like(x) / rand_like(x)                              same, on torch's CUDA generator
    torch.ops.sonar_b200.moments(x) -> double[2]                                   sonar_moments_f32
    torch.ops.sonar_b200.scale_noise(x, factor, normalized) -> Tensor              sonar_moments_f32 + sonar_scale_noise_f32
    torch.ops.sonar_b200.spectral_filter(real, spectrum, mask, H, W, out_scale)    sonar_spectral_filter_f32
    torch.ops.sonar_b200.channel_mix(noise, mixer) -> Tensor                       sonar_channel_mix_f32
    torch.ops.sonar_b200.pyramid_accum(base, levels, weights, H, W, mode, scale)   sonar_pyramid_accum_f32
    torch.ops.sonar_b200.perlin_accum(base, angles, shape, div_fac, blend_mode)    sonar_perlin_accum_f32
    torch.ops.sonar_b200.blend(a, b, t, t_scalar, mode) -> Tensor                  sonar_blend_f32
    torch.ops.sonar_b200.guidance(x, ref, stats_of, kind, blend, factor, sigma, dt) sonar_item_moments_f32 + sonar_guidance_f32
    torch.ops.sonar_b200.wcfg_fused(a, b, filters..., levels, ...) -> Tensor       sonar_wcfg_fused
"""

from __future__ import annotations

from typing import Sequence

import torch
from torch import Tensor

from . import ops
from ._native import SonarStepParams

NAMESPACE = "sonar_b200"
_LIB = torch.library.Library(NAMESPACE, "DEF")
OP_NAMES: list[str] = []


def _register(schema: str, fn) -> None:
    name = schema.split("(", 1)[0]
    _LIB.define(schema)
    _LIB.impl(name, fn, "CUDA")
    OP_NAMES.append(name)


def _contig(t: Tensor | None) -> Tensor | None:
    return None if t is None else t.contiguous()


# ---------------------------------------------------------------------------------------------
def _step(x: Tensor, denoised: Tensor, hist: Tensor | None, noise: Tensor | None, kind: int, mode: int, momentum: float,
          momentum_hist: float, direction: float, sigma: float, c0: float, c1: float, noise_scale: float,
          momentum_active: bool = True, history_active: bool = True, momentum_blend: str = "lerp",
          history_blend: str = "lerp"):  # fmt: skip  (the dispatcher passes only what the caller gave: defaults repeat the schema's)
    """One fused Sonar half step (py/sonar.py:227-320, :460-480, :541-573, :649-735). hist = history_d or None
    (first step, ZERO init); noise = already-normalised ancestral noise or None. Returns (x', history')."""
    x, denoised, hist, noise = _contig(x), _contig(denoised), _contig(hist), _contig(noise)
    p = SonarStepParams()
    x_out = torch.empty_like(x)
    hist_out = torch.empty_like(x)
    p.x, p.denoised, p.x_out, p.hist_out = x.data_ptr(), denoised.data_ptr(), x_out.data_ptr(), hist_out.data_ptr()
    p.hist_in = 0 if hist is None else hist.data_ptr()
    p.hist_state = ops.HIST_NONE if hist is None else ops.HIST_PRESENT
    p.noise = 0 if noise is None else noise.data_ptr()
    p.noise_kind = ops.NOISE_NONE if noise is None else ops.NOISE_TENSOR
    p.n, p.kind, p.mode = x.numel(), kind, mode
    p.momentum_blend, p.history_blend = ops.BLEND_IDS[momentum_blend], ops.BLEND_IDS[history_blend]
    p.momentum_active, p.history_active = int(momentum_active), int(history_active)
    p.momentum, p.sigma, p.c0, p.c1, p.noise_scale = momentum, sigma, c0, c1, noise_scale
    # history_ratios (py/sonar.py:208-219)
    p.hd_ratio = momentum_hist
    p.hd_scale = 1.0 + abs(direction) * (1 - momentum_hist) if direction < 0 else 2.0 - direction
    p.md_scale = direction
    p.hist_in_div, p.noise_threshold_std_devs = 1.0, 2.5
    for t, what in ((x, "x"), (denoised, "denoised"), (hist, "hist"), (noise, "noise")):
        if t is not None and (t.dtype != torch.float32 or t.shape != x.shape):
            raise TypeError(f"sonar_b200::step: {what} must be float32 of x's shape")
    ops.sonar_step(p, x, denoised, hist, noise, x_out, hist_out)
    return x_out, hist_out


_register(
    "step(Tensor x, Tensor denoised, Tensor? hist, Tensor? noise, int kind, int mode, float momentum, float momentum_hist, "
    "float direction, float sigma, float c0, float c1, float noise_scale, bool momentum_active=True, "
    'bool history_active=True, str momentum_blend="lerp", str history_blend="lerp") -> (Tensor, Tensor)',
    _step,
)


# ---------------------------------------------------------------------------------------------
def _philox_fill_(out: Tensor, seed: int, offset: int, p0: float, p1: float, kind: str) -> Tensor:
    if not out.is_contiguous() or out.dtype not in (torch.float32, torch.complex64):
        raise TypeError("sonar_b200 Philox fills take contiguous float32 / complex64 tensors")
    floats = out.numel() * (2 if out.is_complex() else 1)
    grid, inc = ops.philox_policy_cached(out.device.index, floats)
    draw = ops.PhiloxDraw(seed, offset, grid, floats, inc)
    return ops.philox_fill(draw, out, kind=kind, p0=p0, p1=p1)


_register(
    "philox_normal_(Tensor(a!) out, int seed, int offset, float mean=0.0, float std=1.0) -> Tensor(a!)",
    lambda out, seed, offset, mean=0.0, std=1.0: _philox_fill_(out, seed, offset, mean, std, "normal"),
)
_register(
    "philox_uniform_(Tensor(a!) out, int seed, int offset, float low=0.0, float high=1.0) -> Tensor(a!)",
    lambda out, seed, offset, low=0.0, high=1.0: _philox_fill_(out, seed, offset, low, high, "uniform"),
)
_register("randn_like(Tensor x) -> Tensor", lambda x: ops.randn(x.shape, device=x.device, dtype=x.dtype))
_register("rand_like(Tensor x) -> Tensor", lambda x: ops.rand(x.shape, device=x.device))
_register("moments(Tensor x) -> Tensor", lambda x: ops.moments(x.contiguous()))


def _scale_noise(x: Tensor, factor: float = 1.0, normalized: bool = True) -> Tensor:
    from .hostutil import scale_noise

    return scale_noise(x.clone(memory_format=torch.contiguous_format), factor, normalized=normalized)


_register("scale_noise(Tensor x, float factor=1.0, bool normalized=True) -> Tensor", _scale_noise)


def _spectral_filter(real: Tensor | None, spectrum: Tensor | None, mask: Tensor | None, H: int, W: int, out_scale: float):  # noqa: N803
    return ops.spectral_filter(real=_contig(real), spectrum=_contig(spectrum), mask=_contig(mask), hw=(H, W), out_scale=out_scale)


_register("spectral_filter(Tensor? real, Tensor? spectrum, Tensor? mask, int H, int W, float out_scale) -> Tensor", _spectral_filter)
_register(
    "channel_mix(Tensor noise, Tensor mixer) -> Tensor",
    lambda noise, mixer: ops.channel_mix(
        noise.contiguous(), mixer.contiguous(), None, ops.pack_mixer(mixer) if mixer.shape[0] > 8 else None,
    ),
)


def _pyramid_accum(base: Tensor | None, levels: Sequence[Tensor], weights: Sequence[float], H: int, W: int,  # noqa: N803
                   mode: str = "bilinear", base_scale: float = 1.0) -> Tensor:  # fmt: skip
    return ops.pyramid_accumulate(_contig(base), [lv.contiguous() for lv in levels], list(weights), out_hw=(H, W), mode=mode,
                                  base_scale=base_scale)  # fmt: skip


_register(
    'pyramid_accum(Tensor? base, Tensor[] levels, float[] weights, int H, int W, str mode="bilinear", float base_scale=1.0) -> Tensor',
    _pyramid_accum,
)
_register(
    'perlin_accum(Tensor? base, Tensor[] angles, int[] shape, float div_fac=2.0, str blend_mode="lerp") -> Tensor',
    lambda base, angles, shape, div_fac=2.0, blend_mode="lerp": ops.perlin_accumulate(
        _contig(base), [a.contiguous() for a in angles], shape=tuple(shape), div_fac=div_fac, blend_mode=blend_mode,
    ),
)
_register(
    'blend(Tensor a, Tensor b, Tensor? t, float t_scalar=0.5, str mode="lerp") -> Tensor',
    lambda a, b, t, t_scalar=0.5, mode="lerp": ops.blend(a.contiguous(), b.contiguous(), t_scalar if t is None else t, mode=mode),
)


def _guidance(x: Tensor, ref: Tensor, stats_of: Tensor | None, kind: int, blend_mode: str = "lerp", factor: float = 0.0,
              sigma: float = 1.0, dt: float = 0.0):  # fmt: skip
    """guidance_linear / guidance_euler (py/sonar.py:380-411); stats_of = the tensor whose per-item mean / std the
    reference latent takes (None: no shift)."""
    sums = None if stats_of is None else ops.item_moments(stats_of.contiguous())
    return ops.guidance(x.contiguous(), ref.contiguous(), sums, kind=kind, blend_mode=blend_mode, factor=factor, sigma=sigma, dt=dt)


_register(
    'guidance(Tensor x, Tensor ref, Tensor? stats_of, int kind, str blend_mode="lerp", float factor=0.0, float sigma=1.0, '
    "float dt=0.0) -> Tensor",
    _guidance,
)


def _wcfg_fused(a: Tensor, b: Tensor | None, dec_lo: Sequence[float], dec_hi: Sequence[float], rec_lo: Sequence[float],
                rec_hi: Sequence[float], levels: int, mode: str, use_f64: bool, scale_ll: float, scale_hi: Sequence[float],
                addend: Tensor | None, addend_scale: float, x: Tensor | None, x_scale: float, recon_sign: float) -> Tensor:  # fmt: skip
    filters = ops.make_filters(list(dec_lo), list(dec_hi), list(rec_lo), list(rec_hi))
    rows = [list(scale_hi[3 * j : 3 * j + 3]) for j in range(levels)]
    return ops.wcfg_fused(
        a.contiguous(), _contig(b), filters, levels=levels, mode=mode, use_f64=use_f64, scale_ll=scale_ll, scale_hi=rows,
        addend=_contig(addend), addend_scale=addend_scale, x=_contig(x), x_scale=x_scale, recon_sign=recon_sign,
    )  # fmt: skip


_register(
    "wcfg_fused(Tensor a, Tensor? b, float[] dec_lo, float[] dec_hi, float[] rec_lo, float[] rec_hi, int levels, str mode, "
    "bool use_f64, float scale_ll, float[] scale_hi, Tensor? addend, float addend_scale, Tensor? x, float x_scale, "
    "float recon_sign) -> Tensor",
    _wcfg_fused,
)

__all__ = ["NAMESPACE", "OP_NAMES"]
