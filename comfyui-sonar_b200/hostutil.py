"""Host-side helpers shared by the noise graph, the samplers and wavelet CFG.

Mirrors the hot subset of the reference's py/utils.py (BLENDING_MODES :17-21, scale_samples :58-67,
scale_noise :85-106, tensor_to :112-121, normalize_to_scale :452-470, crop_samples :526-568), with
the tensor arithmetic routed to the CUDA kernels in `ops`.
"""

from __future__ import annotations

import math
from typing import Callable, Sequence

import torch

from . import ops, parallel

UPSCALE_METHODS = (
    "bilinear",
    "nearest-exact",
    "nearest",
    "area",
    "bicubic",
    "bislerp",
    "adaptive_avg_pool2d",
)


def fallback(val, default=None):
    return default if val is None else val


def _blend_kernel(mode: str) -> Callable:
    def blend(a: torch.Tensor, b: torch.Tensor, t) -> torch.Tensor:
        return ops.blend(a, b, t, mode=mode)

    blend.__name__ = f"blend_{mode}"
    blend.sonar_blend_mode = mode
    return blend


# Same keys and (a, b, t) calling convention as the reference table; values are kernel launchers.
BLENDING_MODES: dict[str, Callable] = {name: _blend_kernel(name) for name in ("lerp", "inject", "subtract_b")}


def blend_mode_id(fn_or_name) -> int:
    name = fn_or_name if isinstance(fn_or_name, str) else getattr(fn_or_name, "sonar_blend_mode", None)
    if name not in ops.BLEND_IDS:
        raise ValueError(f"Unknown blend mode {fn_or_name!r}; valid: {', '.join(ops.BLEND_IDS)}")
    return ops.BLEND_IDS[name]


def scale_noise(
    noise: torch.Tensor,
    factor: float = 1.0,
    *,
    normalized: bool = True,
    threshold_std_devs: float = 2.5,
    normalize_dims: tuple | None = None,
) -> torch.Tensor:
    """In-place scale_noise (reference py/utils.py:85-106) without the host round trip.

    The moments pass leaves (sum, sum^2) on the device; when the batch is sharded over ranks
    (`parallel.active()`), the two doubles are all-reduced so every rank takes the same global
    decision the un-sharded reference would.
    """
    numel = noise.numel()
    if not normalized or numel == 0:
        return ops.scale(noise, factor) if factor != 1 else noise
    if normalize_dims is not None:
        # Per-dim unconditional variant (:96-99). Not on any configured hot path; kept on the device.
        std = noise.std(dim=normalize_dims, keepdim=True)
        noise = noise / std
        return noise.sub_(noise.mean(dim=normalize_dims, keepdim=True)).mul_(factor)
    if noise.dtype != torch.float32:
        work = noise.float()
        return scale_noise(work, factor, normalized=True, threshold_std_devs=threshold_std_devs).to(noise.dtype)
    if not noise.is_contiguous():
        noise = noise.contiguous()
    sums = ops.attached_sums(noise)  # reduced by the kernel that produced `noise`, if it is still fresh
    if sums is None:
        sums = ops.moments(noise)
    ctx = parallel.active()
    if ctx is not None and ctx.world_size > 1 and ctx.peers is not None:
        # partial sums travel as NVLink stores into every rank's mailbox; the apply kernel waits for them
        epoch = ctx.peers.publish(sums)
        return ops.scale_noise_apply_peers(
            noise, ctx.peers, epoch, parallel.global_numel(numel), factor, threshold_std_devs=threshold_std_devs,
        )
    count = parallel.global_count(numel, sums)
    return ops.scale_noise_apply(noise, sums, count, factor, threshold_std_devs=threshold_std_devs)


def tensor_to(tensor: torch.Tensor, dest) -> torch.Tensor:
    device = dest.device if isinstance(dest, torch.Tensor) else dest
    return tensor.to(dest, non_blocking=torch.device(device).type == "cuda")


def scale_samples(samples: torch.Tensor, width: int, height: int, *, mode: str = "bicubic") -> torch.Tensor:
    """Resize helper of the reference (py/utils.py:58-67). comfy.utils.common_upscale is exactly
    F.interpolate(size=(h, w), mode=mode) for every mode but bislerp; bilinear / nearest-exact /
    area go through our resampling kernel."""
    if mode in {"bilinear", "nearest-exact", "area", "adaptive_avg_pool2d"} and samples.dtype == torch.float32:
        return ops.resample(samples, height, width, mode="area" if mode == "adaptive_avg_pool2d" else mode)
    if mode == "bislerp":
        raise NotImplementedError("bislerp needs ComfyUI's implementation (out of scope, SURVEY.md section 8c)")
    return torch.nn.functional.interpolate(samples, size=(height, width), mode=mode)


def normalize_to_scale(
    latent: torch.Tensor,
    target_min: float,
    target_max: float,
    *,
    dim=(-3, -2, -1),
    eps: float = 1e-07,
) -> torch.Tensor:
    """Min/max rescale per batch item (reference py/utils.py:452-470)."""
    dims = tuple(d % latent.ndim for d in dim)
    if latent.dtype == torch.float32 and dims == tuple(range(1, latent.ndim)) and latent.ndim >= 2:
        return ops.minmax_rescale(latent, target_min, target_max, eps=eps)
    min_val, max_val = latent.amin(dim=dim, keepdim=True), latent.amax(dim=dim, keepdim=True)
    normalized = latent - min_val
    normalized /= (max_val - min_val).add_(eps)
    return normalized.mul_(target_max - target_min).add_(target_min).clamp_(target_min, target_max)


def _shift_slice(s: slice, size: int, offset: int) -> slice:
    if offset == 0:
        return s
    start = 0 if s.start is None else s.start
    stop = size if s.stop is None else s.stop
    if offset < 0:
        adj = min(start, -offset)
        return slice(start - adj, stop - adj)
    adj = min(size - stop, offset)
    return slice(start + adj, stop + adj)


def crop_samples(
    tensor: torch.Tensor,
    width: int,
    height: int,
    *,
    mode: str = "center",
    offset_width: int = 0,
    offset_height: int = 0,
) -> torch.Tensor:
    """Index-only crop (bit-exact by construction; reference py/utils.py:526-568)."""
    if tensor.ndim < 3:
        raise ValueError("Can only handle >= 3 dimensional tensors")
    th, tw = tensor.shape[-2:]
    if (tw, th) == (width, height):
        return tensor
    if tw < width or th < height:
        raise ValueError("Can't crop sample smaller than requested width or height")
    if mode == "center":
        hmode = wmode = "center"
    else:
        parts = mode.split("_")
        if len(parts) != 2:
            raise ValueError("Bad composite mode")
        hmode, wmode = parts
    h_choices = {"top": 0, "center": (th - height) // 2, "bottom": th - height}
    w_choices = {"left": 0, "center": (tw - width) // 2, "right": tw - width}
    if hmode not in h_choices:
        raise ValueError("Bad height mode in composite mode")
    if wmode not in w_choices:
        raise ValueError("Bad width mode in composite mode")
    hslice = _shift_slice(slice(h_choices[hmode], h_choices[hmode] + height), th, offset_height)
    wslice = _shift_slice(slice(w_choices[wmode], w_choices[wmode] + width), tw, offset_width)
    return tensor[..., hslice, wslice]


def clamp_float(val: float, minval: float = 0.0, maxval: float = 1.0) -> float:
    return max(minval, min(val, maxval))


def filter_dict(d: dict, keep: set | Sequence, *, recursive: bool = False) -> dict:
    return {
        k: (filter_dict(v, keep) if recursive and isinstance(v, dict) else v)
        for k, v in d.items()
        if k in keep
    }


def maybe_apply(val, cond, fun):
    return fun(val) if cond else val


def maybe_apply_kwargs(d: dict | None, cond, fun, *, default=None):
    return default if d is None or not cond else fun(**d)


def blend_scalar(a: float, b: float, t: float, *, blend_function: Callable | None = None) -> float:
    """Scalar blend for schedule interpolation (reference py/utils.py:33-55); pure host math."""
    if blend_function is None:
        return a * (1.0 - t) + b * t
    mode = getattr(blend_function, "sonar_blend_mode", None)
    if mode == "inject":
        return b * t + a
    if mode == "subtract_b":
        return a - b * t
    if mode == "lerp":
        return a + t * (b - a) if abs(t) < 0.5 else b - (b - a) * (1.0 - t)
    return float(blend_function(*(torch.tensor((v,), dtype=torch.float64) for v in (a, b, t))).item())


def tensor_item(val, *, collapse_function=torch.max) -> float:
    if isinstance(val, torch.Tensor):
        return float(collapse_function(val).detach().cpu().item())
    return float(val)


def step_from_sigmas(sigma, sigmas: torch.Tensor, *, decimals: int | None = 4, output_decimals: int = 2):
    """Fractional step index of `sigma` within a descending schedule (reference py/utils.py:682-723).
    Host-side schedule logic for wavelet CFG percentages."""
    sigma = tensor_item(sigma)
    sigmas = sigmas.detach().cpu()
    if sigmas.ndim == 2:
        sigmas = sigmas.max(dim=0).values
    elif sigmas.ndim != 1:
        raise ValueError(f"Unexpected number of dimensions in sigmas, should be 1 or 2 but got shape {sigmas.shape}")
    sigmas = sigmas[:-1]
    if not len(sigmas) or torch.any(sigmas <= 0):
        return None
    if decimals is not None:
        sigmas = sigmas.round(decimals=decimals)
        sigma = round(sigma, decimals)
    lo, hi = sigmas.aminmax()
    if not lo <= sigma <= hi:
        return None
    last = len(sigmas) - 1
    idx = int(tensor_item((sigmas - sigma).abs().argmin()))
    at_idx = tensor_item(sigmas[idx])
    if decimals is not None:
        at_idx = round(at_idx, decimals)
    if sigma == at_idx:
        return float(idx)
    idx_low, idx_high = (idx, idx - 1) if sigma > at_idx else (idx + 1, idx)
    if min(idx_low, idx_high) < 0 or max(idx_low, idx_high) > last:
        return None
    s_low, s_high = tensor_item(sigmas[idx_low]), tensor_item(sigmas[idx_high])
    span = s_high - s_low
    if span == 0:
        return float(idx)
    return round(idx_high + (1.0 - (sigma - s_low) / span), output_decimals)


def noise_threshold(numel: int, threshold_std_devs: float = 2.5) -> float:
    return threshold_std_devs / math.sqrt(numel)
