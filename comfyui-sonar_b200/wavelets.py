"""2-D wavelet transforms on the CUDA DWT kernels, with the pytorch_wavelets calling convention.

Mirror of the reference's py/wavelet_functions.py: `Wavelet` (forward -> (yl, yh), inverse),
`expand_yh_scales`, `wavelet_scaling`, `wavelet_blend`. The reference delegates the transform to
pytorch_wavelets.DWTForward/DWTInverse (`wavelet_functions.py:56-79`), an upstream dependency that
is absent from the reference tree and from this image with no pinned version; its algorithm is
restated in csrc/wavelet.cu (pad -> correlate -> decimate / transposed convolution, orientation
order = [high-H/low-W, low-H/high-W, high/high], levels fine -> coarse). Parity for this transform
is therefore anchored on the restated oracle and on size-independent identities (perfect
reconstruction, linearity, orthonormality), not on the library itself: "parity unpinned".

Only the plain 2-D DWT is in scope (no DTCWT / 1-D variants).
"""

from __future__ import annotations

from typing import Callable, Sequence

import torch

from . import ops
from .hostutil import fallback
from .wavelet_tables import DB_DEC_LO

HAVE_WAVELETS = True  # the transform is built in; no third-party package is needed


def filter_bank(wave: str) -> tuple[tuple, tuple, tuple, tuple]:
    """(dec_lo, dec_hi, rec_lo, rec_hi) in pywt's ordering for haar / db1..db12."""
    name = wave.strip().lower()
    if name == "haar":
        order = 1
    elif name.startswith("db") and name[2:].isdigit():
        order = int(name[2:])
    else:
        raise NotImplementedError(
            f"sonar_b200: wavelet {wave!r} has no built-in filter bank (supported: {', '.join(Wavelet.wavelist())})",
        )
    if order not in DB_DEC_LO:
        raise NotImplementedError(f"sonar_b200: wavelet {wave!r}: Daubechies order must be 1..{max(DB_DEC_LO)}")
    dec_lo = tuple(DB_DEC_LO[order])
    n = len(dec_lo)
    rec_lo = tuple(reversed(dec_lo))
    # quadrature mirror: dec_hi[k] = (-1)^(k+1) * dec_lo[L-1-k]
    dec_hi = tuple((-1.0 if k % 2 == 0 else 1.0) * dec_lo[n - 1 - k] for k in range(n))
    rec_hi = tuple(reversed(dec_hi))
    return dec_lo, dec_hi, rec_lo, rec_hi


class Wavelet:
    DEFAULT_MODE = "symmetric"
    DEFAULT_LEVEL = 3
    DEFAULT_WAVE = "db4"
    DEFAULT_USE_1D_DWT = False
    DEFAULT_USE_DTCWT = False
    DEFAULT_QSHIFT = "qshift_a"
    DEFAULT_BIORT = "near_sym_a"

    def __init__(
        self,
        *,
        wave: str = DEFAULT_WAVE,
        level: int = DEFAULT_LEVEL,
        mode: str = DEFAULT_MODE,
        use_1d_dwt: bool = DEFAULT_USE_1D_DWT,
        use_dtcwt: bool = DEFAULT_USE_DTCWT,
        biort: str = DEFAULT_BIORT,
        qshift: str = DEFAULT_QSHIFT,
        inv_wave: str | None = None,
        inv_mode: str | None = None,
        inv_biort: str | None = None,
        inv_qshift=None,
        device=None,
        dtype: torch.dtype = torch.float32,
    ):
        if use_dtcwt or use_1d_dwt:
            raise NotImplementedError("sonar_b200: only the 2-D DWT is in scope (no DTCWT / 1-D DWT kernels)")
        mode = ops.DWT_MODE_ALIASES.get(mode, mode)
        inv_mode = None if inv_mode is None else ops.DWT_MODE_ALIASES.get(inv_mode, inv_mode)
        if mode not in ops.DWT_MODE_IDS:
            raise NotImplementedError(
                f"sonar_b200: padding mode {mode!r} has no kernel (supported: {', '.join(ops.DWT_MODE_IDS)})",
            )
        self.wave, self.level, self.mode = wave, int(level), mode
        self.inv_wave = fallback(inv_wave, wave)
        self.inv_mode = fallback(inv_mode, mode)
        self.filters = ops.make_filters(*filter_bank(wave))
        self.inv_filters = self.filters if self.inv_wave == wave else ops.make_filters(*filter_bank(self.inv_wave))
        self.device = device
        self.dtype = dtype
        _ = (biort, qshift, inv_biort, inv_qshift)

    # ---- pytorch_wavelets-style API ---------------------------------------------------------
    def forward(self, t: torch.Tensor, *, forward_function: Callable | None = None):
        """(yl, yh): yl (B, C, h_J, w_J); yh[j] (B, C, 3, h_j, w_j), fine -> coarse."""
        if forward_function is not None:
            return forward_function(t)
        lead = t.shape[:-2]
        cur = t.reshape(-1, *t.shape[-2:]).to(self.dtype).contiguous()
        yh = []
        for _ in range(self.level):
            cur, hi = ops.dwt2_analysis(cur, None, self.filters, mode=self.mode, coeff_dtype=self.dtype)
            yh.append(hi.reshape(*lead, 3, *hi.shape[-2:]))
        return cur.reshape(*lead, *cur.shape[-2:]), tuple(yh)

    def inverse(
        self,
        yl: torch.Tensor,
        yh: Sequence,
        *,
        inverse_function: Callable | None = None,
        two_step_inverse: bool = False,
        yl_scale: float = 1.0,
        yh_scales=None,
    ):
        """IDWT. `yl_scale` / `yh_scales` (extension) fold `wavelet_scaling` into the synthesis kernels' loads
        instead of a pass per band."""
        if inverse_function is not None:
            return inverse_function((yl, yh))
        yh = tuple(yh)
        band_scales = ((1.0, 1.0, 1.0),) * len(yh) if yh_scales is None else expand_yh_scales(yh, yh_scales=yh_scales)
        lead = yl.shape[:-2]
        ll = yl.reshape(-1, *yl.shape[-2:]).to(self.dtype).contiguous()
        for level in range(len(yh) - 1, -1, -1):
            hi = yh[level]
            hi_p = hi.reshape(-1, 3, *hi.shape[-2:]).to(self.dtype).contiguous()
            sc = band_scales[level] if level < len(band_scales) else (1.0, 1.0, 1.0)
            if isinstance(sc, (int, float)):
                sc = (float(sc),) * 3
            sc = (*sc, 1.0, 1.0, 1.0)[:3]
            scales = (float(yl_scale) if level == len(yh) - 1 else 1.0, *sc)
            if self.inv_mode == "periodization":
                ll = ops.dwt2_synthesis_per(ll, hi_p, scales, self.inv_filters)
            else:
                ll = ops.dwt2_synthesis([(ll, hi_p, scales)], self.inv_filters)
        _ = two_step_inverse  # linear: one pass equals the two-step sum
        return ll.reshape(*lead, *ll.shape[-2:])

    def to(self, *args, copy: bool = False, **kwargs) -> "Wavelet":
        target = Wavelet.__new__(Wavelet) if copy else self
        if copy:
            target.__dict__.update(self.__dict__)
        probe = torch.empty(0).to(*args, **kwargs) if (args or kwargs) else None
        if probe is not None:
            if probe.is_floating_point() and ("dtype" in kwargs or any(isinstance(a, torch.dtype) for a in args)):
                target.dtype = probe.dtype
            if probe.device.type != "cpu" or "device" in kwargs:
                target.device = probe.device
        return target

    @staticmethod
    def wavelist() -> tuple:
        return ("haar", *(f"db{n}" for n in sorted(DB_DEC_LO)))

    @staticmethod
    def biortlist() -> tuple:
        return ()

    @staticmethod
    def qshiftlist() -> tuple:
        return ()

    @staticmethod
    def modelist() -> tuple:
        return tuple(ops.DWT_MODE_IDS)


def expand_yh_scales(yh: Sequence, *, yh_scales: float | Sequence = 1.0) -> tuple:
    """Per-level, per-orientation scale tuples incl. the "fill" shorthand (:148-190).

    `yh` only supplies the number of levels and orientations; a sequence of shapes works too."""
    levels = len(yh)
    shape0 = yh[0].shape if hasattr(yh[0], "shape") else tuple(yh[0])
    orientations = shape0[2] if len(shape0) > 3 else 1
    if isinstance(yh_scales, (float, int)):
        return ((float(yh_scales),) * orientations,) * levels
    ones = (1.0,) * orientations
    expanded = []
    for band in yh_scales:
        if isinstance(band, (float, int)):
            expanded.append((float(band),) * orientations)
        elif isinstance(band, (tuple, list)):
            head = tuple(float(v) for v in band[:orientations])
            expanded.append(head + ones[: orientations - len(head)])
        else:
            expanded.append(band)
    expanded = tuple(expanded)
    if "fill" in expanded:
        at = expanded.index("fill")
        if "fill" in expanded[at + 1 :]:
            raise ValueError("Only one fill allowed.")
        if at == 0 or len(expanded) < 2:
            raise ValueError("Invalid fill value, cannot be in the first position or the only item.")
        if len(expanded) - 1 < levels:
            pad = (expanded[at - 1],) * (levels - (len(expanded) - 1))
            expanded = (*expanded[:at], *pad, *expanded[at + 1 :])
        else:
            expanded = (*expanded[:at], *expanded[at + 1 :])
    return expanded[:levels]


def wavelet_scaling(yl: torch.Tensor, yh: Sequence, yl_scale, yh_scales, *, in_place: bool = False) -> tuple:
    """Scales the approximation and each detail band (:193-216)."""
    if not in_place:
        yl = yl.clone()
        yh = tuple(band.clone() for band in yh)
    if yl_scale != 1.0:
        yl *= yl_scale
    scales = expand_yh_scales(yh, yh_scales=yh_scales if yh_scales is not None else 1.0)
    for band_scale, band in zip(scales, yh):
        if isinstance(band_scale, (int, float)):
            band *= band_scale
            continue
        for o in range(min(band.shape[2], len(band_scale))):
            band[:, :, o] *= band_scale[o]
    return (yl, yh)


def wavelet_blend(a: tuple, b: tuple, *, yl_factor, blend_function: Callable, yh_factor=None, yh_blend_function=None):
    """Band-wise blend of two coefficient sets (:219-238)."""
    if not isinstance(yl_factor, torch.Tensor):
        yl_factor = a[0].new_full((1,), yl_factor)
    if yh_factor is None:
        yh_factor = yl_factor
    elif not isinstance(yh_factor, torch.Tensor):
        yh_factor = a[0].new_full((1,), yh_factor)
    yh_blend_function = fallback(yh_blend_function, blend_function)
    return (
        blend_function(a[0], b[0], yl_factor),
        tuple(yh_blend_function(ta, tb, yh_factor) for ta, tb in zip(a[1], b[1])),
    )
