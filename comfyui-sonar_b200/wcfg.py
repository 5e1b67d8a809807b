"""Wavelet band-split CFG on the fused DWT kernels.

Mirror of the reference's py/wavelet_cfg.py. The rule / schedule objects (host scalars parsed from
YAML) keep the reference's names, fields and defaults; the tensor path -- upstream: 2 forward
transforms, ~8 coefficient passes and 1 inverse transform in eager fp64 (`wavelet_cfg :749-791`,
`process_output :729-747`) -- is restructured around linearity:

    result_w = final * blend(uncond_w', diff_w', t),   uncond_w' = Su*U, diff_w' = Sd*(Sc*C - Su*U)
             = A (.) DWT(cond) + B (.) DWT(uncond)                      (per-band scalars A, B)
             = A (.) DWT(cond - uncond) + (A + B) (.) DWT(uncond)

and when A + B is the same constant c0 in every band (the default configuration: A + B == 1) the
second term is c0 * uncond by perfect reconstruction, so ONE analysis/synthesis pair over
(cond - uncond) suffices; the synthesis kernel applies A on load, and its final level fuses
"+ c0*uncond", the crop to x.shape, the fp32 cast and "x - result".

When the coefficient pyramid of a plane fits one SM's shared memory (db2 / 3 levels / 128x128 in
fp64: 184 KB) that whole pair is ONE launch (`ops.wcfg_fused`, one CTA per plane); otherwise one
launch per level with the coefficients in L2/HBM.
"""

from __future__ import annotations

import math
from enum import Enum, auto
from typing import Callable, NamedTuple, Sequence

import torch
from tqdm import tqdm

from . import hostutil, ops
from .hostutil import clamp_float, filter_dict
from .wavelets import Wavelet, expand_yh_scales


def pretty_non_default(obj: NamedTuple, *, defaults: object | None = None) -> str:
    parts = []
    for name in obj._fields:
        val = getattr(obj, name)
        if defaults is not None and val == getattr(defaults, name):
            continue
        parts.append(f"{name}={val.pretty_non_default()}" if hasattr(val, "pretty_non_default") else f"{name}={val!r}")
    return f"{obj.__class__.__name__}({', '.join(parts)})"


class WCFGSchedule(Enum):
    LINEAR = auto()
    LOGARITHMIC = auto()
    LOG = LOGARITHMIC
    EXPONENTIAL = auto()
    EXP = EXPONENTIAL
    HALF_COSINE = auto()
    SINE = auto()
    SIN = SINE

    def interp(self, val: float) -> float:
        val = clamp_float(val)
        if self == WCFGSchedule.LINEAR:
            return val
        if self == WCFGSchedule.LOGARITHMIC:
            out = 0.0 if val == 0 else math.log(val) + 1.0
        elif self == WCFGSchedule.EXPONENTIAL:
            out = math.exp(val) - 1.0
        elif self == WCFGSchedule.HALF_COSINE:
            out = 1.0 - ((1.0 + math.cos(val * math.pi)) / 2)
        elif self == WCFGSchedule.SINE:
            out = math.sin(val * math.pi)
        else:
            raise ValueError("Bad interpolation schedule!?")
        return clamp_float(out)


class WCFGSchedMode(Enum):
    SAMPLING = auto()
    ENABLED_SAMPLING = auto()
    SIGMAS = auto()
    ENABLED_SIGMAS = auto()
    STEP = auto()
    ENABLED_STEPS = auto()
    MODEL_SAMPLING = SAMPLING
    ENABLED_MODEL_SAMPLING = ENABLED_SAMPLING
    SIGMA_RANGE = SIGMAS
    ENABLED_SIGMA_RANGE = ENABLED_SIGMAS


class WCFGTarget(Enum):
    DENOISED = auto()
    NOISE = auto()
    NOISE_NORM = auto()


class WCFGPercentages(NamedTuple):
    sigma: float
    sigma_min: float
    sigma_max: float
    sigma_first: float | None
    sigma_last: float | None
    steps: int | None
    step: float | None
    step_first: int | None
    step_last: int | None
    pct_sampling: float
    pct_enabled_sampling: float
    pct_sigmas: float | None
    pct_enabled_sigmas: float | None
    pct_steps: float | None
    pct_enabled_steps: float | None

    def invert(self) -> "WCFGPercentages":
        def flip(v):
            return None if v is None else 1.0 - v

        return self._replace(
            pct_sampling=1.0 - self.pct_sampling,
            pct_enabled_sampling=1.0 - self.pct_enabled_sampling,
            pct_sigmas=flip(self.pct_sigmas),
            pct_enabled_sigmas=flip(self.pct_enabled_sigmas),
            pct_steps=flip(self.pct_steps),
            pct_enabled_steps=flip(self.pct_enabled_steps),
        )

    def pct_from_schedmode(self, mode: WCFGSchedMode) -> float | None:
        if mode == WCFGSchedMode.MODEL_SAMPLING:
            return self.pct_sampling
        if mode == WCFGSchedMode.SIGMA_RANGE:
            return self.pct_sigmas
        if mode == WCFGSchedMode.ENABLED_MODEL_SAMPLING:
            return self.pct_enabled_sampling
        if mode == WCFGSchedMode.ENABLED_SIGMA_RANGE:
            return self.pct_enabled_sigmas
        if mode == WCFGSchedMode.STEP:
            if self.pct_steps is None:
                raise RuntimeError("Step percentage not available")
            return self.pct_steps
        raise ValueError("Unknown mode")

    @classmethod
    def build(cls, *, ms: object, start_sigma: float, end_sigma: float, sigma: float, sigmas, **_kwargs):
        """Host-side progress percentages (reference :127-211). `ms` is ComfyUI's model_sampling."""
        if start_sigma < end_sigma:
            raise ValueError("start/end sigmas out of order")
        sigma_max = ms.sigma_max.detach().item()
        sigma_min = ms.sigma_min.detach().item()
        start_sigma = min(sigma_max, start_sigma)
        end_sigma = min(max(sigma_min, end_sigma), sigma_max)
        sigma = min(max(sigma, sigma_min), sigma_max)

        def pct_of(s: float) -> float:
            return 1.0 - (ms.timestep(torch.tensor(s)) / 999).clamp(0, 1).detach().item()

        pct_start, pct_end, pct_curr = pct_of(start_sigma), pct_of(end_sigma), pct_of(sigma)
        pct_range_curr = (pct_curr - pct_start) / (pct_end - pct_start)
        pct_sigmas = pct_enabled_sigmas = step = steps = pct_steps = pct_enabled_steps = None
        sigma_first = sigma_last = step_first = step_last = None
        if sigmas is not None:
            if sigmas.ndim == 2:
                sigmas = sigmas.max(dim=0).values
            elif sigmas.ndim != 1:
                raise ValueError("Unexpected number of dimensions for sample_sigmas")
            sigmas = sigmas.detach().cpu()
            sigma_first, sigma_last = sigmas[0].item(), sigmas[-2].item()
            if sigma_first <= sigma_last:
                raise ValueError("Cannot handle non-descending sigmas (possibly Restart or unsampling)")
            pct_sigmas = (sigma_first - sigma) / (sigma_first - sigma_last)
            start_sigma = min(start_sigma, sigma_first)
            end_sigma = max(end_sigma, sigma_last)
            sigma = min(max(sigma, sigma_last), sigma_first)
            pct_enabled_sigmas = 1.0 if start_sigma == end_sigma else (start_sigma - sigma) / (start_sigma - end_sigma)
            steps = len(sigmas) - 1
            if steps > 1:
                step = hostutil.step_from_sigmas(sigma, sigmas)
                pct_steps = step / (steps - 1) if step is not None else None
                enabled = torch.arange(len(sigmas), dtype=torch.int32)[(sigmas <= start_sigma) & (sigmas >= end_sigma)]
                if len(enabled) > 1:
                    step_first, step_last = enabled[0].item(), enabled[-1].item()
                    pct_enabled_steps = (step - step_first) / (step_last - step_first)
            else:
                step, pct_steps = 0.0, 1.0
        return WCFGPercentages(
            pct_sampling=pct_curr,
            pct_enabled_sampling=pct_range_curr,
            pct_sigmas=pct_sigmas,
            pct_enabled_sigmas=pct_enabled_sigmas,
            pct_steps=pct_steps,
            pct_enabled_steps=pct_enabled_steps,
            sigma=sigma,
            sigma_first=sigma_first,
            sigma_last=sigma_last,
            sigma_min=sigma_min,
            sigma_max=sigma_max,
            steps=steps,
            step=step,
            step_first=step_first,
            step_last=step_last,
        )


class WCFGScales(NamedTuple):
    yl_scale: float = 1.0
    yh_scales: float | Sequence = 1.0

    def get_scales(self, *_args, verbose: bool = False, **_kwargs) -> "WCFGScales":
        if verbose:
            tqdm.write(f"WCFG:     {self.pretty_scales()}")
        return self

    def pretty_yh_scales(self, *, target=None) -> str:
        target = self.yh_scales if target is None else target
        if isinstance(target, float):
            return f"{target:.4f}"
        inner = ", ".join(
            self.pretty_yh_scales(target=v) if isinstance(v, (list, tuple)) else (v if isinstance(v, str) else f"{v:.4f}")
            for v in target
        )
        return f"({inner})"

    def pretty_scales(self) -> str:
        return f"low={self.yl_scale:.4f}, high={self.pretty_yh_scales()}"


class WCFGScheduledScale(NamedTuple):
    schedule: WCFGSchedule = WCFGSchedule.LINEAR
    schedule_mode: WCFGSchedMode = WCFGSchedMode.ENABLED_MODEL_SAMPLING
    schedule_offset: float = 0.0
    schedule_offset_after: float = 0.0
    schedule_multiplier: float = 1.0
    schedule_multiplier_after: float = 1.0
    reverse_schedule: bool = False
    reverse_schedule_after: bool = False
    schedule_min: float = 0.0
    schedule_max: float = 1.0

    @classmethod
    def build(cls, **kwargs) -> "WCFGScheduledScale":
        schedule = kwargs.pop("schedule", DEFAULT_SCHEDULEDSCALE.schedule)
        if isinstance(schedule, str):
            schedule = getattr(WCFGSchedule, schedule.upper())
        schedule_mode = kwargs.pop("schedule_mode", DEFAULT_SCHEDULEDSCALE.schedule_mode)
        if isinstance(schedule_mode, str):
            schedule_mode = getattr(WCFGSchedMode, schedule_mode.upper())
        return WCFGScheduledScale(schedule=schedule, schedule_mode=schedule_mode, **filter_dict(kwargs, cls._fields))

    def get_b_scale(self, pcts: WCFGPercentages) -> float:
        if self.reverse_schedule:
            pcts = pcts.invert()
        pct = pcts.pct_from_schedmode(self.schedule_mode)
        if pct is None:
            raise RuntimeError("Couldn't get percentage")
        shaped = self.schedule.interp(clamp_float((pct + self.schedule_offset) * self.schedule_multiplier))
        pct = clamp_float(
            (shaped + self.schedule_offset_after) * self.schedule_multiplier_after,
            minval=clamp_float(self.schedule_min),
            maxval=clamp_float(self.schedule_max),
        )
        return clamp_float(1.0 - pct) if self.reverse_schedule_after else pct

    def pretty_non_default(self) -> str:
        return pretty_non_default(self, defaults=DEFAULT_SCHEDULEDSCALE)


DEFAULT_SCHEDULEDSCALE = WCFGScheduledScale()


class WCFGScalesRange(NamedTuple):
    scales_start: WCFGScales = WCFGScales()
    scales_end: WCFGScales | None = None
    scheduler: WCFGScheduledScale | None = None
    blend_mode: str = "lerp"

    @classmethod
    def build(cls, **kwargs):
        scales_start = kwargs.pop("scales_start", None)
        if scales_start is None:
            scales_start = {"yl_scale": kwargs.pop("yl_scale", 1.0), "yh_scales": kwargs.pop("yh_scales", 1.0)}
        scales_end = filter_dict(kwargs.pop("scales_end", {}), WCFGScales._fields)
        if not scales_end or scales_end == scales_start:
            return WCFGScales(yl_scale=scales_start.get("yl_scale", 1.0), yh_scales=scales_start.get("yh_scales", 1.0))
        blend_mode = kwargs.pop("blend_mode", "lerp")
        return WCFGScalesRange(
            scales_start=WCFGScales(**scales_start),
            scales_end=WCFGScales(**scales_end),
            scheduler=WCFGScheduledScale.build(**kwargs),
            blend_mode=blend_mode,
        )

    def get_scales(self, pcts: WCFGPercentages, yh: Sequence, *, verbose: bool = False) -> WCFGScales:
        if self.scales_end is None or self.scheduler is None:
            return self.scales_start.get_scales()
        pct = self.scheduler.get_b_scale(pcts)
        if verbose:
            tqdm.write(f"WCFG:   pct={pct:.4f}, percentages: {pcts}")
        start, end = self.scales_start, self.scales_end
        if self.blend_mode == "lerp" and (pct <= 0 or pct >= 1):
            picked = start if pct <= 0 else end
            if verbose:
                tqdm.write(f"WCFG:     {picked.pretty_scales()}")
            return picked
        blend_function = None if self.blend_mode == "lerp" else hostutil.BLENDING_MODES[self.blend_mode]
        s_yh = expand_yh_scales(yh, yh_scales=start.yh_scales)
        e_yh = expand_yh_scales(yh, yh_scales=end.yh_scales)
        result = WCFGScales(
            yl_scale=hostutil.blend_scalar(start.yl_scale, end.yl_scale, pct, blend_function=blend_function),
            yh_scales=tuple(
                tuple(hostutil.blend_scalar(a, b, pct, blend_function=blend_function) for a, b in zip(bs, be))
                for bs, be in zip(s_yh, e_yh)
            ),
        )
        if verbose:
            tqdm.write(f"WCFG:     {result.pretty_scales()}")
        return result

    def pretty_non_default(self) -> str:
        return pretty_non_default(self, defaults=DEFAULT_SCALESRANGE)


DEFAULT_SCALESRANGE = WCFGScalesRange()


class WCFGScheduledFloat(NamedTuple):
    value_start: float
    value_end: float | None = None
    scheduler: WCFGScheduledScale | None = None

    @classmethod
    def build(cls, val, *, default_start=None, default_end=None, **_kwargs) -> "WCFGScheduledFloat":
        if isinstance(val, float):
            return WCFGScheduledFloat(value_start=val)
        if not isinstance(val, dict):
            raise TypeError("Bad type for scheduled float value")
        val = val.copy()
        value_start = val.pop("value_start", default_start)
        value_end = val.pop("value_end", default_end)
        if not isinstance(value_start, (float, int)):
            raise TypeError("Bad type for scheduled float start_value")
        if value_end is None:
            return WCFGScheduledFloat(value_start=val)
        if not isinstance(value_end, (float, int)):
            raise TypeError("Bad type for scheduled float end_value")
        return WCFGScheduledFloat(
            value_start=float(value_start),
            value_end=float(value_end),
            scheduler=WCFGScheduledScale.build(**val),
        )

    def get_value(self, pcts: WCFGPercentages) -> float:
        if self.value_end is None or self.scheduler is None:
            return self.value_start
        pct = self.scheduler.get_b_scale(pcts)
        return (1.0 - pct) * self.value_start + pct * self.value_end


class WCFGWaveletSettings(NamedTuple):
    wave: str = "db4"
    level: int = 5
    padding_mode: str = "symmetric"
    use_1d_dwt: bool = False
    use_dtcwt: bool = False
    biort: str = "near_sym_a"
    qshift: str = "qshift_a"
    inv_wave: str | None = None
    inv_padding_mode: str | None = None
    inv_biort: str | None = None
    inv_qshift: str | None = None

    @classmethod
    def build(cls, **kwargs) -> "WCFGWaveletSettings":
        return WCFGWaveletSettings(**filter_dict(kwargs, cls._fields))

    def make_wavelet(self, **kwargs) -> Wavelet:
        if "per" in (self.padding_mode, self.inv_padding_mode) or "periodization" in (self.padding_mode, self.inv_padding_mode):
            raise NotImplementedError(
                "sonar_b200: wavelet CFG runs on the expansive DWT kernels (symmetric / zero / reflect / periodic); "
                "the non-expansive periodization mode is only built for the wavelet-filtered noise type",
            )
        return Wavelet(
            wave=self.wave,
            level=self.level,
            mode=self.padding_mode,
            use_1d_dwt=self.use_1d_dwt,
            use_dtcwt=self.use_dtcwt,
            biort=self.biort,
            qshift=self.qshift,
            inv_wave=self.inv_wave,
            inv_mode=self.inv_padding_mode,
            inv_biort=self.inv_biort,
            inv_qshift=self.inv_qshift,
            **kwargs,
        )

    def pretty_non_default(self) -> str:
        return pretty_non_default(self, defaults=DEFAULT_WAVELETSETTINGS)


DEFAULT_WAVELETSETTINGS = WCFGWaveletSettings()


class WCFGRule(NamedTuple):
    start_sigma: float = math.inf
    end_sigma: float = 0.0
    verbose: bool = False
    blend_mode: str = "lerp"
    blend_strength: WCFGScheduledFloat = WCFGScheduledFloat(1.0)
    fallback_existing: bool = True
    target_mode: WCFGTarget = WCFGTarget.DENOISED
    diff: WCFGScalesRange | WCFGScales | None = None
    cond: WCFGScalesRange | WCFGScales | None = None
    uncond: WCFGScalesRange | WCFGScales | None = None
    final: WCFGScalesRange | WCFGScales | None = None
    wavelet: WCFGWaveletSettings = DEFAULT_WAVELETSETTINGS
    high_precision_mode: bool = True
    difference_blend_mode: str = "inject"
    difference_blend_strength: WCFGScheduledFloat = WCFGScheduledFloat(1.0)

    @classmethod
    def build(cls, **kwargs) -> "WCFGRule":
        target_mode = kwargs.pop("target_mode", DEFAULT_RULE.target_mode)
        if isinstance(target_mode, str):
            target_mode = getattr(WCFGTarget, target_mode.upper())
        difference = kwargs.pop("diff", None)
        if difference is None:
            difference = kwargs.pop("difference", None)

        def scales(spec):
            return None if spec is None else WCFGScalesRange.build(**spec)

        cond, uncond, final = (scales(kwargs.pop(k, None)) for k in ("cond", "uncond", "final"))
        blend_strength = kwargs.pop("blend_strength", 1.0)
        if not isinstance(blend_strength, (float, int, dict)):
            raise TypeError("Bad type for blend_strength, must be float or dict")
        difference_blend_strength = kwargs.pop("difference_blend_strength", 1.0)
        if not isinstance(difference_blend_strength, (float, int, dict)):
            raise TypeError("Bad type for difference_blend_strength, must be float or dict")
        return WCFGRule(
            target_mode=target_mode,
            diff=scales(difference),
            cond=cond,
            uncond=uncond,
            final=final,
            blend_strength=WCFGScheduledFloat(blend_strength),
            difference_blend_strength=WCFGScheduledFloat(difference_blend_strength),
            wavelet=WCFGWaveletSettings.build(**kwargs),
            **filter_dict(kwargs, cls._fields),
        )

    def make_wavelet(self, **kwargs) -> Wavelet:
        return self.wavelet.make_wavelet(**kwargs)

    def band_scales(self, name: str, pcts: WCFGPercentages, band_shapes: Sequence, *, verbose: bool = False):
        """(yl_scale, per-level orientation tuples) of the `name` scale set, or identity."""
        spec = getattr(self, name)
        levels = len(band_shapes)
        if spec is None:
            return 1.0, ((1.0, 1.0, 1.0),) * levels
        scales = spec.get_scales(pcts, band_shapes)
        if verbose and (scales.yl_scale != 1.0 or scales.yh_scales != 1.0):
            tqdm.write(f"WCFG:     scales({name:>6}): {scales.pretty_scales()}")
        yh = expand_yh_scales(band_shapes, yh_scales=scales.yh_scales if scales.yh_scales is not None else 1.0)
        yh = tuple(yh) + ((1.0, 1.0, 1.0),) * (levels - len(yh))
        return float(scales.yl_scale), yh

    def pretty_non_default(self) -> str:
        return pretty_non_default(self, defaults=DEFAULT_RULE)


DEFAULT_RULE = WCFGRule()


class WCFGRules(NamedTuple):
    rules: Sequence = ()

    def __len__(self) -> int:
        return len(self.rules)

    def __getitem__(self, idx: int) -> WCFGRule:
        return self.rules[idx]

    def __bool__(self) -> bool:
        return bool(self.rules)

    def get_rule(self, sigma: float) -> WCFGRule | None:
        for rule in self.rules:
            if rule.end_sigma <= sigma <= (math.inf if rule.start_sigma < 0 else rule.start_sigma):
                return rule
        return None

    @classmethod
    def build(cls, **params) -> "WCFGRules":
        params = params.copy()
        extra = params.pop("rules", ())
        return WCFGRules(rules=(WCFGRule.build(**params), *(WCFGRule.build(**r) for r in extra)))


class WCFGContext(NamedTuple):
    cond: torch.Tensor
    uncond: torch.Tensor
    x: torch.Tensor
    sigma: torch.Tensor
    wavelet: Wavelet
    dtype: torch.dtype
    op_kwargs: dict


def _linear_band_coefficients(rule: WCFGRule, pcts: WCFGPercentages, band_shapes: Sequence, *, verbose: bool):
    """Per-band scalars (A, B) with result_w = A*cond_w + B*uncond_w, from the four scale sets and
    the difference blend. Index 0 is the approximation band, then (level, orientation)."""
    t = rule.difference_blend_strength.get_value(pcts)
    mode = rule.difference_blend_mode
    if mode not in ops.BLEND_IDS:
        raise KeyError(mode)
    sets = {name: rule.band_scales(name, pcts, band_shapes, verbose=verbose) for name in ("cond", "uncond", "diff", "final")}

    def combine(sc: float, su: float, sd: float, sf: float) -> tuple[float, float]:
        if mode == "inject":  # uncond' + diff' * t
            return sf * t * sd * sc, sf * su * (1.0 - t * sd)
        if mode == "subtract_b":  # uncond' - diff' * t
            return -sf * t * sd * sc, sf * su * (1.0 + t * sd)
        return sf * t * sd * sc, sf * su * ((1.0 - t) - t * sd)  # lerp(uncond', diff', t)

    a_ll, b_ll = combine(*(sets[n][0] for n in ("cond", "uncond", "diff", "final")))
    a_hi, b_hi = [], []
    for lvl in range(len(band_shapes)):
        pairs = [combine(*(sets[n][1][lvl][o] for n in ("cond", "uncond", "diff", "final"))) for o in range(3)]
        a_hi.append(tuple(p[0] for p in pairs))
        b_hi.append(tuple(p[1] for p in pairs))
    return (a_ll, a_hi), (b_ll, b_hi)


class WaveletCFG:
    """The sampler_cfg_function (reference :631-842): fn(args) -> Tensor."""

    def __init__(
        self,
        *,
        existing_cfg: Callable | None,
        rules: WCFGRules,
        operation_cond: Callable | None = None,
        operation_uncond: Callable | None = None,
        operation_fallback_cfg: Callable | None = None,
        operation_wavelet_cfg: Callable | None = None,
        operation_result: Callable | None = None,
    ):
        self.wavelet_cache: dict = {}
        self.rules = rules
        self.fallback_cfg_function = (
            existing_cfg if existing_cfg is not None and (not rules or rules[0].fallback_existing) else self.basic_cfg_function
        )
        self.operation_cond = operation_cond
        self.operation_uncond = operation_uncond
        self.operation_fallback_cfg = operation_fallback_cfg
        self.operation_wavelet_cfg = operation_wavelet_cfg
        self.operation_result = operation_result

    @staticmethod
    def basic_cfg_function(args: dict) -> torch.Tensor:
        x, scale = args["input"], args["cond_scale"]
        uncond, cond = args["uncond_denoised"], args["cond_denoised"]
        return x - (cond - uncond).mul_(scale).add_(uncond)

    @staticmethod
    def maybe_op(t: torch.Tensor, mop: Callable | None, **kwargs) -> torch.Tensor:
        if mop is None:
            return t
        return mop(latent=t, **(kwargs if getattr(mop, "EXTENDED_LATENT_OPERATION", None) else {}))

    def get_context(self, *, rule: WCFGRule, args: dict) -> WCFGContext:
        sigma_orig = sigma = args["sigma"]
        x = args["input"]
        if x.ndim == 3 and not rule.wavelet.use_1d_dwt:
            raise RuntimeError("Enable use_1d_dwt mode for 3D latents.")
        if x.ndim < 3:
            raise RuntimeError("Wavelet CFG can't handle latents with 2 or less dimensions.")
        if sigma.ndim != x.ndim:
            sigma = sigma.reshape(x.shape[0], *((1,) * (x.ndim - sigma.ndim)))
        if rule.target_mode in {WCFGTarget.NOISE, WCFGTarget.NOISE_NORM}:
            cond, uncond = args["cond"], args["uncond"]
            if rule.target_mode == WCFGTarget.NOISE_NORM:
                cond, uncond = cond / sigma, uncond / sigma
        elif rule.target_mode == WCFGTarget.DENOISED:
            cond, uncond = args["cond_denoised"], args["uncond_denoised"]
        else:
            raise ValueError("Bad target mode")
        op_kwargs = {"sigma": sigma_orig, "cond": cond, "uncond": uncond, "cond_scale": args["cond_scale"], "raw_args": args}
        cond = self.maybe_op(cond, self.operation_cond, **op_kwargs)
        uncond = self.maybe_op(uncond, self.operation_uncond, **op_kwargs)
        eff_dtype = torch.float64 if rule.high_precision_mode else x.dtype
        wavelet = self.wavelet_cache.get(id(rule))
        if wavelet is None:
            wavelet = rule.make_wavelet()
            self.wavelet_cache[id(rule)] = wavelet
        wavelet = wavelet.to(device=x.device, dtype=eff_dtype)
        if x.ndim > 4:
            cond = cond.flatten(start_dim=1, end_dim=cond.ndim - 3)
            uncond = uncond.flatten(start_dim=1, end_dim=uncond.ndim - 3)
        return WCFGContext(cond=cond, uncond=uncond, x=x, sigma=sigma, wavelet=wavelet, dtype=eff_dtype, op_kwargs=op_kwargs)

    @classmethod
    def wavelet_cfg(cls, *, rule: WCFGRule, ctx: WCFGContext, pcts: WCFGPercentages, epilogue: dict | None = None):
        """Reconstruction of the band-weighted combination, fp32, shaped (planes, H', W').

        Without `epilogue` the output covers the full reconstruction (like the reference's
        `wavelet.inverse(...)`, possibly 1 px larger than the input); with it the final synthesis
        level also crops to (h, w) and forms x_scale*x + sign*result."""
        wavelet = ctx.wavelet
        coeff_dtype = torch.float64 if ctx.dtype == torch.float64 else torch.float32
        cond = ctx.cond.reshape(-1, *ctx.cond.shape[-2:]).to(torch.float32).contiguous()
        uncond = ctx.uncond.reshape(-1, *ctx.uncond.shape[-2:]).to(torch.float32).contiguous()
        planes, height, width = cond.shape
        taps = wavelet.filters.length
        # band geometry, fine -> coarse
        shapes, hh, ww = [], height, width
        for _ in range(wavelet.level):
            hh, ww = ops.dwt_coeff_len(hh, taps), ops.dwt_coeff_len(ww, taps)
            shapes.append((planes, 1, 3, hh, ww))
        (a_ll, a_hi), (b_ll, b_hi) = _linear_band_coefficients(rule, pcts, shapes, verbose=rule.verbose)
        sums = [a_ll + b_ll] + [a + b for ah, bh in zip(a_hi, b_hi) for a, b in zip(ah, bh)]
        c0 = sums[0]
        full_size = (2 * shapes[0][-2] - taps + 2, 2 * shapes[0][-1] - taps + 2)
        crop = tuple(epilogue["crop"]) if epilogue is not None and epilogue.get("crop") is not None else full_size
        # one transform pair is enough when A + B is band-independent (then (A+B) (.) DWT(u) == c0 * u);
        # the c0 * u term is added at the input resolution, so the output must be cropped to it
        single = all(abs(s - c0) <= 1e-12 * max(1.0, abs(c0)) for s in sums) and crop == (height, width)

        if (
            single
            and wavelet.inv_filters is wavelet.filters
            and ops.wcfg_fused_fits(height, width, taps, wavelet.level, use_f64=coeff_dtype == torch.float64)
        ):
            # everything on chip: one launch, fp32 inputs read once, fp32 result written once
            ep = epilogue or {}
            return ops.wcfg_fused(
                cond, uncond, wavelet.filters, levels=wavelet.level, mode=wavelet.mode,
                use_f64=coeff_dtype == torch.float64, scale_ll=a_ll, scale_hi=a_hi,
                addend=uncond if c0 != 0.0 else None, addend_scale=c0,
                x=ep.get("x"), x_scale=ep.get("x_scale", 1.0), recon_sign=ep.get("recon_sign", 1.0),
            )  # fmt: skip

        def analyse(a, b):
            cur, his = (a, b), []
            for lvl in range(wavelet.level):
                if lvl == 0:
                    ll, hi = ops.dwt2_analysis(cur[0], cur[1], wavelet.filters, mode=wavelet.mode, coeff_dtype=coeff_dtype)
                else:
                    ll, hi = ops.dwt2_analysis(cur, None, wavelet.filters, mode=wavelet.mode, coeff_dtype=coeff_dtype)
                his.append(hi)
                cur = ll
            return cur, his

        ll_d, hi_d = analyse(cond, uncond)
        if single:
            ll_u = hi_u = None
        else:
            ll_u, hi_u = analyse(uncond, None)

        rec = None
        for lvl in reversed(range(wavelet.level)):
            coarsest = lvl == wavelet.level - 1
            sets = [(ll_d if coarsest else rec, hi_d[lvl], (a_ll if coarsest else 1.0, *a_hi[lvl]))]
            if not single:
                ab = tuple(a + b for a, b in zip(a_hi[lvl], b_hi[lvl]))
                sets.append((ll_u if coarsest else rec, hi_u[lvl], ((a_ll + b_ll) if coarsest else 0.0, *ab)))
            final = None
            if lvl == 0:
                final = dict(epilogue) if epilogue is not None else {}
                final["crop"] = crop
                if single and c0 != 0.0:
                    final["addend"], final["addend_scale"] = uncond, c0
            rec = ops.dwt2_synthesis(sets, wavelet.inv_filters, final=final)
        return rec

    def __call__(self, args: dict) -> torch.Tensor:
        sigma = args["sigma"]
        sigma_f = sigma.max().item()
        rule = self.rules.get_rule(sigma_f)
        if rule is None:
            return self.fallback_cfg_function(args)
        if rule.verbose:
            tqdm.write(f"\nWCFG: Rule matched, sigma={sigma_f:.4f}, rule={rule.pretty_non_default()}")
        model = args["model"]
        pcts = WCFGPercentages.build(
            ms=model.model_sampling,
            start_sigma=rule.start_sigma,
            end_sigma=rule.end_sigma,
            sigma=sigma_f,
            sigmas=args.get("model_options", {}).get("transformer_options", {}).get("sample_sigmas"),
        )
        wcfg_blend = rule.blend_strength.get_value(pcts)
        if rule.blend_mode == "lerp" and wcfg_blend == 0:
            return self.maybe_op(
                self.fallback_cfg_function(args),
                self.operation_fallback_cfg,
                sigma=sigma,
                cond=args["cond_denoised"],
                uncond=args["uncond_denoised"],
                raw_args=args,
            )
        ctx = self.get_context(rule=rule, args=args)
        x = ctx.x
        needs_blend = rule.blend_mode != "lerp" or wcfg_blend != 1.0
        simple_tail = (
            not needs_blend
            and self.operation_wavelet_cfg is None
            and rule.target_mode in {WCFGTarget.DENOISED, WCFGTarget.NOISE}
            and x.dtype == torch.float32
        )
        crop = tuple(x.shape[-2:])
        if simple_tail:
            # crop + cast + (x - result) fused into the last synthesis level
            epilogue = {"crop": crop}
            if rule.target_mode == WCFGTarget.DENOISED:
                epilogue |= {"x": x.reshape(-1, *crop).contiguous(), "x_scale": 1.0, "recon_sign": -1.0}
            result = self.wavelet_cfg(rule=rule, ctx=ctx, pcts=pcts, epilogue=epilogue).reshape(x.shape)
        else:
            result = self.wavelet_cfg(rule=rule, ctx=ctx, pcts=pcts, epilogue={"crop": crop}).reshape(x.shape).to(x.dtype)
            if needs_blend:
                normal = self.maybe_op(self.fallback_cfg_function(args), self.operation_fallback_cfg, **ctx.op_kwargs)
                if rule.target_mode == WCFGTarget.DENOISED:
                    normal = x - normal
                elif rule.target_mode == WCFGTarget.NOISE_NORM:
                    normal /= ctx.sigma
                result = hostutil.BLENDING_MODES[rule.blend_mode](normal.contiguous(), result.contiguous(), wcfg_blend)
            if rule.target_mode == WCFGTarget.DENOISED:
                result = x - result
            elif rule.target_mode == WCFGTarget.NOISE_NORM:
                result *= ctx.sigma
            result = self.maybe_op(result, self.operation_wavelet_cfg, **ctx.op_kwargs)
        return self.maybe_op(result, self.operation_result, **ctx.op_kwargs).contiguous()
