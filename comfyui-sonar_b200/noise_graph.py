"""Noise graph: chains, items and combinators above the generators.

Host-side mirror of the in-scope part of the reference's py/noise.py -- `SONAR_CUSTOM_NOISE`
objects with `.clone()/.add()/.rescaled()/.factor/.items/.make_noise_sampler(...)` returning
`ns(sigma, sigma_next) -> Tensor` closures. Every tensor pass inside the
closures is a CUDA kernel launch; the control flow (which child to call, host-RNG draws that pick
flips and rolls) stays in Python exactly as the reference has it.
"""

from __future__ import annotations

import abc
import math
import random
from functools import partial
from typing import Callable

import torch
import yaml

from . import hostutil, ops, parallel
from .generators import *  # noqa: F403  (re-exported like the reference's `from .noise_generation import *`)
from .generators import (
    GaussianNoiseGenerator,
    GreenTestNoiseGenerator,
    HighresPyramidNoiseGenerator,
    MixedNoiseGenerator,
    NoiseType,
    OneFNoiseGenerator,
    PerlinOldNoiseGenerator,
    PowerLawNoiseGenerator,
    PyramidNoiseGenerator,
    PyramidOldNoiseGenerator,
    UniformNoiseGenerator,
)
from .hostutil import fallback, scale_noise

_CLONED_KEYS = frozenset(
    ("custom_noise", "custom_noise_opt", "noise", "noise_opt", "sonar_custom_noise", "sonar_custom_noise_opt"),
)


class CustomNoiseItemBase(abc.ABC):
    """Base of every chain item (reference py/noise.py:30-80)."""

    def __init__(self, factor, *, yaml_parameters=None, **kwargs):
        if yaml_parameters:
            extra = yaml.safe_load(yaml_parameters)
            if extra is not None:
                if not isinstance(extra, dict):
                    raise ValueError("CustomNoiseItem: yaml_parameters must either be null or an object")
                kwargs["ns_kwargs"] = extra
        self.factor = factor
        self.keys = set(kwargs)
        for key, val in kwargs.items():
            if key in _CLONED_KEYS and hasattr(val, "clone"):
                val = val.clone()
            setattr(self, key, val)

    def clone_key(self, k):
        return getattr(self, k)

    def clone(self):
        return self.__class__(self.factor, **{k: self.clone_key(k) for k in self.keys})

    def set_factor(self, factor):
        self.factor = factor
        return self

    def get_normalize(self, k, default=None):
        val = getattr(self, k, None)
        return default if val is None else val

    @abc.abstractmethod
    def make_noise_sampler(self, x, sigma_min=None, sigma_max=None, seed=None, cpu=True, normalized=True, **kwargs):
        raise NotImplementedError


class CustomNoiseItem(CustomNoiseItemBase):
    """A built-in noise type by name (:83-134)."""

    def __init__(self, factor, **kwargs):
        super().__init__(factor, **kwargs)
        if getattr(self, "noise_type", None) is None:
            raise ValueError("Noise type required!")

    @torch.no_grad()
    def make_noise_sampler(self, x, sigma_min=None, sigma_max=None, seed=None, cpu=True, normalized=True, **kwargs):
        ns_kwargs = dict(getattr(self, "ns_kwargs", {}))
        o_sigma, o_sigma_next, o_min, o_max = (
            ns_kwargs.pop(k, None)
            for k in ("override_sigma", "override_sigma_next", "override_sigma_min", "override_sigma_max")
        )
        ns = get_noise_sampler(
            self.noise_type,
            x,
            fallback(o_min, sigma_min),
            fallback(o_max, sigma_max),
            seed=ns_kwargs.pop("seed", seed),
            cpu=ns_kwargs.pop("cpu", cpu),
            factor=self.factor,
            normalized=ns_kwargs.pop("normalized", self.get_normalize("normalize", normalized)),
            **ns_kwargs,
            **kwargs,
        )
        if o_sigma is None and o_sigma_next is None:
            return ns
        return lambda sigma, sigma_next: ns(fallback(o_sigma, sigma), fallback(o_sigma_next, sigma_next))


class CustomNoiseChain:
    """SONAR_CUSTOM_NOISE (:137-196): children are built un-normalised, summed in place, and the
    sum is normalised once with total factor = sum |factor_i|."""

    def __init__(self, items=None):
        self.items = items if items is not None else []

    def clone(self):
        return CustomNoiseChain([i.clone() for i in self.items])

    def add(self, item):
        if item is None:
            raise ValueError("Attempt to add nil item")
        self.items.append(item)

    @property
    def factor(self):
        return sum(abs(i.factor) for i in self.items)

    def rescaled(self, scale=1.0):
        divisor = self.factor / scale
        divisor = divisor if divisor != 0 else 1.0
        result = self.clone()
        if divisor != 1:
            for i in result.items:
                i.set_factor(i.factor / divisor)
        return result

    @torch.no_grad()
    def make_noise_sampler(self, x, sigma_min=None, sigma_max=None, seed=None, cpu=True, normalized=True) -> Callable:
        samplers = tuple(
            i.make_noise_sampler(x, sigma_min, sigma_max, seed=seed, cpu=cpu, normalized=False) for i in self.items
        )
        if not samplers or not all(samplers):
            raise ValueError("Failed to get noise sampler")
        factor = self.factor

        def accumulate(sigma, sigma_next):
            total = None
            for ns in samplers:
                part = ns(sigma, sigma_next)
                total = part if total is None else ops.axpby(total, 1.0, part, 1.0, out=total)
            return total

        def noise_sampler(sigma, sigma_next):
            return scale_noise(accumulate(sigma, sigma_next), factor, normalized=normalized)

        # A consumer that applies scale_noise itself while reading the tensor (the fused Sonar step
        # normalises on load) asks for the un-normalised sum and what is still owed to it.
        noise_sampler.deferred = lambda sigma, sigma_next: (accumulate(sigma, sigma_next), factor, normalized)
        # Look-ahead (see PowerNoiseItem.make_noise_sampler): a chain of ONE sigma-independent item hands its child's
        # batches through; `pending` = what this level still owes each sample, applied by whoever consumes it.
        if len(samplers) == 1 and factor == 1 and not normalized and hasattr(samplers[0], "fused_term"):
            noise_sampler.fusable, noise_sampler.fused_term = samplers[0].fusable, samplers[0].fused_term
        child_lookahead = getattr(samplers[0], "lookahead", None) if len(samplers) == 1 else None
        if child_lookahead is not None:
            noise_sampler.lookahead = child_lookahead
            noise_sampler.lookahead_pending = (factor, normalized)
        return noise_sampler


class NoiseSampler:
    """Wraps a generator factory (:199-257): sigma transform -> generator -> scale_noise -> cast."""

    def __init__(
        self,
        x,
        sigma_min=None,
        sigma_max=None,
        seed=None,
        cpu=False,
        transform: Callable = lambda t: t,
        normalized=False,
        factor: float = 1.0,
        *,
        make_noise_sampler: Callable,
        **kwargs,
    ):
        self.factor = factor
        self.normalized = normalized
        self.transform = transform
        self.device = x.device
        self.dtype = x.dtype
        self.noise_sampler = make_noise_sampler(
            x,
            sigma_min=transform(torch.as_tensor(sigma_min)) if sigma_min is not None else None,
            sigma_max=transform(torch.as_tensor(sigma_max)) if sigma_max is not None else None,
            seed=seed,
            cpu=cpu,
            normalized=False,  # the generator's own flag is forced off; this wrapper normalises (:230)
            **kwargs,
        )

    @classmethod
    def simple(cls, f):
        return lambda *args, **kwargs: cls(
            *args,
            **kwargs,
            make_noise_sampler=lambda x, *_a, **_k: lambda _s, _sn: f(x),
        )

    @classmethod
    def wrap(cls, f):
        return lambda *args, **kwargs: cls(*args, **kwargs, make_noise_sampler=f)

    def __call__(self, *args, **kwargs):
        args = tuple(self.transform(torch.as_tensor(s)) if s is not None else s for s in args)
        noise = self.noise_sampler(*args, **kwargs)
        noise = scale_noise(noise, self.factor, normalized=self.normalized)
        if hasattr(noise, "to"):
            noise = noise.to(dtype=self.dtype, device=self.device)
        return noise

    def fusable(self) -> bool:
        """This sampler is a bare generator sample (nothing owed at this level) that ops.noise_mix can regenerate
        from the Philox stream -- so a parent (BlendedNoise) may fuse it with its sibling instead of reading tensors."""
        gen = self.noise_sampler
        return (
            self.factor == 1
            and not self.normalized
            and self.dtype == torch.float32
            and hasattr(gen, "plan_term")
            and not gen.normalized
            and gen.normalize_dims is None
            and gen.fusable()
        )

    def fused_term(self):
        return self.noise_sampler.plan_term()

    def fused_gaussian(self):
        """If this sampler is plain Gaussian noise, returns (factor, normalized) so a consumer (the
        fused Sonar step) may regenerate the Philox draw in registers instead of reading a tensor;
        otherwise None."""
        gen = self.noise_sampler
        if (
            type(gen) is GaussianNoiseGenerator
            and not gen.normalized
            and gen.normalize_dims is None
            and gen.dtype == torch.float32
            and gen.device_generator() is None
        ):
            return (self.factor, bool(self.normalized), gen.shape)
        return None


class AdvancedNoiseBase(CustomNoiseItemBase):
    ns_factory_arg_keys = ()

    @property
    def ns_factory(self):
        raise NotImplementedError

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        if self.ns_factory is None:
            raise NotImplementedError("ns_factory not implemented")
        picked = {k: getattr(self, k) for k in self.ns_factory_arg_keys if getattr(self, k, None) is not None}
        self.sampler_factory = NoiseSampler.wrap(partial(self.ns_factory, **picked))

    @torch.no_grad()
    def make_noise_sampler(self, *args, **kwargs):
        return self.sampler_factory(*args, factor=self.factor, **kwargs)


class AdvancedPyramidNoise(AdvancedNoiseBase):
    ns_factory_arg_keys = ("discount", "iterations", "upscale_mode")
    pyramid_variants_map = {  # noqa: RUF012
        "pyramid": PyramidNoiseGenerator,
        "pyramid_old": PyramidOldNoiseGenerator,
        "highres_pyramid": HighresPyramidNoiseGenerator,
    }

    @property
    def ns_factory(self):
        return self.pyramid_variants_map[self.variant]


class Advanced1fNoise(AdvancedNoiseBase):
    ns_factory_arg_keys = ("alpha", "hfac", "wfac", "k", "use_sqrt", "base_power")

    @property
    def ns_factory(self):
        return OneFNoiseGenerator


class AdvancedPowerLawNoise(AdvancedNoiseBase):
    ns_factory_arg_keys = ("alpha", "div_max_dims", "use_sign")

    @property
    def ns_factory(self):
        return PowerLawNoiseGenerator


class _ChildHolder(CustomNoiseItemBase):
    """Items that own child chains clone them on copy."""

    child_keys: tuple = ()

    def clone_key(self, k):
        val = getattr(self, k)
        if k in self.child_keys and val is not None and hasattr(val, "clone"):
            return val.clone()
        return val


class CompositeNoise(_ChildHolder):
    """dst * (1 - mask) + src * mask (:470-533)."""

    child_keys = ("mask", "src_noise", "dst_noise")

    def __init__(self, factor, *, dst_noise, src_noise, normalize_dst, normalize_src, normalize_result, mask):
        super().__init__(
            factor,
            dst_noise=dst_noise.clone(),
            src_noise=src_noise.clone(),
            normalize_dst=normalize_dst,
            normalize_src=normalize_src,
            normalize_result=normalize_result,
            mask=mask.clone(),
        )

    def make_noise_sampler(self, x, *args, normalized=True, **kwargs):
        n_src, n_dst, n_result = (self.get_normalize(f"normalize_{k}", normalized) for k in ("src", "dst", "result"))
        nsd = self.dst_noise.make_noise_sampler(x, *args, normalized=n_dst, **kwargs)
        nss = self.src_noise.make_noise_sampler(x, *args, normalized=n_src, **kwargs)
        mask = self.mask.to(x.device, dtype=torch.float32, copy=True)
        mask = mask.reshape((-1, 1, *mask.shape[-2:])).contiguous()
        # F.interpolate(mode="bilinear") == our bilinear resample kernel (setup, not per step)
        mask = ops.resample(mask, x.shape[-2], x.shape[-1], mode="bilinear")
        batch = x.shape[0]
        if mask.shape[0] != batch:  # comfy.utils.repeat_to_batch_size
            if mask.shape[0] > batch:
                mask = mask[:batch]
            else:
                reps = math.ceil(batch / mask.shape[0])
                mask = mask.repeat(reps, 1, 1, 1)[:batch]
        mask = mask.contiguous()
        factor = self.factor

        def noise_sampler(s, sn):
            dst, src = nsd(s, sn), nss(s, sn)
            return scale_noise(ops.composite(dst, src, mask, out=dst), factor, normalized=n_result)

        return noise_sampler


class ScheduledNoise(_ChildHolder):
    """Child noise inside [end_sigma, start_sigma], fallback (or zeros) outside (:626-678)."""

    child_keys = ("noise", "fallback_noise")

    def __init__(self, factor, *, noise, start_sigma, end_sigma, normalize, fallback_noise=None):
        super().__init__(
            factor,
            noise=noise.clone(),
            start_sigma=start_sigma,
            end_sigma=end_sigma,
            normalize=normalize,
            fallback_noise=None if fallback_noise is None else fallback_noise.clone(),
        )

    def make_noise_sampler(self, x, *args, normalized=True, **kwargs):
        factor, start_sigma, end_sigma = self.factor, self.start_sigma, self.end_sigma
        normalize = self.get_normalize("normalize", normalized)
        ns = self.noise.make_noise_sampler(x, *args, normalized=False, **kwargs)
        if self.fallback_noise:
            nsa = self.fallback_noise.make_noise_sampler(x, *args, normalized=False, **kwargs)
        else:

            def nsa(_s, _sn):
                return torch.zeros_like(x)

        def noise_sampler(s, sn):
            if s is None or sn is None:
                raise ValueError("ScheduledNoise requires sigma, sigma_next to be passed")
            chosen = ns if end_sigma <= s <= start_sigma else nsa
            return scale_noise(chosen(s, sn), factor, normalized=normalize)

        return noise_sampler


class RepeatedNoise(_ChildHolder):
    """Caches up to `repeat_length` samples and replays them flipped / rolled / negated, the
    choice driven by a seeded CPU generator (:681-758). Index work only: bit-exact."""

    child_keys = ("noise",)

    def __init__(self, factor, *, noise, **kwargs):
        super().__init__(factor, noise=noise.clone(), **kwargs)

    def make_noise_sampler(self, x, *args, normalized=True, **kwargs):
        factor = self.factor
        repeat_length, max_recycle, permute = self.repeat_length, self.max_recycle, self.permute
        normalize = self.get_normalize("normalize", normalized)
        ns = self.noise.make_noise_sampler(x, *args, normalized=False, **kwargs)
        cache: list = []
        u32_max = 0xFFFF_FFFF
        seed = kwargs.get("seed")
        if seed is None:
            seed = torch.randint(-u32_max, u32_max, (1,), device="cpu", dtype=torch.int64).item()
        gen = torch.Generator(device="cpu")
        gen.manual_seed(seed)
        last_idx = -1

        def noise_sampler(s, sn):
            nonlocal last_idx
            rands = torch.randint(u32_max, (4,), generator=gen, dtype=torch.uint32).tolist()
            skip_permute = permute == "disabled"
            if len(cache) < repeat_length:
                idx = len(cache)
                noise = ns(s, sn)
                cache.append((1, noise))
                skip_permute = permute != "always"
            else:
                idx = rands[0] % repeat_length
                if idx == last_idx:
                    idx = (idx + 1) % repeat_length
                uses, noise = cache[idx]
                if uses >= max_recycle:
                    noise = ns(s, sn)
                    cache[idx] = (1, noise)
                    skip_permute = permute != "always"
                else:
                    cache[idx] = (uses + 1, noise)
            last_idx = idx
            if skip_permute:
                return noise.clone()
            ndim = noise.ndim
            if rands[1] % 2 == 0:
                if rands[2] <= u32_max // 5:
                    noise = noise.clone()
                    if rands[2] & 1 == 1:
                        noise *= -1.0
                else:
                    noise = torch.flip(noise, tuple({rands[2] % ndim, rands[3] % ndim}))
            else:
                dim = rands[2] % ndim
                noise = torch.roll(noise, rands[3] % noise.shape[dim], dims=(dim,)).clone()
            return scale_noise(noise, factor, normalized=normalize)

        return noise_sampler


class WaveletFilteredNoise(_ChildHolder):
    """SonarWaveletFilteredNoise: another chain's noise filtered in the wavelet domain (:1521-1590). Extra
    generator options (wave, level, mode, yl_scale, yh_scales, ...) arrive as `ns_kwargs` from the YAML."""

    child_keys = ("noise", "noise_high")

    def make_noise_sampler(self, x, sigma_min=None, sigma_max=None, *args, normalized=True, **kwargs):
        factor = self.factor
        normalize = self.get_normalize("normalize", normalized)
        internal_ns = internal_ns_high = None
        if getattr(self, "noise", None) is not None:
            internal_ns = self.noise.make_noise_sampler(
                x, *args, sigma_min=sigma_min, sigma_max=sigma_max, normalized=self.normalize_noise, **kwargs,
            )
        if getattr(self, "noise_high", None) is not None:
            internal_ns_high = self.noise_high.make_noise_sampler(
                x, *args, sigma_min=sigma_min, sigma_max=sigma_max, normalized=self.normalize_noise, **kwargs,
            )
        ns_kwargs = dict(getattr(self, "ns_kwargs", {}))
        yl_blend_function = ns_kwargs.pop("yl_blend_function", torch.lerp)
        yh_blend_function = ns_kwargs.pop("yh_blend_function", torch.lerp)
        if isinstance(yl_blend_function, str):
            yl_blend_function = hostutil.BLENDING_MODES[yl_blend_function]
        if isinstance(yh_blend_function, str):
            yh_blend_function = hostutil.BLENDING_MODES[yh_blend_function]
        kwargs |= ns_kwargs
        ns = WaveletFilteredNoiseGenerator(  # noqa: F405
            x, *args, sigma_min=sigma_min, sigma_max=sigma_max, normalized=False, noise_sampler=internal_ns,
            noise_sampler_high=internal_ns_high, yl_blend_function=yl_blend_function, yh_blend_function=yh_blend_function,
            **kwargs,
        )  # fmt: skip

        def noise_sampler(sigma, sigma_next):
            return scale_noise(ns(sigma, sigma_next), factor, normalized=normalize)

        return noise_sampler


class BlendedNoise(_ChildHolder):
    """blend_function(noise_1, noise_2, t); t a constant or a per-element mask noise normalised to
    [0,1] (:1302-1407)."""

    child_keys = ("custom_noise_1", "custom_noise_2", "custom_noise_mask")

    def __init__(
        self,
        factor,
        *,
        normalize,
        blend_function,
        custom_noise_1=None,
        custom_noise_2=None,
        custom_noise_mask=None,
        noise_2_percent=0.5,
    ):
        if custom_noise_1 is None and (custom_noise_mask is not None or noise_2_percent != 1):
            raise ValueError("When custom_noise_1 is not attached noise_2_percent must be set to 1")
        if custom_noise_2 is None and (custom_noise_mask is not None or noise_2_percent != 0):
            raise ValueError("When custom_noise_2 is not attached noise_2_percent must be set to 0")
        if custom_noise_mask is None and noise_2_percent == 1 and custom_noise_1 is None:
            custom_noise_1, custom_noise_2 = custom_noise_2, None
            noise_2_percent = 0.0
        super().__init__(
            factor,
            noise_2_percent=noise_2_percent,
            blend_function=blend_function,
            custom_noise_1=custom_noise_1.clone(),
            custom_noise_2=None if custom_noise_2 is None else custom_noise_2.clone(),
            custom_noise_mask=None if custom_noise_mask is None else custom_noise_mask.clone(),
            normalize=normalize,
        )

    def make_noise_sampler(self, x, *args, normalized=True, **kwargs):
        factor = self.factor
        normalize = self.get_normalize("normalize", normalized)
        blend_function, n2_blend = self.blend_function, self.noise_2_percent
        if isinstance(blend_function, str):
            blend_function = hostutil.BLENDING_MODES[blend_function]

        def child(item):
            return None if item is None else item.make_noise_sampler(x, *args, normalized=False, **kwargs)

        ns_1, ns_2, ns_mask = child(self.custom_noise_1), child(self.custom_noise_2), child(self.custom_noise_mask)

        blend_name = getattr(blend_function, "sonar_blend_mode", None)
        fuse = (
            ns_mask is None
            and blend_name in ops.BLEND_IDS
            and isinstance(n2_blend, (int, float))
            and hasattr(ns_1, "fused_term")
            and hasattr(ns_2, "fused_term")
        )

        def noise_sampler(s, sn):
            if fuse and ns_1.fusable() and ns_2.fusable():
                # both children are generator samples the mix kernel regenerates from the Philox stream: one launch
                # writes blend(noise_1, noise_2, t) (draws reserved child 1 first, like the two calls below)
                term_1, shape_1, begin = ns_1.fused_term()
                term_2, shape_2, _ = ns_2.fused_term()
                if shape_1 == shape_2:
                    mixed = ops.noise_mix(shape_1, term_1, term_2, begin=begin, blend_mode=blend_name, blend_t=n2_blend, device=x.device)
                    return scale_noise(ops.reshape_keep_sums(mixed, x.shape), factor, normalized=normalize)
                raise RuntimeError("fused noise terms disagree on the latent shape")
            noise_1 = ns_1(s, sn)
            if ns_2 is None:
                return scale_noise(noise_1, factor, normalized=normalize)
            noise_2 = ns_2(s, sn)
            if ns_mask is None:
                weight = n2_blend
            else:
                weight = hostutil.normalize_to_scale(ns_mask(s, sn), 0.0, 1.0)
                # (mask + n2_blend).clamp_(0, 1), each step rounded
                weight = ops.affine(weight, n2_blend, 1.0, 0.0).clamp_(0.0, 1.0)
            return scale_noise(blend_function(noise_1, noise_2, weight), factor, normalized=normalize)

        return noise_sampler


class CustomNoiseParametersNoise(_ChildHolder):
    """Shape / dtype / seed adapter around a child chain (:2080-2187). On this path it matters for
    5-D video latents: frames_to_channels folds (B,C,F,H,W) into (B,C*F,H,W) planes."""

    child_keys = ("noise",)

    def make_noise_sampler(self, x, sigma_min=None, sigma_max=None, *args, normalized=True, **kwargs):
        factor = self.factor
        normalize = self.get_normalize("normalize", normalized)
        orig_shape, orig_dtype, orig_device = x.shape, x.dtype, x.device
        override_device = getattr(self, "override_device", None)
        if override_device is not None:
            if torch.device(override_device).type != "cuda":
                raise NotImplementedError("sonar_b200: override_device must be a CUDA device (no CPU generation path)")
            x = x.to(device=override_device)
        if x.ndim == 5 and getattr(self, "frames_to_channels", False):
            x = x.reshape(x.shape[0], x.shape[1] * x.shape[2], *x.shape[3:])
        fix_invalid = getattr(self, "fix_invalid", False)
        override_dtype = getattr(self, "override_dtype", None)
        if override_dtype and x.dtype != override_dtype:
            x = x.to(dtype=override_dtype)
        fixed_aspect = False
        spatdims, height, width = 2, *x.shape[-2:]
        if getattr(self, "ensure_square_aspect_ratio", False):
            if x.ndim == 3:
                height, width, spatdims = 1, x.shape[-1], 1
            side = (height * width) ** 0.5
            if not side.is_integer():
                fixed_aspect = True
                side = math.ceil(side)
                padded = x.new_zeros(*x.shape[:-spatdims], side**2)
                padded[..., : height * width] = x.flatten(start_dim=-spatdims)[..., : height * width]
                x = padded.reshape(*padded.shape[:-1], side, side)
        rng_offset_mode = getattr(self, "rng_offset_mode", "disabled")
        if rng_offset_mode in {"override", "add"}:
            offset = getattr(self, "rng_state_offset", 0)
            kwargs["seed"] = offset if rng_offset_mode == "override" else kwargs.pop("seed", 0) + offset
        rng_mode = getattr(self, "rng_mode", "default")
        if rng_mode != "default":
            raise NotImplementedError("sonar_b200: rng_mode other than 'default' is host RNG-state glue (out of scope)")
        ns = self.noise.make_noise_sampler(x, *args, sigma_min=sigma_min, sigma_max=sigma_max, normalized=False, **kwargs)

        def noise_sampler(sigma, sigma_next) -> torch.Tensor:
            noise = ns(sigma, sigma_next)
            if fix_invalid:
                finite = noise.nan_to_num(0, posinf=0, neginf=0)
                noise = noise.nan_to_num_(0, posinf=finite.max(), neginf=finite.min())
            if fixed_aspect:
                noise = noise.flatten(start_dim=-spatdims)[..., : height * width]
            if noise.shape != orig_shape:
                noise = ops.reshape_keep_sums(noise, orig_shape)
            if noise.dtype != orig_dtype or noise.device != orig_device:
                noise = noise.to(device=orig_device, dtype=orig_dtype)
            return scale_noise(noise, factor, normalized=normalize)

        child_lookahead = getattr(ns, "lookahead", None)
        child_pending = getattr(ns, "lookahead_pending", (1.0, False))
        if (
            child_lookahead is not None
            and child_pending == (1.0, False)
            and factor == 1
            and not normalize
            and not fix_invalid
            and not fixed_aspect
            and x.dtype == orig_dtype
            and x.device == orig_device
        ):
            # a pure reshape around the child: its look-ahead batches pass through as views of the original shape
            def lookahead(count: int, **kw):
                batch = child_lookahead(count, **kw)
                if batch is None:
                    return None
                return type(batch)(((raw.reshape(orig_shape), sums, draw) for raw, sums, draw in batch), batch.table)

            noise_sampler.lookahead = lookahead
        return noise_sampler


class GuidedNoise(_ChildHolder):
    """A child chain's noise (or zeros) pulled towards a reference latent with the samplers' guidance functions
    (reference py/noise.py:536-623): guidance_linear / guidance_euler on the RAW reference
    latent, shifted to the per-item mean / std of the noise ("linear") or of the latent the sampler was built for
    ("euler") when a child chain is attached. Two launches per sample on `csrc/guidance.cu`."""

    child_keys = ("noise", "ref_latent")

    def __init__(self, factor, *, guidance_factor, ref_latent, method, normalize_noise, normalize_result, noise=None):
        super().__init__(
            factor,
            normalize_noise=normalize_noise,
            normalize_result=normalize_result,
            ref_latent=ref_latent.clone(),
            noise=noise.clone() if noise is not None else None,
            method=method,
            guidance_factor=guidance_factor,
        )

    def make_noise_sampler(self, x, *args, normalized=True, **kwargs):
        from .samplers import SonarGuidanceMixin  # (samplers imports this module)

        factor, guidance_factor = self.factor, self.guidance_factor
        normalize_noise, normalize_result = (self.get_normalize(f"normalize_{k}", normalized) for k in ("noise", "result"))
        ns = None if self.noise is None else self.noise.make_noise_sampler(x, *args, normalized=normalize_noise, **kwargs)
        ref_latent = self.ref_latent.to(x, copy=True)
        if ref_latent.shape[-2:] != x.shape[-2:]:  # setup, once per sampler
            ref_latent = torch.nn.functional.interpolate(ref_latent, size=x.shape[-2:], mode="bicubic", align_corners=True)
        ref_latent = ref_latent.to(torch.float32).contiguous()
        latent = x.to(torch.float32).contiguous()
        do_shift = ns is not None

        def base(s, sn):
            return torch.zeros_like(latent) if ns is None else ns(s, sn).to(torch.float32).contiguous()

        if self.method == "linear":

            def noise_sampler(s, sn):
                guided = SonarGuidanceMixin.guidance_linear(base(s, sn), ref_latent, guidance_factor, do_shift=do_shift)
                return scale_noise(guided, factor, normalized=normalize_result)

        elif self.method == "euler":

            def noise_sampler(s, sn):
                guided = SonarGuidanceMixin.guidance_euler(s, sn, base(s, sn), latent, ref_latent, guidance_factor, do_shift=do_shift)
                return scale_noise(guided, factor, normalized=normalize_result)

        else:
            raise ValueError("Bad method")
        return noise_sampler


def _out_of_scope_item(name: str) -> type:
    def make_noise_sampler(self, *args, **kwargs):
        raise NotImplementedError(
            f"sonar_b200: {name} is host-side graph glue / a niche filter outside the B200 hot path "
            "(SURVEY.md section 2); use the reference implementation for it",
        )

    return type(name, (_ChildHolder,), {"make_noise_sampler": make_noise_sampler})


for _name in (
    "ModulatedNoise", "RandomNoise", "ChannelNoise", "RippleFilteredNoise", "NormalizeToScaleNoise",
    "ResizedNoise", "ScatternetFilteredNoise", "LatentOperationFilteredNoise",
    "BlendFilterNoise", "QuantileFilteredNoise", "PerDimNoise", "ShuffledNoise", "PatternBreakNoise", "BlehOpsNoise",
    "AdvancedDistroNoise", "AdvancedCollatzNoise", "AdvancedWaveletNoise", "AdvancedVoronoiNoise",
):  # fmt: skip
    globals()[_name] = _out_of_scope_item(_name)
del _name


def _mixed(name: str, mix: tuple, output_scale: float | None = None) -> Callable:
    return NoiseSampler.wrap(partial(MixedNoiseGenerator, name=name, noise_mix=mix, output_scale=output_scale))


def _todo(noise_type: NoiseType, why: str) -> Callable:
    def factory(*_args, **_kwargs):
        raise NotImplementedError(f"sonar_b200: noise type {noise_type.name.lower()} is out of scope ({why})")

    return factory


# Same keys / parameterisation as the reference table (py/noise.py:2244-2457).
NOISE_SAMPLERS: dict[NoiseType, Callable] = {
    NoiseType.GAUSSIAN: NoiseSampler.wrap(GaussianNoiseGenerator),
    NoiseType.UNIFORM: NoiseSampler.wrap(UniformNoiseGenerator),
    NoiseType.PERLIN: NoiseSampler.wrap(PerlinOldNoiseGenerator),
    NoiseType.ONEF_PINKISH: NoiseSampler.wrap(partial(OneFNoiseGenerator, alpha=-0.5)),
    NoiseType.ONEF_GREENISH: NoiseSampler.wrap(partial(OneFNoiseGenerator, alpha=0.5)),
    NoiseType.ONEF_PINKISHGREENISH: _mixed(
        "onef_pinkishgreenish",
        ((OneFNoiseGenerator, {"alpha": 0.5}, None), (OneFNoiseGenerator, {"alpha": -0.5}, None)),
        0.5,
    ),
    NoiseType.ONEF_PINKISH_MIX: _mixed(
        "onef_pinkish_mix",
        ((OneFNoiseGenerator, {"alpha": -0.5}, -1.0), (OneFNoiseGenerator, {"alpha": -0.5}, None)),
        0.5,
    ),
    NoiseType.ONEF_GREENISH_MIX: _mixed(
        "onef_greenish_mix",
        ((OneFNoiseGenerator, {"alpha": 0.5}, -1.0), (OneFNoiseGenerator, {"alpha": 0.5}, None)),
        0.5,
    ),
    NoiseType.WHITE: NoiseSampler.wrap(partial(PowerLawNoiseGenerator, alpha=0.0, use_sign=True)),
    NoiseType.GREY: NoiseSampler.wrap(partial(PowerLawNoiseGenerator, alpha=0.0, use_sign=False)),
    NoiseType.VELVET: NoiseSampler.wrap(
        partial(PowerLawNoiseGenerator, alpha=1.0, use_sign=True, div_max_dims=(-3, -2, -1)),
    ),
    NoiseType.VIOLET: NoiseSampler.wrap(
        partial(PowerLawNoiseGenerator, alpha=0.5, use_sign=True, div_max_dims=(-3, -2, -1)),
    ),
    NoiseType.HIGHRES_PYRAMID: NoiseSampler.wrap(HighresPyramidNoiseGenerator),
    NoiseType.PYRAMID: NoiseSampler.wrap(PyramidNoiseGenerator),
    NoiseType.RAINBOW_MILD: _mixed(
        "rainbow_mild",
        ((GreenTestNoiseGenerator, {}, 0.55), (GreenTestNoiseGenerator, {}, 0.7)),
        1.15,
    ),
    NoiseType.RAINBOW_INTENSE: _mixed(
        "rainbow_intense",
        ((GreenTestNoiseGenerator, {}, 0.75), (GreenTestNoiseGenerator, {}, 0.5)),
        1.15,
    ),
    NoiseType.GREEN_TEST: NoiseSampler.wrap(GreenTestNoiseGenerator),
    NoiseType.PYRAMID_OLD: NoiseSampler.wrap(PyramidOldNoiseGenerator),
    NoiseType.PYRAMID_AREA: NoiseSampler.wrap(partial(PyramidNoiseGenerator, upscale_mode="area")),
    NoiseType.HIGHRES_PYRAMID_AREA: NoiseSampler.wrap(partial(HighresPyramidNoiseGenerator, upscale_mode="area")),
    NoiseType.PYRAMID_OLD_AREA: NoiseSampler.wrap(partial(PyramidOldNoiseGenerator, upscale_mode="area")),
    NoiseType.PYRAMID_DISCOUNT5: NoiseSampler.wrap(partial(PyramidNoiseGenerator, discount=0.5)),
    NoiseType.PYRAMID_MIX: _mixed(
        "pyramid_mix",
        ((PyramidNoiseGenerator, {"discount": 0.6}, 0.2), (PyramidNoiseGenerator, {"discount": 0.6}, -0.8)),
    ),
    NoiseType.PYRAMID_MIX_AREA: _mixed(
        "pyramid_mix_area",
        (
            (PyramidNoiseGenerator, {"discount": 0.5, "upscale_mode": "area"}, 0.2),
            (PyramidNoiseGenerator, {"discount": 0.5, "upscale_mode": "area"}, -0.8),
        ),
    ),
    # out of scope
    NoiseType.BROWNIAN: _todo(NoiseType.BROWNIAN, "needs torchsde"),
    NoiseType.DISTRO: _todo(NoiseType.DISTRO, "torch.distributions zoo"),
    NoiseType.STUDENTT: _todo(NoiseType.STUDENTT, "torch.distributions sampler"),
    NoiseType.LAPLACIAN: _todo(NoiseType.LAPLACIAN, "torch.distributions sampler"),
    NoiseType.WAVELET: NoiseSampler.wrap(WaveletNoiseGenerator),
    NoiseType.PINK_OLD: _todo(NoiseType.PINK_OLD, "documented as wrong upstream"),
    NoiseType.POWER_OLD: _todo(NoiseType.POWER_OLD, "documented as wrong upstream"),
    NoiseType.PYRAMID_BISLERP: _todo(NoiseType.PYRAMID_BISLERP, "bislerp lives in ComfyUI"),
    NoiseType.HIGHRES_PYRAMID_BISLERP: _todo(NoiseType.HIGHRES_PYRAMID_BISLERP, "bislerp lives in ComfyUI"),
    NoiseType.PYRAMID_OLD_BISLERP: _todo(NoiseType.PYRAMID_OLD_BISLERP, "bislerp lives in ComfyUI"),
    NoiseType.PYRAMID_MIX_BISLERP: _todo(NoiseType.PYRAMID_MIX_BISLERP, "bislerp lives in ComfyUI"),
    NoiseType.COLLATZ: _todo(NoiseType.COLLATZ, "not on the configured hot path"),
    NoiseType.VORONOI_FUZZ: _todo(NoiseType.VORONOI_FUZZ, "not on the configured hot path"),
    NoiseType.VORONOI_MIX: _todo(NoiseType.VORONOI_MIX, "not on the configured hot path"),
}


def get_noise_sampler(
    noise_type,
    x: torch.Tensor,
    sigma_min,
    sigma_max,
    seed: int | None = None,
    cpu: bool = True,
    factor: float = 1.0,
    normalized=False,
    **kwargs,
) -> Callable:
    """Reference py/noise.py:2460-2489."""
    if noise_type is None:
        noise_type = NoiseType.GAUSSIAN
    elif isinstance(noise_type, str):
        noise_type = NoiseType[noise_type.upper()]
    if noise_type == NoiseType.BROWNIAN and (sigma_min is None or sigma_max is None):
        raise ValueError("Must pass sigma min/max when using brownian noise")
    factory = NOISE_SAMPLERS.get(noise_type)
    if factory is None:
        raise ValueError("Unknown noise sampler")
    return factory(x, sigma_min, sigma_max, seed=seed, cpu=cpu, factor=factor, normalized=normalized, **kwargs)


CustomNoise = CustomNoiseChain  # annotation alias used by the sampler config (py/sonar.py:56)
_ = (random, parallel)  # imported for parity with the reference module surface
