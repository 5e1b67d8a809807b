"""Noise generators: the host-side mirror of the reference's in-scope generator classes
(py/noise_generation.py), with every tensor pass replaced by a CUDA kernel launch (`ops`).

Same class names, `ng_params`, call protocol (`gen(sigma, sigma_next) -> Tensor`) and error behaviour
as the reference so the noise graph above it is unchanged. Differences that are inherent to a
device-only path:

* every base draw comes from the torch CUDA Philox stream on `x.device` (`rng.normal/uniform`); the
  `cpu` parameter is accepted for signature compatibility and ignored (there is no CPU path), and a
  CPU `generator` only feeds the *host* draws the reference also makes on the CPU (pyramid level
  sizes, py/noise_generation.py:627-630);
* out-of-scope generator types raise NotImplementedError naming the type.
"""

from __future__ import annotations

import math
from enum import Enum, auto
from typing import Callable, NamedTuple

import torch

import os

from . import hostutil, ops, parallel, rng
from .hostutil import fallback, scale_noise

# Pyramid / Perlin samples (and their blends) are computed element-wise from the Philox stream (csrc/noise_mix.cu)
# instead of materialising every base draw; SONAR_B200_FUSED_NOISE=0 selects the materialising kernels.
FUSED_NOISE = os.environ.get("SONAR_B200_FUSED_NOISE", "1") != "0"
# A generator on its own gains nothing from the fused kernel (B200, 16x16x128x128: pyramid 81 us fused vs 75 us, Perlin
# 44 vs 40 us -- per-element bilinear taps cost more than the row-wise kernel saves in traffic); a blend of two
# generators does (one launch instead of three full-size passes, a quarter of the HBM traffic). Singles: opt-in.
FUSED_SINGLE_GENERATORS = os.environ.get("SONAR_B200_FUSED_SINGLE", "0") == "1"


class NoiseType(Enum):
    # Same members and order as the reference enum (py/noise_generation.py:31-69): the names are
    # the node dropdown values.
    BROWNIAN = auto()
    COLLATZ = auto()
    DISTRO = auto()
    GAUSSIAN = auto()
    GREEN_TEST = auto()
    GREY = auto()
    HIGHRES_PYRAMID = auto()
    HIGHRES_PYRAMID_AREA = auto()
    HIGHRES_PYRAMID_BISLERP = auto()
    LAPLACIAN = auto()
    ONEF_GREENISH = auto()
    ONEF_GREENISH_MIX = auto()
    ONEF_PINKISH = auto()
    ONEF_PINKISH_MIX = auto()
    ONEF_PINKISHGREENISH = auto()
    PERLIN = auto()
    PINK_OLD = auto()
    POWER_OLD = auto()
    PYRAMID = auto()
    PYRAMID_AREA = auto()
    PYRAMID_BISLERP = auto()
    PYRAMID_DISCOUNT5 = auto()
    PYRAMID_MIX = auto()
    PYRAMID_MIX_AREA = auto()
    PYRAMID_MIX_BISLERP = auto()
    PYRAMID_OLD = auto()
    PYRAMID_OLD_AREA = auto()
    PYRAMID_OLD_BISLERP = auto()
    RAINBOW_INTENSE = auto()
    RAINBOW_MILD = auto()
    STUDENTT = auto()
    UNIFORM = auto()
    VELVET = auto()
    VIOLET = auto()
    VORONOI_FUZZ = auto()
    VORONOI_MIX = auto()
    WAVELET = auto()
    WHITE = auto()

    @classmethod
    def get_names(cls, default=GAUSSIAN, skip=None):
        if default is not None:
            if isinstance(default, int):
                default = cls(default)
            yield default.name.lower()
        for nt in cls:
            if nt == default or (skip and nt in skip):
                continue
            yield nt.name.lower()


class NoiseError(Exception):
    pass


class NoiseGenerator:
    """Base class (reference py/noise_generation.py:87-179)."""

    name = "unknown"
    MIN_DIMS = 1
    MAX_DIMS = 0

    def __init__(self, x: torch.Tensor, **kwargs):
        if x.ndim < self.MIN_DIMS:
            raise ValueError(
                f"Noise generator {self.name} requires at least {self.MIN_DIMS} dimension(s) but got input with shape {x.shape}",
            )
        if self.MAX_DIMS > 0 and x.ndim > self.MAX_DIMS:
            raise ValueError(
                f"Noise generator {self.name} requires at most {self.MAX_DIMS} dimension(s) but got input with shape {x.shape}",
            )
        defaults = self.ng_params()
        merged = defaults | kwargs
        for key in defaults:
            setattr(self, key, merged.pop(key))
        self.options = merged
        self.update_x(x)

    @classmethod
    def ng_params(cls) -> dict:
        return {
            "normalized": True,
            "force_normalize": None,
            "normalize_dims": None,
            "cpu": True,
            "generator": None,
        }

    def update_x(self, x: torch.Tensor) -> None:
        if not x.is_cuda:
            raise RuntimeError(
                f"sonar_b200 noise generator {self.name}: latent must live on a CUDA device (got {x.device}); "
                "there is no CPU generation path",
            )
        self.shape = x.shape
        if x.ndim in {4, 5}:
            self.batch, self.channels = x.shape[:2]
            self.height, self.width = x.shape[-2:]
            self.frames = x.shape[-3] if x.ndim == 5 else None
        else:
            self.batch = self.channels = self.frames = self.height = self.width = None
        self.device = x.device
        self.gen_device = x.device
        self.layout = x.layout
        self.dtype = x.dtype

    def device_generator(self):
        """A torch.Generator is only usable for device draws if it is a CUDA generator."""
        gen = self.generator
        return gen if gen is not None and gen.device.type == "cuda" else None

    def rand_like(self, *, fun="normal", shape=None, dtype=None, **_ignored) -> torch.Tensor:
        """Base draw of `shape` (default: the latent shape) on the latent's device.

        `fun` is "normal" / "uniform"; torch.randn / torch.rand are accepted for source
        compatibility with callers written against the reference signature (:133-155)."""
        if fun is torch.randn:
            fun = "normal"
        elif fun is torch.rand:
            fun = "uniform"
        draw = rng.normal if fun == "normal" else rng.uniform
        return draw(
            tuple(fallback(shape, self.shape)),
            device=self.device,
            dtype=fallback(dtype, self.dtype),
            generator=self.device_generator(),
        )

    # ---- RNG-fused generation (csrc/noise_mix.cu) ----
    def fusable(self) -> bool:
        """True when this generator's sample can be computed element-wise from the Philox stream by ops.noise_mix:
        plain float32 draws from the default CUDA generator (no injected / batched draws in flight)."""
        if not FUSED_NOISE or not hasattr(self, "plan_term") or rng._INJECT is not None or rng._PENDING is not None:  # noqa: SLF001
            return False
        if self.dtype != torch.float32 or self.device_generator() is not None or self.width is None or self.width < 2:
            return False
        _, c, h, w = shape = self.get_adjusted_shape()
        total, _ = parallel.global_draw_geometry(shape)
        # index ranges of the fused kernel: 32-bit elements, 24-bit planes and in-plane offsets, 16-bit W and C
        if not (0 < total <= ops.MIX_MAX_ELEMENTS and h * w <= 1 << 24 and total // (h * w) < 1 << 24 and w < 65536 and c < 65536):
            return False
        return self.term_supported()

    def term_supported(self) -> bool:
        return True

    def output_hook(self, noise: torch.Tensor) -> torch.Tensor:
        return scale_noise(
            noise,
            normalized=self.normalized and (self.force_normalize is None or self.force_normalize is True),
            normalize_dims=self.normalize_dims,
        )

    def pre_hook(self) -> None:
        pass

    def generate(self, *args):
        raise NotImplementedError

    def __call__(self, *args, **kwargs) -> torch.Tensor:
        self.pre_hook()
        return self.output_hook(self.generate(*args, **kwargs))

    def __str__(self) -> str:
        pretty = ", ".join(f"{k}={getattr(self, k)!s}" for k in self.ng_params())
        return f"<NoiseGenerator({self.name}): device={self.device}, shape={self.shape}, dtype={self.dtype}, {pretty}>"


class FramesToChannelsNoiseGenerator(NoiseGenerator):
    """5-D (B,C,F,H,W) latents are generated as (B, C*F, H, W) planes (:182-209). Index-only."""

    MIN_DIMS = 4
    MAX_DIMS = 5

    def get_adjusted_shape(self):
        if self.frames:
            return (self.batch, self.channels * self.frames, self.height, self.width)
        return (self.batch, self.channels, self.height, self.width)

    def fix_output_frames(self, noise: torch.Tensor) -> torch.Tensor:
        if not self.frames:
            return noise
        return ops.reshape_keep_sums(noise, (self.batch, self.channels, self.frames, self.height, self.width))

    def rand_like(self, *args, shape=None, **kwargs) -> torch.Tensor:
        noise = super().rand_like(*args, shape=shape, **kwargs)
        if shape is not None:
            return noise
        adjusted = self.get_adjusted_shape()
        return noise.reshape(*adjusted) if noise.shape != adjusted else noise


class GaussianNoiseGenerator(NoiseGenerator):
    name = "gaussian"

    @classmethod
    def ng_params(cls):
        return super().ng_params() | {"normalized": False}

    def generate(self, *_args):
        return self.rand_like()


class UniformNoiseGenerator(NoiseGenerator):
    name = "uniform"

    @classmethod
    def ng_params(cls):
        return super().ng_params() | {"normalized": False, "sub_fac": 0.5, "mul_fac": 3.46, "mean_fac": 0.0}

    def generate(self, *_args):
        # rand.sub_(sub).mul_(mul).add_(mean): one kernel, each step rounded like the op chain
        return ops.affine(self.rand_like(fun="uniform"), -self.sub_fac, self.mul_fac, self.mean_fac)


class PerlinOldNoiseGenerator(FramesToChannelsNoiseGenerator):
    """Perlin with one grid cell per pixel (:289-493): the whole block-position machinery of the
    reference collapses to a 2x2 gradient stencil, shared by every batch item."""

    name = "perlin_old"

    @classmethod
    def ng_params(cls):
        return super().ng_params() | {"div_fac": 2.0, "iterations": 2, "blend_mode": "lerp"}

    def term_supported(self) -> bool:
        return 0 < self.iterations <= ops.MIX_MAX_TABLES and self.blend_mode in ops.BLEND_IDS and self.get_adjusted_shape()[1] <= 65535

    def plan_term(self):
        """Reserves the draws of one sample in the reference's order (base uniform, then one angle grid per iteration)
        and builds the stencil tables; the full-size base draw stays in the Philox stream."""
        b, c, h, w = shape = self.get_adjusted_shape()
        total, begin = parallel.global_draw_geometry(shape)
        base = ops.reserve_draw(total, self.device)
        grids = [ops.reserve_draw(c * (h + 1) * (w + 1), self.device) for _ in range(self.iterations)]
        tables = ops.perlin_tables(grids, c, h, w, blend_mode=self.blend_mode, device=self.device)
        return {"kind": ops.TERM_PERLIN, "base": base, "tables": tables, "div_fac": self.div_fac}, shape, begin

    def generate(self, *_args):
        hostutil.blend_mode_id(self.blend_mode)  # validates like BLENDING_MODES[...] would
        if FUSED_SINGLE_GENERATORS and self.fusable():
            term, shape, begin = self.plan_term()
            return self.fix_output_frames(ops.noise_mix(shape, term, begin=begin, device=self.device))
        with rng.batched():  # base + angle grids: one launch
            base = self.rand_like(fun="uniform")
            b, c, h, w = base.shape
            angles = [
                rng.uniform(
                    (c, h + 1, w + 1),
                    device=self.device,
                    generator=self.device_generator(),
                    low=0.0,
                    high=2.0 * math.pi,
                    batch_sharded=False,
                )
                for _ in range(self.iterations)
            ]
        if base.dtype != torch.float32:
            base = base.float()
        noise = ops.perlin_accumulate(base, angles, shape=(b, c, h, w), div_fac=self.div_fac, blend_mode=self.blend_mode)
        return self.fix_output_frames(noise.to(self.dtype))


class _PyramidBase(FramesToChannelsNoiseGenerator):
    def _accumulate(self, base, levels, weights, orig_h, orig_w):
        mode = self.upscale_mode
        if mode in ops.RESAMPLE_IDS and (base is None or base.dtype == torch.float32):
            return ops.pyramid_accumulate(base, levels, weights, out_hw=(orig_h, orig_w), mode=mode)
        # modes without a fused kernel (bicubic, nearest): resize each level, accumulate with axpby
        noise = base
        for lv, wgt in zip(levels, weights):
            up = hostutil.scale_samples(lv, orig_w, orig_h, mode=mode).contiguous()
            noise = ops.axpby(up, wgt, None) if noise is None else ops.axpby(noise, 1.0, up, wgt, out=noise)
        return noise


class PyramidNoiseGenerator(_PyramidBase):
    name = "pyramid"

    @classmethod
    def ng_params(cls):
        return super().ng_params() | {"discount": 0.7, "upscale_mode": "bilinear", "iterations": 10}

    def term_supported(self) -> bool:
        return self.upscale_mode in ("bilinear", "nearest-exact") and self.iterations <= ops.PYRAMID_MAX_LEVELS

    def plan_term(self):
        """Reserves the draws of one sample in the reference's order (:621-649: base, then per level one CPU-generator
        draw for its size and the level itself). The base and the full-size level 0 stay in the Philox stream; the
        coarse levels (a few per cent of the elements) are materialised by one batched fill launch."""
        b, c, h, w = shape = self.get_adjusted_shape()
        orig_h, orig_w = h, w
        total, begin = parallel.global_draw_geometry(shape)
        base = ops.reserve_draw(total, self.device)
        host_gen = self.generator if self.generator is not None and self.generator.device.type == "cpu" else None
        levels, have_full = [], False
        with rng.batched():
            for i in range(self.iterations):
                r = rng.host_rand(1, host_gen).item() * 2 + 2
                w, h = max(1, int(w / (r**i))), max(1, int(h / (r**i)))
                if (h, w) == (orig_h, orig_w) and not have_full:
                    level, have_full = ops.reserve_draw(total, self.device), True
                else:
                    level = rng.normal((b, c, h, w), device=self.device, dtype=torch.float32)
                levels.append((level, self.discount**i))
                if w == 1 or h == 1:
                    break
        return {"kind": ops.TERM_PYRAMID, "base": base, "levels": levels, "mode": self.upscale_mode}, shape, begin

    def generate(self, *_args):
        if FUSED_SINGLE_GENERATORS and self.fusable():
            term, shape, begin = self.plan_term()
            return self.fix_output_frames(ops.noise_mix(shape, term, begin=begin, device=self.device))
        levels, weights = [], []
        with rng.batched():  # base + every level: one launch
            base = self.rand_like()
            b, c, h, w = base.shape
            orig_h, orig_w = h, w
            host_gen = self.generator if self.generator is not None and self.generator.device.type == "cpu" else None
            for i in range(self.iterations):
                # level size from one CPU-generator draw per level, exactly like the reference (:626-648):
                # r = U(0,1)*2+2, cumulative int division, stop at a 1-pixel side (level 0 is full size)
                r = rng.host_rand(1, host_gen).item() * 2 + 2
                w, h = max(1, int(w / (r**i))), max(1, int(h / (r**i)))
                levels.append(rng.normal((b, c, h, w), device=self.device, dtype=base.dtype, generator=self.device_generator()))
                weights.append(self.discount**i)
                if w == 1 or h == 1:
                    break
        return self.fix_output_frames(self._accumulate(base, levels, weights, orig_h, orig_w))


class HighresPyramidNoiseGenerator(_PyramidBase):
    name = "highres_pyramid"

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        if self.noise_generator is None:
            self.noise_generator = UniformNoiseGenerator(*args, **(kwargs | {"normalized": self.normalize_noise}))

    @classmethod
    def ng_params(cls):
        return super().ng_params() | {
            "normalized": True,
            "discount": 0.7,
            "upscale_mode": "bilinear",
            "iterations": 4,
            "noise_generator": None,
            "normalize_noise": False,
        }

    def generate(self, s, sn):
        b, c, h, w = adjusted = self.get_adjusted_shape()
        orig_h, orig_w = h, w
        base = self.noise_generator(s, sn).reshape(*adjusted)
        host_gen = self.generator if self.generator is not None and self.generator.device.type == "cpu" else None
        rs = rng.host_rand(self.iterations, host_gen) * 2 + 2
        levels, weights = [], []
        with rng.batched():
            for i in range(self.iterations):
                r = rs[i].item()
                h, w = min(orig_h * 15, int(h * (r**i))), min(orig_w * 15, int(w * (r**i)))
                levels.append(rng.normal((b, c, h, w), device=self.device, generator=self.device_generator()))
                weights.append(self.discount**i)
                if h >= orig_h * 15 or w >= orig_w * 15:
                    break
        return self.fix_output_frames(self._accumulate(base.contiguous(), levels, weights, orig_h, orig_w))


class PyramidOldNoiseGenerator(_PyramidBase):
    name = "pyramid_old"

    @classmethod
    def ng_params(cls):
        return super().ng_params() | {
            "discount": 0.8,
            "iterations": 5,
            "upscale_mode": "nearest-exact",
            "normalized": False,
        }

    def generate(self, *_args):
        b, c, h, w = self.get_adjusted_shape()
        levels, weights = [], []
        r = 1
        with rng.batched():
            for i in range(self.iterations):
                r *= 2
                levels.append(
                    rng.normal((b, c, h * r, w * r), device=self.device, generator=self.device_generator(), std=0.5**i),
                )
                weights.append(self.discount**i)
        # the reference accumulates into zeros: 0 + level*w is exact, so no base tensor is needed
        return self.fix_output_frames(self._accumulate(None, levels, weights, h, w).to(self.dtype))


class _SpectralGainGenerator(FramesToChannelsNoiseGenerator):
    """Shared tail of the fft -> real gain -> ifft(.real) generators (OneF, GreenTest)."""

    MIN_DIMS = 4
    MAX_DIMS = 5

    def full_gain(self) -> torch.Tensor:
        """(H, W) real gain on the host (float32), same op sequence as the reference."""
        raise NotImplementedError

    def half_gain(self) -> torch.Tensor:
        cache = getattr(self, "_gain_cache", None)
        if cache is not None:
            return cache
        g = self.full_gain()
        # real(ifft2(fft2(x) * G)) for real x equals filtering with the Hermitian-symmetrised gain
        g_neg = torch.roll(torch.flip(g, dims=(0, 1)), shifts=(1, 1), dims=(0, 1))
        g_sym = (g + g_neg) * 0.5
        half = g_sym[:, : self.width // 2 + 1].contiguous().to(self.device)
        self._gain_cache = half
        return half

    def filtered(self) -> torch.Tensor:
        noise = self.rand_like()
        if noise.dtype != torch.float32:
            noise = noise.float()
        return ops.spectral_filter(
            real=noise,
            mask=self.half_gain(),
            hw=(self.height, self.width),
            out_scale=1.0 / (self.height * self.width),
        )


class GreenTestNoiseGenerator(_SpectralGainGenerator):
    name = "green_test"

    @classmethod
    def ng_params(cls):
        return super().ng_params() | {"scale_fac": 1.0, "x_pow": 2, "y_pow": 2, "power_base": 1}

    def full_gain(self):
        fy = torch.fft.fftfreq(self.height)[:, None] ** self.y_pow
        fx = torch.fft.fftfreq(self.width) ** self.x_pow
        power = torch.sqrt(fy + fx)
        power[0, 0] = self.power_base
        return 1.0 / torch.sqrt(power)

    def generate(self, *_args):
        noise = self.filtered()
        # noise *= scale / noise.std()  (:703; the complex std equals the std of the real part up to
        # the ~1e-8 imaginary residue of a Hermitian-symmetric filter)
        sums = ops.moments(noise)
        count = parallel.global_count(noise.numel(), sums)
        ops.scale_by_std(noise, sums, count, self.scale_fac / (self.width * self.height))
        return self.fix_output_frames(noise.to(self.dtype))


class OneFNoiseGenerator(_SpectralGainGenerator):
    name = "onef"

    @classmethod
    def ng_params(cls):
        return super().ng_params() | {
            "alpha": 2.0,
            "k": 1.0,
            "hfac": 1.0,
            "wfac": 1.0,
            "base_power": 1.0,
            "use_sqrt": True,
        }

    def full_gain(self):
        freq_x = torch.fft.fftfreq(self.height, self.hfac)
        freq_y = torch.fft.fftfreq(self.width, self.wfac)
        fx, fy = torch.meshgrid(freq_x, freq_y, indexing="ij")
        power = (fx**2 + fy**2) ** (-self.alpha / 2.0)
        if self.k != 0:
            power = self.k / power
        power[0, 0] = self.base_power
        if self.use_sqrt and bool((power < 0).any()):
            raise NotImplementedError("onef noise with a negative power spectrum (complex gain) has no kernel")
        return 1.0 / (torch.sqrt(power) if self.use_sqrt else power)

    def generate(self, *_args):
        # fftn over (B,C,H,W) with a gain constant over B,C == per-plane 2-D transform (:750-759)
        return self.fix_output_frames(self.filtered().to(self.dtype))


class WaveletNoiseOctave(NamedTuple):
    octave: int
    height: float
    width: float
    amplitude: float
    total_amplitude: float


class WaveletNoiseGenerator(FramesToChannelsNoiseGenerator):
    """Octave ("wavelet") noise (:2204-2327): per octave, a draw minus its own low-pass (area pool down,
    bilinear up), resized to the latent and accumulated with a geometric amplitude. Built from the
    resampling, blend and axpby kernels; the draws of all octaves share one batched Philox launch."""

    name = "wavelet"
    MIN_DIMS = 4
    MAX_DIMS = 5

    @classmethod
    def ng_params(cls):
        return super().ng_params() | {
            "octave_scale_mode": "adaptive_avg_pool2d",
            "octave_rescale_mode": "bilinear",
            "post_octave_rescale_mode": "bilinear",
            "initial_amplitude": 1.0,
            "persistence": 0.5,
            "octaves": 4,
            "octave_height_factor": 0.5,
            "octave_width_factor": 0.5,
            "height_factor": 2.0,
            "width_factor": 2.0,
            "min_height": 4,
            "min_width": 4,
            "update_blend": 1.0,
            "update_blend_function": "lerp",
            "noise_sampler": None,
        }

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.set_octave_data()

    def set_internal_noise_sampler(self, noise_sampler) -> None:
        self.noise_sampler = noise_sampler

    def set_octave_data(self) -> None:
        """Octave sizes / amplitudes, host arithmetic identical to the reference (:2238-2278)."""
        height, width = self.get_adjusted_shape()[-2:]
        amplitude, total_amplitude = self.initial_amplitude, 0.0
        curr_height, curr_width = height, width
        octave_data = []
        is_reverse = self.octaves < 0
        octaves = range(self.octaves) if not is_reverse else reversed(range(abs(self.octaves)))
        for octave in octaves:
            curr_height /= self.height_factor**octave
            curr_width /= self.width_factor**octave
            if (
                amplitude == 0
                or curr_height < self.min_height
                or curr_width < self.min_width
                or curr_height * self.octave_height_factor < 1
                or curr_width * self.octave_width_factor < 1
            ):
                if is_reverse and not octave_data:
                    curr_height, curr_width = height, width
                    continue
                break
            total_amplitude += abs(amplitude)
            octave_data.append(WaveletNoiseOctave(octave, curr_height, curr_width, amplitude, total_amplitude))
            amplitude *= self.persistence
        if not octave_data or not total_amplitude:
            raise ValueError("Unworkable parameters for wavelet noise")
        self.octave_data = tuple(octave_data)

    def _octave_detail(self, noise: torch.Tensor) -> torch.Tensor:
        height, width = noise.shape[-2:]
        scaled_height = int(max(1, height * self.octave_height_factor))
        scaled_width = int(max(1, width * self.octave_width_factor))
        low = hostutil.scale_samples(noise, scaled_width, scaled_height, mode=self.octave_scale_mode)
        low = hostutil.scale_samples(low.contiguous(), width, height, mode=self.octave_rescale_mode).contiguous()
        detail = ops.axpby(noise, 1.0, low, -1.0)  # noise - scaled_noise
        blend = self.update_blend_function
        if callable(blend) and blend is not torch.lerp:
            return blend(noise, detail, self.update_blend)
        return ops.blend(noise, detail, self.update_blend, mode="lerp" if callable(blend) else blend)

    def generate(self, *args):
        adjusted = self.get_adjusted_shape()
        height, width = adjusted[-2:]
        shapes = [(*adjusted[:-2], int(od.height), int(od.width)) for od in self.octave_data]
        if self.noise_sampler:
            draws = [self.noise_sampler(*args)[..., : shp[-2], : shp[-1]].reshape(shp).contiguous() for shp in shapes]
        else:
            with rng.batched():  # every octave's draw in one launch, reserved in the reference's order
                draws = [self.rand_like(shape=shp) for shp in shapes]
        result = None
        for od, noise in zip(self.octave_data, draws):
            if noise.dtype != torch.float32:
                noise = noise.float()
            octave = self._octave_detail(noise)
            if tuple(octave.shape[-2:]) != (height, width):
                octave = hostutil.scale_samples(octave.contiguous(), width, height, mode=self.post_octave_rescale_mode).contiguous()
            # result += octave_output.mul_(amplitude): product rounded, then added
            result = ops.axpby(octave, od.amplitude, None) if result is None else ops.axpby(result, 1.0, octave, od.amplitude, out=result)
        total = self.octave_data[-1].total_amplitude
        if total != 0:
            result = ops.divide_scalar(result, total)
        return self.fix_output_frames(result.to(self.dtype))


class PowerLawNoiseGenerator(NoiseGenerator):
    name = "powerlaw"

    @classmethod
    def ng_params(cls):
        return super().ng_params() | {"alpha": 2.0, "div_max_dims": None, "use_sign": False, "use_div_max_abs": True}

    def generate(self, *_args):
        noise = self.rand_like()
        if noise.dtype != torch.float32:
            noise = noise.float()
        ops.powerlaw(noise, self.alpha, use_sign=self.use_sign, out=noise)
        if self.div_max_dims is not None:
            dims = tuple(sorted(d % noise.ndim for d in self.div_max_dims))
            if dims == tuple(range(1, noise.ndim)):
                ops.div_item_max(noise, use_abs=self.use_div_max_abs)
            else:
                noise /= torch.amax(
                    torch.abs(noise) if self.use_div_max_abs else noise,
                    keepdim=True,
                    dim=self.div_max_dims,
                )
        return noise.to(self.dtype)


class MixedNoiseGenerator(NoiseGenerator):
    """Weighted sum of child generators (:212-249). `noise_mix` entries are
    (generator class, kwargs, scale or None); `output_scale` multiplies the sum."""

    @classmethod
    def ng_params(cls):
        return super().ng_params() | {
            "name": "mixed_noise",
            "normalized": True,
            "pass_args": frozenset(("cpu",)),
            "noise_mix": (),
            "output_scale": None,
        }

    def __init__(self, x, *args, **kwargs):
        lo = hi = None
        self.name = kwargs["name"]
        for item in kwargs["noise_mix"]:
            ng_class = item[0] if isinstance(item, (tuple, list)) else item
            cmin, cmax = ng_class.MIN_DIMS, ng_class.MAX_DIMS
            lo = max(lo if lo is not None else cmin, cmin)
            hi = min(hi if hi is not None else cmax, cmax)
        self.MIN_DIMS, self.MAX_DIMS = lo, hi
        super().__init__(x, *args, **kwargs)
        # children keep their own class-default `normalized`; only pass_args are forwarded (:236-237)
        forwarded = {k: v for k, v in kwargs.items() if k in self.pass_args}
        self.ng_list = [(ng_class(x, **ng_kwargs, **forwarded), mult) for ng_class, ng_kwargs, mult in self.noise_mix]

    def generate(self, *args):
        total = None
        for ng, mult in self.ng_list:
            part = ng(*args)
            if part.dtype != torch.float32:
                part = part.float()
            if total is None:
                total = part if mult is None else ops.scale(part, mult)
            else:
                ops.axpby(total, 1.0, part, 1.0 if mult is None else mult, out=total)
        if self.output_scale is not None:
            ops.scale(total, self.output_scale)
        return total.to(self.dtype)


class WaveletFilteredNoiseGenerator(FramesToChannelsNoiseGenerator):
    """Noise filtered in the wavelet domain (:1908-2032): DWT of the source noise (optionally a second
    source for the detail bands, blended band-wise), per-band scaling, IDWT, crop. Defaults: haar, 3 levels,
    the non-expansive "periodization" mode. The final `wavelet_scaling` rides on the synthesis kernels' loads;
    `two_step_inverse` is accepted (the transform is linear: inverting the two halves separately and adding
    them is the one-pass inverse)."""

    name = "waveletfilter"
    MIN_DIMS = 4
    MAX_DIMS = 5

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        from .wavelets import Wavelet  # (wavelets imports ops only; late import keeps module init order simple)

        inv_kwargs = {k: self.options[k] for k in ("inv_mode", "inv_biort", "inv_qshift", "inv_wave") if k in self.options}
        self.wavelet = Wavelet(
            wave=self.wave, level=self.level, mode=self.mode, use_1d_dwt=self.use_1d_dwt, use_dtcwt=self.use_dtcwt,
            biort=self.biort, qshift=self.qshift, device=self.gen_device, dtype=torch.float32, **inv_kwargs,
        )  # fmt: skip

    @classmethod
    def ng_params(cls):
        return super().ng_params() | {
            "mode": "periodization",
            "level": 3,
            "wave": "haar",
            "use_1d_dwt": False,
            "use_dtcwt": False,
            "qshift": "qshift_a",
            "biort": "near_sym_a",
            "yl_scale": 1.0,
            "yh_scales": 1.0,
            "two_step_inverse": False,
            "preblend_yl_scale_low": None,
            "preblend_yh_scales_low": None,
            "preblend_yl_scale_high": None,
            "preblend_yh_scales_high": None,
            "yl_blend_function": torch.lerp,
            "yh_blend_function": torch.lerp,
            "yl_blend_high": 0.0,
            "yh_blend_high": 1.0,
            "noise_sampler": None,
            "noise_sampler_high": None,
        }

    def _fix_shape(self, noise, adjusted_shape):
        if noise.shape != adjusted_shape:
            noise = noise.reshape(*adjusted_shape)
        if self.frames:
            noise = noise.reshape(self.batch, self.channels * self.frames, self.height, self.width)
        return noise

    @staticmethod
    def _blend_kernel(fn) -> Callable:
        """torch.lerp (the reference default) and BLENDING_MODES names map onto the blend kernel."""
        if fn is torch.lerp:
            return hostutil.BLENDING_MODES["lerp"]
        if isinstance(fn, str):
            return hostutil.BLENDING_MODES[fn]
        return fn

    def generate(self, *args):
        from .wavelets import wavelet_scaling

        adjusted_shape = self.get_adjusted_shape()
        noise = self.rand_like() if self.noise_sampler is None else self.noise_sampler(*args)
        noise_high = None if self.noise_sampler_high is None else self._fix_shape(self.noise_sampler_high(*args), adjusted_shape)
        noise = self._fix_shape(noise, adjusted_shape)
        if noise.dtype != torch.float32:
            noise = noise.float()
        yl, yh = self.wavelet.forward(noise.contiguous())
        if noise_high is not None:
            yl_high, yh_high = self.wavelet.forward(noise_high.float().contiguous())
            if self.preblend_yl_scale_high is not None or self.preblend_yh_scales_high is not None:
                yl_high, yh_high = wavelet_scaling(
                    yl_high, yh_high, fallback(self.preblend_yl_scale_high, 1.0), fallback(self.preblend_yh_scales_high, 1.0),
                )
            if self.preblend_yl_scale_low is not None or self.preblend_yh_scales_low is not None:
                yl, yh = wavelet_scaling(
                    yl, yh, fallback(self.preblend_yl_scale_low, 1.0), fallback(self.preblend_yh_scales_low, 1.0),
                )
            yl_blend, yh_blend = self._blend_kernel(self.yl_blend_function), self._blend_kernel(self.yh_blend_function)
            yl = yl_blend(yl.contiguous(), yl_high.contiguous(), self.yl_blend_high)
            yh = tuple(yh_blend(a.contiguous(), b.contiguous(), self.yh_blend_high) for a, b in zip(yh, yh_high))
            del noise_high, yl_high, yh_high
        result = self.wavelet.inverse(
            yl, yh, two_step_inverse=self.two_step_inverse, yl_scale=self.yl_scale, yh_scales=self.yh_scales,
        )
        result = self.fix_output_frames(result)
        if result.shape != (target := self.fix_output_frames(noise).shape):
            result = result[tuple(slice(0, dl) for dl in target)]
        return result.contiguous().to(self.dtype)


def _unsupported(name: str, why: str) -> Callable:
    class _Unsupported(NoiseGenerator):
        def __init__(self, *_a, **_k):
            raise NotImplementedError(f"sonar_b200: noise type {name!r} is out of scope for the B200 hot path ({why})")

    _Unsupported.name = name
    _Unsupported.__name__ = f"Unsupported_{name}"
    return _Unsupported


BrownianNoiseGenerator = _unsupported("brownian", "needs torchsde's BrownianTree")
StudentTNoiseGenerator = _unsupported("studentt", "torch.distributions sampler")
LaplacianNoiseGenerator = _unsupported("laplacian", "torch.distributions sampler")
DistroNoiseGenerator = _unsupported("distro", "torch.distributions zoo")
PowerOldNoiseGenerator = _unsupported("power_old", "documented as wrong upstream")
PinkOldNoiseGenerator = _unsupported("pink_old", "documented as wrong upstream")
VoronoiNoiseGenerator = _unsupported("voronoi", "not on the configured hot path")
CollatzNoiseGenerator = _unsupported("collatz", "not on the configured hot path")
ScatternetFilteredNoiseGenerator = _unsupported("scatternet_filtered", "needs pytorch_wavelets ScatLayer")


__all__ = [
    "BrownianNoiseGenerator",
    "CollatzNoiseGenerator",
    "DistroNoiseGenerator",
    "FramesToChannelsNoiseGenerator",
    "GaussianNoiseGenerator",
    "GreenTestNoiseGenerator",
    "HighresPyramidNoiseGenerator",
    "LaplacianNoiseGenerator",
    "MixedNoiseGenerator",
    "NoiseError",
    "NoiseGenerator",
    "NoiseType",
    "OneFNoiseGenerator",
    "PerlinOldNoiseGenerator",
    "PinkOldNoiseGenerator",
    "PowerLawNoiseGenerator",
    "PowerOldNoiseGenerator",
    "PyramidNoiseGenerator",
    "PyramidOldNoiseGenerator",
    "ScatternetFilteredNoiseGenerator",
    "StudentTNoiseGenerator",
    "UniformNoiseGenerator",
    "VoronoiNoiseGenerator",
    "WaveletFilteredNoiseGenerator",
    "WaveletNoiseGenerator",
]
