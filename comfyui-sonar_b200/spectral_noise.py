"""Power-law / band-pass spectral noise ("Advanced Power Noise").

Host-side mirror of the reference's py/nodes/powernoise.py:56-554: `ChannelMixer`, `PowerFilter`,
`PowerNoiseItem`, `PowerFilterNoiseItem`. The filter construction is setup work on an (H, W/2+1)
host tensor and stays in (CPU, float32) torch exactly as upstream; it is cached per
(shape, parameters) because `CustomNOISE` rebuilds its sampler on every call.
The per-sample work -- complex Philox draw, gain, irfft2 [, rfft2 front end], normalisation -- runs
in the shared-memory FFT kernel (`ops.spectral_filter`) and the moments/scale kernels.
"""

from __future__ import annotations

import math

import torch

from . import ops, rng
from .hostutil import scale_noise
from .noise_graph import CustomNoiseItemBase


class ChannelMixer:
    """C x C channel-correlation mixer (:56-104). Identity (skipped) when common_mode == 0."""

    def __init__(self, channel_count, common_mode, channel_correlation):
        self.channel_count = channel_count
        self.common_mode = common_mode
        self.channel_correlation = channel_correlation
        self.mixer = self.build() if common_mode is not None else None
        self.is_identity = self.mixer is not None and bool(
            torch.equal(self.mixer, torch.eye(channel_count, dtype=self.mixer.dtype)),
        )

    def build(self) -> torch.Tensor:
        c, common = self.channel_count, self.common_mode
        pairs = c * (c - 1) // 2
        given = self.channel_correlation[:pairs]
        corr = torch.cat((given * common, torch.full((pairs - given.numel(),), common)))
        m = torch.eye(c).index_put_(tuple(torch.tril_indices(c, c, offset=-1)), corr)
        m += m.tril(-1).mT
        m = torch.linalg.ldl_factor(m).LD
        diag = torch.diagonal_copy(m)
        torch.diagonal(m)[:] = 1.0
        m *= diag.clamp_min(0).sqrt().unsqueeze(0)
        m /= m.norm(dim=1, keepdim=True)
        return m

    def to(self, *args, **kwargs):
        if self.mixer is not None:
            if self.mixer.device.type == "cpu":
                self.mixer_host = self.mixer.to(torch.float32).contiguous()  # small matrices travel by value
            self.mixer = self.mixer.to(*args, **kwargs)
        return self

    def apply(self, noise: torch.Tensor, shape, copy: bool = False) -> torch.Tensor:
        if self.mixer is None:
            return noise.clone() if copy else noise
        b, c, h, w = shape  # 5-D latents must be folded first, as in the reference (:97)
        if c != self.channel_count:
            raise ValueError("Channel count mismatch")
        if self.is_identity:
            return noise  # I @ noise is an exact copy
        # mixer @ noise.swapaxes(0, 1).reshape(c, -1), un-swapped: one kernel on the (B, C, H, W) layout
        mixer = self.mixer
        if mixer.dtype != torch.float32 or not mixer.is_contiguous():
            mixer = self.mixer = mixer.to(torch.float32).contiguous()
        return ops.channel_mix(noise.reshape(b, c, h, w).contiguous(), mixer, getattr(self, "mixer_host", None))

    __call__ = apply


_FILTER_KEYS = (
    "min_freq", "max_freq", "stretch", "rotate", "pnorm", "alpha", "scale", "rel_bw", "oversample", "compose_mode",
)  # fmt: skip


class PowerFilter:
    """Band-pass x 1/f^alpha gain in rfft2 layout (:107-266)."""

    def __init__(
        self,
        *,
        min_freq=0.0,
        max_freq=0.7071,
        stretch=1.0,
        rotate=0.0,
        pnorm=2.0,
        alpha=0.0,
        scale=1.0,
        rel_bw=0.125,
        oversample=4,
        compose_with: "PowerFilter | None" = None,
        compose_mode="max",
    ):
        self.min_freq = min_freq
        self.max_freq = max(max_freq, min_freq)
        self.stretch = stretch
        self.rotate = rotate
        self.pnorm = pnorm
        self.alpha = alpha
        self.scale = scale
        self.rel_bw = rel_bw
        self.oversample = oversample
        self.compose_with = compose_with
        self.compose_mode = compose_mode

    def clone(self):
        kwargs = {k: getattr(self, k) for k in _FILTER_KEYS}
        kwargs["compose_with"] = None if self.compose_with is None else self.compose_with.clone()
        return self.__class__(**kwargs)

    def cache_key(self) -> tuple:
        own = tuple(getattr(self, k) for k in _FILTER_KEYS)
        return (*own, None if self.compose_with is None else self.compose_with.cache_key())

    @classmethod
    def compose(cls, a, b, compose_mode="max"):
        if a.shape != b.shape:
            raise ValueError("Filter compose size mismatch!")
        fn = {"max": torch.max, "min": torch.min, "add": torch.add, "sub": torch.sub, "mul": torch.mul}.get(
            compose_mode,
            torch.max,
        )
        return fn(a, b).clamp_(min=0.0)

    @classmethod
    def normalize(cls, op, shape, mix=1.0, normalization_factor=1.0):
        height, width = shape[-2:]
        if mix < 1.0:
            flat = torch.ones(1, 1, height, width // 2 + 1)
            if mix <= 0.0:
                return flat
        if normalization_factor != 0:
            # RMS-normalise, blended toward "no change" by normalization_factor
            op *= torch.lerp(torch.scalar_tensor(1.0), 1.0 / op.square().mean().sqrt(), normalization_factor)
        if mix < 1.0:
            op = torch.lerp(flat, op, mix, out=op)
        return op

    def _gain(self, d: torch.Tensor) -> torch.Tensor:
        """1/f^alpha inside [min_freq, max_freq), Gaussian skirts outside."""
        gain = torch.empty_like(d)
        above_min = d >= self.min_freq
        below_max = d < self.max_freq
        band = above_min & below_max
        gain[band] = d[band].pow(-self.alpha)
        over = ~below_max
        gain[over] = math.pow(self.max_freq, -self.alpha) * torch.exp(
            -(d[over] - self.max_freq).square() / (self.rel_bw * self.max_freq) ** 2,
        )
        if self.min_freq > 0.0:
            under = ~above_min
            gain[under] = math.pow(self.min_freq, -self.alpha) * torch.exp(
                -(d[under] - self.min_freq).square() / (self.rel_bw * self.min_freq) ** 2,
            )
        return gain

    def build(self, shape, override_oversample=None, composed=True) -> torch.Tensor:
        oversample = self.oversample if override_oversample is None else override_oversample
        height, width = shape[-2:]
        bins = width // 2 + 1
        # oversampled fftshift(rfft2freq) grid; complex numbers only as 2-D points for the rotation
        grid = torch.complex(
            torch.linspace(0, 0.5, oversample * bins),
            torch.linspace(-(height // 2) / height, ((height - 1) // 2) / height, oversample * height).unsqueeze(1),
        )
        if abs(self.rotate) >= 1e-3:
            grid *= torch.exp(1.0j * torch.deg2rad(torch.scalar_tensor(self.rotate)))
        if self.stretch > 1.0:
            grid.real *= self.stretch
        else:
            grid.imag *= 1.0 / self.stretch
        if abs(self.pnorm - 2.0) < 1e-3:
            dist = grid.abs()
        else:
            dist = torch.view_as_real(grid).abs().pow(self.pnorm).sum(-1).pow(1.0 / self.pnorm)
        op = torch.nn.functional.interpolate(
            self._gain(dist)[None, None, ...],
            (height, bins),
            mode="bilinear",
            align_corners=True,
        )
        op = op.roll(-(height // 2), -2)  # ifftshift along H
        if self.alpha > 0:
            op[..., 0, 0] = 0  # the gain diverges at DC
        if self.scale != 1.0:
            op *= self.scale
        if composed and self.compose_with is not None:
            return self.compose(op, self.compose_with.build(shape, override_oversample=override_oversample), self.compose_mode)
        return op


_FILTER_CACHE: dict[tuple, torch.Tensor] = {}


def _parse_correlation(channel_correlation):
    if isinstance(channel_correlation, str):
        vals = tuple(float(v) for v in (p.strip() for p in channel_correlation.split(",")) if v)
        return torch.tensor(vals, device="cpu", dtype=torch.float)
    return channel_correlation


class PowerNoiseItem(CustomNoiseItemBase):
    """SonarPowerNoise item (:297-408)."""

    def __init__(self, factor, *, channel_correlation, power_filter=None, **kwargs):
        channel_correlation = _parse_correlation(channel_correlation)
        if power_filter is None:
            fargs = {
                k: kwargs.pop(k) for k in ("min_freq", "max_freq", "stretch", "rotate", "pnorm", "alpha") if k in kwargs
            }
            power_filter = PowerFilter(**fargs)
        super().__init__(factor, power_filter=power_filter, channel_correlation=channel_correlation, **kwargs)

    def make_filter(self, shape, oversample=None) -> torch.Tensor:
        norm_factor = getattr(self, "filter_norm_factor", 1.0)
        key = (tuple(shape[-2:]), oversample, self.mix, norm_factor, self.power_filter.cache_key())
        cached = _FILTER_CACHE.get(key)
        if cached is None:
            cached = PowerFilter.normalize(
                self.power_filter.build(shape, override_oversample=oversample),
                shape,
                mix=self.mix,
                normalization_factor=norm_factor,
            )
            if len(_FILTER_CACHE) > 64:
                _FILTER_CACHE.clear()
            _FILTER_CACHE[key] = cached
        return cached.clone()

    def make_noise_sampler_internal(self, x, noise_sampler, filter_rfft, normalized=True, *, spectral_input=False):
        shape = x.shape
        if x.ndim != 4:
            # the reference fails in ChannelMixer.apply with the same exception type (`b, c, h, w = shape`)
            raise ValueError(f"PowerNoise requires a 4-D latent (got {tuple(shape)}); fold video frames first")
        height, width = shape[-2:]
        mask = filter_rfft.reshape(height, width // 2 + 1).to(device=x.device, dtype=torch.float32).contiguous()
        mixer = ChannelMixer(shape[1], self.common_mode, self.channel_correlation).to(x.device, non_blocking=True)
        ortho = 1.0 / math.sqrt(height * width)
        factor = self.factor

        def sampler(sigma, sigma_next):
            drawn = noise_sampler(sigma, sigma_next)
            if spectral_input:
                # half spectrum drawn directly in frequency space: gain + irfft2(norm="ortho")
                noise = ops.spectral_filter(spectrum=drawn, mask=mask, hw=(height, width), out_scale=ortho)
            else:
                # spatial noise: rfft2(ortho) -> gain -> irfft2(ortho) in one kernel
                if drawn.dtype != torch.float32:
                    drawn = drawn.float()
                noise = ops.spectral_filter(
                    real=drawn.contiguous(), mask=mask, hw=(height, width), out_scale=ortho * ortho,
                )
            noise = mixer(noise, shape)
            return scale_noise(noise, factor, normalized=normalized)

        return sampler

    def make_noise_sampler(self, x, sigma_min=None, sigma_max=None, *, seed=None, cpu=True, normalized=True):
        shape, device = x.shape, x.device
        filter_rfft = self.make_filter(shape)
        if self.time_brownian:
            if sigma_min is None:
                raise ValueError("time correlated brownian mode is valid only for stochastic samplers")
            raise NotImplementedError("sonar_b200: time_brownian needs torchsde's BrownianTree (out of scope)")
        bins = filter_rfft.shape[-1]

        def spectrum_sampler(_s, _sn):
            # complex64 randn on x.device from the global generator, ignoring cpu and seed (:396-401)
            return rng.normal((*shape[:-1], bins), device=device, dtype=torch.complex64)

        return self.make_noise_sampler_internal(x, spectrum_sampler, filter_rfft, normalized=normalized, spectral_input=True)


class PowerFilterNoiseItem(PowerNoiseItem):
    """SonarPowerFilterNoise: power filter applied to another chain's (spatial) noise (:471-522)."""

    def __init__(self, factor, *, noise, normalize_noise, normalize_result, **kwargs):
        super().__init__(
            factor,
            noise=noise.clone(),
            normalize_noise=normalize_noise,
            normalize_result=normalize_result,
            **kwargs,
        )

    def clone_key(self, k):
        if k == "noise":
            return self.noise.clone()
        return super().clone_key(k)

    def make_noise_sampler(self, x, sigma_min=None, sigma_max=None, *, seed=None, cpu=True, normalized=True):
        normalize_noise = self.get_normalize("normalize_noise", False)  # noqa: FBT003
        normalize_result = self.get_normalize("normalize_result", normalized)
        filter_rfft = self.make_filter(x.shape)
        child = self.noise.make_noise_sampler(x, sigma_min, sigma_max, seed, cpu, normalized=normalize_noise)
        return self.make_noise_sampler_internal(x, child, filter_rfft, normalized=normalize_result, spectral_input=False)
