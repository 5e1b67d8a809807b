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
import os

import torch

from . import ops, parallel, rng
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

    _built: dict = {}  # (channels, common_mode, correlation values) -> mixer: a pure function of its arguments

    def build(self) -> torch.Tensor:
        """The reference's construction (:63-90), memoised: a sampler is built per sampling run, and for a folded
        video latent (C = 528) the LDL factorisation alone is ~1 ms of host time per run."""
        key = (self.channel_count, float(self.common_mode), tuple(float(v) for v in self.channel_correlation.flatten().tolist()))
        hit = self._built.get(key)
        if hit is None:
            if len(self._built) > 16:
                self._built.clear()
            hit = self._built[key] = self._build()
        return hit.clone()

    def _build(self) -> torch.Tensor:
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
        if self.mixer is not None and not self.is_identity:  # (an identity mixer is never applied: nothing to move)
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
        packed = getattr(self, "mixer_packed", None)
        if c > 8 and (packed is None or packed.device != mixer.device):
            packed = self.mixer_packed = ops.pack_mixer(mixer)  # once per sampler: pre-tiled for the bulk-async GEMM
        return ops.channel_mix(noise.reshape(b, c, h, w).contiguous(), mixer, getattr(self, "mixer_host", None), packed)

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

        def finish(noise):
            return scale_noise(mixer(noise, shape), factor, normalized=normalized)

        def sampler(sigma, sigma_next):
            drawn = noise_sampler(sigma, sigma_next)
            if spectral_input:
                if isinstance(drawn, _PhiloxSpectrum):
                    # small draw: the half spectrum is regenerated from the Philox stream inside the FFT kernel
                    return finish(ops.spectral_filter(philox=drawn.args, mask=mask, hw=(height, width), out_scale=ortho))
                # half spectrum drawn directly in frequency space: gain + irfft2(norm="ortho")
                noise = ops.spectral_filter(spectrum=drawn, mask=mask, hw=(height, width), out_scale=ortho)
            else:
                # spatial noise: rfft2(ortho) -> gain -> irfft2(ortho) in one kernel
                if drawn.dtype != torch.float32:
                    drawn = drawn.float()
                noise = ops.spectral_filter(
                    real=drawn.contiguous(), mask=mask, hw=(height, width), out_scale=ortho * ortho,
                )
            return finish(noise)

        sampler.spectral = (mask, ortho, mixer, factor, normalized)  # for the look-ahead path of make_noise_sampler
        return sampler

    def make_noise_sampler(self, x, sigma_min=None, sigma_max=None, *, seed=None, cpu=True, normalized=True):
        shape, device = x.shape, x.device
        filter_rfft = self.make_filter(shape)
        if self.time_brownian:
            if sigma_min is None:
                raise ValueError("time correlated brownian mode is valid only for stochastic samplers")
            raise NotImplementedError("sonar_b200: time_brownian needs torchsde's BrownianTree (out of scope)")
        bins = filter_rfft.shape[-1]
        spec_shape = (*shape[:-1], bins)
        height, width = shape[-2:]
        planes = math.prod(shape[:-2])
        in_kernel_ok = x.ndim == 4 and ops.spectral_plan_batched(height, width, planes)
        std = 1.0 / math.sqrt(2.0)  # torch's complex normal: real and imaginary part each N(0, 1/2)

        def spectrum_sampler(_s, _sn):
            # complex64 randn on x.device from the global generator, ignoring cpu and seed (:396-401)
            if in_kernel_ok and IN_KERNEL_PHILOX and rng._INJECT is None and rng._PENDING is None:  # noqa: SLF001
                total, begin = parallel.global_draw_geometry(spec_shape)
                grid, _ = ops.philox_policy_cached(device.index if device.index is not None else torch.cuda.current_device(), 2 * total)
                if 2 * total <= 256 * grid:
                    # a single ATen row: torch's own kernel spends one Philox call per value too, so regenerating the
                    # draw inside the FFT kernel costs no extra instructions and saves the spectrum's write, read and
                    # launch (but see IN_KERNEL_PHILOX: such draws have too few planes to fill the GPU)
                    draw = ops.reserve_draw(2 * total, device)
                    return _PhiloxSpectrum((draw, 2 * begin, std, spec_shape[:-2], device))
            return rng.normal(spec_shape, device=device, dtype=torch.complex64)

        sampler = self.make_noise_sampler_internal(x, spectrum_sampler, filter_rfft, normalized=normalized, spectral_input=True)
        mask, ortho, mixer, factor, _ = sampler.spectral
        if x.ndim == 4 and factor == 1 and not normalized and (mixer.mixer is None or mixer.is_identity):

            state: dict = {}  # the sampler's look-ahead slab (see below)

            def lookahead(count: int, *, chunk: int | None = None, slot: int | None = None, grid_limits: tuple | None = None):
                """The next `count` samples in ONE Philox launch + ONE FFT launch (per-sample statistics): a sample is a
                function of the generator state only, so a consumer that knows it will ask again (a sampler with
                `noise_draws_left`) can have them made together -- small launches run at a fraction of the throughput of
                large ones (528 planes of 90x160: 36 us, 4224 planes: 22 us per 528). Returns [(raw sample, its {sum,
                sum^2} row, draw)], or None when not applicable. The torch generator is NOT advanced: the consumer
                advances it draw by draw (and drops the rest if somebody else drew in between).
                chunk / slot (pipelined consumers, samplers._prefetch_noise): make exactly min(count, chunk) samples into
                output buffer `slot` (0 / 1) of the slab, so that the consumer can read one buffer on its stream while
                the next samples are written into the other one on a second stream. grid_limits = (fill, fft): CTAs per
                SM the two producer launches may occupy (sonar_set_grid_limit; 0 = no limit)."""
                if rng._INJECT is not None or rng._PENDING is not None or LOOKAHEAD_BYTES <= 0:  # noqa: SLF001
                    return None
                numel = math.prod(shape)
                if chunk is None:
                    cap = min(ops.FILL_BATCH_MAX, max(1, LOOKAHEAD_BYTES // (12 * numel)))
                    count = -(-count // -(-count // cap))  # equal batches: 18 draws under a cap of 11 are made 9 + 9
                    if count < 2:
                        return None
                    room = count
                else:
                    room = max(1, min(chunk, ops.FILL_BATCH_MAX))  # the slab is laid out for full chunks
                    count = max(1, min(count, room))
                total, begin = parallel.global_draw_geometry(spec_shape)
                draws = ops.peek_draws(2 * total, device, count)
                # One slab per sampler for spectrum + samples + their statistics, taken from (and returned to) a pool of
                # our own: a batch is made only when the previous one has been consumed by kernels already enqueued, so
                # the slab is simply overwritten. Going through torch's caching allocator for ~1 GB blocks nine times per
                # run fragments it (smaller tensors get carved out of the freed slab) and ends in cudaMalloc / cudaFree
                # stalls of 20-180 ms inside a sampler step.
                spec_bytes = 8 * room * math.prod(spec_shape)
                out_bytes = 4 * room * numel
                tab_bytes = -(-16 * room // 256) * 256
                n_slots = 1 if slot is None else 2
                need = spec_bytes + n_slots * (out_bytes + tab_bytes)
                slab = state.get("slab")
                if slab is None or slab.tensor.numel() < need or state.get("layout") != (room, n_slots):
                    slab = state["slab"] = _Slab.acquire(device, need)
                    state["layout"] = (room, n_slots)
                base = spec_bytes + (slot or 0) * (out_bytes + tab_bytes)
                spec = slab.tensor[: 8 * count * math.prod(spec_shape)].view(torch.complex64).reshape(count, *spec_shape)
                out = slab.tensor[base : base + 4 * count * numel].view(torch.float32).reshape(count * planes, height, width)
                table = slab.tensor[base + out_bytes : base + out_bytes + 16 * count].view(torch.float64).reshape(count, 2)
                table.zero_()
                fill_limit, fft_limit = grid_limits or (0, 0)
                try:
                    ops.set_grid_limit(fill_limit)
                    ops.philox_fill_batch([(d, spec[j], "normal", 0.0, std, 2 * begin) for j, d in enumerate(draws)])
                    ops.set_grid_limit(fft_limit)
                    out, table = ops.spectral_filter(
                        spectrum=spec.reshape(count * planes, height, bins), mask=mask, hw=(height, width), out_scale=ortho,
                        segment_planes=planes, out=out, table=table,
                    )
                finally:
                    ops.set_grid_limit(0)
                out = out.reshape(count, *shape)
                return NoiseBatch(((out[j], table[j], draws[j]) for j in range(count)), table)

            sampler.lookahead = lookahead
        return sampler


LOOKAHEAD_BYTES = int(os.environ.get("SONAR_B200_LOOKAHEAD_BYTES", 2 << 30))  # spectrum + samples made ahead of time by one look-ahead batch
# Regenerating a single-row complex draw inside the FFT kernel (ops.spectral_filter(philox=...)) is instruction-neutral
# and saves a launch, but a single-row draw has at most ~70 planes of 64x64: the Philox work then runs on as many CTAs
# instead of the ~66 x 256 threads of the fill kernel. Measured on B200, C1 (1x4x64x64): 39 us in-kernel vs 7 + 17.5 us
# as two launches. Kept behind this switch (and tested), off by default.
IN_KERNEL_PHILOX = os.environ.get("SONAR_B200_SPECTRAL_PHILOX") == "1"


class NoiseBatch(list):
    """[(raw sample, its {sum, sum^2} row, draw)] of one look-ahead batch + the (count, 2) statistics table the rows are
    views of (what a sharded consumer all-reduces, once per batch)."""

    def __init__(self, items, table):
        super().__init__(items)
        self.table = table


class _Slab:
    """A look-ahead buffer owned by one noise sampler; returns to the pool (keyed by device, stream and size) when the
    sampler is garbage-collected, so the next sampler of the same shape on the same stream reuses it without an
    allocator round trip."""

    _pool: dict = {}

    def __init__(self, key, tensor):
        self.key, self.tensor = key, tensor

    @classmethod
    def acquire(cls, device, nbytes: int) -> "_Slab":
        idx = device.index if device.index is not None else torch.cuda.current_device()
        key = (idx, ops._raw_stream(idx), int(nbytes))  # noqa: SLF001
        free = cls._pool.get(key)
        tensor = free.pop() if free else torch.empty(int(nbytes), device=device, dtype=torch.uint8)
        return cls(key, tensor)

    def __del__(self):
        pool = self._pool.setdefault(self.key, [])
        if len(pool) < 2:  # (at most two idle slabs per shape: the rest goes back to torch's allocator)
            pool.append(self.tensor)


class _PhiloxSpectrum:
    """A reserved complex normal draw that the FFT kernel regenerates in registers (ops.spectral_filter(philox=...))."""

    __slots__ = ("args",)

    def __init__(self, args):
        self.args = args


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
