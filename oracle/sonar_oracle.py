"""CPU oracle for the sonar_b200 hot path.  TEST INFRASTRUCTURE -- NOT A PRODUCT PATH.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may import
this module. It restates, in eager CPU PyTorch / numpy, the algorithms of blepping/ComfyUI-sonar
that the CUDA kernels implement; each function cites the reference file:line it follows.

Pinning: the reference ships no tests or golden vectors. The oracle is pinned
against outputs of the reference itself, imported in the build container with recorded base draws
(tests/golden/make_golden.py -> tests/golden/*.pt, checked by tests/test_oracle_golden.py, and
directly against /root/reference when that tree is present). Exception -- the 2-D wavelet
transform: the reference delegates it to pytorch_wavelets, which is absent from the reference tree
and from this image (no pinned version), so `dwt2_forward` / `dwt2_inverse` restate that library's
published algorithm and their parity is UNPINNED (anchored on perfect reconstruction, orthonormality,
the equal-scales identity and the library-free known answers of tests/test_oracle_wavelet_known_answers.py:
closed-form and published filter values, PyWavelets' documentation examples, hand-derived band order and signs). The same holds for the "periodization" mode used by the
wavelet-filtered noise type (`_afb1d_per` / `_sfb1d_per`: perfect reconstruction incl. odd sizes, Parseval,
the haar block transform). FreeU-Extreme (`freeu_*`) IS pinned: tests/golden/freeu.pt.

Random inputs are always *injected*: generator functions take `draws`, the base tensors in the
order the reference would draw them.
"""

from __future__ import annotations

import math
from typing import Iterator, Sequence

import numpy as np
import torch
import torch.nn.functional as F  # noqa: N812

# =============================================================================================
# Philox4x32-10 + ATen's CUDA element mapping (integer work: bit-exact)
# =============================================================================================
PHILOX_M0, PHILOX_M1 = 0xD2511F53, 0xCD9E8D57
PHILOX_W0, PHILOX_W1 = 0x9E3779B9, 0xBB67AE85
MASK32 = 0xFFFFFFFF


def philox4x32_10(counter: np.ndarray, key: np.ndarray) -> np.ndarray:
    """Philox4x32-10 (Salmon et al. 2011; curand_philox4x32_x.h curand_Philox4x32_10).
    counter: (..., 4) uint32, key: (..., 2) uint32 -> (..., 4) uint32."""
    c = [counter[..., i].astype(np.uint64) for i in range(4)]
    k0 = key[..., 0].astype(np.uint64)
    k1 = key[..., 1].astype(np.uint64)
    for _ in range(10):
        p0 = np.uint64(PHILOX_M0) * c[0]
        p1 = np.uint64(PHILOX_M1) * c[2]
        hi0, lo0 = p0 >> np.uint64(32), p0 & np.uint64(MASK32)
        hi1, lo1 = p1 >> np.uint64(32), p1 & np.uint64(MASK32)
        c = [hi1 ^ c[1] ^ k0, lo1, hi0 ^ c[3] ^ k1, lo0]
        k0 = (k0 + np.uint64(PHILOX_W0)) & np.uint64(MASK32)
        k1 = (k1 + np.uint64(PHILOX_W1)) & np.uint64(MASK32)
    return np.stack([x.astype(np.uint32) for x in c], axis=-1)


def aten_policy(numel: int, sm_count: int = 148, max_threads_per_sm: int = 2048) -> tuple[int, int]:
    """(grid_blocks, counter_offset): ATen/native/cuda/DistributionTemplates.h:50-62, unroll 4."""
    if numel <= 0:
        return 0, 0
    grid = min(sm_count * (max_threads_per_sm // 256), (numel + 255) // 256)
    return grid, ((numel - 1) // (256 * grid * 4) + 1) * 4


def aten_raw_u32(numel: int, seed: int, offset: int, grid_blocks: int) -> np.ndarray:
    """The uint32 each element of a CUDA draw is derived from: element li belongs to thread
    t = li % T, Philox call k = (li // T) // 4, lane (li // T) % 4 (DistributionTemplates.h:66-86).
    For normals only the lane PAIR matters (Box-Muller), see `aten_normal`."""
    t_total = grid_blocks * 256
    li = np.arange(numel, dtype=np.int64)
    t = li % t_total
    r = li // t_total
    blocks = _philox_blocks(t, r // 4, seed, offset)
    return blocks[np.arange(numel), r % 4]


def _philox_blocks(thread: np.ndarray, call: np.ndarray, seed: int, offset: int) -> np.ndarray:
    ctr64 = (np.uint64(offset // 4) + call.astype(np.uint64))
    counter = np.stack(
        [
            (ctr64 & np.uint64(MASK32)).astype(np.uint32),
            (ctr64 >> np.uint64(32)).astype(np.uint32),
            thread.astype(np.uint32),
            np.zeros_like(thread, dtype=np.uint32),
        ],
        axis=-1,
    )
    key = np.broadcast_to(np.array([seed & MASK32, (seed >> 32) & MASK32], dtype=np.uint32), (*thread.shape, 2))
    return philox4x32_10(counter, key)


def aten_uniform(numel: int, seed: int, offset: int, grid_blocks: int, low: float = 0.0, high: float = 1.0) -> np.ndarray:
    """torch.rand / uniform_ on CUDA, bit-exact: curand_uniform4 (x * 2^-32 + 2^-33, one fused
    rounding) then rand * (to - from) + from with the `== to -> from` reversal (:487-501)."""
    raw = aten_raw_u32(numel, seed, offset, grid_blocks)
    inv = np.longdouble(np.float32(2.3283064e-10))
    # `x * CURAND_2POW32_INV + CURAND_2POW32_INV/2`: the uint32 is first converted to float (rounded to
    # 24 bits), then one fused multiply-add; the long-double product/sum below is exact -> 1 rounding
    u = (raw.astype(np.float32).astype(np.longdouble) * inv + inv / 2).astype(np.float32)
    lo, hi = np.float32(low), np.float32(high)
    rng = np.float32(hi - lo)
    val = (u.astype(np.longdouble) * np.longdouble(rng) + np.longdouble(lo)).astype(np.float32)
    return np.where(val == hi, lo, val).astype(np.float32)


def aten_normal(numel: int, seed: int, offset: int, grid_blocks: int, std: float = 1.0) -> np.ndarray:
    """torch.randn on CUDA up to transcendental rounding: curand_box_muller4 in float64.
    (The bit-exact check of the normal stream is against torch.randn itself on the GPU box.)"""
    t_total = grid_blocks * 256
    li = np.arange(numel, dtype=np.int64)
    t, r = li % t_total, li // t_total
    blocks = _philox_blocks(t, r // 4, seed, offset).astype(np.float64)
    lane = r % 4
    pair = lane // 2
    x = np.where(pair == 0, blocks[:, 0], blocks[:, 2])
    y = np.where(pair == 0, blocks[:, 1], blocks[:, 3])
    inv = float(np.float32(2.3283064e-10))
    inv_2pi = float(np.float32(np.float32(2.3283064e-10) * np.float32(6.2831855)))
    u = x * inv + inv / 2
    v = y * inv_2pi + inv_2pi / 2
    s = np.sqrt(-2.0 * np.log(u))
    out = np.where(lane % 2 == 0, np.sin(v) * s, np.cos(v) * s)
    return (out * std).astype(np.float32)


# =============================================================================================
# utils (reference py/utils.py)
# =============================================================================================
def torch_lerp(a, b, w):
    return torch.lerp(a, b, w)


BLENDING_MODES = {  # py/utils.py:17-21
    "lerp": torch.lerp,
    "inject": lambda a, b, t: (b * t).add_(a),
    "subtract_b": lambda a, b, t: a - b * t,
}


def scale_noise(noise: torch.Tensor, factor: float = 1.0, *, normalized: bool = True, threshold_std_devs: float = 2.5):
    """py/utils.py:85-106 (global, conditional normalisation; in place)."""
    numel = noise.numel()
    if not normalized or numel == 0:
        return noise.mul_(factor) if factor != 1 else noise
    mean, std = noise.mean().item(), noise.std().item()
    threshold = threshold_std_devs / math.sqrt(numel)
    if abs(mean) > threshold:
        noise -= mean
    if abs(1.0 - std) > threshold:
        noise /= std
    return noise.mul_(factor) if factor != 1 else noise


def normalize_to_scale(latent, target_min, target_max, *, dim=(-3, -2, -1), eps=1e-07):
    """py/utils.py:452-470."""
    mn, mx = latent.amin(dim=dim, keepdim=True), latent.amax(dim=dim, keepdim=True)
    out = latent - mn
    out /= (mx - mn).add_(eps)
    return out.mul_(target_max - target_min).add_(target_min).clamp_(target_min, target_max)


def get_ancestral_step(sigma_from, sigma_to, eta=1.0):
    """k-diffusion get_ancestral_step ([upstream]; call sites py/sonar.py:547, :678, :714)."""
    if not eta:
        return sigma_to, 0.0
    sigma_up = min(sigma_to, eta * (sigma_to**2 * (sigma_from**2 - sigma_to**2) / sigma_from**2) ** 0.5)
    sigma_down = (sigma_to**2 - sigma_up**2) ** 0.5
    return sigma_down, sigma_up


# =============================================================================================
# generators (reference py/noise_generation.py); `draws` = iterator over injected base tensors
# =============================================================================================
def _next(draws: Iterator[torch.Tensor], shape=None) -> torch.Tensor:
    t = next(draws).clone()
    if shape is not None and tuple(t.shape) != tuple(shape):
        if t.numel() != math.prod(shape):
            raise AssertionError(f"oracle draw shape {tuple(t.shape)} != expected {tuple(shape)}")
        t = t.reshape(shape)  # frames-to-channels generators draw 5-D and fold (noise_generation.py:202-209)
    return t


def uniform_noise(draws, sub_fac=0.5, mul_fac=3.46, mean_fac=0.0):
    """UniformNoiseGenerator.generate :508-514."""
    return _next(draws).sub_(sub_fac).mul_(mul_fac).add_(mean_fac)


def perlin_noise(draws, shape, div_fac=2.0, iterations=2, blend="lerp"):
    """PerlinOldNoiseGenerator.generate :478-493 with perlin_noise :424-476 / perlin_noise_tensor
    :353-421 specialised to what the reference actually calls: grid == output size, i.e. one cell
    per pixel, positions == (0.5, 0.5), smooth_step(0.5) == 0.5; corner order TL, TR, BL, BR
    (:331-337); the (C,H,W) result is broadcast over the batch (:484)."""
    b, c, h, w = shape
    blend_fn = BLENDING_MODES[blend]
    noise = _next(draws, shape).div_(div_fac)
    half = torch.full((1,), 0.5, dtype=noise.dtype)
    for _ in range(iterations):
        ang = _next(draws, (c, h + 1, w + 1))
        cos, sin = torch.cos(ang), torch.sin(ang)

        def dot(gc, gs, px, py):
            return gc * px + gs * py

        tl = dot(cos[:, :-1, :-1], sin[:, :-1, :-1], 0.5, 0.5)
        tr = dot(cos[:, :-1, 1:], sin[:, :-1, 1:], 0.5 - 1, 0.5)
        bl = dot(cos[:, 1:, :-1], sin[:, 1:, :-1], 0.5, 0.5 - 1)
        br = dot(cos[:, 1:, 1:], sin[:, 1:, 1:], 0.5 - 1, 0.5 - 1)
        row0 = blend_fn(tl, tr, half)
        row1 = blend_fn(bl, br, half)
        noise += blend_fn(row0, row1, half)
    return noise


def pyramid_level_sizes(h: int, w: int, iterations: int, rs: Sequence[float]) -> list[tuple[int, int]]:
    """Level sizes of PyramidNoiseGenerator.generate :626-648 given the host draws rs[i] in [0,1)."""
    sizes = []
    for i in range(iterations):
        r = rs[i] * 2 + 2
        w, h = max(1, int(w / (r**i))), max(1, int(h / (r**i)))
        sizes.append((h, w))
        if w == 1 or h == 1:
            break
    return sizes


def pyramid_noise(draws, shape, sizes: Sequence[tuple[int, int]], discount=0.7, mode="bilinear"):
    """PyramidNoiseGenerator.generate :621-649 (scale_samples -> F.interpolate, py/utils.py:58-67)."""
    b, c, h, w = shape
    noise = _next(draws, shape)
    for i, (lh, lw) in enumerate(sizes):
        level = _next(draws, (b, c, lh, lw))
        noise += F.interpolate(level, size=(h, w), mode=mode).mul_(discount**i)
    return noise


def highres_pyramid_noise(draws, shape, sizes, discount=0.7, mode="bilinear"):
    """HighresPyramidNoiseGenerator.generate :539-564; base = Uniform generator (un-normalised)."""
    b, c, h, w = shape
    noise = uniform_noise(draws).reshape(shape)
    for i, (lh, lw) in enumerate(sizes):
        level = _next(draws, (b, c, lh, lw))
        noise += F.interpolate(level, size=(h, w), mode=mode).mul_(discount**i)
    return noise


def pyramid_old_noise(draws, shape, discount=0.8, iterations=5, mode="nearest-exact"):
    """PyramidOldNoiseGenerator.generate :579-606 (draws already carry std = 0.5**i)."""
    b, c, h, w = shape
    noise = torch.zeros(shape)
    r = 1
    for i in range(iterations):
        r *= 2
        level = _next(draws, (b, c, h * r, w * r))
        noise += F.interpolate(level, size=(h, w), mode=mode).mul_(discount**i)
    return noise


def onef_noise(draws, shape, alpha=2.0, k=1.0, hfac=1.0, wfac=1.0, base_power=1.0, use_sqrt=True):
    """OneFNoiseGenerator.generate :737-759 (fftn over ALL dims, as the reference does)."""
    b, _c, h, w = shape
    noise = _next(draws, shape)
    fx, fy = torch.meshgrid(torch.fft.fftfreq(h, hfac), torch.fft.fftfreq(w, wfac), indexing="ij")
    power = (fx**2 + fy**2) ** (-alpha / 2.0)
    if k != 0:
        power = k / power
    power[0, 0] = base_power
    power = power.unsqueeze(0).expand(b, 1, h, w)
    noise_fft = torch.fft.fftn(noise)
    noise_fft /= torch.sqrt(power.to(noise_fft.dtype)) if use_sqrt else power.to(noise_fft.dtype)
    return torch.fft.ifftn(noise_fft).real


def green_test_noise(draws, shape, scale_fac=1.0, x_pow=2, y_pow=2, power_base=1):
    """GreenTestNoiseGenerator.generate :694-704."""
    _b, _c, h, w = shape
    noise = _next(draws, shape)
    scale = scale_fac / (w * h)
    fy = torch.fft.fftfreq(h)[:, None] ** y_pow
    fx = torch.fft.fftfreq(w) ** x_pow
    power = torch.sqrt(fy + fx)
    power[0, 0] = power_base
    noise = torch.fft.ifft2(torch.fft.fft2(noise) / torch.sqrt(power))
    noise *= scale / noise.std()
    return torch.real(noise)


def wavelet_octaves(height, width, *, octaves=4, initial_amplitude=1.0, persistence=0.5, height_factor=2.0,
                    width_factor=2.0, min_height=4, min_width=4, octave_height_factor=0.5, octave_width_factor=0.5):
    """WaveletNoiseGenerator.set_octave_data py/noise_generation.py:2238-2278 (forward octave order)."""
    amp, total, ch, cw, out = initial_amplitude, 0.0, height, width, []
    for octave in range(octaves):
        ch /= height_factor**octave
        cw /= width_factor**octave
        if amp == 0 or ch < min_height or cw < min_width or ch * octave_height_factor < 1 or cw * octave_width_factor < 1:
            break
        total += abs(amp)
        out.append((int(ch), int(cw), amp, total))
        amp *= persistence
    return out


def wavelet_noise(draws, shape, **kw):
    """WaveletNoiseGenerator.generate / _generate_octave py/noise_generation.py:2280-2327, defaults:
    adaptive_avg_pool2d down by 0.5, bilinear up, detail = noise - low-pass, update_blend 1."""
    b, c, h, w = shape
    ohf, owf = kw.get("octave_height_factor", 0.5), kw.get("octave_width_factor", 0.5)
    octs = wavelet_octaves(h, w, **kw)
    result = torch.zeros(shape)
    for oh, ow, amp, _total in octs:
        noise = _next(draws, (b, c, oh, ow))
        sh, sw = int(max(1, oh * ohf)), int(max(1, ow * owf))
        low = F.interpolate(F.adaptive_avg_pool2d(noise, (sh, sw)), size=(oh, ow), mode="bilinear")
        octave = torch.lerp(noise, noise - low, 1.0)
        if octave.shape != result.shape:
            octave = F.interpolate(octave, size=(h, w), mode="bilinear")
        result += octave.mul_(amp)
    return result / octs[-1][3]


def powerlaw_noise(draws, alpha=2.0, div_max_dims=None, use_sign=False, use_div_max_abs=True):
    """PowerLawNoiseGenerator.generate :775-786."""
    noise = _next(draws)
    modulation = torch.abs(noise) ** alpha
    noise = (torch.sign(noise) if use_sign else noise).mul_(modulation)
    if div_max_dims is not None:
        noise /= torch.amax(torch.abs(noise) if use_div_max_abs else noise, keepdim=True, dim=div_max_dims)
    return noise


# =============================================================================================
# power noise (reference py/nodes/powernoise.py)
# =============================================================================================
def power_filter(
    shape,
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
    mix=1.0,
    normalization_factor=1.0,
) -> torch.Tensor:
    """PowerFilter.build :189-266 followed by PowerFilter.normalize :169-187 -> (1,1,H,W/2+1)."""
    max_freq = max(max_freq, min_freq)
    height, width = shape[-2:]
    bins = width // 2 + 1
    fc = torch.complex(
        torch.linspace(0, 0.5, oversample * bins),
        torch.linspace(-(height // 2) / height, ((height - 1) // 2) / height, oversample * height).unsqueeze(1),
    )
    if abs(rotate) >= 1e-3:
        fc *= torch.exp(1.0j * torch.deg2rad(torch.scalar_tensor(rotate)))
    if stretch > 1.0:
        fc.real *= stretch
    else:
        fc.imag *= 1.0 / stretch
    d = fc.abs() if abs(pnorm - 2.0) < 1e-3 else torch.view_as_real(fc).abs().pow(pnorm).sum(-1).pow(1.0 / pnorm)
    op = torch.empty_like(d)
    hp, lp = d >= min_freq, d < max_freq
    band = hp & lp
    op[band] = d[band].pow(-alpha)
    op[~lp] = math.pow(max_freq, -alpha) * torch.exp(-(d[~lp] - max_freq).square() / (rel_bw * max_freq) ** 2)
    if min_freq > 0.0:
        op[~hp] = math.pow(min_freq, -alpha) * torch.exp(-(d[~hp] - min_freq).square() / (rel_bw * min_freq) ** 2)
    op = F.interpolate(op[None, None, ...], (height, bins), mode="bilinear", align_corners=True)
    op = op.roll(-(height // 2), -2)
    if alpha > 0:
        op[..., 0, 0] = 0
    if scale != 1.0:
        op *= scale
    # normalize
    if mix < 1.0:
        flat = torch.ones(1, 1, height, bins)
        if mix <= 0.0:
            return flat
    if normalization_factor != 0:
        op *= torch.lerp(torch.scalar_tensor(1.0), 1.0 / op.square().mean().sqrt(), normalization_factor)
    if mix < 1.0:
        op = torch.lerp(flat, op, mix, out=op)
    return op


def channel_mixer(channels: int, common_mode: float, channel_correlation) -> torch.Tensor:
    """ChannelMixer.build py/nodes/powernoise.py:63-88: unit-row-norm factor L*sqrt(D) of the LDL decomposition of
    the channel correlation matrix (off-diagonals = given correlations x common_mode, missing ones = common_mode)."""
    if isinstance(channel_correlation, str):
        channel_correlation = torch.tensor([float(v) for v in channel_correlation.split(",") if v.strip()], dtype=torch.float)
    pairs = channels * (channels - 1) // 2
    given = channel_correlation[:pairs]
    corr = torch.cat((given * common_mode, torch.full((pairs - given.numel(),), common_mode)))
    m = torch.eye(channels).index_put_(tuple(torch.tril_indices(channels, channels, offset=-1)), corr)
    m += m.tril(-1).mT
    m = torch.linalg.ldl_factor(m).LD
    diag = torch.diagonal_copy(m)
    torch.diagonal(m)[:] = 1.0
    m *= diag.clamp_min(0).sqrt().unsqueeze(0)
    m /= m.norm(dim=1, keepdim=True)
    return m


def channel_mix(noise: torch.Tensor, mixer: torch.Tensor) -> torch.Tensor:
    """ChannelMixer.apply :94-101: out[b, c] = sum_c' mixer[c, c'] noise[b, c'] per pixel."""
    b, c, h, w = noise.shape
    return (mixer @ noise.swapaxes(0, 1).reshape(c, -1)).reshape(c, b, h, w).swapaxes(1, 0)


def power_noise(draws, shape, filter_rfft, *, factor=1.0, normalized=True, spectral_input=True, mixer=None):
    """PowerNoiseItem sampler :355-366; `mixer` = channel_mixer(...) or None (identity, common_mode == 0)."""
    drawn = _next(draws)
    spec = drawn if spectral_input else torch.fft.rfft2(drawn, norm="ortho")
    noise = torch.fft.irfft2(spec.mul_(filter_rfft), s=shape[-2:], norm="ortho")
    if mixer is not None:
        noise = channel_mix(noise, mixer)
    return scale_noise(noise, factor, normalized=normalized)


# =============================================================================================
# Sonar samplers (reference py/sonar.py) -- elementwise recurrences, float32 like the reference
# =============================================================================================
# =============================================================================================
# Reference-latent guidance (py/sonar.py:323-411)
# =============================================================================================
def freeu_ffilter(x: torch.Tensor, filter_kwargs: dict, normalization_factor: float = 1.0) -> torch.Tensor:
    """ffilter py/nodes/freeu_extreme.py:10-29: irfft2(rfft2(x, ortho) * normalised PowerFilter, ortho)."""
    filt = power_filter(x.shape, normalization_factor=normalization_factor, **filter_kwargs)
    spec = torch.fft.rfft2(x.to(torch.float32), norm="ortho")
    return torch.fft.irfft2(spec.mul_(filt), s=x.shape[-2:], norm="ortho").to(x.dtype)


def freeu_scale(h: torch.Tensor, scale: float, hidden_mean: bool):
    """FreeUExtremeConfig.get_scale py/nodes/freeu_extreme.py:183-194."""
    if not hidden_mean:
        return scale
    hmean = h.mean(1).unsqueeze(1)
    flat = hmean.view(hmean.shape[0], -1)
    hmax, hmin = flat.max(dim=-1, keepdim=True)[0], flat.min(dim=-1, keepdim=True)[0]
    hmean = (hmean - hmin.unsqueeze(2).unsqueeze(3)) / (hmax - hmin).unsqueeze(2).unsqueeze(3)
    return 1.0 + (scale - 1.0) * hmean


def freeu_apply(x: torch.Tensor, *, scale=1.0, hidden_mean=True, slice=1.0, slice_offset=0.0, blend=1.0,  # noqa: A002
                blend_mode=None, filter=None, filter_norm=1.0) -> torch.Tensor:  # noqa: A002
    """FreeUExtremeConfig.apply py/nodes/freeu_extreme.py:203-227 (+ apply_filter :229-246); returns a new tensor."""
    x = x.clone()
    features = x.shape[1]
    sc = freeu_scale(x, scale, hidden_mean)
    size, offs = int(features * slice), int(features * slice_offset)
    part = x[:, offs : offs + size]
    filtered = part if filter is None else freeu_ffilter(part, filter, filter_norm)
    xslice = filtered * sc
    if blend != 1.0:
        modes = {"lerp": torch_lerp, "inject": lambda a, b, t: b * t + a, "subtract_b": lambda a, b, t: a - b * t}
        xslice = modes[blend_mode](part, xslice, blend)
    x[:, offs : offs + size] = xslice
    return x


def prepare_ref_latent(latent: torch.Tensor) -> torch.Tensor:
    """:335-341 -- per-plane standardisation of the reference latent."""
    avg = latent.mean(dim=(-2, -1), keepdim=True)
    std = latent.std(dim=(-2, -1), keepdim=True)
    return (latent - avg).div_(std).to(latent.dtype)


def guidance_shift(t: torch.Tensor, ref: torch.Tensor) -> torch.Tensor:
    """:372-378 -- give the reference the per-batch-item mean / (unbiased) std of `t`."""
    dim = tuple(range(-(t.ndim - 1), 0))
    return (ref * t.std(dim=dim, keepdim=True)).add_(t.mean(dim=dim, keepdim=True))


def guidance_linear(x, ref, factor, blend=None, do_shift=True):
    """:400-411"""
    return (blend or torch_lerp)(x, guidance_shift(x, ref) if do_shift else ref, factor)


def guidance_euler(sigma, sigma_next, x, denoised, ref, factor, do_shift=True):
    """:380-398 -- an Euler step of size (sigma_next - sigma) * factor towards the shifted reference."""
    if torch.equal(torch.as_tensor(sigma), torch.as_tensor(sigma_next)):
        return guidance_linear(x, ref, factor, do_shift=do_shift)
    d = (x - (guidance_shift(denoised, ref) if do_shift else ref)) / sigma
    return (d * ((sigma_next - sigma) * factor)).add_(x)


def guided_noise(draws, x, ref_latent, *, method, guidance_factor, factor=1.0, has_noise=True, sigma=None, sigma_next=None,
                 normalize_noise=True, normalize_result=True):
    """GuidedNoise.make_noise_sampler py/noise.py:565-623 around a Gaussian child chain: the child noise (or zeros)
    is pulled towards the RAW reference latent (bicubic-resized to the latent when needed); method "euler" uses the
    latent `x` the sampler was built for as the `denoised` argument (:608-616)."""
    ref = ref_latent.to(x, copy=True)
    if ref.shape[-2:] != x.shape[-2:]:
        ref = F.interpolate(ref, size=x.shape[-2:], mode="bicubic", align_corners=True)
    if has_noise:
        # child chain (normalized=normalize_noise): one Gaussian item, summed, normalised once at the chain level
        noise = scale_noise(_next(draws).clone(), 1.0, normalized=normalize_noise)
    else:
        noise = torch.zeros_like(x)
    if method == "linear":
        out = guidance_linear(noise, ref, guidance_factor, do_shift=has_noise)
    elif method == "euler":
        out = guidance_euler(sigma, sigma_next, noise, x, ref, guidance_factor, do_shift=has_noise)
    else:
        raise ValueError("Bad method")
    return scale_noise(out, factor, normalized=normalize_result)


class SonarOracle:
    """Restates SonarBase :70-320, SonarEuler.step :460-480, SonarEulerAncestral.step :541-573 and
    SonarDPMPPSDE.momentum_step :649-735. Noise tensors are passed in (already normalised)."""

    def __init__(
        self,
        *,
        momentum=0.95,
        momentum_hist=0.75,
        direction=1.0,
        mode="new",
        init="zero",
        momentum_start_step=0,
        momentum_end_step=9999,
        always_update_history=True,
        blend_mode="lerp",
        momentum_blend_mode=None,
        history_blend_mode=None,
        guidance_blend_mode=None,
        init_noise=None,
        guidance=None,
    ):
        # guidance: dict(guidance_type="LINEAR"|"EULER", factor, start_step, end_step, ref=prepared latent)
        self.guidance = guidance
        self.gblend = BLENDING_MODES[guidance_blend_mode or blend_mode]
        self.m, self.mh, self.direction = momentum, momentum_hist, direction
        self.mode, self.init = mode, init
        self.start, self.end, self.always = momentum_start_step, momentum_end_step, always_update_history
        self.mblend = BLENDING_MODES[momentum_blend_mode or blend_mode]
        self.hblend = BLENDING_MODES[history_blend_mode or blend_mode]
        self.hist = None
        self.init_noise = init_noise

    def ratios(self):  # :208-219
        d, mh = self.direction, self.mh
        return (mh, 1.0 + abs(d) * (1 - mh) if d < 0 else 2.0 - d, d)

    def check(self, step, is_history=False):  # :221-225
        if is_history and self.always:
            return True
        return self.start <= step <= self.end

    def init_hist(self, x, den, sigma, step):  # :169-206
        if self.hist is not None or not self.check(step, True):
            return
        src = den if self.mode == "denoised" else x
        if self.init == "sample":
            self.hist = src
        elif self.init == "sample_norm":
            self.hist = src / sigma
        elif self.init == "rand":
            self.hist = self.init_noise.clone()

    def update(self, v, step):  # :227-236
        if self.mh == 1 or not self.check(step, True):
            return
        hr, hs, ms = self.ratios()
        self.hist = v if self.hist is None else self.hblend(v * ms, self.hist * hs, hr)

    def mix(self, hist, item, sigma, is_denoised=False):  # :238-260
        den_mode = self.mode == "denoised"
        if self.m == 1 or hist is None or (den_mode and not is_denoised) or (not den_mode and is_denoised):
            return item
        return self.mblend(hist * sigma if is_denoised else hist, item, self.m)

    def momentum_denoised(self, x, den, sigma, step):  # :262-283
        out = self.mix(self.hist, den, sigma, is_denoised=True)
        self.init_hist(x, den, sigma, step)
        self.update(den / sigma, step)
        return out if self.check(step) else den

    def momentum_d(self, x, den, sigma, step, d=None):  # :285-307
        hd = self.hist
        d = (x - den) / sigma if d is None else d
        if self.m == 1 or self.mode == "denoised":
            return d
        md = self.mix(hd, d, sigma)
        self.init_hist(x, den, sigma, step)
        self.update(d if self.mode == "new" else md, step)
        return md if self.check(step) else d

    def guide(self, step, x, den, sigma, sigma_next):  # guidance_step :343-369
        g = self.guidance
        if g is None or g["factor"] == 0.0 or not g["start_step"] <= step <= g["end_step"]:
            return x
        if g["guidance_type"] == "LINEAR":
            return guidance_linear(x, g["ref"], g["factor"], blend=self.gblend)
        return guidance_euler(sigma, sigma_next, x, den, g["ref"], g["factor"])

    def euler_step(self, step, x, den, sigma, sigma_next):  # SonarEuler.step :460-480
        out = self.euler(step, x, den, sigma, sigma_next)
        return self.guide(step, out, den, sigma, sigma_next) if sigma_next > 0 else out

    def euler(self, step, x, den, sigma, sigma_down):  # :309-320
        dt = sigma_down - sigma
        den = self.momentum_denoised(x, den, sigma, step)
        return (self.momentum_d(x, den, sigma, step) * dt).add_(x)

    def euler_ancestral(self, step, x, den, sigma, sigma_next, noise, eta=1.0, s_noise=1.0):  # :541-573
        sigma_down, sigma_up = get_ancestral_step(sigma, sigma_next, eta=eta)
        out = self.euler(step, x, den, sigma, sigma_down)
        if sigma_next > 0:
            out = self.guide(step, out, den, sigma, sigma_next)
            out = out + noise * (s_noise * sigma_up)
        return out

    def dpmpp_sde(self, step, x, den, sigma, sigma_next, model, noise1, noise2, eta=1.0, s_noise=1.0):  # :649-735
        def sigma_fn(t):
            return t.neg().exp()

        def t_fn(s):
            return s.log().neg()

        if sigma_next == 0:
            sigma_down, _ = get_ancestral_step(sigma, sigma_next, eta=eta)
            return self.euler(step, x, den, sigma, sigma_down)
        r = 1 / 2
        t, t_next = t_fn(sigma), t_fn(sigma_next)
        h = t_next - t
        s = t + h * r
        fac = 1 / (2 * r)
        s_t, s_s = sigma_fn(t), sigma_fn(s)
        sd, su = get_ancestral_step(s_t, s_s, eta)
        s_ = t_fn(sd)
        md1 = self.momentum_denoised(x, den, sigma, step)
        diff_2 = (t - s_).expm1() * md1
        mdv = self.momentum_d(x, md1, sigma, step, d=diff_2)
        x_2 = ((sigma_fn(s_) / s_t) * x).sub_(mdv)
        x_2 += noise1.clone().mul_(s_noise * su)
        den2 = model(x_2, s_s)
        md2 = self.momentum_denoised(x, den2, s_s, step)
        s_t_next = sigma_fn(t_next)
        sd, su = get_ancestral_step(s_t, s_t_next, eta)
        t_down = t_fn(sd)
        denoised_d = (1 - fac) * md1 + fac * md2
        diff_1 = (t - t_down).expm1() * denoised_d
        mdv = self.momentum_d(x, md2, s_s, step, d=diff_1)
        out = ((sigma_fn(t_down) / s_t) * x).sub_(mdv)
        out = self.guide(step, out, denoised_d, sigma, sigma_next)  # :731 (guided by denoised_d, not the raw output)
        out += noise2.clone().mul_(s_noise * su)
        return out


# =============================================================================================
# 2-D DWT, pytorch_wavelets semantics ([upstream], restated; parity unpinned)
# =============================================================================================
def db_filters(dec_lo: Sequence[float]):
    """(dec_lo, dec_hi, rec_lo, rec_hi) with pywt's quadrature-mirror convention."""
    dec_lo = [float(v) for v in dec_lo]
    n = len(dec_lo)
    rec_lo = dec_lo[::-1]
    dec_hi = [(-1.0 if k % 2 == 0 else 1.0) * dec_lo[n - 1 - k] for k in range(n)]
    return dec_lo, dec_hi, rec_lo, dec_hi[::-1]


def _reflect(idx: np.ndarray, minx: float, maxx: float) -> np.ndarray:
    rng = maxx - minx
    mod = np.fmod(idx - minx, 2 * rng)
    mod = np.where(mod < 0, mod + 2 * rng, mod)
    return np.array(np.where(mod >= rng, 2 * rng - mod, mod) + minx, dtype=idx.dtype)


def _pad_indices(n: int, before: int, after: int, mode: str) -> np.ndarray:
    idx = np.arange(-before, n + after, dtype="int32")
    if mode == "symmetric":
        return _reflect(idx, -0.5, n - 0.5)
    if mode == "reflect":
        return _reflect(idx, 0, n - 1)
    if mode == "periodic":
        return np.mod(idx, n)
    raise ValueError(mode)


def _afb1d(x: torch.Tensor, lo: torch.Tensor, hi: torch.Tensor, mode: str, dim: int) -> torch.Tensor:
    """pytorch_wavelets.dwt.lowlevel.afb1d: pad, correlate with the reversed filters, stride 2."""
    c = x.shape[1]
    d = dim % 4
    n = x.shape[d]
    taps = lo.numel()
    out = (n + taps - 1) // 2
    p = 2 * (out - 1) - n + taps
    stride = (2, 1) if d == 2 else (1, 2)
    shape = [1, 1, 1, 1]
    shape[d] = taps
    filt = torch.cat([lo.flip(0).reshape(shape), hi.flip(0).reshape(shape)] * c, dim=0)
    if mode == "zero":
        pad = (0, 0, p // 2, (p + 1) // 2) if d == 2 else (p // 2, (p + 1) // 2, 0, 0)
        x = F.pad(x, pad)
    else:
        idx = torch.from_numpy(_pad_indices(n, p // 2, (p + 1) // 2, mode)).long()
        x = x.index_select(d, idx)
    return F.conv2d(x, filt, stride=stride, groups=c)


def _afb1d_per(x: torch.Tensor, lo: torch.Tensor, hi: torch.Tensor, dim: int) -> torch.Tensor:
    """pytorch_wavelets lowlevel.afb1d, mode "periodization" ([upstream], restated from memory): odd lengths
    repeat the last sample, roll by -L/2, zero-pad L-1, correlate with the reversed filters at stride 2, fold
    the L/2 overhanging outputs back onto the head, keep N/2."""
    c = x.shape[1]
    d = dim % 4
    taps = lo.numel()
    if x.shape[d] % 2 == 1:
        x = torch.cat((x, x.narrow(d, x.shape[d] - 1, 1)), dim=d)
    n = x.shape[d]
    x = torch.roll(x, -(taps // 2), dims=d)
    shape = [1, 1, 1, 1]
    shape[d] = taps
    filt = torch.cat([lo.flip(0).reshape(shape), hi.flip(0).reshape(shape)] * c, dim=0)
    stride = (2, 1) if d == 2 else (1, 2)
    pad = (taps - 1, 0) if d == 2 else (0, taps - 1)
    lohi = F.conv2d(x, filt, padding=pad, stride=stride, groups=c)
    n2, l2 = n // 2, taps // 2
    head = lohi.narrow(d, 0, l2) + lohi.narrow(d, n2, l2)
    return torch.cat((head, lohi.narrow(d, l2, n2 - l2)), dim=d) if n2 > l2 else head.narrow(d, 0, n2)


def _sfb1d_per(lo: torch.Tensor, hi: torch.Tensor, g0: torch.Tensor, g1: torch.Tensor, dim: int) -> torch.Tensor:
    """pytorch_wavelets lowlevel.sfb1d, mode "periodization": transposed convolution, the L-2 tail wrapped onto
    the head, N = 2 * len kept, rolled by 1 - L/2."""
    c = lo.shape[1]
    d = dim % 4
    taps = g0.numel()
    shape = [1, 1, 1, 1]
    shape[d] = taps
    stride = (2, 1) if d == 2 else (1, 2)
    f0 = torch.cat([g0.reshape(shape)] * c, dim=0)
    f1 = torch.cat([g1.reshape(shape)] * c, dim=0)
    y = F.conv_transpose2d(lo, f0, stride=stride, groups=c) + F.conv_transpose2d(hi, f1, stride=stride, groups=c)
    n = 2 * lo.shape[d]
    if taps > 2:
        head = y.narrow(d, 0, taps - 2) + y.narrow(d, n, taps - 2)
        y = torch.cat((head, y.narrow(d, taps - 2, n - (taps - 2))), dim=d)
    else:
        y = y.narrow(d, 0, n)
    return torch.roll(y, 1 - taps // 2, dims=d)


def dwt2_forward(x: torch.Tensor, filters, level: int, mode: str = "symmetric"):
    """pytorch_wavelets.DWTForward(J, wave, mode): (yl, [yh_1 (finest) .. yh_J])."""
    dec_lo, dec_hi, _, _ = filters
    lo = torch.tensor(dec_lo, dtype=x.dtype)
    hi = torch.tensor(dec_hi, dtype=x.dtype)
    yh, ll = [], x
    per = mode in ("per", "periodization")
    for _ in range(level):
        lohi = _afb1d_per(ll, lo, hi, dim=3) if per else _afb1d(ll, lo, hi, mode, dim=3)
        y = _afb1d_per(lohi, lo, hi, dim=2) if per else _afb1d(lohi, lo, hi, mode, dim=2)
        s = y.shape
        y = y.reshape(s[0], -1, 4, s[-2], s[-1])
        ll = y[:, :, 0].contiguous()
        yh.append(y[:, :, 1:].contiguous())
    return ll, yh


def _sfb1d(lo: torch.Tensor, hi: torch.Tensor, g0: torch.Tensor, g1: torch.Tensor, dim: int) -> torch.Tensor:
    """pytorch_wavelets lowlevel.sfb1d (non-periodized modes): transposed conv, pad L-2."""
    c = lo.shape[1]
    d = dim % 4
    taps = g0.numel()
    shape = [1, 1, 1, 1]
    shape[d] = taps
    stride = (2, 1) if d == 2 else (1, 2)
    pad = (taps - 2, 0) if d == 2 else (0, taps - 2)
    f0 = torch.cat([g0.reshape(shape)] * c, dim=0)
    f1 = torch.cat([g1.reshape(shape)] * c, dim=0)
    return F.conv_transpose2d(lo, f0, stride=stride, padding=pad, groups=c) + F.conv_transpose2d(
        hi, f1, stride=stride, padding=pad, groups=c,
    )


def dwt2_inverse(yl: torch.Tensor, yh: Sequence[torch.Tensor], filters, mode: str = "symmetric"):
    """pytorch_wavelets.DWTInverse(wave, mode); the expansive modes share one synthesis, periodization has its own."""
    sfb = _sfb1d_per if mode in ("per", "periodization") else _sfb1d
    _, _, rec_lo, rec_hi = filters
    g0 = torch.tensor(rec_lo, dtype=yl.dtype)
    g1 = torch.tensor(rec_hi, dtype=yl.dtype)
    ll = yl
    for h in reversed(list(yh)):
        if ll.shape[-2] > h.shape[-2]:
            ll = ll[..., :-1, :]
        if ll.shape[-1] > h.shape[-1]:
            ll = ll[..., :-1]
        lh, hl, hh = torch.unbind(h, dim=2)
        lo = sfb(ll, lh, g0, g1, dim=2)
        hi = sfb(hl, hh, g0, g1, dim=2)
        ll = sfb(lo, hi, g0, g1, dim=3)
    return ll


def expand_scales(levels: int, yh_scales) -> list[tuple[float, float, float]]:
    """expand_yh_scales py/wavelet_functions.py:148-190 for numeric inputs (no 'fill')."""
    if isinstance(yh_scales, (int, float)):
        return [(float(yh_scales),) * 3] * levels
    out = []
    for band in yh_scales:
        if isinstance(band, (int, float)):
            out.append((float(band),) * 3)
        else:
            vals = [float(v) for v in band[:3]]
            out.append(tuple(vals + [1.0] * (3 - len(vals))))
    return (out + [(1.0, 1.0, 1.0)] * levels)[:levels]


def wavelet_cfg(
    cond: torch.Tensor,
    uncond: torch.Tensor,
    x: torch.Tensor,
    filters,
    *,
    level: int,
    mode: str = "symmetric",
    diff=None,
    cond_scales=None,
    uncond_scales=None,
    final=None,
    difference_blend_mode: str = "inject",
    difference_blend_strength: float = 1.0,
    high_precision: bool = True,
    denoised_target: bool = True,
):
    """WaveletCFG.wavelet_cfg py/wavelet_cfg.py:749-791 + process_output :729-747, op for op.
    Each scales argument is None or (yl_scale, yh_scales)."""
    dtype = torch.float64 if high_precision else x.dtype

    def apply(coeffs, scales):  # wavelet_scaling py/wavelet_functions.py:193-216
        if scales is None:
            return coeffs
        yl, yh = coeffs[0].clone(), [b.clone() for b in coeffs[1]]
        if scales[0] != 1.0:
            yl *= scales[0]
        for sc, band in zip(expand_scales(len(yh), scales[1]), yh):
            for o in range(3):
                band[:, :, o] *= sc[o]
        return yl, yh

    c4 = cond.reshape(cond.shape[0], -1, *cond.shape[-2:]) if cond.ndim > 4 else cond
    u4 = uncond.reshape(uncond.shape[0], -1, *uncond.shape[-2:]) if uncond.ndim > 4 else uncond
    condw = apply(dwt2_forward(c4.to(dtype), filters, level, mode), cond_scales)
    uncondw = apply(dwt2_forward(u4.to(dtype), filters, level, mode), uncond_scales)
    diffw = apply((condw[0] - uncondw[0], [a - b for a, b in zip(condw[1], uncondw[1])]), diff)
    blend = BLENDING_MODES[difference_blend_mode]
    t = condw[0].new_full((1,), difference_blend_strength)
    resultw = apply((blend(uncondw[0], diffw[0], t), [blend(a, b, t) for a, b in zip(uncondw[1], diffw[1])]), final)
    result = dwt2_inverse(resultw[0], resultw[1], filters).to(dtype=x.dtype)
    if x.ndim > 4:
        result = result[..., : x.shape[-2], : x.shape[-1]].reshape(x.shape)
    else:
        result = result[tuple(slice(None, sz) for sz in x.shape)]
    return x - result if denoised_target else result
