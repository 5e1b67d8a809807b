"""Parity of the CUDA noise path (generators, graph, power noise) with the reference: the golden
fixtures recorded from the unmodified reference are reproduced from the same injected draws, and
the oracle is matched at the BASELINE.json sizes."""
from __future__ import annotations

import math

import pytest
import torch

from helpers import NOISE_TYPE_NAMES, assert_close, oracle_noise_type
from oracle import sonar_oracle as orc

pytestmark = pytest.mark.gpu


def _run_type(sb, cuda, name, case):
    x = torch.zeros(tuple(case["shape"]), device=cuda)
    with sb.rng.injected(case["draws"]) as left:
        ns = sb.noise_graph.get_noise_sampler(name, x, None, None, seed=0, cpu=True, normalized=True)
        out = ns(None, None)
        assert not left, "generator consumed fewer draws than the reference"
    assert out.device == x.device and out.dtype == x.dtype and out.shape == x.shape
    return out


@pytest.mark.parametrize("name", NOISE_TYPE_NAMES)
def test_noise_types_golden(sb, cuda, golden, name):
    case = golden("noise_types")[name]
    out = _run_type(sb, cuda, {"pyramid_5d": "pyramid", "wavelet_odd": "wavelet"}.get(name, name), case)
    assert_close(out, case["out"], what=name)


def test_scale_noise_decisions(sb, cuda):
    """Both branches of the conditional normalisation (py/utils.py:100-106), in place."""
    torch.manual_seed(0)
    for mean, std in ((0.0, 1.0), (0.3, 1.0), (0.0, 1.7), (-0.2, 0.4)):
        base = torch.randn(4, 4, 32, 32) * std + mean
        want = orc.scale_noise(base.clone(), 0.8, normalized=True)
        dev = base.to(cuda)
        got = sb.hostutil.scale_noise(dev, 0.8, normalized=True)
        assert got.data_ptr() == dev.data_ptr()
        assert_close(got, want, what=f"mean {mean} std {std}")
    x = torch.randn(3, 5, 7)
    assert_close(sb.hostutil.scale_noise(x.to(cuda), 2.0, normalized=False), x * 2.0)
    assert sb.hostutil.scale_noise(torch.zeros(0, 4, device=cuda)).numel() == 0


POWER_CASES = ["white_33x40", "band_18x20", "rot_stretch_26x38", "odd_15x21"]
POWER_DEFAULTS = {
    "time_brownian": False, "alpha": 0.0, "max_freq": 0.7071, "min_freq": 0.0, "stretch": 1.0, "rotate": 0.0,
    "pnorm": 2.0, "mix": 1.0, "common_mode": 0.0, "channel_correlation": "1, 1, 1, 1, 1, 1",
}  # fmt: skip


def _power_chain(sb, **kw):
    chain = sb.noise_graph.CustomNoiseChain()
    chain.add(sb.spectral_noise.PowerNoiseItem(1.0, **(POWER_DEFAULTS | kw)))
    return chain


@pytest.mark.parametrize("name", POWER_CASES)
def test_power_noise_golden(sb, cuda, golden, name):
    case = golden("power_noise")[name]
    item = sb.spectral_noise.PowerNoiseItem(1.0, **(POWER_DEFAULTS | case["params"]))
    assert_close(item.make_filter(case["shape"]), case["filter"], what=f"{name} filter")
    x = torch.zeros(case["shape"], device=cuda)
    with sb.rng.injected(case["draws"]):
        out = _power_chain(sb, **case["params"]).make_noise_sampler(x, None, None, seed=0, normalized=True)(None, None)
    assert_close(out, case["out"], what=name)


def test_power_filter_noise_golden(sb, cuda, golden):
    case = golden("power_noise")["filter_noise_24x20"]
    inner = sb.noise_graph.CustomNoiseChain()
    inner.add(sb.noise_graph.CustomNoiseItem(1.0, noise_type="gaussian"))
    item = sb.spectral_noise.PowerFilterNoiseItem(
        1.0, noise=inner, normalize_noise=None, normalize_result=None,
        power_filter=sb.spectral_noise.PowerFilter(alpha=1.0), mix=1.0, common_mode=0.0,
        channel_correlation="1,1,1,1,1,1", time_brownian=True, filter_norm_factor=1.0,
    )  # fmt: skip
    chain = sb.noise_graph.CustomNoiseChain()
    chain.add(item)
    x = torch.zeros(case["shape"], device=cuda)
    with sb.rng.injected(case["draws"]):
        out = chain.make_noise_sampler(x, None, None, seed=0, normalized=True)(None, None)
    assert_close(out, case["out"], what="filter_noise")


def test_power_noise_video_5d_golden(sb, cuda, golden):
    case = golden("power_noise")["video_5d_18x20"]
    params = sb.noise_graph.CustomNoiseParametersNoise(
        1.0, noise=_power_chain(sb, alpha=1.0), normalize=None, override_device=None, override_dtype=None,
        frames_to_channels=True, ensure_square_aspect_ratio=False, fix_invalid=False, rng_mode="default",
        rng_offset_mode="disabled", rng_state_offset=0,
    )  # fmt: skip
    chain = sb.noise_graph.CustomNoiseChain()
    chain.add(params)
    x = torch.zeros(case["shape"], device=cuda)
    with sb.rng.injected(case["draws"]):
        out = chain.make_noise_sampler(x, None, None, seed=0, normalized=True)(None, None)
    assert_close(out, case["out"], what="video 5d")
    # a raw 5-D latent is rejected like the reference's ChannelMixer does
    with pytest.raises(ValueError):
        _power_chain(sb, alpha=1.0).make_noise_sampler(x, None, None, seed=0)


@pytest.mark.parametrize("hw", [(64, 64), (128, 128), (90, 160), (104, 152), (31, 47), (256, 256), (1, 8), (8, 1)])
def test_spectral_kernel_vs_torch_fft(sb, cuda, hw):
    """irfft2 of a NON-Hermitian half spectrum and the rfft2 round trip, any mixed-radix size."""
    h, w = hw
    torch.manual_seed(h * 1000 + w)
    spec = torch.randn(3, h, w // 2 + 1, dtype=torch.complex64)
    mask = torch.rand(h, w // 2 + 1) + 0.5
    want = torch.fft.irfft2(spec * mask, s=(h, w), norm="ortho")
    got = sb.ops.spectral_filter(spectrum=spec.to(cuda), mask=mask.to(cuda), hw=hw, out_scale=1.0 / math.sqrt(h * w))
    assert_close(got, want, what=f"irfft2 {hw}")
    real = torch.randn(3, h, w)
    want = torch.fft.irfft2(torch.fft.rfft2(real, norm="ortho") * mask, s=(h, w), norm="ortho")
    got = sb.ops.spectral_filter(real=real.to(cuda), mask=mask.to(cuda), hw=hw, out_scale=1.0 / (h * w))
    assert_close(got, want, what=f"rfft2-irfft2 {hw}")
    ident = sb.ops.spectral_filter(real=real.to(cuda), mask=None, hw=hw, out_scale=1.0 / (h * w))
    assert_close(ident, real, what=f"identity {hw}")


@pytest.mark.parametrize(
    ("hw", "planes"),
    [((32, 32), 301), ((16, 16), 700), ((18, 20), 5), ((160, 90), 3), ((48, 96), 7), ((100, 200), 2), ((64, 64), 149), ((8, 4), 33),
     # the c2r fold fused into the first row stage: first row radix 3 / 5 / 3 / 5 / 3 (three stages) / 4, m even
     ((24, 24), 200), ((30, 80), 9), ((50, 48), 4), ((20, 100), 6), ((12, 192), 5), ((16, 64), 40)],
)
def test_spectral_batched_groups_and_radices(sb, cuda, hw, planes):
    """Plane groups (several small planes per CTA, ragged last group) and every radix of the batched kernel
    (16, 10, 9, 8, 5, 4, 3, 2), spectrum and real input."""
    h, w = hw
    torch.manual_seed(h * 977 + w + planes)
    spec = torch.randn(planes, h, w // 2 + 1, dtype=torch.complex64)
    mask = torch.rand(h, w // 2 + 1) + 0.5
    want = torch.fft.irfft2(spec * mask, s=(h, w), norm="ortho")
    got = sb.ops.spectral_filter(spectrum=spec.to(cuda), mask=mask.to(cuda), hw=hw, out_scale=1.0 / math.sqrt(h * w))
    assert_close(got, want, what=f"irfft2 {hw} x{planes}")
    real = torch.randn(planes, h, w)
    want = torch.fft.irfft2(torch.fft.rfft2(real, norm="ortho") * mask, s=(h, w), norm="ortho")
    got = sb.ops.spectral_filter(real=real.to(cuda), mask=mask.to(cuda), hw=hw, out_scale=1.0 / (h * w))
    assert_close(got, want, what=f"rfft2-irfft2 {hw} x{planes}")
    sums = got._sonar_sums if hasattr(got, "_sonar_sums") else None
    if sums is not None:
        torch.cuda.synchronize()


def test_graph_golden(sb, cuda, golden):
    ng = sb.noise_graph
    g = golden("noise_graph")

    def chain_of(noise_type, factor=1.0):
        c = ng.CustomNoiseChain()
        c.add(ng.CustomNoiseItem(factor, noise_type=noise_type))
        return c

    shape = tuple(g["chain_two"]["shape"])
    x = torch.zeros(shape, device=cuda)

    # config C3 (scaled down)
    blended = ng.CustomNoiseChain()
    blended.add(ng.BlendedNoise(1.0, normalize=None, blend_function=sb.hostutil.BLENDING_MODES["lerp"],
                                custom_noise_1=chain_of("pyramid"), custom_noise_2=chain_of("perlin"), noise_2_percent=0.5))
    sched = ng.CustomNoiseChain()
    sched.add(ng.ScheduledNoise(1.0, noise=blended, start_sigma=10.0, end_sigma=1.0, normalize=None,
                                fallback_noise=chain_of("gaussian")))
    for tag in ("in_range", "fallback"):
        case = g[f"c3_scheduled_{tag}"]
        with sb.rng.injected(case["draws"]):
            ns = sched.make_noise_sampler(x, torch.tensor(0.03), torch.tensor(14.6), seed=0, normalized=True)
            out = ns(torch.tensor(case["sigma"]), torch.tensor(case["sigma"] * 0.9))
        assert_close(out, case["out"], what=f"c3 {tag}")
    with pytest.raises(ValueError):
        sched.make_noise_sampler(x, None, None, seed=0)(None, None)

    two = ng.CustomNoiseChain()
    two.add(ng.CustomNoiseItem(0.6, noise_type="gaussian"))
    two.add(ng.CustomNoiseItem(-0.4, noise_type="uniform"))
    with sb.rng.injected(g["chain_two"]["draws"]):
        out = two.make_noise_sampler(x, None, None, seed=0, normalized=True)(None, None)
    assert_close(out, g["chain_two"]["out"], what="chain_two")
    with sb.rng.injected(g["chain_rescaled"]["draws"]):
        out = two.rescaled(2.0).make_noise_sampler(x, None, None, seed=0, normalized=False)(None, None)
    assert_close(out, g["chain_rescaled"]["out"], what="chain_rescaled")

    case = g["composite"]
    comp = ng.CustomNoiseChain()
    comp.add(ng.CompositeNoise(1.0, dst_noise=chain_of("gaussian"), src_noise=chain_of("uniform"), normalize_dst=None,
                               normalize_src=None, normalize_result=None, mask=case["mask"]))
    with sb.rng.injected(case["draws"]):
        out = comp.make_noise_sampler(x, None, None, seed=0, normalized=True)(None, None)
    assert_close(out, case["out"], what="composite")

    case = g["blended_mask"]
    bm = ng.CustomNoiseChain()
    bm.add(ng.BlendedNoise(1.0, normalize=None, blend_function=sb.hostutil.BLENDING_MODES["lerp"],
                           custom_noise_1=chain_of("gaussian"), custom_noise_2=chain_of("uniform"),
                           custom_noise_mask=chain_of("gaussian"), noise_2_percent=0.25))
    with sb.rng.injected(case["draws"]):
        out = bm.make_noise_sampler(x, None, None, seed=0, normalized=True)(None, None)
    assert_close(out, case["out"], what="blended_mask")


def test_repeated_noise_bit_exact(sb, cuda, golden):
    """Index / flip / roll / negate work must be bit-exact (north_star)."""
    ng = sb.noise_graph
    case = golden("noise_graph")["repeated"]
    inner = ng.CustomNoiseChain()
    inner.add(ng.CustomNoiseItem(1.0, noise_type="gaussian"))
    rep = ng.CustomNoiseChain()
    rep.add(ng.RepeatedNoise(1.0, noise=inner, repeat_length=2, max_recycle=3, permute="enabled", normalize=None))
    x = torch.zeros(tuple(case["shape"]), device=cuda)
    with sb.rng.injected(case["draws"]):
        ns = rep.make_noise_sampler(x, None, None, seed=case["seed"], normalized=False)
        outs = torch.stack([ns(None, None) for _ in range(6)])
    assert torch.equal(outs.cpu(), case["out"])


def test_crop_samples_bit_exact(sb, cuda):
    t = torch.arange(2 * 3 * 10 * 12, dtype=torch.float32, device=cuda).reshape(2, 3, 10, 12)
    assert torch.equal(sb.hostutil.crop_samples(t, 6, 4), t[..., 3:7, 3:9])
    assert torch.equal(sb.hostutil.crop_samples(t, 6, 4, mode="top_left"), t[..., :4, :6])
    assert torch.equal(sb.hostutil.crop_samples(t, 6, 4, mode="bottom_right", offset_width=-2), t[..., 6:, 4:10])
    with pytest.raises(ValueError):
        sb.hostutil.crop_samples(t, 20, 4)


@pytest.mark.parametrize("mode", ["bilinear", "nearest-exact", "area"])
@pytest.mark.parametrize("src,dst", [((7, 5), (32, 48)), ((1, 1), (16, 16)), ((36, 36), (128, 128)), ((120, 90), (8, 12))])
def test_resample_matches_interpolate(sb, cuda, mode, src, dst):
    torch.manual_seed(1)
    x = torch.randn(2, 3, *src)
    want = torch.nn.functional.interpolate(x, size=dst, mode=mode)
    assert_close(sb.ops.resample(x.to(cuda), *dst, mode=mode), want, what=f"{mode} {src}->{dst}")


def test_full_size_configs_vs_oracle(sb, cuda):
    """BASELINE.json config C3 shape (16x16x128x128): pyramid + perlin blend vs the oracle."""
    shape = (16, 16, 128, 128)
    torch.manual_seed(0)
    sizes = orc.pyramid_level_sizes(128, 128, 10, [0.4963, 0.7682, 0.0885, 0.1320, 0.3074, 0.6341, 0.4901, 0.8964, 0.4556, 0.6323])
    host = [torch.tensor([v]) for v in [0.4963, 0.7682, 0.0885, 0.1320, 0.3074, 0.6341, 0.4901, 0.8964, 0.4556, 0.6323]][: len(sizes)]
    dev = [torch.randn(shape)] + [torch.randn(16, 16, h, w) for h, w in sizes]
    draws = [dev[0]]
    for hdraw, lvl in zip(host, dev[1:]):
        draws += [hdraw, lvl]
    want = orc.scale_noise(orc.pyramid_noise(iter(dev), shape, sizes), 1.0, normalized=True)
    x = torch.zeros(shape, device=cuda)
    with sb.rng.injected(draws):
        got = sb.noise_graph.get_noise_sampler("pyramid", x, None, None, normalized=True)(None, None)
    assert_close(got, want, what="pyramid C3")
    pdraws = [torch.rand(shape), torch.rand(16, 129, 129) * 2 * math.pi, torch.rand(16, 129, 129) * 2 * math.pi]
    want = orc.scale_noise(orc.perlin_noise(iter(pdraws), shape), 1.0, normalized=True)
    with sb.rng.injected(pdraws):
        got = sb.noise_graph.get_noise_sampler("perlin", x, None, None, normalized=True)(None, None)
    assert_close(got, want, what="perlin C3")


def test_producer_kernels_hand_over_their_moments(sb, cuda):
    """blend / axpby / pyramid / perlin / spectral reduce {sum, sum^2} of their output in the same
    launch; scale_noise uses them only while the tensor is unmodified."""
    torch.manual_seed(4)
    a, b = torch.randn(3, 4, 33, 40, device=cuda), torch.randn(3, 4, 33, 40, device=cuda)

    def check(t, what):
        slot = sb.ops.attached_sums(t)
        assert slot is not None, f"{what}: no statistics attached"
        d = t.double()
        want = torch.stack((d.sum(), (d * d).sum()))
        torch.testing.assert_close(slot, want, rtol=1e-6, atol=1e-4, msg=lambda m: f"{what}: {m}")

    check(sb.ops.blend(a, b, 0.3), "blend")
    check(sb.ops.axpby(a, 0.5, b, 2.0), "axpby")
    check(sb.ops.axpby(a[..., 1:].contiguous().flatten()[1:], 1.0, None), "axpby unaligned")
    levels = [torch.randn(12, 33, 40, device=cuda), torch.randn(12, 9, 11, device=cuda)]
    check(sb.ops.pyramid_accumulate(a.reshape(12, 33, 40), levels, [1.0, 0.7], out_hw=(33, 40)), "pyramid")
    angles = [torch.rand(4, 34, 41, device=cuda) * 6.28 for _ in range(2)]
    check(sb.ops.perlin_accumulate(torch.rand(3, 4, 33, 40, device=cuda), angles, shape=(3, 4, 33, 40)), "perlin")
    spec = torch.randn(5, 18, 11, dtype=torch.complex64, device=cuda)
    check(sb.ops.spectral_filter(spectrum=spec, mask=None, hw=(18, 20), out_scale=1.0 / math.sqrt(360)), "spectral")

    # scale_noise with handed-over statistics == scale_noise from scratch
    t = sb.ops.blend(a * 1.5 + 0.2, b, 0.3)
    want = orc.scale_noise(t.cpu().clone(), 1.0, normalized=True)
    assert_close(sb.hostutil.scale_noise(t, 1.0, normalized=True), want, what="scale_noise with attached sums")
    # a torch in-place op invalidates them (tensor version changes); so do our own in-place kernels
    t = sb.ops.blend(a, b, 0.3)
    t.mul_(3.0).add_(1.0)
    assert sb.ops.attached_sums(t) is None
    want = orc.scale_noise(t.cpu().clone(), 1.0, normalized=True)
    assert_close(sb.hostutil.scale_noise(t, 1.0, normalized=True), want, what="scale_noise after in-place edit")
    t = sb.ops.blend(a, b, 0.3)
    sb.ops.scale(t, 2.0)
    assert sb.ops.attached_sums(t) is None
    # a view of the same elements keeps them
    t = sb.ops.blend(a, b, 0.3)
    v = sb.ops.reshape_keep_sums(t, (3, 4, 33 * 40))
    assert sb.ops.attached_sums(v) is not None


def test_attached_sums_expire_when_the_ring_wraps(sb, cuda):
    """the producer-side statistics live in a 256-slot ring; a tensor held across a full turn of the ring
    must not be normalised with another tensor's sums."""
    a = torch.randn(4, 256, device=cuda)
    held = sb.ops.axpby(a, 2.0, None)
    slot = sb.ops.attached_sums(held)
    assert slot is not None
    want = torch.stack((held.double().sum(), held.double().square().sum()))
    torch.testing.assert_close(slot, want, rtol=1e-6, atol=1e-4)
    for _ in range(260):
        sb.ops.axpby(a, 1.0, None)
    assert sb.ops.attached_sums(held) is None
    got = sb.hostutil.scale_noise(held.clone(), 1.0, normalized=True)  # falls back to its own moments pass
    assert_close(got, orc.scale_noise(held.cpu().clone(), 1.0, normalized=True), what="stale tag ignored")
    # a second stream gets a ring of its own
    side = torch.cuda.Stream(device=cuda)
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        other = sb.ops.axpby(a, 3.0, None)
        s2 = sb.ops.attached_sums(other)
    side.synchronize()
    torch.testing.assert_close(s2, torch.stack((other.double().sum(), other.double().square().sum())), rtol=1e-6, atol=1e-4)


@pytest.mark.parametrize(("hw", "planes"), [((256, 256), 3), ((256, 256), 151), ((192, 320), 5), ((320, 192), 2), ((240, 288), 75)])
def test_spectral_cluster_kernel(sb, cuda, hw, planes):
    """Planes whose half spectrum exceeds one SM's shared memory (256x256: 264 KB) run on a cluster of two CTAs with
    the plane in distributed shared memory: spectrum and real input, more planes than clusters, output moments."""
    h, w = hw
    torch.manual_seed(h + w + planes)
    spec = torch.randn(planes, h, w // 2 + 1, dtype=torch.complex64)
    mask = torch.rand(h, w // 2 + 1) + 0.5
    want = torch.fft.irfft2(spec * mask, s=(h, w), norm="ortho")
    got = sb.ops.spectral_filter(spectrum=spec.to(cuda), mask=mask.to(cuda), hw=hw, out_scale=1.0 / math.sqrt(h * w))
    assert_close(got, want, what=f"cluster irfft2 {hw} x{planes}")
    sums = sb.ops.attached_sums(got)
    torch.testing.assert_close(sums.cpu(), torch.stack((want.double().sum(), want.double().square().sum())), rtol=1e-5, atol=1e-2)
    real = torch.randn(planes, h, w)
    want = torch.fft.irfft2(torch.fft.rfft2(real, norm="ortho") * mask, s=(h, w), norm="ortho")
    got = sb.ops.spectral_filter(real=real.to(cuda), mask=mask.to(cuda), hw=hw, out_scale=1.0 / (h * w))
    assert_close(got, want, what=f"cluster rfft2-irfft2 {hw} x{planes}")
    ident = sb.ops.spectral_filter(real=real.to(cuda), mask=None, hw=hw, out_scale=1.0 / (h * w))
    assert_close(ident, real, what=f"cluster identity {hw}")
