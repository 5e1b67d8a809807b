"""CUDA parity for the round-2 fixtures recorded from the unmodified reference: north-star config 5 as specified
(frames_to_channels power noise as the custom noise of sonar_dpmpp_sde), the non-identity ChannelMixer kernel and
GuidedNoise; plus config 5 at its full per-GPU shard size against the oracle."""
from __future__ import annotations

import pytest
import torch

from helpers import assert_close, stub_model
from oracle import sonar_oracle as orc
from test_round2_oracle import GUIDED, POWER_DEFAULTS, c5_oracle_run

pytestmark = pytest.mark.gpu

POWER_KW = {"time_brownian": False, "common_mode": 0.0, "channel_correlation": "1, 1, 1, 1, 1, 1"} | POWER_DEFAULTS


def power_chain(sb, **kw):
    chain = sb.noise_graph.CustomNoiseChain()
    chain.add(sb.spectral_noise.PowerNoiseItem(1.0, **(POWER_KW | kw)))
    return chain


def video_chain(sb, **kw):
    item = sb.noise_graph.CustomNoiseParametersNoise(
        1.0, noise=power_chain(sb, **kw), normalize=None, override_device=None, override_dtype=None,
        frames_to_channels=True, ensure_square_aspect_ratio=False, fix_invalid=False, rng_mode="default",
        rng_offset_mode="disabled", rng_state_offset=0,
    )  # fmt: skip
    chain = sb.noise_graph.CustomNoiseChain()
    chain.add(item)
    return chain


@pytest.mark.parametrize("variant", ["default", "classic"])
def test_c5_job_golden(sb, cuda, golden, variant):
    """Config 5 exactly as north_star states it, scaled down: the reference's own output for
    sonar_dpmpp_sde(custom_noise = SonarCustomNoiseParameters(frames_to_channels) o SonarPowerNoise(alpha=1))."""
    g = golden("round2")["c5"]
    case = g[variant]
    steps = []
    with sb.rng.injected(case["draws"]) as left:
        out = sb.samplers.SonarDPMPPSDE.sampler(
            stub_model, g["x0"].to(cuda), g["sigmas"].to(cuda), extra_args={"seed": 0}, disable=True,
            sonar_params=dict(case["params"]) | {"custom_noise": video_chain(sb, alpha=1.0)},
            callback=lambda d: steps.append(d["x"].clone()), eta=1.0, s_noise=1.0,
        )  # fmt: skip
        assert not left, "unused recorded draws"
    assert_close(torch.stack(steps), case["steps"], what=f"c5 {variant}")
    assert_close(out, case["out"], what=f"c5 {variant} final")


@pytest.mark.parametrize("pipelined", [False, True])
def test_c5_job_shard_size_vs_oracle(sb, cuda, monkeypatch, pipelined):
    """Config 5 at the per-GPU shard size 1x16x33x90x160 with the DEVICE Philox stream (no injection): the product
    draws its complex spectra from torch's CUDA generator state; the oracle is fed torch.randn(complex64,
    device='cuda') from the same seed. pipelined: the schedule of larger shards forced at this size -- producers on the
    second stream, the 48-register FFT form, the step launch in two parts."""
    monkeypatch.setattr(sb.samplers, "PIPELINE_MIN_NUMEL", 0 if pipelined else 1 << 62)
    sigmas = torch.tensor([14.6, 5.0, 1.2, 0.0])
    torch.manual_seed(5)
    x0 = torch.randn(1, 16, 33, 90, 160) * sigmas[0]

    def model(x, sigma, **_kw):
        return x * 0.9

    torch.manual_seed(1234)
    launches = sb.ops.LAUNCH_COUNT
    got = sb.samplers.SonarDPMPPSDE.sampler(
        model, x0.to(cuda), sigmas.to(cuda), extra_args={"seed": 0}, disable=True,
        sonar_params={"custom_noise": video_chain(sb, alpha=1.0)},
    )
    offset_after = torch.cuda.default_generators[0].get_offset()
    launches = sb.ops.LAUNCH_COUNT - launches
    torch.manual_seed(1234)
    draws = [torch.randn((1, 528, 90, 81), dtype=torch.complex64, device=cuda).cpu() for _ in range(4)]
    assert torch.cuda.default_generators[0].get_offset() == offset_after  # generator advanced exactly like torch.randn
    b, c, f, h, w = x0.shape
    filt = orc.power_filter((b, c * f, h, w), alpha=1.0)
    it = iter(draws)

    def noise():
        return orc.scale_noise(orc.power_noise(it, (b, c * f, h, w), filt, normalized=False).reshape(x0.shape), 1.0, normalized=True)

    o, x = orc.SonarOracle(), x0.clone()
    for i in range(3):
        den = model(x, sigmas[i])
        if sigmas[i + 1] == 0:
            x = o.dpmpp_sde(i, x, den, sigmas[i], sigmas[i + 1], model, None, None)
        else:
            n1, n2 = noise(), noise()
            x = o.dpmpp_sde(i, x, den, sigmas[i], sigmas[i + 1], model, n1, n2)
    assert_close(got, x, what="c5 shard")
    # 2 steps x 2 half steps x (one noise launch + one fused step launch) + the final Euler step
    if not pipelined:
        assert launches <= 2 * 2 * 2 + 1, launches


@pytest.mark.parametrize("name", ["c4_common", "c4_corr", "c3_neg", "c16_short_corr"])
def test_channel_mixer_golden(sb, cuda, golden, name):
    case = golden("round2")["mixer"][name]
    x = torch.zeros(case["shape"], device=cuda)
    with sb.rng.injected(case["draws"]) as left:
        out = power_chain(sb, **case["params"]).make_noise_sampler(x, None, None, seed=0, cpu=True, normalized=True)(None, None)
        assert not left
    assert_close(out, case["out"], what=name)


def test_channel_mixer_video_golden(sb, cuda, golden):
    case = golden("round2")["mixer"]["video_c132"]
    x = torch.zeros(case["shape"], device=cuda)
    with sb.rng.injected(case["draws"]) as left:
        out = video_chain(sb, **case["params"]).make_noise_sampler(x, None, None, seed=0, cpu=True, normalized=True)(None, None)
        assert not left
    assert_close(out, case["out"], what="video_c132")


def test_channel_mixer_filter_noise_golden(sb, cuda, golden):
    case = golden("round2")["mixer"]["filter_noise_c4"]
    ng, sn = sb.noise_graph, sb.spectral_noise
    inner = ng.CustomNoiseChain()
    inner.add(ng.CustomNoiseItem(1.0, noise_type="gaussian"))
    item = sn.PowerFilterNoiseItem(
        1.0, noise=inner, normalize_noise=None, normalize_result=None, power_filter=sn.PowerFilter(alpha=1.0),
        mix=1.0, common_mode=0.4, channel_correlation="1,-0.5,0.5,1,1,0.2", time_brownian=True, filter_norm_factor=1.0,
    )  # fmt: skip
    chain = ng.CustomNoiseChain()
    chain.add(item)
    x = torch.zeros(case["shape"], device=cuda)
    with sb.rng.injected(case["draws"]) as left:
        out = chain.make_noise_sampler(x, None, None, seed=0, cpu=True, normalized=True)(None, None)
        assert not left
    assert_close(out, case["out"], what="filter_noise_c4")


@pytest.mark.parametrize(
    "shape", [(1, 528, 90, 160), (2, 132, 10, 12), (3, 5, 7, 9), (2, 8, 16, 16), (1, 70, 9, 13), (2, 16, 32, 32), (4, 4, 128, 128)],
)
def test_channel_mix_kernel_vs_matmul(sb, cuda, shape):
    """out[b, c] = sum_k M[c, k] in[b, k] on both kernel forms (per-pixel mat-vec for C <= 8, tiled GEMM above),
    ragged sizes included; the fused output moments must equal a separate reduction."""
    torch.manual_seed(shape[1])
    b, c, h, w = shape
    noise = torch.randn(shape)
    mixer = orc.channel_mixer(c, 0.15, "1, 0.5, -0.25, 0.75").contiguous()  # (ldl_factor returns column-major storage)
    want = orc.channel_mix(noise.double(), mixer.double())
    got = sb.ops.channel_mix(noise.to(cuda), mixer.to(cuda), mixer, sb.ops.pack_mixer(mixer.to(cuda)) if c > 8 else None)
    assert_close(got, want.float(), what=f"channel_mix {shape}")
    sums = sb.ops.attached_sums(got)
    assert sums is not None
    ref = torch.stack((want.sum(), want.square().sum()))
    torch.testing.assert_close(sums.cpu(), ref, rtol=1e-6, atol=1e-3)


@pytest.mark.parametrize("name", GUIDED)
def test_guided_noise_golden(sb, cuda, golden, name):
    g = golden("round2")["guided"]
    case = g[name]
    cfg = case["config"]
    ng = sb.noise_graph
    inner = None
    if cfg["noise"]:
        inner = ng.CustomNoiseChain()
        inner.add(ng.CustomNoiseItem(1.0, noise_type="gaussian"))
    item = ng.GuidedNoise(
        cfg.get("factor", 1.0), guidance_factor=cfg["guidance_factor"], ref_latent=case["ref"].clone(), method=cfg["method"],
        normalize_noise=cfg.get("normalize_noise"), normalize_result=cfg.get("normalize_result"), noise=inner,
    )  # fmt: skip
    chain = ng.CustomNoiseChain()
    chain.add(item)
    s, sn = case["sigmas"]
    with sb.rng.injected(case["draws"]) as left:
        ns = chain.make_noise_sampler(g["x"].to(cuda), torch.tensor(0.03), torch.tensor(14.6), seed=0, cpu=True, normalized=True)
        out = ns(torch.tensor(s), torch.tensor(sn))
        assert not left
    assert_close(out, case["out"], what=name)


@pytest.mark.parametrize("shape", [(1, 4, 64, 64), (2, 4, 32, 48), (8, 4, 128, 128)])
@pytest.mark.parametrize("in_kernel", [False, True])
def test_power_noise_device_philox_vs_oracle(sb, cuda, shape, in_kernel, monkeypatch):
    """Config C1 with the DEVICE generator: small draws (a single ATen row) regenerate the complex spectrum from the
    Philox stream inside the FFT kernel; larger ones materialise it. Either way the sample equals the oracle fed
    torch.randn(complex64, device='cuda') from the same generator state, and the generator advances identically."""
    b, c, h, w = shape
    monkeypatch.setattr(sb.spectral_noise, "IN_KERNEL_PHILOX", in_kernel)
    x = torch.zeros(shape, device=cuda)
    ns = power_chain(sb, alpha=1.0).make_noise_sampler(x, None, None, seed=0, cpu=True, normalized=True)
    torch.manual_seed(321)
    launches = sb.ops.LAUNCH_COUNT
    got = [ns(None, None) for _ in range(2)]
    launches = sb.ops.LAUNCH_COUNT - launches
    offset = torch.cuda.default_generators[0].get_offset()
    torch.manual_seed(321)
    draws = [torch.randn((b, c, h, w // 2 + 1), dtype=torch.complex64, device=cuda).cpu() for _ in range(2)]
    assert torch.cuda.default_generators[0].get_offset() == offset
    filt = orc.power_filter(shape, alpha=1.0)
    for j in range(2):
        want = orc.scale_noise(orc.power_noise(iter([draws[j]]), shape, filt, normalized=False), 1.0, normalized=True)
        assert_close(got[j], want, what=f"power noise {shape} sample {j}")
    single_row = 2 * b * c * h * (w // 2 + 1) <= 256 * sb.ops.philox_policy(2 * b * c * h * (w // 2 + 1))[0]
    assert launches == (4 if single_row and in_kernel else 6), launches  # (Philox fill +) FFT + scale_noise per sample


def test_noise_lookahead_equals_draw_by_draw(sb, cuda, monkeypatch):
    """Look-ahead batches (several power-noise samples per Philox / FFT launch) change nothing but the launch count:
    same samples, same generator advance -- also when the model itself consumes random numbers between the draws."""
    sigmas = torch.cat((torch.linspace(14.6, 0.5, 5), torch.zeros(1))).to(cuda)
    torch.manual_seed(9)
    x0 = (torch.randn(2, 4, 3, 18, 20) * 14.6).to(cuda)

    def plain(x, sigma, **_kw):
        return x * 0.9

    def noisy(x, sigma, **_kw):
        return x * 0.9 + torch.randn(3, device=x.device).sum() * 0.0

    def run(model):
        torch.manual_seed(77)
        launches = sb.ops.LAUNCH_COUNT
        out = sb.samplers.SonarDPMPPSDE.sampler(
            model, x0.clone(), sigmas, extra_args={"seed": 0}, disable=True, sonar_params={"custom_noise": video_chain(sb, alpha=1.0)},
        )
        return out, torch.cuda.default_generators[0].get_offset(), sb.ops.LAUNCH_COUNT - launches

    for model in (plain, noisy):
        ahead, off_a, launches_a = run(model)
        monkeypatch.setattr(sb.spectral_noise, "LOOKAHEAD_BYTES", 0)
        single, off_b, launches_b = run(model)
        monkeypatch.undo()
        assert off_a == off_b
        assert_close(ahead, single, what=f"look-ahead vs draw by draw ({model.__name__})", rtol=1e-6, atol=1e-6)
        if model is plain:
            assert launches_a < launches_b  # 8 draws: 1 + 1 launches instead of 8 + 8


@pytest.mark.parametrize("chunk", [1, 3])
def test_noise_pipeline_equals_batched(sb, cuda, monkeypatch, chunk):
    """Pipelined production (next sample made on a second stream beside the fused step, co-scheduling grid limits,
    ping-pong buffers) hands out the very same samples as the batched look-ahead: identical result bits and generator
    advance -- also when the model draws random numbers in between (every speculative sample is then dropped)."""
    sigmas = torch.cat((torch.linspace(14.6, 0.5, 6), torch.zeros(1))).to(cuda)
    torch.manual_seed(10)
    x0 = (torch.randn(2, 4, 3, 18, 20) * 14.6).to(cuda)

    def plain(x, sigma, **_kw):
        return x * 0.9

    def noisy(x, sigma, **_kw):
        return x * 0.9 + torch.randn(3, device=x.device).sum() * 0.0

    def run(model, pipelined):
        monkeypatch.setattr(sb.samplers, "NOISE_PIPELINE", pipelined)
        monkeypatch.setattr(sb.samplers, "PIPELINE_MIN_NUMEL", 0)
        monkeypatch.setattr(sb.samplers, "NOISE_PIPELINE_CHUNK", chunk)
        torch.manual_seed(78)
        out = sb.samplers.SonarDPMPPSDE.sampler(
            model, x0.clone(), sigmas, extra_args={"seed": 0}, disable=True, sonar_params={"custom_noise": video_chain(sb, alpha=1.0)},
        )
        torch.cuda.synchronize()
        return out, torch.cuda.default_generators[0].get_offset()

    for model in (plain, noisy):
        piped, off_a = run(model, True)
        batched, off_b = run(model, False)
        assert off_a == off_b
        assert torch.equal(piped, batched), f"pipelined vs batched ({model.__name__})"
    assert sb.ops._GRID_LIMIT == 0  # noqa: SLF001  (the co-scheduling hint never outlives a launch)
