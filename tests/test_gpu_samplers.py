"""Parity of the fused Sonar step / samplers with the reference (golden fixtures, oracle) and the
identities the reference code implies."""
from __future__ import annotations

import pytest
import torch

from helpers import assert_close, sampler_oracle_run, stub_model
from oracle import sonar_oracle as orc

pytestmark = pytest.mark.gpu

TOL = {"rtol": 1e-5, "atol": 1e-5}  # the north-star fp32 tolerance


def _sampler(sb, kind):
    s = sb.samplers
    return {
        "euler": s.SonarEuler.sampler,
        "euler_ancestral": s.SonarEulerAncestral.sampler,
        "euler_ancestral_eta": s.SonarEulerAncestral.sampler,
        "dpmpp_sde": s.SonarDPMPPSDE.sampler,
    }[kind]


def _cases(golden, kind):
    g = golden("samplers")
    return g, [(k, c) for k, c in g["cases"].items() if k.split("/")[0] == kind]


@pytest.mark.parametrize("kind", ["euler", "euler_ancestral", "euler_ancestral_eta", "dpmpp_sde"])
def test_samplers_golden(sb, cuda, golden, kind):
    g, cases = _cases(golden, kind)
    assert len(cases) >= 2
    for key, case in cases:
        steps = []
        with sb.rng.injected(case["draws"]) as left:
            out = _sampler(sb, kind)(
                stub_model, g["x0"].to(cuda), g["sigmas"].to(cuda), extra_args={"seed": 0}, disable=True,
                sonar_params=dict(case["params"]), callback=lambda d: steps.append(d["x"].clone()),
                **case["sampler_kwargs"],
            )  # fmt: skip
            assert not left, f"{key}: unused recorded draws"
        assert_close(torch.stack(steps), case["steps"], what=key, **TOL)
        assert_close(out, case["out"], what=key + " final", **TOL)


def test_callback_payload_and_sigmas_on_cpu(sb, cuda, golden):
    g = golden("samplers")
    seen = []
    out = sb.samplers.SonarEuler.sampler(
        stub_model, g["x0"].to(cuda), g["sigmas"], extra_args={"seed": 0}, disable=True, callback=seen.append,
    )
    assert out.device.type == "cuda" and len(seen) == len(g["sigmas"]) - 1
    assert set(seen[0]) == {"x", "i", "sigma", "sigma_hat", "denoised"}
    assert_close(out, g["cases"]["euler/default"]["out"], what="cpu sigmas", **TOL)


def test_momentum_one_is_plain_euler(sb, cuda):
    """momentum == 1 short-circuits to the plain Euler update (py/sonar.py:250, :301)."""
    torch.manual_seed(0)
    sigmas = torch.cat((torch.linspace(10.0, 0.1, 9), torch.zeros(1)))
    x = torch.randn(2, 4, 32, 32) * sigmas[0]
    got = sb.samplers.SonarEuler.sampler(
        stub_model, x.to(cuda), sigmas.to(cuda), extra_args={}, disable=True, sonar_params={"momentum": 1.0},
    )
    want = x.clone()
    for i in range(len(sigmas) - 1):
        d = (want - stub_model(want, sigmas[i])) / sigmas[i]
        want = want + d * (sigmas[i + 1] - sigmas[i])
    assert_close(got, want, what="plain euler", **TOL)


def test_fused_philox_noise_equals_tensor_noise(sb, cuda):
    """The in-register Philox noise (look-ahead statistics batch + one fused launch per step) equals
    drawing torch.randn on the GPU, normalising it with scale_noise and adding it -- and advances
    torch's generator the same."""
    torch.manual_seed(0)
    sigmas = torch.cat((torch.linspace(14.6, 0.03, 30), torch.zeros(1)))
    x0 = (torch.randn(8, 4, 128, 128) * sigmas[0]).to(cuda)

    def run(force_tensor, model=stub_model):
        torch.manual_seed(77)
        ns = None
        if force_tensor:
            def ns(_s, _sn):
                n = torch.randn(x0.shape, device=cuda)
                return sb.hostutil.scale_noise(n, 1.0, normalized=True)
        launches = sb.ops.LAUNCH_COUNT
        out = sb.samplers.SonarEulerAncestral.sampler(
            model, x0.clone(), sigmas.to(cuda), extra_args={"seed": 0}, disable=True, noise_sampler=ns,
        )
        return out, torch.cuda.default_generators[0].get_offset(), sb.ops.LAUNCH_COUNT - launches

    plain, off_b, _ = run(True)
    fused, off_a, launches = run(False)
    assert off_a == off_b
    assert launches == 30 + 2  # one fused launch per step + ONE statistics launch (+ its decisions) for all 29 draws
    assert_close(fused, plain, what="fused vs tensor noise", rtol=1e-6, atol=1e-5)

    # a denoiser that consumes random numbers itself invalidates the predicted generator offsets:
    # the sampler must notice, re-plan, and still produce exactly the tensor-noise result
    def noisy_model(x, sigma, **_kw):
        return x * 0.9 + torch.randn(3, device=x.device).sum() * 0.0

    plain, off_b, _ = run(True, noisy_model)
    fused, off_a, launches = run(False, noisy_model)
    assert off_a == off_b
    assert launches > 30 + 2  # statistics were re-planned
    assert_close(fused, plain, what="fused vs tensor noise, generator shared with the model", rtol=1e-6, atol=1e-5)


def test_fused_noise_large_tensor_path(sb, cuda):
    """Config C5 per-GPU shard shape, DPM++ SDE (two noise draws per step, several Philox calls per
    thread): same values and generator advance as torch.randn + scale_noise."""
    sigmas = torch.tensor([14.6, 6.0, 1.5, 0.0])
    torch.manual_seed(0)
    x0 = (torch.randn(1, 16, 33, 90, 160) * sigmas[0]).to(cuda)

    def run(explicit):
        torch.manual_seed(5)
        ns = None
        if explicit:
            def ns(_s, _sn):
                return sb.hostutil.scale_noise(torch.randn(x0.shape, device=cuda), 1.0, normalized=True)
        out = sb.samplers.SonarDPMPPSDE.sampler(
            lambda x, s, **k: x * 0.9, x0.clone(), sigmas.to(cuda), extra_args={"seed": 0}, disable=True,
            sonar_params={"noise_type": "gaussian"}, noise_sampler=ns,
        )
        return out, torch.cuda.default_generators[0].get_offset()

    fused, off_a = run(False)
    plain, off_b = run(True)
    assert off_a == off_b
    assert_close(fused, plain, what="large fused vs tensor noise", rtol=1e-6, atol=1e-5)


def test_c2_full_size_vs_oracle(sb, cuda):
    """BASELINE.json config C2: sonar_euler_ancestral, SDXL latents 8x4x128x128, 30 steps."""
    torch.manual_seed(1)
    sigmas = torch.cat((torch.linspace(14.6, 0.03, 30), torch.zeros(1)))
    x0 = torch.randn(8, 4, 128, 128) * sigmas[0]
    draws = [torch.randn(x0.shape) for _ in range(29)]
    case = {"params": {}, "sampler_kwargs": {"eta": 1.0, "s_noise": 1.0}, "draws": draws}
    want = sampler_oracle_run("euler_ancestral", case, x0, sigmas, lambda x, s: x * 0.9)
    with sb.rng.injected(draws):
        got = sb.samplers.SonarEulerAncestral.sampler(
            lambda x, s, **k: x * 0.9, x0.to(cuda), sigmas.to(cuda), extra_args={"seed": 0}, disable=True,
        )
    assert_close(got, want[-1], what="C2", rtol=1e-5, atol=1e-5)


def test_c5_shard_dpmpp_vs_oracle(sb, cuda):
    """Config C5 per-GPU shard (1x16x33x90x160 video latent), 3 DPM++ SDE steps, injected noise."""
    torch.manual_seed(2)
    sigmas = torch.tensor([14.6, 7.0, 2.0, 0.0])
    shape = (1, 16, 33, 90, 160)
    x0 = torch.randn(shape) * sigmas[0]
    draws = [torch.randn(shape) for _ in range(4)]
    case = {"params": {"noise_type": "gaussian"}, "sampler_kwargs": {"eta": 1.0, "s_noise": 1.0}, "draws": draws}
    want = sampler_oracle_run("dpmpp_sde", case, x0, sigmas, lambda x, s: x * 0.9)
    with sb.rng.injected(draws):
        got = sb.samplers.SonarDPMPPSDE.sampler(
            lambda x, s, **k: x * 0.9, x0.to(cuda), sigmas.to(cuda), extra_args={"seed": 0}, disable=True,
            sonar_params={"noise_type": "gaussian"},
        )
    assert_close(got, want[-1], what="C5 shard", rtol=1e-5, atol=1e-5)


def test_step_kernel_unaligned_and_odd_sizes(sb, cuda):
    """Scalar tail / unaligned-pointer paths of the fused step."""
    torch.manual_seed(3)
    for n in (1, 3, 1023, 4099):
        base = torch.randn(n + 1)
        x, den, hist, nz = (torch.randn(n + 1, device=cuda)[1:] for _ in range(4))
        o = orc.SonarOracle()
        o.hist = hist.cpu().clone()
        sigma, sigma_next = torch.tensor(5.0), torch.tensor(3.0)
        want = o.euler_ancestral(3, x.cpu(), den.cpu(), sigma, sigma_next, nz.cpu())
        s = sb.samplers.SonarBase(sb.samplers.SonarConfig())
        s.history_d = hist.clone()
        sd, su = sb.kdiff.get_ancestral_step(sigma, sigma_next)
        got = s.momentum_step(3, x.contiguous(), den.contiguous(), 5.0, float(sd), noise_tensor=nz.contiguous(), noise_scale=float(su))
        assert_close(got, want, what=f"n={n}")
        assert_close(s.history_d, o.hist, what=f"hist n={n}")
        _ = base


def test_config_errors_match_reference(sb, golden):
    for params, (exc_name, msg) in golden("host_logic")["config_errors"].items():
        with pytest.raises((ValueError, TypeError)) as info:
            sb.samplers.SonarBase.get_config(None, eval(params))  # noqa: S307 - fixture literal
        assert type(info.value).__name__ == exc_name
        assert str(info.value) == msg


def test_deferred_chain_normalisation_equals_explicit_scale_noise(sb, cuda):
    """Chain noise handed to the step un-normalised (scale_noise applied on load from the producer's
    statistics) == the same chain normalised by its own scale_noise pass first."""
    ng = sb.noise_graph
    chain = ng.CustomNoiseChain()
    chain.add(ng.CustomNoiseItem(1.0, noise_type="pyramid"))
    chain.add(ng.CustomNoiseItem(0.5, noise_type="perlin"))
    sigmas = torch.cat((torch.linspace(12.0, 0.2, 7), torch.zeros(1))).to(cuda)
    torch.manual_seed(11)
    x0 = (torch.randn(3, 4, 40, 48) * 12.0).to(cuda)

    def run(hide_deferred):
        torch.manual_seed(42)
        base = chain.make_noise_sampler(x0, sigmas[sigmas > 0].min().cpu(), sigmas.max().cpu(), seed=0)
        assert hasattr(base, "deferred")
        ns = (lambda s, sn: base(s, sn)) if hide_deferred else base
        return sb.samplers.SonarEulerAncestral.sampler(
            stub_model, x0.clone(), sigmas, extra_args={"seed": 0}, disable=True, noise_sampler=ns,
        )

    assert_close(run(False), run(True), what="deferred vs explicit", rtol=1e-6, atol=1e-5)


@pytest.mark.parametrize("kind", ["euler", "euler_ancestral", "dpmpp_sde"])
def test_guidance_golden(sb, cuda, golden, kind):
    """Reference-latent guidance (py/sonar.py:323-411) on the CUDA kernels vs the recorded reference:
    LINEAR (lerp / inject) and EULER, a step window, a batch-broadcast latent, DENOISED momentum mode."""
    g = golden("guidance")
    cases = [(k, c) for k, c in g["cases"].items() if k.split("/")[0] == kind]
    assert len(cases) == 5
    s = sb.samplers
    for key, case in cases:
        gk = case["guidance"]
        guidance = s.GuidanceConfig(
            guidance_type=s.GuidanceType[gk["guidance_type"]], factor=gk["factor"], start_step=gk["start_step"],
            end_step=gk["end_step"], latent=case["latent"].clone(),
        )  # fmt: skip
        steps = []
        with sb.rng.injected(case["draws"]) as left:
            out = _sampler(sb, kind)(
                stub_model, g["x0"].to(cuda), g["sigmas"].to(cuda), extra_args={"seed": 0}, disable=True,
                sonar_params=dict(case["params"]) | {"guidance": guidance}, callback=lambda d: steps.append(d["x"].clone()),
                **case["sampler_kwargs"],
            )  # fmt: skip
            assert not left, f"{key}: unused recorded draws"
        assert_close(torch.stack(steps), case["steps"], what=key, **TOL)
        assert_close(out, case["out"], what=key + " final", **TOL)


def test_guidance_kernels_vs_oracle(sb, cuda):
    """item_moments + guidance kernels vs the oracle on odd sizes (scalar tails, unaligned views)."""
    torch.manual_seed(12)
    for shape, ref_batch in (((3, 4, 9, 7), 3), ((2, 5, 16, 16), 1), ((1, 3, 33, 5), 1)):
        x, den = torch.randn(shape) * 3 + 0.5, torch.randn(shape) * 2 - 1
        ref = orc.prepare_ref_latent(torch.randn(ref_batch, *shape[1:]))
        xd, dd, rd = x.to(cuda), den.to(cuda), ref.to(cuda)
        sums = sb.ops.item_moments(xd)
        d64 = x.double().flatten(1)
        torch.testing.assert_close(sums.cpu(), torch.stack((d64.sum(1), (d64 * d64).sum(1)), dim=1), rtol=1e-9, atol=1e-7)
        for mode in ("lerp", "inject", "subtract_b"):
            want = orc.guidance_linear(x, ref, 0.15, blend=orc.BLENDING_MODES[mode])
            got = sb.ops.guidance(xd, rd, sums, kind=sb.ops.GUIDANCE_LINEAR, blend_mode=mode, factor=0.15)
            assert_close(got, want, what=f"linear {mode} {shape}")
        sigma, sigma_next = torch.tensor(7.0), torch.tensor(4.5)
        want = orc.guidance_euler(sigma, sigma_next, x, den, ref, 0.3)
        got = sb.ops.guidance(xd, rd, sb.ops.item_moments(dd), kind=sb.ops.GUIDANCE_EULER, sigma=7.0,
                              dt=float((sigma_next - sigma) * 0.3))
        assert_close(got, want, what=f"euler {shape}")


def test_half_precision_latents_are_sampled_in_fp32(sb, cuda, golden):
    """fp16 / bf16 latents (what ComfyUI hands over for half-precision models) run through the fp32 kernels and come
    back in their own dtype instead of raising."""
    g = golden("samplers")
    want = sb.samplers.SonarEuler.sampler(stub_model, g["x0"].to(cuda).half().float(), g["sigmas"].to(cuda), extra_args={}, disable=True)
    got = sb.samplers.SonarEuler.sampler(stub_model, g["x0"].to(cuda).half(), g["sigmas"].to(cuda), extra_args={}, disable=True)
    assert got.dtype == torch.float16
    assert torch.equal(got, want.half())
