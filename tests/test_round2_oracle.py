"""Pins the oracle against the round-2 fixtures recorded from the unmodified reference
(tests/golden/make_golden.py::gen_round2): north-star config 5 as specified (frames_to_channels power noise as the
custom noise of sonar_dpmpp_sde), non-identity ChannelMixer, GuidedNoise. CPU only."""
from __future__ import annotations

import pytest
import torch

from helpers import assert_close, stub_model
from oracle import sonar_oracle as orc

POWER_DEFAULTS = {"alpha": 0.0, "max_freq": 0.7071, "min_freq": 0.0, "stretch": 1.0, "rotate": 0.0, "pnorm": 2.0, "mix": 1.0}


def filter_kwargs(params: dict) -> dict:
    return {k: v for k, v in (POWER_DEFAULTS | params).items() if k in POWER_DEFAULTS}


def c5_oracle_run(case: dict, x0: torch.Tensor, sigmas: torch.Tensor) -> torch.Tensor:
    """sonar_dpmpp_sde with custom_noise = CustomNoiseParametersNoise(frames_to_channels) o PowerNoise(alpha=1):
    two power-noise samples per step, folded to (B, C*F, H, W), un-folded, normalised once at the outer chain."""
    b, c, f, h, w = x0.shape
    folded = (b, c * f, h, w)
    filt = orc.power_filter(folded, alpha=1.0)
    draws = iter(case["draws"])

    def noise():
        raw = orc.power_noise(draws, folded, filt, normalized=False)
        return orc.scale_noise(raw.reshape(x0.shape), 1.0, normalized=True)

    o = orc.SonarOracle(mode=case["params"].get("momentum_mode", "new"))
    x, steps = x0.clone(), []
    for i in range(len(sigmas) - 1):
        sigma, sigma_next = sigmas[i], sigmas[i + 1]
        den = stub_model(x, sigma)
        if sigma_next == 0:
            x = o.dpmpp_sde(i, x, den, sigma, sigma_next, stub_model, None, None)
        else:
            n1, n2 = noise(), noise()
            x = o.dpmpp_sde(i, x, den, sigma, sigma_next, stub_model, n1, n2)
        steps.append(x.clone())
    assert next(draws, None) is None, "unused recorded draws"
    return torch.stack(steps)


@pytest.mark.parametrize("variant", ["default", "classic"])
def test_c5_job_oracle(golden, variant):
    g = golden("round2")["c5"]
    steps = c5_oracle_run(g[variant], g["x0"], g["sigmas"])
    assert_close(steps, g[variant]["steps"], what=f"c5 {variant}", rtol=0, atol=0)
    assert_close(steps[-1], g[variant]["out"], what=f"c5 {variant} final")


@pytest.mark.parametrize("name", ["c4_common", "c4_corr", "c3_neg", "c16_short_corr"])
def test_channel_mixer_oracle(golden, name):
    case = golden("round2")["mixer"][name]
    shape, params = case["shape"], case["params"]
    mixer = orc.channel_mixer(shape[1], params["common_mode"], params.get("channel_correlation", "1, 1, 1, 1, 1, 1"))
    assert torch.equal(mixer, case["mixer"]), "mixer matrix"
    assert not torch.equal(mixer, torch.eye(shape[1])), "fixture must exercise a non-identity mixer"
    filt = orc.power_filter(shape, **filter_kwargs(params))
    out = orc.power_noise(iter(case["draws"]), shape, filt, normalized=False, mixer=mixer)
    assert_close(orc.scale_noise(out, 1.0, normalized=True), case["out"], what=name)


def test_channel_mixer_video_oracle(golden):
    case = golden("round2")["mixer"]["video_c132"]
    b, c, f, h, w = case["shape"]
    folded, params = (b, c * f, h, w), case["params"]
    mixer = orc.channel_mixer(c * f, params["common_mode"], params["channel_correlation"])
    out = orc.power_noise(iter(case["draws"]), folded, orc.power_filter(folded, alpha=1.0), normalized=False, mixer=mixer)
    assert_close(orc.scale_noise(out.reshape(case["shape"]), 1.0, normalized=True), case["out"], what="video_c132")


def test_channel_mixer_filter_noise_oracle(golden):
    case = golden("round2")["mixer"]["filter_noise_c4"]
    shape = case["shape"]
    mixer = orc.channel_mixer(4, 0.4, "1,-0.5,0.5,1,1,0.2")
    out = orc.power_noise(iter(case["draws"]), shape, orc.power_filter(shape, alpha=1.0), normalized=False,
                          spectral_input=False, mixer=mixer)  # fmt: skip
    assert_close(orc.scale_noise(out, 1.0, normalized=True), case["out"], what="filter_noise_c4")


GUIDED = ["linear", "linear_one_ref", "linear_no_noise", "linear_resized_ref", "euler", "euler_no_noise",
          "euler_equal_sigmas", "linear_unnormalized"]  # fmt: skip


def guided_oracle(g: dict, name: str) -> torch.Tensor:
    case = g[name]
    cfg = case["config"]
    s, sn = case["sigmas"]
    # item-level flags default to the chain's normalized=False for its children (py/noise.py:181)
    nn = cfg.get("normalize_noise")
    nr = cfg.get("normalize_result")
    out = orc.guided_noise(
        iter(case["draws"]), g["x"], case["ref"], method=cfg["method"], guidance_factor=cfg["guidance_factor"],
        factor=1.0, has_noise=cfg["noise"], sigma=torch.tensor(s), sigma_next=torch.tensor(sn),
        normalize_noise=False if nn is None else nn, normalize_result=False if nr is None else nr,
    )  # fmt: skip
    # the chain multiplies by nothing (children carry their factor) and normalises once with factor = sum |factor_i|
    return orc.scale_noise(out.mul_(cfg.get("factor", 1.0)) if cfg.get("factor", 1.0) != 1 else out,
                           abs(cfg.get("factor", 1.0)), normalized=True)  # fmt: skip


@pytest.mark.parametrize("name", GUIDED)
def test_guided_noise_oracle(golden, name):
    g = golden("round2")["guided"]
    assert_close(guided_oracle(g, name), g[name]["out"], what=name, rtol=0, atol=0)
