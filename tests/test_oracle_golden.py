"""Pins the CPU oracle against the reference: every golden fixture (recorded by running the
unmodified reference, tests/golden/make_golden.py) must be reproduced by the oracle restatement
from the same recorded draws. Runs without a GPU."""
from __future__ import annotations

import hashlib

import numpy as np
import pytest
import torch

from helpers import NOISE_TYPE_NAMES, assert_close, oracle_noise_type, sampler_oracle_run, stub_model
from oracle import sonar_oracle as orc


def test_philox_known_answers():
    """Random123 kat_vectors for philox4x32, 10 rounds."""
    kats = [
        ([0, 0, 0, 0], [0, 0], [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]),
        ([0xFFFFFFFF] * 4, [0xFFFFFFFF] * 2, [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]),
        (
            [0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344],
            [0xA4093822, 0x299F31D0],
            [0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1],
        ),
    ]
    for ctr, key, want in kats:
        got = orc.philox4x32_10(np.array(ctr, dtype=np.uint32), np.array(key, dtype=np.uint32))
        assert [int(v) for v in got] == want


def test_aten_policy_matches_survey():
    # grid = min(SMs * (2048/256), ceil(n/256)), offset = ((n-1)/(256*grid*4)+1)*4
    assert orc.aten_policy(16384) == (64, 4)
    assert orc.aten_policy(524288) == (1184, 4)
    assert orc.aten_policy(60_825_600) == (1184, ((60_825_600 - 1) // (256 * 1184 * 4) + 1) * 4)
    assert orc.aten_policy(0) == (0, 0)


def test_c1_golden_matches_survey_probe(golden):
    """SURVEY.md section 4.4 recorded the same vector independently."""
    out = golden("power_noise")["c1_pink"]["out"]
    assert out.shape == (1, 4, 64, 64) and out.device.type == "cpu"
    first = out.flatten()[:6].tolist()
    want = [-0.04554218, 0.66203880, 0.07913631, 0.10460906, 0.53216058, -0.68568218]
    assert np.allclose(first, want, atol=2e-7)
    assert hashlib.sha256(out.numpy().tobytes()).hexdigest()[:16] == "d9c84f5041958d51"


@pytest.mark.parametrize("name", NOISE_TYPE_NAMES)
def test_noise_types(golden, name):
    case = golden("noise_types")[name]
    assert_close(oracle_noise_type(name, case), case["out"], what=name)


@pytest.mark.parametrize("name", ["white_33x40", "band_18x20", "rot_stretch_26x38", "odd_15x21"])
def test_power_filter_and_noise(golden, name):
    case = golden("power_noise")[name]
    filt = orc.power_filter(case["shape"], **case["params"])
    assert_close(filt, case["filter"], what=f"{name} filter")
    out = orc.power_noise(iter(case["draws"]), case["shape"], filt, normalized=False)
    out = orc.scale_noise(out, 1.0, normalized=True)  # chain-level normalisation
    assert_close(out, case["out"], what=name)


def test_power_noise_c1(golden):
    case = golden("power_noise")["c1_pink"]
    filt = orc.power_filter(case["shape"], alpha=1.0)
    out = orc.power_noise(iter(case["draws"]), case["shape"], filt, normalized=False)
    assert_close(orc.scale_noise(out, 1.0, normalized=True), case["out"], what="c1")


def test_power_filter_noise_rfft_front_end(golden):
    case = golden("power_noise")["filter_noise_24x20"]
    filt = orc.power_filter(case["shape"], alpha=1.0)
    out = orc.power_noise(iter(case["draws"]), case["shape"], filt, normalized=False, spectral_input=False)
    assert_close(orc.scale_noise(out, 1.0, normalized=True), case["out"], what="filter_noise")


def test_power_noise_video_5d(golden):
    case = golden("power_noise")["video_5d_18x20"]
    b, c, f, h, w = case["shape"]
    filt = orc.power_filter((b, c * f, h, w), alpha=1.0)
    out = orc.power_noise(iter(case["draws"]), (b, c * f, h, w), filt, normalized=False)
    out = out.reshape(case["shape"])  # CustomNoiseParametersNoise un-folds, then the chain normalises
    assert_close(orc.scale_noise(out, 1.0, normalized=True), case["out"], what="video")


def _gaussian(it):
    return next(it).clone()


def test_graph_chain_two(golden):
    case = golden("noise_graph")["chain_two"]
    it = iter(case["draws"])
    total = _gaussian(it).mul_(0.6)
    total.add_(orc.uniform_noise(it).mul_(-0.4))
    assert_close(orc.scale_noise(total, 1.0, normalized=True), case["out"], what="chain_two")


def test_graph_chain_rescaled(golden):
    case = golden("noise_graph")["chain_rescaled"]
    it = iter(case["draws"])
    total = _gaussian(it).mul_(0.6 / 0.5)
    total.add_(orc.uniform_noise(it).mul_(-0.4 / 0.5))
    assert_close(orc.scale_noise(total, 2.0, normalized=False), case["out"], what="chain_rescaled")


def _pyramid_from(it_dev, it_host, shape):
    sizes, (h, w) = [], shape[-2:]
    for i in range(10):
        r = next(it_host).item() * 2 + 2
        w, h = max(1, int(w / (r**i))), max(1, int(h / (r**i)))
        sizes.append((h, w))
        if w == 1 or h == 1:
            break
    return orc.pyramid_noise(it_dev, shape, sizes)


def test_graph_c3_scheduled(golden):
    g = golden("noise_graph")
    case = g["c3_scheduled_in_range"]
    shape = tuple(case["shape"])
    host = iter([d for d in case["draws"] if d.ndim == 1])
    dev = iter([d for d in case["draws"] if d.ndim != 1])
    n1 = _pyramid_from(dev, host, shape)
    n2 = orc.perlin_noise(dev, shape)
    out = torch.lerp(n1, n2, torch.full((1,), 0.5))
    assert_close(orc.scale_noise(out, 1.0, normalized=True), case["out"], what="c3 in range")
    case = g["c3_scheduled_fallback"]
    out = orc.scale_noise(case["draws"][0].clone(), 1.0, normalized=True)
    assert_close(out, case["out"], what="c3 fallback")


def test_graph_composite(golden):
    case = golden("noise_graph")["composite"]
    it = iter(case["draws"])
    shape = tuple(case["shape"])
    mask = torch.nn.functional.interpolate(case["mask"].reshape(-1, 1, 8, 8), size=shape[-2:], mode="bilinear")
    mask = mask.repeat(shape[0], 1, 1, 1)
    dst = orc.scale_noise(_gaussian(it))  # children inherit normalized=False from the chain -> no-op below
    dst = case["draws"][0].clone().mul_(1 - mask)
    src = orc.uniform_noise(iter(case["draws"][1:])).mul_(mask)
    assert_close(orc.scale_noise(dst.add_(src), 1.0, normalized=True), case["out"], what="composite")


def test_graph_blended_mask(golden):
    case = golden("noise_graph")["blended_mask"]
    it = iter(case["draws"])
    n1, n2, nm = _gaussian(it), orc.uniform_noise(it), _gaussian(it)
    t = (orc.normalize_to_scale(nm, 0.0, 1.0) + 0.25).clamp_(0.0, 1.0)
    assert_close(orc.scale_noise(torch.lerp(n1, n2, t), 1.0, normalized=True), case["out"], what="blended_mask")


CASES = None


def _sampler_cases(golden):
    return golden("samplers")


@pytest.mark.parametrize("sname", ["euler", "euler_ancestral", "euler_ancestral_eta", "dpmpp_sde"])
def test_samplers(golden, sname):
    g = _sampler_cases(golden)
    ran = 0
    for key, case in g["cases"].items():
        kind, variant = key.split("/")
        if kind != sname:
            continue
        steps = sampler_oracle_run(kind, case, g["x0"], g["sigmas"], stub_model)
        assert_close(steps, case["steps"], what=key, rtol=0, atol=0)
        assert_close(steps[-1], case["out"], what=key + " final", rtol=0, atol=0)
        ran += 1
    assert ran >= 2


@pytest.mark.parametrize("sname", ["euler", "euler_ancestral", "dpmpp_sde"])
def test_guidance(golden, sname):
    """Reference-latent guidance (py/sonar.py:323-411): LINEAR / EULER, windows, blends, broadcast latent."""
    g = golden("guidance")
    ran = 0
    for key, case in g["cases"].items():
        kind, variant = key.split("/")
        if kind != sname:
            continue
        steps = sampler_oracle_run(kind, case, g["x0"], g["sigmas"], stub_model)
        assert_close(steps, case["steps"], what=key, rtol=0, atol=0)
        ran += 1
    assert ran == 5


def test_expand_scales(golden):
    for spec, want in golden("host_logic")["expand"]:
        if any(isinstance(v, str) for v in (spec if isinstance(spec, list) else [])):
            continue  # "fill" handling lives in the product host code (tests/test_host_logic.py)
        got = orc.expand_scales(4, spec)
        want4 = list(want) + [(1.0, 1.0, 1.0)] * (4 - len(want))
        assert [tuple(g) for g in got] == [tuple(w) for w in want4]
