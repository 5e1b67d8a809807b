"""Drop-in boundary: the node surface matches the reference's (schemas dumped from its INPUT_TYPES())."""
from __future__ import annotations

import json

import pytest
import torch
import yaml

from conftest import GOLDEN
from helpers import assert_close

SCHEMAS = json.loads((GOLDEN / "node_schemas.json").read_text())


def _normalise(input_types: dict) -> dict:
    out = {}
    for section, fields in input_types.items():
        out[section] = {}
        for key, spec in fields.items():
            opts = {k: v for k, v in (spec[1] if len(spec) > 1 else {}).items() if k != "tooltip"}
            kind = list(spec[0]) if isinstance(spec[0], (tuple, list)) else str(spec[0])
            out[section][key] = [kind, json.loads(json.dumps(opts))]
    return out


@pytest.mark.parametrize("name", sorted(SCHEMAS))
def test_node_schema_matches_reference(sb, name):
    assert sb.HAVE_COMFY, "tests/shim/comfy must be importable"
    cls = sb.NODE_CLASS_MAPPINGS[name]
    want = SCHEMAS[name]
    assert list(cls.RETURN_TYPES) == want["return_types"]
    assert cls.FUNCTION == want["function"] and cls.CATEGORY == want["category"]
    assert callable(getattr(cls, cls.FUNCTION))
    got = _normalise(cls.INPUT_TYPES())
    ref = want["input_types"]
    if name == "SonarWaveletCFG":
        # the default YAML is documentation text; what must agree is what it parses to
        mine = got["required"]["yaml_parameters"][1].pop("default")
        theirs = ref["required"]["yaml_parameters"][1].pop("default")
        assert yaml.safe_load(mine) == yaml.safe_load(theirs)
    if name == "SonarWaveletFilteredNoise":
        # the placeholder is a commented option template; what must agree is the options it documents
        mine = got["optional"]["yaml_parameters"][1].pop("placeholder")
        theirs = ref["optional"]["yaml_parameters"][1].pop("placeholder")
        assert yaml.safe_load(mine) == yaml.safe_load(theirs)
    for section in ("required", "optional"):
        assert list(got.get(section, {})) == list(ref.get(section, {})), f"{name}.{section} field order"
        assert got.get(section, {}) == ref.get(section, {}), f"{name}.{section}"


def test_samplers_registered_with_comfy(sb):
    import comfy.samplers as cs

    for name in ("sonar_euler", "sonar_euler_ancestral", "sonar_dpmpp_sde"):
        assert name in cs.KSampler.SAMPLERS
        assert callable(getattr(cs.k_diffusion_sampling, f"sample_{name}"))


def test_chain_building_nodes(sb):
    n = sb.nodes
    (chain,) = n.SonarCustomNoiseNode().go(factor=0.5, rescale=0.0, noise_type="pyramid")
    (chain,) = n.SonarAdvanced1fNoiseNode().go(
        factor=1.5, rescale=1.0, alpha=0.25, k=1.0, vertical_factor=1.0, horizontal_factor=1.0, use_sqrt=True,
        sonar_custom_noise_opt=chain,
    )  # fmt: skip
    assert [type(i).__name__ for i in chain.items] == ["CustomNoiseItem", "Advanced1fNoise"]
    assert chain.factor == pytest.approx(1.0) and chain.items[1].hfac == 1.0
    (empty,) = n.SonarCustomNoiseNode().go(factor=0.0, rescale=0.0, noise_type="gaussian")
    assert empty.items == []
    (pyr,) = n.SonarAdvancedPyramidNoiseNode().go(
        factor=1.0, rescale=0.0, variant="pyramid", iterations=-1, discount=0.0, upscale_mode="default",
    )  # fmt: skip
    assert pyr.items[0].iterations is None and pyr.items[0].discount is None and pyr.items[0].upscale_mode is None
    # the reference node swaps normalize_src / normalize_dst (nodes/noise_filters.py:246-247)
    (comp,) = n.SonarCompositeNoiseNode().go(
        factor=1.0, sonar_custom_noise_dst=chain, sonar_custom_noise_src=chain, normalize_src="forced",
        normalize_dst="disabled", normalize_result="default", mask=torch.zeros(1, 4, 4),
    )  # fmt: skip
    item = comp.items[0]
    assert item.normalize_dst is True and item.normalize_src is False and item.normalize_result is None
    with pytest.raises(ValueError):
        n.SonarBlendedNoiseNode().go(factor=1.0, rescale=0.0, normalize="default", noise_2_percent=0.5, blend_mode="nope")
    (filt,) = n.SonarPowerFilterNode.go(alpha=1.0, blur=0.2)
    assert filt.alpha == 1.0 and filt.rel_bw == 0.2
    sampler = n.SamplerNodeSonarEulerAncestral.get_sampler(
        momentum=0.9, momentum_hist=0.7, momentum_init="ZERO", direction=1.0, rand_init_noise_type="gaussian",
        noise_type="gaussian", eta=0.8, s_noise=1.0,
    )[0]  # fmt: skip
    assert sampler.sampler_function == sb.samplers.SonarEulerAncestral.sampler
    assert sampler.extra_options["eta"] == 0.8 and sampler.extra_options["sonar_config"].momentum == 0.9


@pytest.mark.gpu
def test_config_c1_noise_object_golden(sb, golden):
    """BASELINE.json config C1 end to end through the node surface: SonarPowerNoise(alpha=1) ->
    'SONAR_CUSTOM_NOISE to NOISE' -> generate_noise on a CPU latent -> CPU noise."""
    case = golden("power_noise")["c1_pink"]
    n = sb.nodes
    kwargs = {k: (v[1]["default"] if len(v) > 1 and "default" in v[1] else None) for k, v in n.SonarPowerNoiseNode.INPUT_TYPES()["required"].items()}
    kwargs |= {"alpha": 1.0}
    (chain,) = n.SonarPowerNoiseNode().go(**kwargs)
    (noise_obj,) = n.SonarToComfyNOISENode.go(custom_noise=chain, seed=0)
    assert noise_obj.seed == 0
    with sb.rng.injected(case["draws"]):
        out = noise_obj.generate_noise({"samples": torch.zeros(1, 4, 64, 64)})
    assert out.device.type == "cpu" and out.dtype == torch.float32
    assert_close(out, case["out"], what="C1 NOISE object")
    zero = n.CustomNOISE(chain, 0, multiplier=0.0).generate_noise({"samples": torch.zeros(1, 4, 8, 8)})
    assert zero.abs().max() == 0
    # batch_index: one draw per index with seed + idx
    out = n.CustomNOISE(chain, 3).generate_noise({"samples": torch.zeros(2, 4, 16, 16), "batch_index": [1, 1, 0]})
    assert out.shape == (3, 4, 16, 16) and torch.equal(out[0], out[1]) and not torch.equal(out[0], out[2])


@pytest.mark.gpu
def test_noisy_latent_like_golden(sb, golden):
    """NoisyLatentLike end to end against the reference node (fixture recorded by make_golden.py gen_noisy_latent):
    built-in and custom noise, repeat_batch, add_to_latent, the sigma / model multiplier in both max_denoise branches."""
    fx = golden("noisy_latent")
    n = sb.nodes

    class _MS:
        sigma_max = torch.tensor(fx["sigma_max"])

    class _LF:
        scale_factor = fx["scale_factor"]

    class _Inner:
        model_sampling, latent_format = _MS(), _LF()

    class _Model:
        model = _Inner()

    kwargs = {k: (v[1]["default"] if len(v) > 1 and "default" in v[1] else None) for k, v in n.SonarPowerNoiseNode.INPUT_TYPES()["required"].items()}
    (chain,) = n.SonarPowerNoiseNode().go(**(kwargs | {"alpha": 1.0}))
    for name, case in fx["cases"].items():
        with sb.rng.injected(case["draws"]) as left:
            (res,) = n.NoisyLatentLikeNode.go(
                latent={"samples": fx["latent"].clone()}, cpu_noise=True, custom_noise_opt=chain if case["custom"] else None,
                mul_by_sigmas_opt=case["sigmas"], model_opt=_Model() if case["sigmas"] is not None else None, **case["kwargs"],
            )
            assert not left, f"{name}: recorded draws not consumed"
        out = res["samples"]
        assert out.device.type == "cpu" and out.dtype == torch.float32
        # 1e-5 relative to the tensor's scale: the sigma cases multiply unit noise by ~110 before adding the latent
        scale = max(1.0, float(case["out"].abs().max()))
        assert_close(out / scale, case["out"] / scale, what=f"NoisyLatentLike {name}")


@pytest.mark.gpu
def test_noisy_latent_like_and_config_override(sb, cuda):
    n = sb.nodes
    latent = {"samples": torch.zeros(2, 4, 16, 16)}
    before = torch.random.get_rng_state()
    (res,) = n.NoisyLatentLikeNode.go(noise_type="pyramid", seed=5, latent=latent, multiplier=2.0, repeat_batch=2)
    assert torch.equal(torch.random.get_rng_state(), before)
    out = res["samples"]
    assert out.shape == (4, 4, 16, 16) and out.device.type == "cpu"
    assert float(out.std()) == pytest.approx(2.0, rel=0.05)
    base = n.SamplerNodeSonarEuler.get_sampler(
        momentum=0.95, momentum_hist=0.75, momentum_init="ZERO", direction=1.0, rand_init_noise_type="gaussian",
    )[0]  # fmt: skip
    (wrapped,) = n.SamplerNodeConfigOverride().get_sampler(
        sampler=base, eta=1.0, s_noise=1.0, s_churn=0.0, r=0.5, sde_solver="midpoint", noise_type="perlin",
        yaml_parameters="sonar_params:\n  momentum: 1.0\n",
    )  # fmt: skip
    sigmas = torch.tensor([5.0, 2.0, 0.0])
    x = torch.randn(1, 4, 8, 8, device=cuda) * 5
    got = wrapped.sampler_function(lambda x, s, **k: x * 0.5, x, sigmas.to(cuda), extra_args={}, disable=True, **wrapped.extra_options)
    want = x.clone()
    for i in range(2):
        want = want + (want - want * 0.5) / sigmas[i] * (sigmas[i + 1] - sigmas[i])
    assert_close(got, want, what="override -> momentum 1 euler")
