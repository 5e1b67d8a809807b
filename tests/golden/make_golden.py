"""Generates the golden fixtures in tests/golden/*.pt by running the UNMODIFIED reference
(blepping/ComfyUI-sonar at /root/reference) on CPU with its random draws recorded.

    python tests/golden/make_golden.py            # needs /root/reference; writes tests/golden/*.pt

Each fixture holds the recorded base draws (in call order), the inputs and the reference output, so
that the oracle (tests/test_oracle_golden.py, CPU) and the CUDA path (tests/test_gpu_*.py) can be fed
the very same random tensors ("injected" parity). The reference ships no golden
vectors of its own; this script is how the oracle is pinned. `/root/reference` does not exist on the
GPU box, which is why the outputs are committed.
"""

from __future__ import annotations

import contextlib
import importlib.util
import sys
from pathlib import Path

import torch

HERE = Path(__file__).resolve().parent
REPO = HERE.parent.parent
REFERENCE = Path("/root/reference")


def load_reference():
    sys.path.insert(0, str(REPO / "tests" / "shim"))
    if "sonar_ref" in sys.modules:
        return sys.modules["sonar_ref"]
    spec = importlib.util.spec_from_file_location(
        "sonar_ref", REFERENCE / "__init__.py", submodule_search_locations=[str(REFERENCE)],
    )
    mod = importlib.util.module_from_spec(spec)
    sys.modules["sonar_ref"] = mod
    spec.loader.exec_module(mod)
    return mod


@contextlib.contextmanager
def record_draws():
    """Records every tensor produced by torch.randn / rand / normal / Tensor.uniform_."""
    draws: list[torch.Tensor] = []
    orig = {"randn": torch.randn, "rand": torch.rand, "normal": torch.normal, "uniform_": torch.Tensor.uniform_}

    def wrap(fn):
        def inner(*args, **kwargs):
            out = fn(*args, **kwargs)
            draws.append(out.detach().clone())
            return out

        return inner

    def uniform_(self, *args, **kwargs):
        out = orig["uniform_"](self, *args, **kwargs)
        draws.append(out.detach().clone())
        return out

    torch.randn, torch.rand, torch.normal = wrap(orig["randn"]), wrap(orig["rand"]), wrap(orig["normal"])
    torch.Tensor.uniform_ = uniform_
    # NoiseGenerator.rand_like binds `fun=torch.randn` as a keyword default at import time
    # (py/noise_generation.py:136): re-point that default at the recording wrapper as well.
    rand_like = sys.modules["sonar_ref.py.noise_generation"].NoiseGenerator.rand_like
    rand_like.__kwdefaults__["fun"] = torch.randn
    try:
        yield draws
    finally:
        torch.randn, torch.rand, torch.normal = orig["randn"], orig["rand"], orig["normal"]
        torch.Tensor.uniform_ = orig["uniform_"]
        rand_like.__kwdefaults__["fun"] = orig["randn"]


def stub_model(x, sigma, **_kwargs):
    """Deterministic non-linear denoiser stand-in (the UNet is not part of the hot path)."""
    return x * 0.9 - 0.05 * torch.tanh(x)


def save(name: str, payload: dict) -> None:
    path = HERE / f"{name}.pt"
    torch.save(payload, path)
    print(f"{name:38s} {path.stat().st_size / 1024:8.1f} KiB  draws={len(payload.get('draws', []))}")


# ---------------------------------------------------------------------------------------------
def gen_noise_types(ref) -> None:
    noise = ref.py.noise
    cases = {
        "gaussian": (2, 4, 16, 16),
        "uniform": (2, 4, 16, 16),
        "perlin": (2, 3, 16, 20),
        "pyramid": (2, 3, 32, 32),
        "pyramid_discount5": (1, 2, 24, 40),
        "pyramid_area": (1, 2, 24, 24),
        "pyramid_mix": (1, 2, 32, 32),
        "pyramid_old": (1, 2, 4, 4),
        "highres_pyramid": (1, 2, 8, 12),
        "onef_pinkish": (2, 3, 16, 24),
        "onef_greenish": (1, 2, 18, 20),
        "onef_pinkish_mix": (1, 2, 16, 16),
        "onef_pinkishgreenish": (1, 2, 16, 16),
        "green_test": (2, 2, 16, 16),
        "rainbow_mild": (1, 2, 16, 16),
        "white": (2, 4, 8, 8),
        "grey": (2, 4, 8, 8),
        "velvet": (2, 4, 8, 8),
        "violet": (2, 4, 8, 8),
        "wavelet": (2, 3, 32, 32),
        "wavelet_odd": (1, 2, 36, 52),
    }
    out = {}
    for i, (name, shape) in enumerate(cases.items()):
        torch.manual_seed(100 + i)
        x = torch.zeros(shape)
        with record_draws() as draws:
            ns = noise.get_noise_sampler(name.split("_odd")[0], x, None, None, seed=0, cpu=True, normalized=True)
            result = ns(None, None)
        out[name] = {"shape": shape, "draws": draws, "out": result.clone()}
    # 5-D latent through frames-to-channels generators
    torch.manual_seed(7)
    x5 = torch.zeros(1, 2, 3, 16, 16)
    with record_draws() as draws:
        result = noise.get_noise_sampler("pyramid", x5, None, None, seed=0, cpu=True, normalized=True)(None, None)
    out["pyramid_5d"] = {"shape": tuple(x5.shape), "draws": draws, "out": result.clone()}
    save("noise_types", out)


def gen_power_noise(ref) -> None:
    noise, misc, pn = ref.py.noise, ref.py.nodes.misc, ref.py.nodes.powernoise
    defaults = {
        "time_brownian": False, "alpha": 0.0, "max_freq": 0.7071, "min_freq": 0.0, "stretch": 1.0, "rotate": 0.0,
        "pnorm": 2.0, "mix": 1.0, "common_mode": 0.0, "channel_correlation": "1, 1, 1, 1, 1, 1",
    }  # fmt: skip

    def chain_with(**kw):
        chain = noise.CustomNoiseChain()
        chain.add(pn.PowerNoiseItem(1.0, **(defaults | kw)))
        return chain

    out = {}
    # config C1: pink (alpha=1) on an SD1.5 latent through the NOISE object, seed 0
    with record_draws() as draws:
        result = misc.CustomNOISE(chain_with(alpha=1.0), 0).generate_noise({"samples": torch.zeros(1, 4, 64, 64)})
    out["c1_pink"] = {"shape": (1, 4, 64, 64), "alpha": 1.0, "draws": draws, "out": result.clone()}
    variants = {
        "white_33x40": ((1, 2, 33, 40), {}),
        "band_18x20": ((2, 3, 18, 20), {"alpha": 0.5, "min_freq": 0.1, "max_freq": 0.4}),
        "rot_stretch_26x38": ((1, 2, 26, 38), {"alpha": 1.5, "rotate": 30.0, "stretch": 2.0, "pnorm": 1.5, "mix": 0.75}),
        "odd_15x21": ((1, 2, 15, 21), {"alpha": 1.0}),
    }
    for i, (name, (shape, kw)) in enumerate(variants.items()):
        torch.manual_seed(200 + i)
        x = torch.zeros(shape)
        with record_draws() as draws:
            ns = chain_with(**kw).make_noise_sampler(x, None, None, seed=0, cpu=True, normalized=True)
            result = ns(None, None)
        filt = pn.PowerNoiseItem(1.0, **(defaults | kw)).make_filter(shape)
        out[name] = {"shape": shape, "params": kw, "draws": draws, "filter": filt.clone(), "out": result.clone()}
    # PowerFilterNoiseItem: rfft2 -> filter -> irfft2 on spatial gaussian noise
    torch.manual_seed(210)
    x = torch.zeros(2, 3, 24, 20)
    inner = noise.CustomNoiseChain()
    inner.add(noise.CustomNoiseItem(1.0, noise_type="gaussian"))
    item = pn.PowerFilterNoiseItem(
        1.0, noise=inner, normalize_noise=None, normalize_result=None, power_filter=pn.PowerFilter(alpha=1.0),
        mix=1.0, common_mode=0.0, channel_correlation="1,1,1,1,1,1", time_brownian=True, filter_norm_factor=1.0,
    )  # fmt: skip
    chain = noise.CustomNoiseChain()
    chain.add(item)
    with record_draws() as draws:
        result = chain.make_noise_sampler(x, None, None, seed=0, cpu=True, normalized=True)(None, None)
    out["filter_noise_24x20"] = {"shape": (2, 3, 24, 20), "draws": draws, "out": result.clone()}
    # 5-D video latent via SonarCustomNoiseParameters(frames_to_channels=True) (config C5, scaled down)
    torch.manual_seed(211)
    x5 = torch.zeros(1, 2, 3, 18, 20)
    params_item = noise.CustomNoiseParametersNoise(
        1.0, noise=chain_with(alpha=1.0), normalize=None, override_device=None, override_dtype=None,
        frames_to_channels=True, ensure_square_aspect_ratio=False, fix_invalid=False, rng_mode="default",
        rng_offset_mode="disabled", rng_state_offset=0,
    )  # fmt: skip
    chain5 = noise.CustomNoiseChain()
    chain5.add(params_item)
    with record_draws() as draws:
        result = chain5.make_noise_sampler(x5, None, None, seed=0, cpu=True, normalized=True)(None, None)
    out["video_5d_18x20"] = {"shape": tuple(x5.shape), "draws": draws, "out": result.clone()}
    save("power_noise", out)


def gen_graph(ref) -> None:
    noise = ref.py.noise

    def chain_of(noise_type, factor=1.0):
        c = noise.CustomNoiseChain()
        c.add(noise.CustomNoiseItem(factor, noise_type=noise_type))
        return c

    out = {}
    shape = (2, 4, 16, 16)
    x = torch.zeros(shape)
    # config C3 (scaled down): Scheduled(Blended(lerp 0.5, pyramid, perlin), fallback gaussian)
    blended = noise.CustomNoiseChain()
    blended.add(
        noise.BlendedNoise(
            1.0, normalize=None, blend_function=torch.lerp, custom_noise_1=chain_of("pyramid"),
            custom_noise_2=chain_of("perlin"), noise_2_percent=0.5,
        ),  # fmt: skip
    )
    sched = noise.CustomNoiseChain()
    sched.add(noise.ScheduledNoise(1.0, noise=blended, start_sigma=10.0, end_sigma=1.0, normalize=None, fallback_noise=chain_of("gaussian")))
    for tag, sig in (("in_range", 5.0), ("fallback", 12.0)):
        torch.manual_seed(300)
        with record_draws() as draws:
            ns = sched.make_noise_sampler(x, torch.tensor(0.03), torch.tensor(14.6), seed=0, cpu=True, normalized=True)
            result = ns(torch.tensor(sig), torch.tensor(sig * 0.9))
        out[f"c3_scheduled_{tag}"] = {"shape": shape, "sigma": sig, "draws": draws, "out": result.clone()}
    # chain of two items with factors, normalised at the chain level
    torch.manual_seed(301)
    two = noise.CustomNoiseChain()
    two.add(noise.CustomNoiseItem(0.6, noise_type="gaussian"))
    two.add(noise.CustomNoiseItem(-0.4, noise_type="uniform"))
    with record_draws() as draws:
        result = two.make_noise_sampler(x, None, None, seed=0, cpu=True, normalized=True)(None, None)
    out["chain_two"] = {"shape": shape, "draws": draws, "out": result.clone()}
    # rescaled chain
    torch.manual_seed(302)
    with record_draws() as draws:
        result = two.rescaled(2.0).make_noise_sampler(x, None, None, seed=0, cpu=True, normalized=False)(None, None)
    out["chain_rescaled"] = {"shape": shape, "draws": draws, "out": result.clone()}
    # composite with a mask
    torch.manual_seed(303)
    mask = torch.zeros(1, 8, 8)
    mask[:, 2:6, 1:5] = 1.0
    comp = noise.CustomNoiseChain()
    comp.add(
        noise.CompositeNoise(
            1.0, dst_noise=chain_of("gaussian"), src_noise=chain_of("uniform"), normalize_dst=None,
            normalize_src=None, normalize_result=None, mask=mask,
        ),  # fmt: skip
    )
    with record_draws() as draws:
        result = comp.make_noise_sampler(x, None, None, seed=0, cpu=True, normalized=True)(None, None)
    out["composite"] = {"shape": shape, "mask": mask, "draws": draws, "out": result.clone()}
    # blended with a mask noise
    torch.manual_seed(304)
    bm = noise.CustomNoiseChain()
    bm.add(
        noise.BlendedNoise(
            1.0, normalize=None, blend_function=torch.lerp, custom_noise_1=chain_of("gaussian"),
            custom_noise_2=chain_of("uniform"), custom_noise_mask=chain_of("gaussian"), noise_2_percent=0.25,
        ),  # fmt: skip
    )
    with record_draws() as draws:
        result = bm.make_noise_sampler(x, None, None, seed=0, cpu=True, normalized=True)(None, None)
    out["blended_mask"] = {"shape": shape, "draws": draws, "out": result.clone()}
    # repeated noise: 6 calls, flips / rolls driven by the seeded CPU generator (bit-exact index work)
    torch.manual_seed(305)
    rep = noise.CustomNoiseChain()
    rep.add(noise.RepeatedNoise(1.0, noise=chain_of("gaussian"), repeat_length=2, max_recycle=3, permute="enabled", normalize=None))
    with record_draws() as draws:
        ns = rep.make_noise_sampler(x, None, None, seed=1234, cpu=True, normalized=False)
        results = [ns(None, None).clone() for _ in range(6)]
    out["repeated"] = {"shape": shape, "seed": 1234, "draws": draws, "out": torch.stack(results)}
    save("noise_graph", out)


def gen_samplers(ref) -> None:
    sonar = ref.py.sonar
    x0 = None
    out = {}
    sigmas = torch.cat((torch.linspace(14.6, 0.03, 7), torch.zeros(1)))
    variants = {
        "default": {},
        "classic": {"momentum_mode": "classic"},
        "denoised": {"momentum_mode": "denoised"},
        "dir_neg": {"direction": -0.5},
        "init_sample": {"init": "sample"},
        "init_sample_norm": {"init": "sample_norm", "momentum_mode": "classic"},
        "init_sample_denoised": {"init": "sample", "momentum_mode": "denoised"},
        "init_rand": {"init": "rand", "rand_init_noise_type": "gaussian", "rand_init_noise_multiplier": 0.5},
        "momentum_one": {"momentum": 1.0},
        "hist_one": {"momentum_hist": 1.0, "init": "sample"},
        "window": {"momentum_start_step": 2, "momentum_end_step": 4, "always_update_history": False},
        "blend_inject": {"momentum_blend_mode": "inject", "history_blend_mode": "subtract_b"},
    }
    samplers = {
        "euler": (sonar.SonarEuler.sampler, {}),
        "euler_ancestral": (sonar.SonarEulerAncestral.sampler, {"eta": 1.0, "s_noise": 1.0}),
        "euler_ancestral_eta": (sonar.SonarEulerAncestral.sampler, {"eta": 0.6, "s_noise": 1.1}),
        "dpmpp_sde": (sonar.SonarDPMPPSDE.sampler, {"eta": 1.0, "s_noise": 1.0}),
    }
    torch.manual_seed(400)
    x0 = torch.randn(2, 4, 8, 12) * sigmas[0]
    for sname, (fn, skw) in samplers.items():
        for vname, params in variants.items():
            if sname == "euler_ancestral_eta" and vname not in {"default", "classic"}:
                continue
            params = dict(params)
            if sname.startswith("dpmpp"):
                params.setdefault("noise_type", "gaussian")  # default Brownian needs torchsde
            steps = []
            torch.manual_seed(401)
            with record_draws() as draws:
                result = fn(
                    stub_model, x0.clone(), sigmas, extra_args={"seed": 0}, disable=True, sonar_params=params,
                    callback=lambda d: steps.append(d["x"].clone()), **skw,
                )  # fmt: skip
            out[f"{sname}/{vname}"] = {
                "params": params, "sampler_kwargs": skw, "draws": draws, "out": result.clone(), "steps": torch.stack(steps),
            }  # fmt: skip
    save("samplers", {"x0": x0, "sigmas": sigmas, "cases": out})


def gen_guidance(ref) -> None:
    """Reference-latent guidance (py/sonar.py:323-411) through all three samplers: LINEAR (lerp and
    inject blends) and EULER, a step window, a batch-broadcast reference latent."""
    sonar = ref.py.sonar
    sigmas = torch.cat((torch.linspace(12.0, 0.05, 6), torch.zeros(1)))
    torch.manual_seed(500)
    x0 = torch.randn(3, 4, 8, 12) * sigmas[0]
    ref_full = torch.randn(3, 4, 8, 12) * 2.0 + 0.3
    ref_one = torch.randn(1, 4, 8, 12) - 0.5
    variants = {
        "linear": ({"guidance_type": "LINEAR", "factor": 0.05, "start_step": 0, "end_step": 9999}, ref_full, {}),
        "linear_window_one": ({"guidance_type": "LINEAR", "factor": 0.1, "start_step": 1, "end_step": 3}, ref_one, {}),
        "linear_inject": ({"guidance_type": "LINEAR", "factor": 0.02, "start_step": 0, "end_step": 9999}, ref_full,
                          {"guidance_blend_mode": "inject"}),
        "euler": ({"guidance_type": "EULER", "factor": 0.2, "start_step": 0, "end_step": 9999}, ref_full, {}),
        "euler_denoised_mode": ({"guidance_type": "EULER", "factor": 0.3, "start_step": 1, "end_step": 9999}, ref_one,
                                {"momentum_mode": "denoised"}),
    }
    samplers = {
        "euler": (sonar.SonarEuler.sampler, {}),
        "euler_ancestral": (sonar.SonarEulerAncestral.sampler, {"eta": 1.0, "s_noise": 1.0}),
        "dpmpp_sde": (sonar.SonarDPMPPSDE.sampler, {"eta": 1.0, "s_noise": 1.0}),
    }
    out = {}
    for sname, (fn, skw) in samplers.items():
        for vname, (gkw, latent, extra) in variants.items():
            guidance = sonar.GuidanceConfig(
                guidance_type=sonar.GuidanceType[gkw["guidance_type"]], factor=gkw["factor"],
                start_step=gkw["start_step"], end_step=gkw["end_step"], latent=latent.clone(),
            )  # fmt: skip
            params = dict(extra)
            if sname.startswith("dpmpp"):
                params.setdefault("noise_type", "gaussian")
            steps = []
            torch.manual_seed(501)
            with record_draws() as draws:
                result = fn(
                    stub_model, x0.clone(), sigmas, extra_args={"seed": 0}, disable=True,
                    sonar_params=params | {"guidance": guidance}, callback=lambda d: steps.append(d["x"].clone()), **skw,
                )  # fmt: skip
            out[f"{sname}/{vname}"] = {
                "params": params, "guidance": gkw, "latent": latent.clone(), "sampler_kwargs": skw, "draws": draws,
                "out": result.clone(), "steps": torch.stack(steps),
            }  # fmt: skip
    save("guidance", {"x0": x0, "sigmas": sigmas, "cases": out})


def gen_host_logic(ref) -> None:
    """Host-only behaviour: config merging errors, yh scale expansion, rule parsing."""
    sonar, wf = ref.py.sonar, ref.py.wavelet_functions
    out = {"config_errors": {}, "expand": [], "history_ratios": []}
    for params in ({"momentum_mode": "bogus"}, {"init": 3}, {"noise_type": "nope"}):
        try:
            sonar.SonarBase.get_config(None, params)
        except (ValueError, TypeError) as exc:
            out["config_errors"][repr(params)] = (type(exc).__name__, str(exc))
    for direction, mh in ((1.0, 0.75), (-0.5, 0.75), (0.3, 0.2), (-1.0, 1.0)):
        cfg = sonar.SonarConfig(direction=direction, momentum_hist=mh)
        out["history_ratios"].append(((direction, mh), tuple(sonar.SonarBase(cfg).history_ratios)))
    shapes = [torch.zeros(1, 1, 3, 4, 4)] * 4
    for spec in (2.0, [1, 2], [[3, 4, 5], 2.0], [1.0, "fill", 9.0], [[1], [2, 3], "fill"], [1, 2, 3, 4, 5, 6]):
        out["expand"].append((spec, wf.expand_yh_scales(shapes, yh_scales=spec)))
    save("host_logic", out)


def gen_freeu(ref) -> None:
    """FreeU-Extreme: ffilter and FreeUExtremeConfig.apply on UNet-like activations,
    and the patched-model handlers of FreeUExtremeNode."""
    fx, pn = ref.py.nodes.freeu_extreme, ref.py.nodes.powernoise
    out = {"ffilter": {}, "apply": {}, "node": {}}
    torch.manual_seed(600)
    for name, shape, fkw, norm in (
        ("lowpass_16x20", (2, 6, 16, 20), {"alpha": 1.0}, 1.0),
        ("band_32x32", (1, 8, 32, 32), {"alpha": 0.0, "min_freq": 0.1, "max_freq": 0.35}, 0.0),
        ("odd_15x18", (1, 3, 15, 18), {"alpha": 0.5, "scale": 0.9}, 0.5),
    ):
        x = torch.randn(shape)
        # (the reference's ffilter only works with a cache: without one `filter_rfft` is unbound, :12-15)
        got = fx.ffilter(x.clone(), pn.PowerFilter(**fkw), normalization_factor=norm, cfg_idx=0, filter_cache={})
        out["ffilter"][name] = {"x": x, "filter": fkw, "norm": norm, "out": got.clone()}
    base = {"target": "backbone", "stage_1": True, "stage_2": True, "stage_3": True}
    cases = {
        "v2_full": {"scale": 1.3, "hidden_mean": True, "filter": {"alpha": 1.0}},
        "plain_scale": {"scale": 0.8, "hidden_mean": False, "filter": None},
        "slice_lerp": {"scale": 1.2, "hidden_mean": True, "slice": 0.5, "slice_offset": 0.25, "blend": 0.6,
                       "blend_mode": "lerp", "filter": {"alpha": 0.5, "min_freq": 0.05}, "filter_norm": 1.0},
        "slice_inject": {"scale": 1.1, "hidden_mean": False, "slice": 0.34, "slice_offset": 0.5, "blend": 0.25,
                         "blend_mode": "inject", "filter": {"alpha": 1.0}, "filter_norm": 0.5},
        "subtract": {"scale": 1.4, "hidden_mean": True, "blend": 0.5, "blend_mode": "subtract_b", "filter": None},
    }  # fmt: skip
    for i, (name, kw) in enumerate(cases.items()):
        torch.manual_seed(610 + i)
        x = torch.randn(2, 12, 16, 24)
        kw = dict(kw)
        fkw = kw.pop("filter")
        cfg = fx.FreeUExtremeConfig(**base, **kw, sonar_power_filter_opt=None if fkw is None else pn.PowerFilter(**fkw))
        got = cfg.apply(0, x.clone(), {})
        out["apply"][name] = {"x": x, "config": kw, "filter": fkw, "out": got.clone()}

    # the node: a stand-in model that records the patches; stage lookup by channel count, percent window
    class _MS:
        @staticmethod
        def timestep(sigma):
            return sigma * 100.0

    class _Model:
        def __init__(self):
            self.patches = {}
            self.model = type("M", (), {"model_config": type("C", (), {"unet_config": {"model_channels": 4}})()})()

        def clone(self):
            return self

        def get_model_object(self, _name):
            return _MS()

        def set_model_input_block_patch(self, fn):
            self.patches["input"] = fn

        def set_model_patch(self, fn, name):
            self.patches[name] = fn

        def set_model_output_block_patch(self, fn):
            self.patches["output"] = fn

    second = fx.FreeUExtremeConfig(
        target="skip", stage_2=True, start=0.0, end=1.0, scale=0.7, hidden_mean=False, final=True,
        sonar_power_filter_opt=pn.PowerFilter(alpha=1.0), filter_norm=1.0,
    )
    first = fx.FreeUExtremeConfig(
        target="backbone", stage_1=True, stage_2=True, start=0.2, end=0.9, scale=1.25, hidden_mean=True, final=False,
        sonar_power_filter_opt=pn.PowerFilter(alpha=0.5), filter_norm=1.0, frux_config_opt=second,
    )
    (model,) = fx.FreeUExtremeNode.go(_Model(), False, input_config=first, middle_config=first, output_config=first)
    torch.manual_seed(630)
    h16, h8, hsp8, h5 = torch.randn(1, 16, 8, 8), torch.randn(2, 8, 12, 10), torch.randn(2, 8, 12, 10), torch.randn(1, 5, 8, 8)
    rec = {"inputs": {"h16": h16, "h8": h8, "hsp8": hsp8, "h5": h5}, "calls": []}
    for sigma in (5.0, 0.5, 9.5):  # pct = 1 - sigma / 9.99: inside, inside, outside the [0.2, 0.9] window
        topt = {"sigmas": torch.tensor([sigma, sigma])}
        a = model.patches["input"](h16.clone(), topt)
        b = model.patches["middle_block_patch"](h5.clone(), topt)
        c, d = model.patches["output"](h8.clone(), hsp8.clone(), topt)
        rec["calls"].append({"sigma": sigma, "input": a.clone(), "middle": b.clone(), "out_h": c.clone(), "out_hsp": d.clone()})
    out["node"] = rec
    save("freeu", out)


def gen_round2(ref) -> None:
    """Round-2 fixtures: (1) north-star config 5 as specified, scaled down -- frames_to_channels power noise as
    the custom noise of sonar_dpmpp_sde on a 5-D video latent; (2) SonarPowerNoise with a non-identity
    ChannelMixer (common_mode != 0, non-unit channel_correlation), incl. the filter-noise form and a
    frames_to_channels latent with a large channel count; (3) GuidedNoise (py/noise.py:565-623)."""
    noise, pn, sonar = ref.py.noise, ref.py.nodes.powernoise, ref.py.sonar
    defaults = {
        "time_brownian": False, "alpha": 0.0, "max_freq": 0.7071, "min_freq": 0.0, "stretch": 1.0, "rotate": 0.0,
        "pnorm": 2.0, "mix": 1.0, "common_mode": 0.0, "channel_correlation": "1, 1, 1, 1, 1, 1",
    }  # fmt: skip

    def power_chain(**kw):
        chain = noise.CustomNoiseChain()
        chain.add(pn.PowerNoiseItem(1.0, **(defaults | kw)))
        return chain

    def video_chain(**kw):
        item = noise.CustomNoiseParametersNoise(
            1.0, noise=power_chain(**kw), normalize=None, override_device=None, override_dtype=None,
            frames_to_channels=True, ensure_square_aspect_ratio=False, fix_invalid=False, rng_mode="default",
            rng_offset_mode="disabled", rng_state_offset=0,
        )  # fmt: skip
        chain = noise.CustomNoiseChain()
        chain.add(item)
        return chain

    out = {"c5": {}, "mixer": {}, "guided": {}}
    # ---- (1) config 5: CustomNoiseParametersNoise(frames_to_channels) o PowerNoise(alpha=1) -> sonar_dpmpp_sde ----
    sigmas = torch.cat((torch.linspace(14.6, 0.03, 5), torch.zeros(1)))
    torch.manual_seed(700)
    x0 = torch.randn(2, 3, 4, 18, 20) * sigmas[0]
    for vname, params in (("default", {}), ("classic", {"momentum_mode": "classic"})):
        steps = []
        torch.manual_seed(701)
        with record_draws() as draws:
            result = sonar.SonarDPMPPSDE.sampler(
                stub_model, x0.clone(), sigmas, extra_args={"seed": 0}, disable=True,
                sonar_params=params | {"custom_noise": video_chain(alpha=1.0)},
                callback=lambda d: steps.append(d["x"].clone()), eta=1.0, s_noise=1.0,
            )  # fmt: skip
        out["c5"][vname] = {"params": params, "draws": draws, "out": result.clone(), "steps": torch.stack(steps)}
    out["c5"]["x0"], out["c5"]["sigmas"] = x0, sigmas

    # ---- (2) non-identity ChannelMixer ----
    variants = {
        "c4_common": ((2, 4, 16, 20), {"alpha": 1.0, "common_mode": 0.3}),
        "c4_corr": ((2, 4, 18, 16), {"alpha": 0.5, "common_mode": 0.5, "channel_correlation": "0.9, 0.5, -0.3, 0.2, 0.7, 1"}),
        "c3_neg": ((1, 3, 16, 16), {"alpha": 0.0, "common_mode": -0.25, "channel_correlation": "1, 0.5"}),
        "c16_short_corr": ((2, 16, 12, 16), {"alpha": 1.0, "common_mode": 0.2, "channel_correlation": "1, 0.3, 0.8"}),
    }
    for i, (name, (shape, kw)) in enumerate(variants.items()):
        torch.manual_seed(710 + i)
        x = torch.zeros(shape)
        with record_draws() as draws:
            result = power_chain(**kw).make_noise_sampler(x, None, None, seed=0, cpu=True, normalized=True)(None, None)
        item = pn.PowerNoiseItem(1.0, **(defaults | kw))
        mixer = pn.ChannelMixer(shape[1], item.common_mode, item.channel_correlation).mixer
        out["mixer"][name] = {"shape": shape, "params": kw, "draws": draws, "mixer": mixer.clone(), "out": result.clone()}
    # video latent: 12 x 11 = 132 folded channels, the (C x C) @ (C x B*H*W) product is a real GEMM
    torch.manual_seed(720)
    x5 = torch.zeros(2, 12, 11, 10, 12)
    kw = {"alpha": 1.0, "common_mode": 0.1, "channel_correlation": "1, 0.5, 0.25"}
    with record_draws() as draws:
        result = video_chain(**kw).make_noise_sampler(x5, None, None, seed=0, cpu=True, normalized=True)(None, None)
    out["mixer"]["video_c132"] = {"shape": tuple(x5.shape), "params": kw, "draws": draws, "out": result.clone()}
    # the rfft2 front end (PowerFilterNoiseItem) with a mixer
    torch.manual_seed(721)
    x = torch.zeros(2, 4, 20, 24)
    inner = noise.CustomNoiseChain()
    inner.add(noise.CustomNoiseItem(1.0, noise_type="gaussian"))
    item = pn.PowerFilterNoiseItem(
        1.0, noise=inner, normalize_noise=None, normalize_result=None, power_filter=pn.PowerFilter(alpha=1.0),
        mix=1.0, common_mode=0.4, channel_correlation="1,-0.5,0.5,1,1,0.2", time_brownian=True, filter_norm_factor=1.0,
    )  # fmt: skip
    chain = noise.CustomNoiseChain()
    chain.add(item)
    with record_draws() as draws:
        result = chain.make_noise_sampler(x, None, None, seed=0, cpu=True, normalized=True)(None, None)
    out["mixer"]["filter_noise_c4"] = {"shape": tuple(x.shape), "draws": draws, "out": result.clone()}

    # ---- (3) GuidedNoise ----
    def gaussian_chain():
        c = noise.CustomNoiseChain()
        c.add(noise.CustomNoiseItem(1.0, noise_type="gaussian"))
        return c

    shape = (3, 4, 12, 16)
    torch.manual_seed(730)
    x = torch.randn(shape) * 3.0 + 0.5
    ref_same = torch.randn(shape) * 2.0 - 0.3
    ref_one = torch.randn(1, 4, 12, 16) + 1.0
    ref_small = torch.randn(3, 4, 6, 8)
    gcases = {
        "linear": {"method": "linear", "guidance_factor": 0.3, "ref": ref_same, "noise": True},
        "linear_one_ref": {"method": "linear", "guidance_factor": 0.1, "ref": ref_one, "noise": True, "factor": 0.5},
        "linear_no_noise": {"method": "linear", "guidance_factor": 0.25, "ref": ref_same, "noise": False},
        "linear_resized_ref": {"method": "linear", "guidance_factor": 0.2, "ref": ref_small, "noise": True},
        "euler": {"method": "euler", "guidance_factor": 0.2, "ref": ref_same, "noise": True},
        "euler_no_noise": {"method": "euler", "guidance_factor": 0.4, "ref": ref_one, "noise": False},
        "euler_equal_sigmas": {"method": "euler", "guidance_factor": 0.2, "ref": ref_same, "noise": True, "sigmas": (3.0, 3.0)},
        "linear_unnormalized": {"method": "linear", "guidance_factor": 0.3, "ref": ref_same, "noise": True,
                                "normalize_noise": False, "normalize_result": False},
    }  # fmt: skip
    for i, (name, kw) in enumerate(gcases.items()):
        torch.manual_seed(740 + i)
        item = noise.GuidedNoise(
            kw.get("factor", 1.0), guidance_factor=kw["guidance_factor"], ref_latent=kw["ref"].clone(), method=kw["method"],
            normalize_noise=kw.get("normalize_noise"), normalize_result=kw.get("normalize_result"),
            noise=gaussian_chain() if kw["noise"] else None,
        )  # fmt: skip
        chain = noise.CustomNoiseChain()
        chain.add(item)
        s, sn = kw.get("sigmas", (5.0, 3.5))
        with record_draws() as draws:
            ns = chain.make_noise_sampler(x.clone(), torch.tensor(0.03), torch.tensor(14.6), seed=0, cpu=True, normalized=True)
            result = ns(torch.tensor(s), torch.tensor(sn))
        out["guided"][name] = {
            "config": {k: v for k, v in kw.items() if k != "ref"}, "ref": kw["ref"].clone(), "sigmas": (s, sn),
            "draws": draws, "out": result.clone(),
        }  # fmt: skip
    out["guided"]["x"] = x
    save("round2", out)


IN_SCOPE_NODES = (
    "SamplerSonarEuler", "SamplerSonarEulerA", "SamplerSonarDPMPPSDE", "SonarGuidanceConfig", "SonarCustomNoise",
    "SonarCustomNoiseAdv", "SonarPowerNoise", "SonarPowerFilterNoise", "SonarPowerFilter", "SonarAdvancedPyramidNoise",
    "SonarAdvanced1fNoise", "SonarAdvancedPowerLawNoise", "SonarCompositeNoise", "SonarScheduledNoise",
    "SonarBlendedNoise", "SonarRepeatedNoise", "SonarCustomNoiseParameters", "SONAR_CUSTOM_NOISE to NOISE",
    "SamplerConfigOverride", "SonarWaveletCFG", "NoisyLatentLike", "FreeUExtremeConfig", "FreeUExtreme",
    "SonarWaveletFilteredNoise", "SonarGuidedNoise",
)  # fmt: skip


def gen_noisy_latent(ref) -> None:
    """NoisyLatentLike (py/nodes/misc.py:55-150): the node end to end on the reference -- built-in
    noise type and custom chain, repeat_batch, add_to_latent, the sigma / model multiplier (both max_denoise branches)."""
    misc, noise, pn = ref.py.nodes.misc, ref.py.noise, ref.py.nodes.powernoise
    defaults = {
        "time_brownian": False, "alpha": 1.0, "max_freq": 0.7071, "min_freq": 0.0, "stretch": 1.0, "rotate": 0.0,
        "pnorm": 2.0, "mix": 1.0, "common_mode": 0.0, "channel_correlation": "1, 1, 1, 1, 1, 1",
    }  # fmt: skip
    chain = noise.CustomNoiseChain()
    chain.add(pn.PowerNoiseItem(1.0, **defaults))

    class _MS:
        sigma_max = torch.tensor(14.6)

    class _LF:
        scale_factor = 0.13025

    class _Inner:
        model_sampling, latent_format = _MS(), _LF()

    class _Model:
        model = _Inner()

    torch.manual_seed(900)
    latent = torch.randn(2, 4, 16, 24)
    cases = {
        "gaussian_repeat_add": {"noise_type": "gaussian", "seed": 11, "multiplier": 1.5, "add_to_latent": True, "repeat_batch": 2},
        "power_custom": {"noise_type": "gaussian", "seed": 12, "multiplier": 0.75, "repeat_batch": 3, "custom": True},
        "sigmas_max_denoise": {"noise_type": "gaussian", "seed": 13, "multiplier": 1.0, "add_to_latent": True,
                               "sigmas": torch.tensor([14.6, 7.0, 1.0, 0.0])},
        "sigmas_partial": {"noise_type": "gaussian", "seed": 14, "multiplier": 2.0, "normalize": False,
                           "sigmas": torch.tensor([5.0, 2.0, 0.0])},
    }  # fmt: skip
    out = {"latent": latent, "scale_factor": _LF.scale_factor, "sigma_max": 14.6, "cases": {}}
    for name, kw in cases.items():
        kw = dict(kw)
        custom, sigmas = kw.pop("custom", False), kw.pop("sigmas", None)
        with record_draws() as draws:
            (res,) = misc.NoisyLatentLikeNode.go(
                latent={"samples": latent.clone()}, cpu_noise=True, custom_noise_opt=chain if custom else None,
                mul_by_sigmas_opt=sigmas, model_opt=_Model() if sigmas is not None else None, **kw,
            )
        out["cases"][name] = {"kwargs": kw, "custom": custom, "sigmas": sigmas, "draws": draws, "out": res["samples"].clone()}
    save("noisy_latent", out)


def gen_node_schemas(ref) -> None:
    """INPUT_TYPES / RETURN_TYPES / FUNCTION / CATEGORY of the in-scope nodes (tooltips dropped)."""
    import json

    out = {}
    for name in IN_SCOPE_NODES:
        cls = ref.NODE_CLASS_MAPPINGS[name]
        schema = {}
        for section, fields in cls.INPUT_TYPES().items():
            schema[section] = {}
            for key, spec in fields.items():
                opts = {k: v for k, v in (spec[1] if len(spec) > 1 else {}).items() if k != "tooltip"}
                kind = list(spec[0]) if isinstance(spec[0], (tuple, list)) else str(spec[0])
                schema[section][key] = [kind, opts]
        out[name] = {
            "input_types": schema,
            "return_types": list(cls.RETURN_TYPES),
            "function": cls.FUNCTION,
            "category": getattr(cls, "CATEGORY", None),
        }
    path = HERE / "node_schemas.json"
    path.write_text(json.dumps(out, indent=1))
    print(f"{'node_schemas':38s} {path.stat().st_size / 1024:8.1f} KiB")


def main() -> None:
    if not REFERENCE.exists():
        raise SystemExit("make_golden.py needs the reference at /root/reference")
    torch.set_num_threads(1)
    ref = load_reference()
    only = sys.argv[1:]  # e.g. `make_golden.py guidance` regenerates one fixture
    if only:
        for name in only:
            globals()[f"gen_{name}"](ref)
        return
    gen_noise_types(ref)
    gen_power_noise(ref)
    gen_graph(ref)
    gen_samplers(ref)
    gen_guidance(ref)
    gen_host_logic(ref)
    gen_freeu(ref)
    gen_round2(ref)
    gen_noisy_latent(ref)
    gen_node_schemas(ref)


if __name__ == "__main__":
    main()
