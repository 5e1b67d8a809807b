"""Shared helpers for the parity tests: how each golden case maps onto oracle calls."""
from __future__ import annotations

import torch

from oracle import sonar_oracle as orc

RTOL = ATOL = 1e-5  # fp32 tolerance stated by BASELINE.json north_star


def assert_close(actual: torch.Tensor, expected: torch.Tensor, *, rtol=RTOL, atol=ATOL, what=""):
    actual, expected = actual.detach().cpu(), expected.detach().cpu()
    assert actual.shape == expected.shape, f"{what}: shape {tuple(actual.shape)} vs {tuple(expected.shape)}"
    torch.testing.assert_close(actual, expected, rtol=rtol, atol=atol, msg=lambda m: f"{what}: {m}")


def split_draws(draws):
    """Separates the (n,)-shaped host draws (pyramid level sizing) from the tensor draws."""
    host = [d for d in draws if d.ndim == 1]
    dev = [d for d in draws if d.ndim != 1]
    return host, dev


def oracle_noise_type(name: str, case: dict) -> torch.Tensor:
    """Reference pipeline for get_noise_sampler(name, x, normalized=True)(None, None) restated with
    the oracle: generator -> (its own output_hook) -> NoiseSampler scale_noise."""
    shape = tuple(case["shape"])
    shape4 = shape if len(shape) == 4 else (shape[0], shape[1] * shape[2], *shape[3:])
    host, dev = split_draws(case["draws"])
    it = iter(dev)

    def pyramid(discount=0.7, mode="bilinear", host_iter=None):
        rs = []
        sizes = []
        h, w = shape4[-2:]
        # host draws arrive one per level until a side hits 1 (or 10 iterations)
        hi = host_iter
        for i in range(10):
            r = next(hi).item()
            rs.append(r)
            rr = r * 2 + 2
            w, h = max(1, int(w / (rr**i))), max(1, int(h / (rr**i)))
            sizes.append((h, w))
            if w == 1 or h == 1:
                break
        return orc.pyramid_noise(it, shape4, sizes, discount=discount, mode=mode)

    hi = iter(host)
    if name == "gaussian":
        out = next(it).clone()
    elif name == "uniform":
        out = orc.uniform_noise(it)
    elif name == "perlin":
        out = orc.perlin_noise(it, shape4)
    elif name in {"pyramid", "pyramid_5d"}:
        out = pyramid(host_iter=hi)
    elif name == "pyramid_discount5":
        out = pyramid(discount=0.5, host_iter=hi)
    elif name == "pyramid_area":
        out = pyramid(mode="area", host_iter=hi)
    elif name == "pyramid_mix":
        a = orc.scale_noise(pyramid(discount=0.6, host_iter=hi)).mul_(0.2)
        b = orc.scale_noise(pyramid(discount=0.6, host_iter=hi)).mul_(-0.8)
        out = a.add_(b)
    elif name == "pyramid_old":
        out = orc.pyramid_old_noise(it, shape4)
    elif name == "highres_pyramid":
        rs = next(hi) * 2 + 2
        sizes, (h, w) = [], shape4[-2:]
        oh, ow = h, w
        for i in range(4):
            r = rs[i].item()
            h, w = min(oh * 15, int(h * (r**i))), min(ow * 15, int(w * (r**i)))
            sizes.append((h, w))
            if h >= oh * 15 or w >= ow * 15:
                break
        # reference draw order: base uniform, then the host rs, then the levels
        out = orc.highres_pyramid_noise(it, shape4, sizes)
    elif name == "onef_pinkish":
        out = orc.onef_noise(it, shape4, alpha=-0.5)
    elif name == "onef_greenish":
        out = orc.onef_noise(it, shape4, alpha=0.5)
    elif name == "onef_pinkish_mix":
        a = orc.scale_noise(orc.onef_noise(it, shape4, alpha=-0.5)).mul_(-1.0)
        b = orc.scale_noise(orc.onef_noise(it, shape4, alpha=-0.5))
        out = a.add_(b).mul_(0.5)
    elif name == "onef_pinkishgreenish":
        a = orc.scale_noise(orc.onef_noise(it, shape4, alpha=0.5))
        b = orc.scale_noise(orc.onef_noise(it, shape4, alpha=-0.5))
        out = a.add_(b).mul_(0.5)
    elif name == "green_test":
        out = orc.green_test_noise(it, shape4)
    elif name == "rainbow_mild":
        a = orc.scale_noise(orc.green_test_noise(it, shape4)).mul_(0.55)
        b = orc.scale_noise(orc.green_test_noise(it, shape4)).mul_(0.7)
        out = a.add_(b).mul_(1.15)
    elif name in {"wavelet", "wavelet_odd"}:
        out = orc.wavelet_noise(it, shape4)
    elif name == "white":
        out = orc.powerlaw_noise(it, alpha=0.0, use_sign=True)
    elif name == "grey":
        out = orc.powerlaw_noise(it, alpha=0.0, use_sign=False)
    elif name == "velvet":
        out = orc.powerlaw_noise(it, alpha=1.0, use_sign=True, div_max_dims=(-3, -2, -1))
    elif name == "violet":
        out = orc.powerlaw_noise(it, alpha=0.5, use_sign=True, div_max_dims=(-3, -2, -1))
    else:
        raise KeyError(name)
    # NoiseSampler.__init__ forces the generator's own `normalized` off (noise.py:230) and
    # normalises once itself (:254); mixed generators' children keep their class default.
    out = orc.scale_noise(out.reshape(shape4), 1.0, normalized=True)
    return out.reshape(shape)


NOISE_TYPE_NAMES = (
    "gaussian", "uniform", "perlin", "pyramid", "pyramid_discount5", "pyramid_area", "pyramid_mix", "pyramid_old",
    "highres_pyramid", "onef_pinkish", "onef_greenish", "onef_pinkish_mix", "onef_pinkishgreenish", "green_test",
    "rainbow_mild", "white", "grey", "velvet", "violet", "wavelet", "wavelet_odd", "pyramid_5d",
)  # fmt: skip


def sampler_oracle_run(kind: str, case: dict, x0: torch.Tensor, sigmas: torch.Tensor, model):
    """Re-runs a golden sampler case with the oracle; returns the per-step x stack. Draw order of
    the reference per step: [RAND history init, first step only] -> noise 1 [-> noise 2 for DPM++]."""
    params = dict(case["params"])
    skw = case["sampler_kwargs"]
    draws = iter(case["draws"])
    o = orc.SonarOracle(
        momentum=params.get("momentum", 0.95),
        momentum_hist=params.get("momentum_hist", 0.75),
        direction=params.get("direction", 1.0),
        mode=params.get("momentum_mode", "new"),
        init=params.get("init", "zero"),
        momentum_start_step=params.get("momentum_start_step", 0),
        momentum_end_step=params.get("momentum_end_step", 9999),
        always_update_history=params.get("always_update_history", True),
        momentum_blend_mode=params.get("momentum_blend_mode"),
        history_blend_mode=params.get("history_blend_mode"),
        guidance_blend_mode=params.get("guidance_blend_mode"),
        guidance=None if "guidance" not in case else case["guidance"] | {"ref": orc.prepare_ref_latent(case["latent"])},
    )
    mult = params.get("rand_init_noise_multiplier", 1.0)

    def normalised(t):  # get_noise_sampler(GAUSSIAN, normalized=True): one conditional normalisation
        return orc.scale_noise(t.clone(), 1.0, normalized=True)

    x = x0.clone()
    steps = []
    eta, s_noise = skw.get("eta", 1.0), skw.get("s_noise", 1.0)
    for i in range(len(sigmas) - 1):
        sigma, sigma_next = sigmas[i], sigmas[i + 1]
        den = model(x, sigma)
        if o.init == "rand" and o.hist is None and o.init_noise is None:
            o.init_noise = normalised(next(draws)) * mult
        if kind == "euler":
            x = o.euler_step(i, x, den, sigma, sigma_next)
        elif kind.startswith("euler_ancestral"):
            noise = normalised(next(draws)) if sigma_next > 0 else None
            x = o.euler_ancestral(i, x, den, sigma, sigma_next, noise, eta=eta, s_noise=s_noise)
        elif sigma_next == 0:
            x = o.dpmpp_sde(i, x, den, sigma, sigma_next, model, None, None, eta=eta, s_noise=s_noise)
        else:
            # noise 2 is drawn after the second model call upstream, but the draw order does not
            # depend on the model, so fetching both eagerly is equivalent
            n1, n2 = normalised(next(draws)), normalised(next(draws))
            x = o.dpmpp_sde(i, x, den, sigma, sigma_next, model, n1, n2, eta=eta, s_noise=s_noise)
        steps.append(x.clone())
    return torch.stack(steps)


def stub_model(x, sigma, **_kwargs):
    """Same denoiser stand-in as tests/golden/make_golden.py."""
    return x * 0.9 - 0.05 * torch.tanh(x)
