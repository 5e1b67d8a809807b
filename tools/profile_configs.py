"""Runs one BASELINE.json config a few times on cuda:0 -- the target of the ncu captures under
profiles/ (see profiles/README.md for the exact commands).

    python tools/profile_configs.py c2|c2_large|c1|c3|c4|c5_noise|c5_dpm [reps]
"""
import math
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import sonar_b200 as sb  # noqa: E402

dev = torch.device("cuda", 0)
which = sys.argv[1] if len(sys.argv) > 1 else "c2"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
ng, sn = sb.noise_graph, sb.spectral_noise
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def model(x, sigma, **_kw):
    flush.zero_()
    return x * 0.9


def power_chain():
    c = ng.CustomNoiseChain()
    c.add(sn.PowerNoiseItem(1.0, time_brownian=False, alpha=1.0, max_freq=0.7071, min_freq=0.0, stretch=1.0, rotate=0.0,
                            pnorm=2.0, mix=1.0, common_mode=0.0, channel_correlation="1, 1, 1, 1, 1, 1"))
    return c


def chain_of(t):
    c = ng.CustomNoiseChain()
    c.add(ng.CustomNoiseItem(1.0, noise_type=t))
    return c


if which in {"c2", "c2_large"}:
    shape = (8, 4, 128, 128) if which == "c2" else (1, 16, 33, 90, 160)
    sig = torch.cat((torch.linspace(14.6, 0.03, 30 if which == "c2" else 4), torch.zeros(1))).to(dev)
    x0 = torch.randn(shape, device=dev) * 14.6
    fn = lambda: sb.samplers.SonarEulerAncestral.sampler(model, x0, sig, extra_args={"seed": 0}, disable=True)  # noqa: E731
elif which == "c1":
    x = torch.zeros(1, 4, 64, 64, device=dev)
    ns = power_chain().make_noise_sampler(x, None, None, seed=0)
    fn = lambda: ns(None, None)  # noqa: E731
elif which == "c3":
    blended = ng.CustomNoiseChain()
    blended.add(ng.BlendedNoise(1.0, normalize=None, blend_function=sb.hostutil.BLENDING_MODES["lerp"],
                                custom_noise_1=chain_of("pyramid"), custom_noise_2=chain_of("perlin"), noise_2_percent=0.5))
    sched = ng.CustomNoiseChain()
    sched.add(ng.ScheduledNoise(1.0, noise=blended, start_sigma=10.0, end_sigma=1.0, normalize=None, fallback_noise=chain_of("gaussian")))
    x = torch.zeros(16, 16, 128, 128, device=dev)
    ns = sched.make_noise_sampler(x, torch.tensor(0.03), torch.tensor(14.6), seed=0)

    def fn():
        torch.manual_seed(0)
        return ns(torch.tensor(5.0), torch.tensor(4.5))
elif which == "c4":
    class _MS:
        sigma_min, sigma_max = torch.tensor(0.03), torch.tensor(14.6)

        @staticmethod
        def timestep(sg):
            return (sg.log() - math.log(0.03)) / (math.log(14.6) - math.log(0.03)) * 999

    class _Model:
        model_sampling = _MS()

    cond, uncond, xin = (torch.randn(16, 4, 128, 128, device=dev) for _ in range(3))
    cfg = sb.wcfg.WaveletCFG(existing_cfg=None, rules=sb.wcfg.WCFGRules.build(
        wave="db2", level=3, diff={"yl_scale": 5, "yh_scales": [[3, 4, 5]] * 3}))
    wargs = {"sigma": torch.full((16,), 5.0, device=dev), "input": xin, "cond_denoised": cond, "uncond_denoised": uncond,
             "cond_scale": 7.0, "model": _Model(), "model_options": {}}
    fn = lambda: cfg(wargs)  # noqa: E731
elif which == "c5_noise":
    x5 = torch.zeros(1, 16, 33, 90, 160, device=dev)
    params = ng.CustomNoiseParametersNoise(
        1.0, noise=power_chain(), normalize=None, override_device=None, override_dtype=None, frames_to_channels=True,
        ensure_square_aspect_ratio=False, fix_invalid=False, rng_mode="default", rng_offset_mode="disabled", rng_state_offset=0)
    c5 = ng.CustomNoiseChain()
    c5.add(params)
    ns5 = c5.make_noise_sampler(x5, None, None, seed=0)
    fn = lambda: ns5(None, None)  # noqa: E731
elif which == "c5_dpm":
    sig5 = torch.tensor([14.6, 7.0, 2.0, 0.7], device=dev)
    xv = torch.randn(1, 16, 33, 90, 160, device=dev) * 14.6
    fn = lambda: sb.samplers.SonarDPMPPSDE.sampler(  # noqa: E731
        model, xv, sig5, extra_args={"seed": 0}, disable=True, sonar_params={"noise_type": "gaussian"})
else:
    raise SystemExit(f"unknown config {which}")

for _ in range(reps):
    fn()
torch.cuda.synchronize()
print(f"{which}: {reps} reps, {sb.ops.LAUNCH_COUNT} sonar_b200 launches")
