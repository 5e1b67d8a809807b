"""RNG-fused pyramid / Perlin / blend (csrc/noise_mix.cu): the samples computed element-wise from the Philox stream
equal the materialising kernels (whose parity with the reference the golden fixtures pin) for the same generator
state, and advance torch's CUDA generator and the CPU generator identically."""
from __future__ import annotations

import pytest
import torch

from helpers import assert_close

pytestmark = pytest.mark.gpu


def chain_of(sb, noise_type, **item_kw):
    c = sb.noise_graph.CustomNoiseChain()
    c.add(sb.noise_graph.CustomNoiseItem(1.0, noise_type=noise_type, **item_kw))
    return c


def c3_graph(sb, t=0.5, mode="lerp"):
    ng = sb.noise_graph
    blended = ng.CustomNoiseChain()
    blended.add(ng.BlendedNoise(1.0, normalize=None, blend_function=sb.hostutil.BLENDING_MODES[mode],
                                custom_noise_1=chain_of(sb, "pyramid"), custom_noise_2=chain_of(sb, "perlin"), noise_2_percent=t))
    sched = ng.CustomNoiseChain()
    sched.add(ng.ScheduledNoise(1.0, noise=blended, start_sigma=10.0, end_sigma=1.0, normalize=None,
                                fallback_noise=chain_of(sb, "gaussian")))
    return sched


def both_ways(sb, monkeypatch, make, n_calls=2):
    """Runs `make()`'s sampler with the fused and the materialising kernels from the same generator states."""
    res = {}
    for fused in (True, False):
        monkeypatch.setattr(sb.generators, "FUSED_NOISE", fused)
        monkeypatch.setattr(sb.generators, "FUSED_SINGLE_GENERATORS", fused)
        torch.manual_seed(4242)
        ns = make()
        launches = sb.ops.LAUNCH_COUNT
        outs = [ns(torch.tensor(5.0), torch.tensor(4.5)) for _ in range(n_calls)]
        res[fused] = (outs, torch.cuda.default_generators[0].get_offset(), torch.rand(1).item(), sb.ops.LAUNCH_COUNT - launches)
    monkeypatch.undo()
    return res


@pytest.mark.parametrize("shape", [(16, 16, 128, 128), (2, 3, 32, 32), (1, 2, 24, 40), (3, 5, 17, 23), (1, 2, 3, 16, 16)])
@pytest.mark.parametrize("kind", ["pyramid", "perlin", "c3"])
def test_fused_equals_materialised(sb, cuda, monkeypatch, kind, shape):
    x = torch.zeros(shape, device=cuda)

    def make():
        chain = c3_graph(sb) if kind == "c3" else chain_of(sb, kind)
        return chain.make_noise_sampler(x, torch.tensor(0.03), torch.tensor(14.6), seed=0, cpu=True, normalized=True)

    res = both_ways(sb, monkeypatch, make)
    (fused, off_f, cpu_f, launches_f), (plain, off_p, cpu_p, launches_p) = res[True], res[False]
    assert off_f == off_p and cpu_f == cpu_p  # CUDA generator offset and CPU generator state advanced identically
    for j, (a, b) in enumerate(zip(fused, plain)):
        assert a.shape == x.shape
        assert_close(a, b, what=f"{kind} {shape} sample {j}", rtol=1e-6, atol=2e-6)
    assert launches_f <= launches_p and (launches_f < launches_p or kind != "c3")


def test_c3_launches_and_traffic(sb, cuda):
    """Config C3: Scheduled(Blended(lerp 0.5, pyramid, perlin)) on 16x16x128x128 is the coarse-level fill, the Perlin
    tables, ONE fused element-wise launch and the normalisation: 4 launches (5 materialising ones before, 3 of them
    full-size passes)."""
    x = torch.zeros(16, 16, 128, 128, device=cuda)
    ns = c3_graph(sb).make_noise_sampler(x, torch.tensor(0.03), torch.tensor(14.6), seed=0)
    torch.manual_seed(0)
    ns(torch.tensor(5.0), torch.tensor(4.5))
    launches = sb.ops.LAUNCH_COUNT
    out = ns(torch.tensor(5.0), torch.tensor(4.5))
    assert sb.ops.LAUNCH_COUNT - launches == 4
    assert abs(float(out.std()) - 1.0) < 1e-3 and abs(float(out.mean())) < 1e-3


@pytest.mark.parametrize("mode", ["inject", "subtract_b"])
def test_fused_blend_modes_and_nearest(sb, cuda, monkeypatch, mode):
    x = torch.zeros(2, 4, 48, 40, device=cuda)

    def make():
        ng = sb.noise_graph
        blended = ng.CustomNoiseChain()
        pyr = ng.CustomNoiseChain()
        pyr.add(ng.AdvancedPyramidNoise(1.0, variant="pyramid", discount=0.6, iterations=6, upscale_mode="nearest-exact"))
        blended.add(ng.BlendedNoise(1.0, normalize=None, blend_function=sb.hostutil.BLENDING_MODES[mode],
                                    custom_noise_1=chain_of(sb, "perlin"), custom_noise_2=pyr, noise_2_percent=0.3))
        return blended.make_noise_sampler(x, None, None, seed=0, cpu=True, normalized=True)

    res = both_ways(sb, monkeypatch, make)
    assert res[True][1] == res[False][1] and res[True][2] == res[False][2]
    for a, b in zip(res[True][0], res[False][0]):
        assert_close(a, b, what=f"blend {mode}", rtol=1e-6, atol=2e-6)
