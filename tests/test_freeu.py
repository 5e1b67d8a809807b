"""FreeU-Extreme: the oracle against the fixtures recorded from the reference
(CPU), and the CUDA path against the same fixtures (GPU). tests/golden/freeu.pt holds ffilter outputs,
FreeUExtremeConfig.apply outputs and the patched-model handler outputs of the reference node."""
from __future__ import annotations

import pytest
import torch

from helpers import assert_close
from oracle import sonar_oracle as orc

FFILTER = ("lowpass_16x20", "band_32x32", "odd_15x18")
APPLY = ("v2_full", "plain_scale", "slice_lerp", "slice_inject", "subtract")


@pytest.mark.parametrize("name", FFILTER)
def test_oracle_ffilter(golden, name):
    case = golden("freeu")["ffilter"][name]
    assert_close(orc.freeu_ffilter(case["x"], case["filter"], case["norm"]), case["out"], what=name)


@pytest.mark.parametrize("name", APPLY)
def test_oracle_apply(golden, name):
    case = golden("freeu")["apply"][name]
    got = orc.freeu_apply(case["x"], filter=case["filter"], **case["config"])
    assert_close(got, case["out"], what=name)


def test_config_list_and_matching(sb):
    fx = sb.freeu
    tail = fx.FreeUExtremeConfig(target="skip", stage_2=True, blend=0.0)  # dropped: blend == 0
    mid = fx.FreeUExtremeConfig(target="both", stage_3=True, start=0.25, end=0.75, frux_config_opt=tail)
    head = fx.FreeUExtremeConfig(target="backbone", stage_1=True, frux_config_opt=mid)
    assert head.get_config_list() == [mid, head]
    assert mid.check_match(0.5, 3) and mid.check_match(0.5, 3, is_skip=True)
    assert not mid.check_match(0.8, 3) and not mid.check_match(0.5, 1)
    assert head.check_match(0.0, 1) and not head.check_match(0.0, 1, is_skip=True)
    clone = head.clone()
    assert clone is not head and clone.frux_config is mid and clone.scale == head.scale


@pytest.mark.gpu
@pytest.mark.parametrize("name", FFILTER)
def test_gpu_ffilter(sb, cuda, golden, name):
    case = golden("freeu")["ffilter"][name]
    cache = {}
    pf = sb.spectral_noise.PowerFilter(**case["filter"])
    got = sb.freeu.ffilter(case["x"].to(cuda), pf, normalization_factor=case["norm"], cfg_idx=0, filter_cache=cache)
    assert_close(got, case["out"], what=name)
    assert list(cache) == [(0, case["x"].shape[-2:])]
    # the cached filter is reused, and the call works without a cache as well
    again = sb.freeu.ffilter(case["x"].to(cuda), pf, normalization_factor=case["norm"], cfg_idx=0, filter_cache=cache)
    assert torch.equal(again, got)
    assert_close(sb.freeu.ffilter(case["x"].to(cuda), pf, normalization_factor=case["norm"]), case["out"], what=f"{name} no cache")


@pytest.mark.gpu
@pytest.mark.parametrize("name", APPLY)
def test_gpu_apply(sb, cuda, golden, name):
    case = golden("freeu")["apply"][name]
    pf = None if case["filter"] is None else sb.spectral_noise.PowerFilter(**case["filter"])
    cfg = sb.freeu.FreeUExtremeConfig(
        target="backbone", stage_1=True, stage_2=True, stage_3=True, sonar_power_filter_opt=pf, **case["config"],
    )
    x = case["x"].to(cuda)
    got = cfg.apply(0, x, {})
    assert got is x  # in place, like the reference's slice assignment
    assert_close(got, case["out"], what=name)
    # half-precision activations go through fp32 and come back in their own dtype
    xh = case["x"].to(cuda, torch.float16)
    want = orc.freeu_apply(xh.float().cpu(), filter=case["filter"], **case["config"])
    goth = cfg.apply(0, xh, {})
    assert goth.dtype == torch.float16
    assert_close(goth.float(), want, what=f"{name} fp16", rtol=2e-3, atol=2e-3)


@pytest.mark.gpu
def test_gpu_get_scale_matches_oracle(sb, cuda):
    torch.manual_seed(5)
    h = torch.randn(3, 10, 12, 18)
    cfg = sb.freeu.FreeUExtremeConfig(target="backbone", scale=1.4, hidden_mean=True)
    assert_close(cfg.get_scale(h.to(cuda)), orc.freeu_scale(h, 1.4, True), what="hidden mean scale")
    assert sb.freeu.FreeUExtremeConfig(target="backbone", scale=0.7, hidden_mean=False).get_scale(h.to(cuda)) == 0.7
    hu = torch.randn(2, 5, 7, 9)  # hw not a multiple of 4: scalar kernels
    assert_close(cfg.get_scale(hu.to(cuda)), orc.freeu_scale(hu, 1.4, True), what="hidden mean scale (unaligned)")


@pytest.mark.gpu
def test_gpu_node_handlers_golden(sb, cuda, golden):
    """FreeUExtremeNode.go on a stand-in model: stage lookup, percent window, final / stacked configs."""
    rec = golden("freeu")["node"]
    fx, pf = sb.freeu, sb.spectral_noise.PowerFilter

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
        sonar_power_filter_opt=pf(alpha=1.0), filter_norm=1.0,
    )
    first = fx.FreeUExtremeConfig(
        target="backbone", stage_1=True, stage_2=True, start=0.2, end=0.9, scale=1.25, hidden_mean=True, final=False,
        sonar_power_filter_opt=pf(alpha=0.5), filter_norm=1.0, frux_config_opt=second,
    )
    (model,) = fx.FreeUExtremeNode.go(_Model(), False, input_config=first, middle_config=first, output_config=first)
    inp = {k: v.to(cuda) for k, v in rec["inputs"].items()}
    for call in rec["calls"]:
        topt = {"sigmas": torch.tensor([call["sigma"], call["sigma"]], device=cuda)}
        a = model.patches["input"](inp["h16"].clone(), topt)
        b = model.patches["middle_block_patch"](inp["h5"].clone(), topt)
        c, d = model.patches["output"](inp["h8"].clone(), inp["hsp8"].clone(), topt)
        for got, key in ((a, "input"), (b, "middle"), (c, "out_h"), (d, "out_hsp")):
            assert_close(got, call[key], what=f"sigma {call['sigma']} {key}")


@pytest.mark.gpu
def test_gpu_freeu_unet_sized_activation(sb, cuda):
    """SDXL stage-1 sized activation (2 x 1280 x 32 x 32): several planes per CTA in the spectral kernel."""
    torch.manual_seed(11)
    x = torch.randn(2, 1280, 32, 32)
    kw = {"scale": 1.3, "hidden_mean": True, "slice": 0.5, "slice_offset": 0.0, "blend": 1.0}
    cfg = sb.freeu.FreeUExtremeConfig(
        target="backbone", stage_1=True, sonar_power_filter_opt=sb.spectral_noise.PowerFilter(alpha=1.0), filter_norm=1.0, **kw,
    )
    got = cfg.apply(0, x.to(cuda), {})
    want = orc.freeu_apply(x, filter={"alpha": 1.0}, filter_norm=1.0, **kw)
    assert_close(got, want, what="unet stage 1")
