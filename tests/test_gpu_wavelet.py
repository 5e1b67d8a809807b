"""DWT kernels and wavelet CFG vs the oracle restatement (pytorch_wavelets semantics; this part of
the oracle is UNPINNED -- see oracle/sonar_oracle.py) plus size-independent identities."""
from __future__ import annotations

import pytest
import torch

from helpers import assert_close
from oracle import sonar_oracle as orc

pytestmark = pytest.mark.gpu


def _filters(sb, wave):
    dec_lo, dec_hi, rec_lo, rec_hi = sb.wavelets.filter_bank(wave)
    return (list(dec_lo), list(dec_hi), list(rec_lo), list(rec_hi))


@pytest.mark.parametrize("wave,level,shape", [("db2", 3, (2, 3, 128, 128)), ("db4", 5, (1, 4, 128, 128)), ("haar", 2, (1, 2, 33, 90)), ("db3", 2, (2, 1, 17, 23))])
@pytest.mark.parametrize("mode", ["symmetric", "zero", "reflect", "periodic"])
def test_forward_inverse_vs_oracle(sb, cuda, wave, level, shape, mode):
    torch.manual_seed(0)
    x = torch.randn(shape, dtype=torch.float64)
    f = _filters(sb, wave)
    yl_w, yh_w = orc.dwt2_forward(x, f, level, mode)
    wv = sb.wavelets.Wavelet(wave=wave, level=level, mode=mode, dtype=torch.float64)
    yl, yh = wv.forward(x.to(cuda))
    assert_close(yl, yl_w, what="yl", rtol=1e-11, atol=1e-11)
    assert len(yh) == level
    for a, b in zip(yh, yh_w):
        assert_close(a, b, what="yh", rtol=1e-11, atol=1e-11)
    rec = wv.inverse(yl, yh)
    assert_close(rec, orc.dwt2_inverse(yl_w, yh_w, f), what="inverse", rtol=1e-11, atol=1e-11)
    if mode != "zero" or True:
        # perfect reconstruction (crop the symmetric-mode overhang)
        assert_close(rec[..., : shape[-2], : shape[-1]], x, what="PR", rtol=1e-10, atol=1e-10)


def test_fp32_transform(sb, cuda):
    torch.manual_seed(1)
    x = torch.randn(2, 4, 64, 64)
    f = _filters(sb, "db2")
    yl_w, yh_w = orc.dwt2_forward(x.double(), f, 3)
    yl, yh = sb.wavelets.Wavelet(wave="db2", level=3, dtype=torch.float32).forward(x.to(cuda))
    assert yl.dtype == torch.float32
    assert_close(yl.double(), yl_w, what="yl fp32", rtol=1e-5, atol=1e-5)
    assert_close(yh[0].double(), yh_w[0], what="yh fp32", rtol=1e-5, atol=1e-5)


class _MS:
    sigma_min = torch.tensor(0.03)
    sigma_max = torch.tensor(14.6)

    @staticmethod
    def timestep(sigma):
        return (sigma.log() - torch.tensor(0.03).log()) / (torch.tensor(14.6).log() - torch.tensor(0.03).log()) * 999


class _Model:
    model_sampling = _MS()


def _args(cond, uncond, x, sigma=5.0, scale=7.0):
    return {
        "sigma": torch.full((x.shape[0],), sigma, device=x.device), "input": x, "cond_denoised": cond,
        "uncond_denoised": uncond, "cond": x - cond, "uncond": x - uncond, "cond_scale": scale, "model": _Model(),
        "model_options": {},
    }  # fmt: skip


C4_RULE = {"wave": "db2", "level": 3, "diff": {"yl_scale": 5, "yh_scales": [[3, 4, 5], [3, 4, 5], [3, 4, 5]]}}


def test_c4_wavelet_cfg_vs_oracle(sb, cuda):
    """BASELINE.json config C4: db2, 3 levels, separate H/V scales, SDXL batch 16x4x128x128."""
    torch.manual_seed(4)
    cond, uncond, x = (torch.randn(16, 4, 128, 128) for _ in range(3))
    want = orc.wavelet_cfg(cond, uncond, x, _filters(sb, "db2"), level=3, diff=(5, [[3, 4, 5]] * 3))
    fn = sb.wcfg.WaveletCFG(existing_cfg=None, rules=sb.wcfg.WCFGRules.build(**C4_RULE))
    got = fn(_args(cond.to(cuda), uncond.to(cuda), x.to(cuda)))
    assert got.is_contiguous() and got.dtype == torch.float32 and got.shape == x.shape
    assert_close(got, want, what="C4")


@pytest.mark.parametrize(
    "rule,kw",
    [
        ({"wave": "db4", "level": 5, "diff": {"yl_scale": 2.0, "yh_scales": 3.0}}, {"diff": (2.0, 3.0), "level": 5, "wave": "db4"}),
        (
            {"wave": "db2", "level": 2, "cond": {"yl_scale": 1.5, "yh_scales": [1.0, 2.0]}, "uncond": {"yl_scale": 0.5, "yh_scales": 0.75},
             "final": {"yl_scale": 1.1, "yh_scales": [[1.0, 0.9, 0.8]]}, "diff": {"yl_scale": 3.0, "yh_scales": [[2, 3, 4], 1.5]}},
            {"level": 2, "wave": "db2", "cond_scales": (1.5, [1.0, 2.0]), "uncond_scales": (0.5, 0.75), "final": (1.1, [[1.0, 0.9, 0.8]]),
             "diff": (3.0, [[2, 3, 4], 1.5])},
        ),
        (
            {"wave": "haar", "level": 3, "difference_blend_mode": "lerp", "difference_blend_strength": 0.7, "diff": {"yl_scale": 4.0, "yh_scales": 2.0}},
            {"level": 3, "wave": "haar", "difference_blend_mode": "lerp", "difference_blend_strength": 0.7, "diff": (4.0, 2.0)},
        ),
        (
            {"wave": "db2", "level": 3, "high_precision_mode": False, "diff": {"yl_scale": 5.0, "yh_scales": 3.0}},
            {"level": 3, "wave": "db2", "high_precision": False, "diff": (5.0, 3.0)},
        ),
        (
            {"wave": "db2", "level": 2, "target_mode": "noise", "diff": {"yl_scale": 2.0, "yh_scales": 2.5}},
            {"level": 2, "wave": "db2", "diff": (2.0, 2.5), "denoised_target": False, "use_noise": True},
        ),
    ],
)
def test_wavelet_cfg_rule_variants(sb, cuda, rule, kw):
    torch.manual_seed(5)
    shape = (2, 4, 48, 40)
    cond, uncond, x = (torch.randn(shape) for _ in range(3))
    kw = dict(kw)
    wave = kw.pop("wave")
    use_noise = kw.pop("use_noise", False)
    c_in, u_in = (x - cond, x - uncond) if use_noise else (cond, uncond)
    want = orc.wavelet_cfg(c_in, u_in, x, _filters(sb, wave), **kw)
    fn = sb.wcfg.WaveletCFG(existing_cfg=None, rules=sb.wcfg.WCFGRules.build(**rule))
    got = fn(_args(cond.to(cuda), uncond.to(cuda), x.to(cuda)))
    tol = 1e-5 if kw.get("high_precision", True) else 5e-5
    assert_close(got, want, what=str(rule)[:60], rtol=tol, atol=tol)


def test_equal_scales_is_plain_cfg(sb, cuda):
    """All band scales == s  =>  uncond + s*(cond - uncond) by linearity."""
    torch.manual_seed(6)
    cond, uncond, x = (torch.randn(2, 4, 64, 64, device=cuda) for _ in range(3))
    s = 6.5
    fn = sb.wcfg.WaveletCFG(existing_cfg=None, rules=sb.wcfg.WCFGRules.build(wave="db2", level=3, diff={"yl_scale": s, "yh_scales": s}))
    got = fn(_args(cond, uncond, x))
    assert_close(got, x - (uncond + s * (cond - uncond)), what="equal scales")


def test_video_5d_and_odd_sizes(sb, cuda):
    torch.manual_seed(7)
    shape = (1, 4, 3, 33, 45)
    cond, uncond, x = (torch.randn(shape) for _ in range(3))
    want = orc.wavelet_cfg(cond, uncond, x, _filters(sb, "db2"), level=3, diff=(5, [[3, 4, 5]] * 3))
    fn = sb.wcfg.WaveletCFG(existing_cfg=None, rules=sb.wcfg.WCFGRules.build(**C4_RULE))
    got = fn(_args(cond.to(cuda), uncond.to(cuda), x.to(cuda)))
    assert_close(got, want, what="5d odd")


def test_rule_out_of_range_falls_back(sb, cuda):
    torch.manual_seed(8)
    cond, uncond, x = (torch.randn(1, 4, 16, 16, device=cuda) for _ in range(3))
    fn = sb.wcfg.WaveletCFG(existing_cfg=None, rules=sb.wcfg.WCFGRules.build(start_sigma=3.0, end_sigma=1.0, **C4_RULE))
    got = fn(_args(cond, uncond, x, sigma=5.0, scale=7.0))
    assert_close(got, x - ((cond - uncond) * 7.0 + uncond), what="fallback cfg")


@pytest.mark.parametrize(
    "rule,shape",
    [
        (C4_RULE, (3, 4, 128, 128)),    # 12 planes: two CTAs (a cluster) per plane
        (C4_RULE, (20, 4, 128, 128)),   # 80 planes: more than half the SMs -> one CTA per plane
        (C4_RULE, (5, 3, 127, 90)),     # cluster path, odd height: uneven halves
        ({"wave": "db4", "level": 3, "diff": {"yl_scale": 2.0, "yh_scales": [[1, 2, 3], 2.5, 0.5]}}, (2, 3, 33, 47)),
        ({"wave": "haar", "level": 4, "high_precision_mode": False, "diff": {"yl_scale": 4.0, "yh_scales": 2.0}}, (1, 4, 64, 96)),
        ({"wave": "db2", "level": 2, "padding_mode": "periodic", "diff": {"yl_scale": 3.0, "yh_scales": 1.5}}, (2, 2, 40, 24)),
        ({"wave": "db3", "level": 1, "padding_mode": "reflect", "diff": {"yl_scale": 3.0, "yh_scales": 1.5}}, (150, 1, 18, 20)),
    ],
)
def test_fused_wcfg_launch_equals_per_level_path(sb, cuda, rule, shape, monkeypatch):
    """The one-launch, all-in-shared-memory path computes exactly what the per-level kernels do."""
    torch.manual_seed(9)
    cond, uncond, x = (torch.randn(shape, device=cuda) for _ in range(3))
    fn = sb.wcfg.WaveletCFG(existing_cfg=None, rules=sb.wcfg.WCFGRules.build(**rule))
    seen = []
    real = sb.ops.wcfg_fused
    monkeypatch.setattr(sb.ops, "wcfg_fused", lambda *a, **k: (seen.append(1), real(*a, **k))[1])
    fused = fn(_args(cond, uncond, x))
    assert seen, "fused path not taken"
    monkeypatch.setattr(sb.ops, "wcfg_fused_fits", lambda *a, **k: False)
    per_level = fn(_args(cond, uncond, x))
    assert len(seen) == 1
    tol = 1e-6 if rule.get("high_precision_mode", True) else 2e-5
    assert_close(fused, per_level, what="fused vs per-level", rtol=tol, atol=tol)


def test_fused_wcfg_declines_what_does_not_fit(sb, cuda):
    """256x256 fp64 planes exceed one SM's shared memory: the per-level path runs (and is correct)."""
    assert sb.ops.wcfg_fused_fits(128, 128, 4, 3, use_f64=True)
    assert not sb.ops.wcfg_fused_fits(256, 256, 4, 3, use_f64=True)
    torch.manual_seed(10)
    cond, uncond, x = (torch.randn(1, 2, 256, 256) for _ in range(3))
    want = orc.wavelet_cfg(cond, uncond, x, _filters(sb, "db2"), level=3, diff=(5, [[3, 4, 5]] * 3))
    fn = sb.wcfg.WaveletCFG(existing_cfg=None, rules=sb.wcfg.WCFGRules.build(**C4_RULE))
    got = fn(_args(cond.to(cuda), uncond.to(cuda), x.to(cuda)))
    assert_close(got, want, what="256x256")
