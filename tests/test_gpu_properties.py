"""Size-independent properties at BASELINE.json's full sizes (where the CPU oracle would take too long
to be the checker) and the edge cases: empty, single-element, ragged and unaligned inputs."""
from __future__ import annotations

import math

import pytest
import torch

from helpers import assert_close

pytestmark = pytest.mark.gpu

C5_SHARD = (1, 16, 33, 90, 160)


def test_spectral_round_trip_and_linearity_at_c5_size(sb, cuda):
    """rfft2 -> (unit gain) -> irfft2 is the identity; the shaped transform is linear in its input."""
    torch.manual_seed(0)
    planes = C5_SHARD[1] * C5_SHARD[2]
    x = torch.randn(planes, 90, 160, device=cuda)
    back = sb.ops.spectral_filter(real=x, mask=None, hw=(90, 160), out_scale=1.0 / (90 * 160))
    assert_close(back, x, what="rfft2/irfft2 identity")
    # half-spectrum path (the batched kernel): linear, and equal to torch.fft on a Hermitian-free spectrum
    s1 = torch.randn(planes, 90, 81, dtype=torch.complex64, device=cuda)
    s2 = torch.randn(planes, 90, 81, dtype=torch.complex64, device=cuda)
    mask = torch.rand(90, 81, device=cuda) + 0.5
    ortho = 1.0 / math.sqrt(90 * 160)
    f = lambda s: sb.ops.spectral_filter(spectrum=s, mask=mask, hw=(90, 160), out_scale=ortho)  # noqa: E731
    assert_close(f(s1 * 2.0 + s2), f(s1) * 2.0 + f(s2), what="linearity", rtol=1e-5, atol=5e-5)
    want = torch.fft.irfft2(s1[:8] * mask, s=(90, 160), norm="ortho")
    assert_close(f(s1)[:8], want, what="vs torch.fft (same device)")


def test_wavelet_perfect_reconstruction_at_c4_size(sb, cuda):
    """Unit band scales: IDWT(DWT(c - u)) + u == c, through the fused single-launch path."""
    torch.manual_seed(1)
    cond, uncond = (torch.randn(64, 128, 128, device=cuda) for _ in range(2))
    bank = sb.ops.make_filters(*sb.wavelets.filter_bank("db2"))
    out = sb.ops.wcfg_fused(cond, uncond, bank, levels=3, mode="symmetric", use_f64=True, scale_ll=1.0,
                            scale_hi=[[1.0] * 3] * 3, addend=uncond, addend_scale=1.0)
    assert_close(out, cond, what="perfect reconstruction", rtol=1e-6, atol=1e-6)
    # homogeneity: scaling every band by s scales the reconstructed difference by s
    s = 3.5
    scaled = sb.ops.wcfg_fused(cond, uncond, bank, levels=3, mode="symmetric", use_f64=True, scale_ll=s,
                               scale_hi=[[s] * 3] * 3)
    assert_close(scaled, (cond - uncond) * s, what="equal scales", rtol=1e-5, atol=1e-5)


def test_step_linearity_and_noise_statistics_at_c5_size(sb, cuda):
    """The fused step is affine in (x, denoised, history); the regenerated, normalised Gaussian noise it
    injects has |mean| and |1 - std| inside scale_noise's own threshold."""
    torch.manual_seed(2)
    n = math.prod(C5_SHARD)
    x, den, hist = (torch.randn(C5_SHARD, device=cuda) for _ in range(3))

    def step(xv, dv, hv, noise_kw=None):
        s = sb.samplers.SonarBase(sb.samplers.SonarConfig())
        s.history_d = hv.clone()
        out = s.fused_step(3, xv, dv, 5.0, kind=sb.ops.STEP_EULER, c0=-1.5, noise_scale=1.0, noise_philox=noise_kw)
        return out, s.history_d

    a, ha = step(x, den, hist)
    b, hb = step(x * 2, den * 2, hist * 2)
    assert_close(b, a * 2, what="homogeneity x", rtol=1e-5, atol=1e-5)
    assert_close(hb, ha * 2, what="homogeneity history", rtol=1e-5, atol=1e-5)
    draw = sb.ops.reserve_draw(n, cuda)
    sums = sb.ops.philox_normal_moments_batch(draw, [draw.offset], begin=0, count=n, device=cuda)
    dec = sb.ops.norm_decisions(sums, n)
    kw = {"draw": draw, "factor": 1.0, "normalized": True, "begin": 0, "sums": (sums, dec), "sums_ptr": sums.data_ptr(),
          "decision_ptr": dec.data_ptr(), "count": n}
    noisy, _ = step(x, den, hist, kw)
    noise = (noisy - a).double()
    thr = 2.5 / math.sqrt(n)
    assert abs(noise.mean().item()) <= thr * 1.01 + 1e-6
    assert abs(1.0 - noise.std().item()) <= thr * 1.01 + 1e-5


def test_empty_and_tiny_inputs(sb, cuda):
    """n = 0 launches nothing and returns empty tensors; n = 1 goes through the scalar tails."""
    e = torch.zeros(0, 4, 8, 8, device=cuda)
    assert sb.ops.blend(e, e, 0.5).numel() == 0
    assert sb.ops.axpby(e, 1.0, e, 1.0).numel() == 0
    assert sb.hostutil.scale_noise(e).numel() == 0
    assert sb.ops.randn((0, 3), device=cuda).shape == (0, 3)
    assert sb.ops.spectral_filter(spectrum=torch.zeros(0, 8, 5, dtype=torch.complex64, device=cuda), mask=None,
                                  hw=(8, 8), out_scale=1.0).shape == (0, 8, 8)
    assert sb.ops.item_moments(torch.zeros(0, 5, device=cuda)).shape == (0, 2)
    s = sb.samplers.SonarBase(sb.samplers.SonarConfig())
    assert s.fused_step(0, e, e, 1.0, kind=sb.ops.STEP_EULER, c0=-0.5).numel() == 0
    one = torch.tensor([2.0], device=cuda)
    assert_close(sb.ops.blend(one, one * 3, 0.25), torch.tensor([3.0]))
    assert_close(sb.ops.axpby(one, 2.0, one, 0.5), torch.tensor([5.0]))
    # the statistics ring survives empty producers (their slot is handed out again untouched)
    t = sb.ops.blend(torch.ones(5, device=cuda), torch.ones(5, device=cuda) * 3, 0.5)
    torch.testing.assert_close(sb.ops.attached_sums(t).cpu(), torch.tensor([10.0, 20.0], dtype=torch.float64))


def test_ragged_and_unaligned_views(sb, cuda):
    """Odd sizes and 4-byte-aligned (not 16-byte-aligned) base pointers take the scalar paths and
    give the same values as the vector paths."""
    torch.manual_seed(3)
    base = torch.randn(3 * 5 * 7 * 9 + 1, device=cuda)
    a = base[1:].reshape(3, 5, 7, 9)  # storage offset 1 element: not 16-byte aligned
    b = torch.randn(3, 5, 7, 9, device=cuda)
    assert_close(sb.ops.blend(a.contiguous(), b, 0.3), torch.lerp(a, b, 0.3), what="blend unaligned")
    got = sb.hostutil.scale_noise((a * 1.7 + 0.4).contiguous(), 0.9, normalized=True)
    ref = (a * 1.7 + 0.4)
    ref = (ref - ref.mean()) / ref.std() * 0.9
    assert_close(got, ref, what="scale_noise ragged", rtol=1e-5, atol=1e-5)
    up = sb.ops.resample(a.contiguous(), 13, 11, mode="bilinear")
    assert_close(up, torch.nn.functional.interpolate(a, size=(13, 11), mode="bilinear"), what="resample ragged")
