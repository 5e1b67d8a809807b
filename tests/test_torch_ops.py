"""The torch.library custom-op layer (north_star: "thin C-ABI / torch.library custom-op layer")."""
from __future__ import annotations

import pytest
import torch

from helpers import assert_close
from oracle import sonar_oracle as orc


def test_ops_are_registered_for_cuda_only(sb):
    names = sb.torch_ops.OP_NAMES
    assert {"step", "philox_normal_", "spectral_filter", "wcfg_fused", "scale_noise", "channel_mix", "pyramid_accum"} <= set(names)
    for name in names:
        assert hasattr(torch.ops.sonar_b200, name)
        assert torch._C._dispatch_has_kernel_for_dispatch_key(f"sonar_b200::{name}", "CUDA")  # noqa: SLF001
        assert not torch._C._dispatch_has_kernel_for_dispatch_key(f"sonar_b200::{name}", "CPU")  # noqa: SLF001


def test_cpu_tensors_are_rejected_by_the_dispatcher(sb):
    """No CPU fallback: the dispatcher has nothing to run for CPU tensors."""
    with pytest.raises((NotImplementedError, RuntimeError)):
        torch.ops.sonar_b200.moments(torch.zeros(8))
    with pytest.raises((NotImplementedError, RuntimeError)):
        torch.ops.sonar_b200.scale_noise(torch.zeros(8), 1.0, True)


@pytest.mark.gpu
def test_step_op_matches_oracle(sb, cuda):
    torch.manual_seed(0)
    x, den, hist, nz = (torch.randn(2, 4, 32, 32) for _ in range(4))
    o = orc.SonarOracle()
    o.hist = hist.clone()
    sigma, sigma_next = torch.tensor(5.0), torch.tensor(3.0)
    want = o.euler_ancestral(3, x, den, sigma, sigma_next, nz)
    sd, su = orc.get_ancestral_step(sigma, sigma_next, 1.0)
    got, h = torch.ops.sonar_b200.step(
        x.to(cuda), den.to(cuda), hist.to(cuda), nz.to(cuda), sb.ops.STEP_EULER, sb.ops.MODE_NEW, 0.95, 0.75, 1.0, 5.0,
        float(sd - sigma), 0.0, float(su),
    )
    assert_close(got, want, what="step op x")
    assert_close(h, o.hist, what="step op history")


@pytest.mark.gpu
def test_philox_and_scale_noise_ops(sb, cuda):
    torch.manual_seed(42)
    want = torch.randn(3, 5, 64, 64, device=cuda)
    out = torch.empty(3, 5, 64, 64, device=cuda)
    torch.ops.sonar_b200.philox_normal_(out, 42, 0)
    assert torch.equal(out, want)
    torch.manual_seed(42)
    assert torch.equal(torch.ops.sonar_b200.randn_like(out), want)
    raw = want * 1.7 + 0.3
    assert_close(torch.ops.sonar_b200.scale_noise(raw, 2.0, True), orc.scale_noise(raw.cpu().clone(), 2.0, normalized=True), what="scale_noise op")
    sums = torch.ops.sonar_b200.moments(raw)
    torch.testing.assert_close(sums.cpu(), torch.stack((raw.double().sum(), raw.double().square().sum())).cpu(), rtol=1e-9, atol=1e-6)


@pytest.mark.gpu
def test_spectral_and_mixer_ops(sb, cuda):
    torch.manual_seed(1)
    spec = torch.randn(6, 64, 33, dtype=torch.complex64)
    mask = torch.rand(64, 33) + 0.5
    want = torch.fft.irfft2(spec * mask, s=(64, 64), norm="ortho")
    got = torch.ops.sonar_b200.spectral_filter(None, spec.to(cuda), mask.to(cuda), 64, 64, 1.0 / 64.0)
    assert_close(got, want, what="spectral_filter op")
    noise = torch.randn(2, 12, 8, 8)
    mixer = orc.channel_mixer(12, 0.2, "1, 0.5").contiguous()
    assert_close(torch.ops.sonar_b200.channel_mix(noise.to(cuda), mixer.to(cuda)), orc.channel_mix(noise, mixer), what="channel_mix op")


@pytest.mark.gpu
def test_wcfg_fused_op(sb, cuda):
    torch.manual_seed(2)
    cond, uncond = torch.randn(8, 64, 64), torch.randn(8, 64, 64)
    filt = orc.db_filters(sb.wavelets.DB_DEC_LO[2])
    dec_lo, dec_hi, rec_lo, rec_hi = filt
    scales = [3.0, 4.0, 5.0] * 2
    got = torch.ops.sonar_b200.wcfg_fused(
        cond.to(cuda), uncond.to(cuda), dec_lo, dec_hi, rec_lo, rec_hi, 2, "symmetric", True, 5.0, scales, None, 1.0, None, 0.0, 1.0,
    )
    yl, yh = orc.dwt2_forward((cond - uncond).double()[None], filt, 2)
    yh = [b * torch.tensor([3.0, 4.0, 5.0], dtype=torch.float64).view(1, 1, 3, 1, 1) for b in yh]
    want = orc.dwt2_inverse(yl * 5.0, yh, filt)[0, :, :64, :64].float()
    assert_close(got, want, what="wcfg_fused op")
