"""The statistics scale_noise needs from a power-noise sample follow from its half spectrum.

PowerNoiseItem (reference py/nodes/powernoise.py:355-366) returns irfft2(noise_rfft * filter, s=(H, W), norm="ortho") of
a NON-Hermitian half spectrum and scale_noise (py/utils.py:85-106) normalises it with the sample's global mean and std.
torch's irfft2 runs a complex inverse transform along H and a c2r transform along W that ignores the imaginary parts
of the k = 0 and k = W/2 columns after the column transform, i.e. it uses the Hermitian-symmetrised version of those
two columns. With that, Parseval gives both sums without transforming anything:

    sum(x)   = sqrt(H W) * Re Y[0, 0]
    sum(x^2) = sum_ky |Ysym[ky, 0]|^2 + |Ysym[ky, M]|^2 + 2 sum_{0 < kx < M} |Y[ky, kx]|^2,   Ysym[ky] = (Y[ky] + conj Y[-ky]) / 2

(CPU only: this pins the identity a future kernel may rely on to fuse the sampler step into the FFT epilogue.)"""
from __future__ import annotations

import math

import pytest
import torch


def parseval_sums(spectrum: torch.Tensor, height: int, width: int) -> tuple[torch.Tensor, torch.Tensor]:
    m = width // 2
    mirror = (-torch.arange(height)) % height

    def symmetrised(col):
        return 0.5 * (col + col[:, mirror].conj())

    s1 = math.sqrt(height * width) * spectrum[:, 0, 0].real
    s2 = (
        (symmetrised(spectrum[:, :, 0]).abs() ** 2).sum(1)
        + (symmetrised(spectrum[:, :, m]).abs() ** 2).sum(1)
        + 2.0 * (spectrum[:, :, 1:m].abs() ** 2).sum((1, 2))
    )
    return s1, s2


@pytest.mark.parametrize("hw", [(90, 160), (64, 64), (18, 20), (7, 12), (33, 90)])
def test_sums_of_irfft2_follow_from_the_half_spectrum(hw):
    height, width = hw
    torch.manual_seed(height * 1000 + width)
    spectrum = torch.randn(3, height, width // 2 + 1, dtype=torch.complex128)
    gain = torch.rand(height, width // 2 + 1, dtype=torch.float64) + 0.1
    gain[0, 0] = 0.0 if height % 2 else gain[0, 0]  # both a zeroed and a live DC bin
    shaped = spectrum * gain
    x = torch.fft.irfft2(shaped, s=(height, width), norm="ortho")
    s1, s2 = parseval_sums(shaped, height, width)
    torch.testing.assert_close(x.sum((1, 2)), s1, rtol=1e-12, atol=1e-11)
    torch.testing.assert_close((x * x).sum((1, 2)), s2, rtol=1e-12, atol=1e-11)
