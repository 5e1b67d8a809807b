"""Wavelet-filtered noise: the non-expansive "periodization" DWT and
WaveletFilteredNoiseGenerator. The reference delegates the transform to pytorch_wavelets (absent from
the reference tree and from this image, no pinned version): PARITY UNPINNED -- the oracle restates the
library's published algorithm and is anchored on size-independent identities (perfect reconstruction,
orthonormality, the haar block transform); the CUDA path is checked against the oracle on the same draws."""
from __future__ import annotations

import pytest
import torch

from helpers import assert_close
from oracle import sonar_oracle as orc

CASES = [("haar", (16, 24)), ("db2", (16, 24)), ("db4", (32, 40)), ("db2", (15, 21)), ("haar", (7, 9)), ("db3", (18, 20))]


def _filters(sb, wave):
    dec_lo, dec_hi, rec_lo, rec_hi = sb.wavelets.filter_bank(wave)
    return list(dec_lo), list(dec_hi), list(rec_lo), list(rec_hi)


@pytest.mark.parametrize(("wave", "hw"), CASES)
def test_oracle_periodization_identities(sb, wave, hw):
    torch.manual_seed(1)
    f = _filters(sb, wave)
    h, w = hw
    x = torch.randn(2, 3, h, w, dtype=torch.float64)
    yl, yh = orc.dwt2_forward(x, f, 2, "periodization")
    assert yl.shape[-2:] == ((((h + 1) // 2) + 1) // 2, (((w + 1) // 2) + 1) // 2)
    assert [tuple(b.shape[-2:]) for b in yh][0] == ((h + 1) // 2, (w + 1) // 2)
    rec = orc.dwt2_inverse(yl, yh, f, "periodization")
    assert_close(rec[..., :h, :w], x, what="perfect reconstruction", rtol=1e-12, atol=1e-12)
    if h % 4 == 0 and w % 4 == 0:  # even at every level: the transform is orthonormal
        energy = yl.square().sum() + sum(b.square().sum() for b in yh)
        assert_close(energy, x.square().sum(), what="Parseval", rtol=1e-12, atol=1e-12)


def test_oracle_haar_periodization_is_the_block_transform(sb):
    torch.manual_seed(2)
    x = torch.randn(1, 2, 6, 8, dtype=torch.float64)
    yl, yh = orc.dwt2_forward(x, _filters(sb, "haar"), 1, "periodization")
    a, b, c, d = x[..., 0::2, 0::2], x[..., 0::2, 1::2], x[..., 1::2, 0::2], x[..., 1::2, 1::2]
    assert_close(yl, (a + b + c + d) / 2, what="ll", rtol=1e-12, atol=1e-12)
    assert_close(yh[0][:, :, 2], (a - b - c + d) / 2, what="hh", rtol=1e-12, atol=1e-12)


@pytest.mark.gpu
@pytest.mark.parametrize(("wave", "hw"), CASES)
@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
def test_gpu_periodization_dwt_vs_oracle(sb, cuda, wave, hw, dtype):
    torch.manual_seed(3)
    h, w = hw
    x = torch.randn(2, 3, h, w, dtype=dtype)
    wl = sb.wavelets.Wavelet(wave=wave, level=2, mode="periodization", dtype=dtype)
    yl, yh = wl.forward(x.to(cuda))
    want_yl, want_yh = orc.dwt2_forward(x.double(), _filters(sb, wave), 2, "periodization")
    tol = {"rtol": 1e-5, "atol": 1e-5} if dtype == torch.float32 else {"rtol": 1e-11, "atol": 1e-11}
    assert_close(yl.double(), want_yl, what="yl", **tol)
    for got, want in zip(yh, want_yh):
        assert_close(got.double(), want, what="yh", **tol)
    rec = wl.inverse(yl, yh)
    assert_close(rec[..., :h, :w].double(), x.double(), what="round trip", **tol)
    # scales folded into the synthesis loads == scaling the coefficients first
    scaled = wl.inverse(yl, yh, yl_scale=0.5, yh_scales=[[1.5, 0.25, 2.0], 0.75])
    syl, syh = sb.wavelets.wavelet_scaling(yl, yh, 0.5, [[1.5, 0.25, 2.0], 0.75])
    assert_close(scaled, wl.inverse(syl, syh), what="folded scales", **tol)


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(2, 4, 32, 48), (1, 3, 17, 30), (1, 2, 3, 16, 24)])
def test_gpu_wavelet_filtered_noise_vs_oracle(sb, cuda, shape):
    """Default generator (haar, 3 levels, periodization) with band scales, fed an injected Gaussian draw."""
    torch.manual_seed(4)
    base = torch.randn(shape)
    x = torch.zeros(shape, device=cuda)
    yh_scales = [[1.5, 0.5, 2.0], 0.25, "fill"]
    with sb.rng.injected([base]):
        gen = sb.generators.WaveletFilteredNoiseGenerator(x, normalized=False, yl_scale=0.3, yh_scales=yh_scales)
        got = gen()
    planes = base.reshape(shape[0], -1, *shape[-2:]).double()
    f = _filters(sb, "haar")
    yl, yh = orc.dwt2_forward(planes, f, 3, "periodization")
    scales = sb.wavelets.expand_yh_scales([tuple(b.shape) for b in yh], yh_scales=yh_scales)
    yh = [b * torch.tensor(sc, dtype=torch.float64).view(1, 1, 3, 1, 1) for b, sc in zip(yh, scales)]
    want = orc.dwt2_inverse(yl * 0.3, yh, f, "periodization")[..., : shape[-2], : shape[-1]].reshape(shape)
    assert got.shape == x.shape and got.dtype == x.dtype
    assert_close(got, want.float(), what=f"filtered noise {shape}")
    # all scales 1: the filter is the identity (perfect reconstruction)
    with sb.rng.injected([base]):
        ident = sb.generators.WaveletFilteredNoiseGenerator(x, normalized=False)()
    assert_close(ident, base, what="identity filter")


@pytest.mark.gpu
def test_gpu_wavelet_filtered_noise_item_and_node(sb, cuda):
    """Chain item with separate low / high sources (blended band-wise) through the node, symmetric db2 variant."""
    ng = sb.noise_graph
    low = ng.CustomNoiseChain()
    low.add(ng.CustomNoiseItem(1.0, noise_type="gaussian"))
    (chain,) = sb.nodes.SonarWaveletFilteredNoiseNode().go(
        factor=1.0, rescale=0.0, normalize="disabled", normalize_noise=False, custom_noise=low,
        yaml_parameters="wave: db2\nlevel: 2\nmode: symmetric\nyl_scale: 2.0\nyh_scales: [0.5, 0.25]\n",
    )
    item = chain.items[0]
    assert type(item).__name__ == "WaveletFilteredNoise" and item.noise_high is not item.noise
    clone = chain.clone().items[0]
    assert clone.noise is not item.noise and clone.ns_kwargs == item.ns_kwargs
    torch.manual_seed(5)
    shape = (1, 2, 20, 28)
    n_low, n_high = torch.randn(shape), torch.randn(shape)
    x = torch.zeros(shape, device=cuda)
    with sb.rng.injected([n_low, n_high]):
        got = chain.make_noise_sampler(x, None, None, seed=0, normalized=False)(None, None)
    f = _filters(sb, "db2")
    yl, _ = orc.dwt2_forward(n_low.double(), f, 2, "symmetric")
    _, yh = orc.dwt2_forward(n_high.double(), f, 2, "symmetric")  # yl from the low source, yh from the high one
    yh = [yh[0] * 0.5, yh[1] * 0.25]
    want = orc.dwt2_inverse(yl * 2.0, yh, f)[..., :20, :28]
    assert_close(got, want.float(), what="item")
    with pytest.raises(NotImplementedError):
        sb.generators.WaveletFilteredNoiseGenerator(x, use_dtcwt=True)
