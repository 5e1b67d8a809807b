"""Host logic of the batched spectral kernel, checked without a GPU.

`sonar_spectral_plan` (C ABI, host only) reports the plan `sonar_spectral_filter_f32` would use. The tests
check the plan itself (radices multiply to the axis length, shared memory fits, CTAs per SM x threads fill the
register file) and then run a numpy MODEL of the kernel's algorithm with exactly those radices -- in-place
decimation-in-time / decimation-in-frequency stages on digit-reversed slots, the half-length Hermitian pair pass,
the conjugation trick for the forward transform -- against numpy.fft for every size class the planner produces.
The model is test infrastructure (it restates csrc/spectral.cu's index arithmetic); the CUDA kernel itself is
checked against torch.fft in tests/test_gpu_noise.py."""
from __future__ import annotations

import ctypes

import numpy as np
import pytest


def _plan(sb, h, w, planes=64, real=False):
    info = sb._native.SonarSpectralPlanInfo()
    lib = sb._native.load(build_if_missing=False)
    assert lib.sonar_spectral_plan(h, w, planes, int(real), ctypes.byref(info)) == 0
    return info


def _radices(info):
    return list(info.col_radix[: info.n_col_stages]), list(info.row_radix[: info.n_row_stages])


# ---------------------------------------------------------------------------------------------
# numpy model of one axis (inverse-sign, unnormalised), in place on `buf` along axis 0
# ---------------------------------------------------------------------------------------------
def _sub_lengths(n, radices):
    m, block = [], n
    for r in radices:
        block //= r
        m.append(block)
    return m


def _positions(n, radices):
    """pos[k] = slot of index k after DIF / slot DIT reads it from: sum of digit_f * m_f, k = d0 + R0 (d1 + R1 ...)."""
    m = _sub_lengths(n, radices)
    pos = np.zeros(n, dtype=np.int64)
    for k in range(n):
        rem, slot = k, 0
        for r, mf in zip(radices, m):
            slot += (rem % r) * mf
            rem //= r
        pos[k] = slot
    return pos


def _stage(buf, n, radices, f, *, dit):
    r, m = radices[f], _sub_lengths(n, radices)[f]
    block = m * r
    t = np.arange(r)
    dft = np.exp(2j * np.pi * np.outer(t, t) / r)  # inverse-sign butterfly
    for j in range(n // r):
        q, i = divmod(j, m)
        slots = q * block + i + t * m
        tw = np.exp(2j * np.pi * i * t / block).reshape(-1, *([1] * (buf.ndim - 1)))
        v = buf[slots]
        if dit:
            v = v * tw
        v = np.tensordot(dft, v, axes=(1, 0))
        if not dit:
            v = v * tw
        buf[slots] = v


def _dit(buf, n, radices):  # digit-reversed slots in -> natural order out
    for f in reversed(range(len(radices))):
        _stage(buf, n, radices, f, dit=True)


def _dif(buf, n, radices):  # natural in -> index k at slot pos[k]
    for f in range(len(radices)):
        _stage(buf, n, radices, f, dit=False)


def _pair_pass(rows, m, w, *, unfold):
    """rows: (..., m + 1) along the LAST axis, natural order, Nyquist at m. In place."""
    out = rows.copy()
    tw = np.exp(2j * np.pi * np.arange(m) / w)
    for k in range(m // 2 + 1):
        if k == 0:
            if unfold:
                y0 = rows[..., 0]
                out[..., 0] = 2 * (y0.real - y0.imag)
                out[..., m] = 2 * (y0.real + y0.imag)
            else:
                out[..., 0] = (rows[..., 0].real + rows[..., m].real) + 1j * (rows[..., 0].real - rows[..., m].real)
            continue
        ak, am = rows[..., k], rows[..., m - k]
        s, d = ak + np.conj(am), ak - np.conj(am)
        wd = tw[k] * d
        out[..., k] = s + 1j * wd
        if 2 * k != m:
            out[..., m - k] = np.conj(s) + 1j * np.conj(wd)
    rows[...] = out


def _model_irfft2(spec, mask, h, w, col_r, row_r):
    """Unnormalised c2r 2-D inverse of a (possibly non-Hermitian) half spectrum, the kernel's way."""
    m = w // 2
    pos_h, pos_m = _positions(h, col_r), _positions(m, row_r)
    buf = np.zeros((h, m + 1), dtype=np.complex128)
    buf[pos_h] = spec * mask  # first column stage gathers row ky into slot pos_h[ky]
    _dit(buf, h, col_r)
    _pair_pass(buf, m, w, unfold=False)
    rows = np.ascontiguousarray(buf[:, :m].T)  # (m, h): transform axis first
    _dif(rows, m, row_r)
    z = rows[pos_m]  # the last row stage stores slot pos_m[n] to output index n
    out = np.empty((h, w))
    out[:, 0::2], out[:, 1::2] = z.real.T, z.imag.T
    return out


def _model_filter_real(x, mask, h, w, col_r, row_r):
    m = w // 2
    pos_h, pos_m = _positions(h, col_r), _positions(m, row_r)
    packed = np.conj(x[:, 0::2] + 1j * x[:, 1::2])  # conj trick: FFT(z) = conj(IFFT(conj z))
    rows = np.zeros((m, h), dtype=np.complex128)
    rows[pos_m] = packed.T  # first forward row stage gathers index n into slot pos_m[n]
    _dit(rows, m, row_r)
    buf = np.zeros((h, m + 1), dtype=np.complex128)
    buf[:, :m] = rows.T
    _pair_pass(buf, m, w, unfold=True)  # 2 conj rfft along W
    _dif(buf, h, col_r)  # 2 conj rfft2, ky at slot pos_h[ky]
    idx_h = np.argsort(pos_h)  # slot -> ky
    buf = np.conj(buf) * mask[idx_h]  # gain on the loads of the inverse column transform
    _dit(buf, h, col_r)
    _pair_pass(buf, m, w, unfold=False)
    rows = np.ascontiguousarray(buf[:, :m].T)
    _dif(rows, m, row_r)
    z = rows[pos_m]
    out = np.empty((h, w))
    out[:, 0::2], out[:, 1::2] = z.real.T, z.imag.T
    return out * 0.5


def _ref_irfft2_unnormalised(spec, h, w):
    """torch.fft.irfft2 semantics for a non-Hermitian half spectrum: complex inverse along H, then c2r along W."""
    cols = np.fft.ifft(spec, axis=0) * h
    return np.fft.irfft(cols, n=w, axis=1) * w


SIZES = [(64, 64), (128, 128), (90, 160), (32, 32), (18, 20), (160, 90), (48, 96), (100, 200), (8, 4), (45, 50),
         (120, 72), (36, 40), (27, 54), (2, 4), (200, 36), (25, 30)]


@pytest.mark.parametrize("hw", SIZES)
def test_plan_properties_and_model_vs_numpy_fft(sb, hw):
    h, w = hw
    info = _plan(sb, h, w)
    assert info.batched == 1, f"{hw} should take the batched kernel"
    col_r, row_r = _radices(info)
    assert int(np.prod(col_r)) == h and int(np.prod(row_r)) == w // 2
    assert set(col_r + row_r) <= {2, 3, 4, 5, 8, 9, 10, 16}
    assert row_r == sorted(row_r), "row axis: largest radix last (bank-friendly global stage)"
    assert info.group >= 1 and info.threads in (256, 320, 512) and 1 <= info.ctas_per_sm <= 4
    assert info.threads * info.ctas_per_sm <= 1024  # 64 registers per thread
    assert (info.smem_bytes + 1024) * info.ctas_per_sm <= 227 * 1024 + 1024
    assert info.grid * info.group >= min(64, 148 * info.ctas_per_sm * info.group)
    rng = np.random.default_rng(h * 1000 + w)
    spec = rng.standard_normal((h, w // 2 + 1)) + 1j * rng.standard_normal((h, w // 2 + 1))
    mask = rng.random((h, w // 2 + 1)) + 0.5
    got = _model_irfft2(spec, mask, h, w, col_r, row_r)
    np.testing.assert_allclose(got, _ref_irfft2_unnormalised(spec * mask, h, w), rtol=1e-9, atol=1e-9 * h * w)
    x = rng.standard_normal((h, w))
    got = _model_filter_real(x, mask, h, w, col_r, row_r) / (h * w)
    want = np.fft.irfft2(np.fft.rfft2(x) * mask, s=(h, w))
    np.testing.assert_allclose(got, want, rtol=1e-9, atol=1e-9)


def test_every_plannable_axis_length_transforms_correctly(sb):
    """1-D check of the in-place DIT / DIF pair for every length <= 256 the planner accepts (W = 2 n)."""
    rng = np.random.default_rng(7)
    planned = 0
    for n in range(2, 257):
        info = _plan(sb, 4, 2 * n)
        rest = n
        for p in (2, 3, 5):
            while rest % p == 0:
                rest //= p
        assert bool(info.batched) == (rest == 1), f"length {n}: only 2^a 3^b 5^c lengths are planned"
        if not info.batched:
            continue
        planned += 1
        _, radices = _radices(info)
        assert int(np.prod(radices)) == n
        x = rng.standard_normal(n) + 1j * rng.standard_normal(n)
        want = np.fft.ifft(x) * n
        pos = _positions(n, radices)
        buf = np.zeros(n, dtype=np.complex128)
        buf[pos] = x
        _dit(buf, n, radices)
        np.testing.assert_allclose(buf, want, rtol=1e-9, atol=1e-9 * n)
        buf = x.copy()
        _dif(buf, n, radices)
        np.testing.assert_allclose(buf[pos], want, rtol=1e-9, atol=1e-9 * n)
    assert planned == 51  # the 5-smooth lengths in [2, 256]


def test_plan_declines_what_the_generic_kernel_handles(sb):
    assert _plan(sb, 31, 47).batched == 0       # odd width
    assert _plan(sb, 64, 2 * 7 * 8).batched == 0  # prime factor 7
    big = _plan(sb, 256, 256, planes=64)
    assert big.batched == 0 and big.cluster == 1  # 264 KB half spectrum: distributed shared memory of a 2-CTA cluster
    assert (big.threads, big.ctas_per_sm, big.grid) == (1024, 1, 128) and big.smem_bytes <= 227 * 1024
    assert _plan(sb, 512, 512).batched == 0 and _plan(sb, 512, 512).cluster == 0  # beyond two SMs: generic kernel
    small = _plan(sb, 32, 32, planes=2560)
    assert small.batched == 1 and small.group > 1, "UNet-sized planes are grouped per CTA"
    assert _plan(sb, 32, 32, planes=4).group == 1, "few planes: one per CTA so more SMs work"
    big = _plan(sb, 90, 160, planes=528, real=True)
    assert big.ctas_per_sm == 3 and big.threads == 320 and big.grid == 444


# ---------------------------------------------------------------------------------------------
# the c2r fold fused into the first row stage (csrc/spectral.cu run_stage_fold)
# ---------------------------------------------------------------------------------------------
def _fold_pair(ak, am, wk):
    s, d = ak + np.conj(am), ak - np.conj(am)
    wd = wk * d
    return s + 1j * wd, np.conj(s) + 1j * np.conj(wd)


def _fused_fold_first_stage(buf, m_len, w, radices):
    """buf: (h, m_len + 1) rows in natural order, Nyquist at slot m_len. Runs the fold and the first DIF row stage
    the way the kernel's items do: butterfly j of that stage owns slots j + t m, whose mirrors are the slots of
    butterfly m - j; one item folds both in 'registers' and runs the two butterflies. In place on buf[:, :m_len]."""
    r = radices[0]
    m = m_len // r
    assert m % 2 == 0 and len(radices) >= 2
    tw_w = np.exp(2j * np.pi * np.arange(m_len) / w)
    t = np.arange(r)
    dft = np.exp(2j * np.pi * np.outer(t, t) / r)

    def finish(row, j, v):
        v = dft @ v
        v = v * np.exp(2j * np.pi * j * t / m_len)  # DIF: twiddle the outputs (block length = the whole row)
        row[j + t * m] = v

    for row in buf:
        for jj in range(m // 2 + 1):
            if jj in (0, m // 2):
                v = row[jj + t * m].copy()
                if jj == 0:
                    x0, xm = row[0].real, row[m_len].real
                    v[0] = (x0 + xm) + 1j * (x0 - xm)
                    for tt in range(1, r):
                        if 2 * tt < r:
                            v[tt], v[r - tt] = _fold_pair(v[tt], v[r - tt], tw_w[tt * m])
                    if r % 2 == 0:
                        v[r // 2] = _fold_pair(v[r // 2], v[r // 2], tw_w[(r // 2) * m])[0]
                else:
                    for tt in range(r):
                        if 2 * tt < r - 1:
                            v[tt], v[r - 1 - tt] = _fold_pair(v[tt], v[r - 1 - tt], tw_w[jj + tt * m])
                    if r % 2 == 1:
                        mid = (r - 1) // 2
                        v[mid] = _fold_pair(v[mid], v[mid], tw_w[jj + mid * m])[0]
                finish(row, jj, v)
            else:
                j2 = m - jj
                va, vb = row[jj + t * m].copy(), row[j2 + t * m].copy()
                for tt in range(r):
                    va[tt], vb[r - 1 - tt] = _fold_pair(va[tt], vb[r - 1 - tt], tw_w[jj + tt * m])
                finish(row, jj, va)
                finish(row, j2, vb)


def test_fused_fold_equals_fold_pass_then_first_row_stage(sb):
    """For every plannable width whose row plan lets the fold ride on the first row stage (two stages or more, first
    radix <= 8, m even) the fused items compute exactly what pair_pass followed by that stage computes."""
    rng = np.random.default_rng(11)
    fused = 0
    for n in range(4, 257):
        info = _plan(sb, 4, 2 * n)
        if not info.batched:
            continue
        _, radices = _radices(info)
        if len(radices) < 2 or radices[0] > 8 or (n // radices[0]) % 2:
            continue
        fused += 1
        h, w = 3, 2 * n
        a = rng.standard_normal((h, n + 1)) + 1j * rng.standard_normal((h, n + 1))
        want = a.copy()
        _pair_pass(want, n, w, unfold=False)
        rows = np.ascontiguousarray(want[:, :n].T)
        _stage(rows, n, radices, 0, dit=False)
        got = a.copy()
        _fused_fold_first_stage(got, n, w, radices)
        np.testing.assert_allclose(got[:, :n], rows.T, rtol=1e-12, atol=1e-12, err_msg=f"row length {n}, radices {radices}")
    assert fused >= 15, fused  # e.g. 12, 24, 32, 40, 48, 64, 80 (the C5 width), 96, 128, ...
