"""Bit-exact parity of the device Philox generators with torch's own CUDA generators."""
from __future__ import annotations

import math

import numpy as np
import pytest
import torch

from oracle import sonar_oracle as orc

pytestmark = pytest.mark.gpu

SHAPES = [(1,), (7,), (1, 4, 64, 64), (8, 4, 128, 128), (3, 5, 33, 90), (303104 * 4 + 13,), (2, 16, 33, 90, 160)]


@pytest.mark.parametrize("shape", SHAPES)
@pytest.mark.parametrize("seed", [0, 1234567891011])
def test_randn_matches_torch(sb, cuda, shape, seed):
    torch.manual_seed(seed)
    _ = torch.rand(5, device=cuda)  # non-zero starting offset
    want = torch.randn(shape, device=cuda)
    off_want = torch.cuda.default_generators[0].get_offset()
    want_next = torch.randn(9, device=cuda)
    torch.manual_seed(seed)
    _ = torch.rand(5, device=cuda)
    got = sb.ops.randn(shape, device=cuda)
    assert torch.equal(got, want)
    assert torch.cuda.default_generators[0].get_offset() == off_want
    # torch's own stream continues identically afterwards
    assert torch.equal(torch.randn(9, device=cuda), want_next)


@pytest.mark.parametrize("shape", [(1, 4, 64, 33), (2, 3, 18, 11), (8, 528, 90, 81)])
def test_complex_randn_matches_torch(sb, cuda, shape):
    torch.manual_seed(3)
    want = torch.randn(shape, device=cuda, dtype=torch.complex64)
    off = torch.cuda.default_generators[0].get_offset()
    torch.manual_seed(3)
    got = sb.ops.randn(shape, device=cuda, dtype=torch.complex64)
    assert torch.equal(torch.view_as_real(got), torch.view_as_real(want))
    assert torch.cuda.default_generators[0].get_offset() == off


@pytest.mark.parametrize("shape", [(5,), (16, 129, 129), (16, 16, 128, 128)])
def test_uniform_matches_torch(sb, cuda, shape):
    torch.manual_seed(11)
    want01 = torch.rand(shape, device=cuda)
    want2pi = torch.empty(shape, device=cuda).uniform_(to=2.0 * np.pi)
    torch.manual_seed(11)
    got01 = sb.ops.rand(shape, device=cuda)
    got2pi = sb.ops.rand(shape, device=cuda, low=0.0, high=2.0 * np.pi)
    assert torch.equal(got01, want01)
    assert torch.equal(got2pi, want2pi)


def test_uniform_matches_cpu_oracle(sb, cuda):
    """The integer Philox stream + uniform transform restated on the CPU is bit-identical."""
    n = 70000
    torch.manual_seed(99)
    gen = torch.cuda.default_generators[0]
    seed, offset = gen.initial_seed(), gen.get_offset()
    got = sb.ops.rand((n,), device=cuda).cpu().numpy()
    props = torch.cuda.get_device_properties(0)
    grid, _ = orc.aten_policy(n, props.multi_processor_count, props.max_threads_per_multi_processor)
    want = orc.aten_uniform(n, seed, offset, grid)
    assert np.array_equal(got, want)
    torch.manual_seed(99)
    normal = sb.ops.randn((n,), device=cuda).cpu().numpy()
    np.testing.assert_allclose(normal, orc.aten_normal(n, seed, offset, grid), rtol=0, atol=1e-5)  # __sincosf is the fast intrinsic


def test_generator_argument(sb, cuda):
    g1 = torch.Generator(device=cuda).manual_seed(42)
    g2 = torch.Generator(device=cuda).manual_seed(42)
    assert torch.equal(sb.ops.randn((3, 1000), device=cuda, generator=g1), torch.randn((3, 1000), device=cuda, generator=g2))
    assert g1.get_offset() == g2.get_offset()


@pytest.mark.parametrize("world", [2, 4, 8])
def test_sharded_slices_reassemble(sb, cuda, world):
    """Each rank fills only its batch slice of the global draw; concatenated they equal torch.randn."""
    shape = (8, 4, 33, 20)
    torch.manual_seed(5)
    want = torch.randn(shape, device=cuda)
    parts = []
    for rank in range(world):
        torch.manual_seed(5)
        with sb.parallel.sharded(shape[0], rank=rank, world_size=world) as ctx:
            local = (ctx.local_batch, *shape[1:])
            parts.append(sb.rng.normal(local, device=cuda))
    assert torch.equal(torch.cat(parts), want)


def test_philox_moments_match_materialised(sb, cuda):
    n = 8 * 4 * 128 * 128
    torch.manual_seed(8)
    draw = sb.ops.reserve_draw(n, cuda)
    sums = sb.ops.philox_normal_moments(draw, begin=0, count=n, sums=sb.ops.new_sums(cuda))
    torch.manual_seed(8)
    x = torch.randn(n, device=cuda).double()
    want = torch.stack((x.sum(), (x * x).sum()))
    torch.testing.assert_close(sums, want, rtol=1e-6, atol=1e-3)  # per-thread partials are fp32


def test_empty_and_errors(sb, cuda):
    assert sb.ops.randn((0, 4), device=cuda).shape == (0, 4)
    with pytest.raises(RuntimeError):
        sb.ops.moments(torch.zeros(4))  # CPU tensor: no fallback
    with pytest.raises(TypeError):
        sb.ops.randn((4,), device=cuda, dtype=torch.float16)


def test_batched_draws_equal_separate_torch_draws(sb, cuda):
    """rng.batched(): several draws reserved in order, materialised by ONE launch, still bit-identical
    to the sequence of torch calls (different sizes, kinds, transforms; more draws than one launch holds)."""
    n_draws = sb.ops.FILL_BATCH_MAX + 8
    shapes = [(16, 16, 128, 128), (16, 16, 36, 36), (16, 16, 7, 7), (16, 16, 1, 1), (3, 5), (16, 129, 129)]
    torch.manual_seed(2024)
    want = []
    for i in range(n_draws):
        shp = shapes[i % len(shapes)]
        if i % 3 == 0:
            want.append(torch.empty(shp, device=cuda).uniform_(0.0, 2.0 * math.pi))
        elif i % 3 == 1:
            want.append(torch.randn(shp, device=cuda))
        else:
            want.append(torch.randn(shp, device=cuda, dtype=torch.complex64))
    off_want = torch.cuda.default_generators[0].get_offset()
    torch.manual_seed(2024)
    launches = sb.ops.LAUNCH_COUNT
    got = []
    with sb.rng.batched():
        for i in range(n_draws):
            shp = shapes[i % len(shapes)]
            if i % 3 == 0:
                got.append(sb.rng.uniform(shp, device=cuda, low=0.0, high=2.0 * math.pi))
            elif i % 3 == 1:
                got.append(sb.rng.normal(shp, device=cuda))
            else:
                got.append(sb.rng.normal(shp, device=cuda, dtype=torch.complex64))
    assert sb.ops.LAUNCH_COUNT - launches == 2  # FILL_BATCH_MAX + 8 draws
    assert torch.cuda.default_generators[0].get_offset() == off_want
    for i, (g, w) in enumerate(zip(got, want)):
        assert torch.equal(g, w), f"draw {i}"
