"""CPU-only checks: C-ABI surface, host logic mirrored from the reference, sharding helpers."""
from __future__ import annotations

import ctypes
import importlib
import re
import subprocess
import sys
from pathlib import Path

import pytest
import torch

REPO = Path(__file__).resolve().parent.parent


def _header_symbols() -> set[str]:
    text = (REPO / "include" / "sonar_b200.h").read_text()
    return set(re.findall(r"^\s*(?:int|int64_t)\s+(sonar_\w+)\s*\(", text, flags=re.M))


def test_ctypes_signatures_have_the_header_arity(sb):
    """Every prototype in include/sonar_b200.h takes as many parameters as its ctypes argtypes entry, and the
    functions that do not return an error code are the ones registered with their own restype."""
    text = re.sub(r"/\*.*?\*/", "", (REPO / "include" / "sonar_b200.h").read_text(), flags=re.S)
    protos = re.findall(r"^\s*(int|int64_t)\s+(sonar_\w+)\s*\(([^;{]*?)\)\s*;", text, flags=re.M | re.S)
    assert len(protos) == len(sb._native.SIGNATURES)
    for ret, name, args in protos:
        args = args.strip()
        n_args = 0 if args in ("", "void") else args.count(",") + 1
        assert n_args == len(sb._native.SIGNATURES[name]), f"{name}: header has {n_args} parameters"
        if ret == "int64_t":
            assert name in sb._native.RESTYPES, f"{name} returns int64_t: needs a ctypes restype"


def test_library_exports_every_declared_symbol(sb):
    lib = sb._native.load()
    declared = _header_symbols()
    assert declared, "no symbols parsed from include/sonar_b200.h"
    assert declared == set(sb._native.SIGNATURES), "ctypes table and header disagree"
    for name in declared:
        assert hasattr(lib, name), f"libsonar_b200.so does not export {name}"
    assert lib.sonar_abi_version() == sb._native.ABI_VERSION


def test_struct_layouts_match_the_header(sb, tmp_path):
    """sizeof of every by-value struct, as compiled by the C compiler, equals the ctypes mirror."""
    names = ["SonarStepParams", "SonarPyramidParams", "SonarPerlinParams", "SonarSpectralParams", "SonarSpectralPlanInfo",
             "SonarWaveletFilters", "SonarDwtAnalysisParams", "SonarDwtSynthesisParams", "SonarWcfgFusedParams",
             "SonarFreeuParams", "SonarGuidanceParams", "SonarFillBatch", "SonarMixTerm", "SonarNoiseMixParams"]  # fmt: skip
    names = [n for n in names if hasattr(sb._native, n)]
    assert len(names) >= 10
    src = tmp_path / "sz.c"
    body = "".join(f'printf("%zu\\n", sizeof({n}));' for n in names)
    src.write_text(f'#include <stdio.h>\n#include "{REPO}/include/sonar_b200.h"\nint main(void){{{body}return 0;}}\n')
    exe = tmp_path / "sz"
    subprocess.run(["gcc", str(src), "-o", str(exe)], check=True)
    sizes = [int(v) for v in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    for n, size in zip(names, sizes):
        assert ctypes.sizeof(getattr(sb._native, n)) == size, n
    # ... and every field sits at the offset the C compiler gives it (same names in the header and the mirror)
    fields = [(n, f[0]) for n in names for f in getattr(sb._native, n)._fields_]
    body = "".join(f'printf("%zu\\n", offsetof({n}, {f}));' for n, f in fields)
    src.write_text(
        f'#include <stdio.h>\n#include <stddef.h>\n#include "{REPO}/include/sonar_b200.h"\nint main(void){{{body}return 0;}}\n',
    )
    subprocess.run(["gcc", str(src), "-o", str(exe)], check=True)
    offsets = [int(v) for v in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    for (n, f), off in zip(fields, offsets):
        assert getattr(getattr(sb._native, n), f).offset == off, f"{n}.{f}"


def test_no_cpu_fallback(sb):
    with pytest.raises(RuntimeError, match="CUDA"):
        sb.ops.moments(torch.zeros(8))
    with pytest.raises(RuntimeError, match="CUDA"):
        sb.noise_graph.get_noise_sampler("gaussian", torch.zeros(1, 4, 8, 8), None, None)
    with pytest.raises(RuntimeError):
        sb.hostutil.scale_noise(torch.ones(4, 4), 1.0, normalized=True)


def test_product_does_not_import_the_oracle():
    pkg = REPO / "comfyui-sonar_b200"
    for path in pkg.glob("*.py"):
        text = path.read_text()
        assert "import oracle" not in text and "from oracle" not in text, path.name
        assert "/root/reference" not in text, path.name


def test_philox_policy_matches_oracle(sb):
    from oracle import sonar_oracle as orc

    for n in (1, 255, 256, 16384, 524288, 303104 * 4, 303104 * 4 + 1, 60_825_600):
        assert sb.ops.philox_policy(n) == orc.aten_policy(n)


def test_config_merge_and_errors(sb, golden):
    s = sb.samplers
    cfg = s.SonarBase.get_config(None, {"momentum": 0.5, "momentum_mode": " classic ", "init": "RAND", "noise_type": "pyramid"})
    assert cfg.momentum == 0.5 and cfg.momentum_mode == s.MomentumMode.CLASSIC
    assert cfg.init == s.HistoryType.RAND and cfg.noise_type == sb.noise_graph.NoiseType.PYRAMID
    merged = s.SonarBase.get_config(cfg, {"direction": -1.0})
    assert merged.direction == -1.0 and merged.momentum == 0.5
    for params, (exc_name, msg) in golden("host_logic")["config_errors"].items():
        with pytest.raises((ValueError, TypeError)) as info:
            s.SonarBase.get_config(None, eval(params))  # noqa: S307 - fixture literal
        assert type(info.value).__name__ == exc_name and str(info.value) == msg
    with pytest.raises(TypeError):
        s.SonarBase.get_config(None, {"not_a_field": 1})


def test_history_ratios(sb, golden):
    s = sb.samplers
    for (direction, mh), want in golden("host_logic")["history_ratios"]:
        got = s.SonarBase(s.SonarConfig(direction=direction, momentum_hist=mh)).history_ratios
        assert tuple(got) == tuple(want)


def test_expand_yh_scales(sb, golden):
    shapes = [(1, 1, 3, 4, 4)] * 4
    for spec, want in golden("host_logic")["expand"]:
        assert sb.wavelets.expand_yh_scales(shapes, yh_scales=spec) == want
    with pytest.raises(ValueError):
        sb.wavelets.expand_yh_scales(shapes, yh_scales=["fill", 1.0])
    with pytest.raises(ValueError):
        sb.wavelets.expand_yh_scales(shapes, yh_scales=[1.0, "fill", "fill"])


def test_noise_type_enum_matches_reference_names(sb):
    names = list(sb.noise_graph.NoiseType.get_names())
    assert names[0] == "gaussian" and len(names) == 38 and len(set(names)) == 38
    assert set(sb.noise_graph.NOISE_SAMPLERS) == set(sb.noise_graph.NoiseType)


def test_chain_factor_and_rescale(sb):
    ng = sb.noise_graph
    chain = ng.CustomNoiseChain()
    chain.add(ng.CustomNoiseItem(0.6, noise_type="gaussian"))
    chain.add(ng.CustomNoiseItem(-0.4, noise_type="uniform"))
    assert chain.factor == pytest.approx(1.0)
    scaled = chain.rescaled(2.0)
    assert [i.factor for i in scaled.items] == pytest.approx([1.2, -0.8]) and chain.items[0].factor == 0.6
    with pytest.raises(ValueError):
        chain.add(None)
    with pytest.raises(ValueError):
        ng.CustomNoiseItem(1.0)
    with pytest.raises(ValueError):
        ng.CustomNoiseItem(1.0, noise_type="gaussian", yaml_parameters="[1, 2]")
    with pytest.raises(ValueError):
        ng.BlendedNoise(1.0, normalize=None, blend_function=None, custom_noise_2=chain, noise_2_percent=0.5)


def test_wavelet_filter_banks_are_orthonormal(sb):
    for wave in sb.wavelets.Wavelet.wavelist():
        dec_lo, dec_hi, rec_lo, rec_hi = (torch.tensor(f, dtype=torch.float64) for f in sb.wavelets.filter_bank(wave))
        n = len(dec_lo)
        assert float(dec_lo.sum()) == pytest.approx(2**0.5, abs=1e-12)
        assert float(dec_hi.sum()) == pytest.approx(0.0, abs=1e-12)
        for shift in range(0, n, 2):
            want = 1.0 if shift == 0 else 0.0
            assert float((dec_lo[: n - shift] * dec_lo[shift:]).sum()) == pytest.approx(want, abs=1e-12)
            assert float((dec_lo[: n - shift] * dec_hi[shift:]).sum()) == pytest.approx(0.0, abs=1e-12)
        assert torch.equal(rec_lo, dec_lo.flip(0)) and torch.equal(rec_hi, dec_hi.flip(0))
    with pytest.raises(NotImplementedError):
        sb.wavelets.filter_bank("sym4")


def test_wcfg_rule_parsing(sb):
    w = sb.wcfg
    rules = w.WCFGRules.build(
        wave="db2", level=3, diff={"yl_scale": 5, "yh_scales": [[3, 4, 5]]}, target_mode="noise",
        rules=[{"start_sigma": 2.0, "end_sigma": 0.5, "difference": {"yl_scale": 2.0}}],
    )
    assert len(rules) == 2 and rules[0].wavelet.wave == "db2" and rules[0].target_mode == w.WCFGTarget.NOISE
    assert rules.get_rule(1.0) is rules[0] and rules[1].diff.yl_scale == 2.0
    only = w.WCFGRules.build(start_sigma=2.0, end_sigma=0.5)
    assert only.get_rule(3.0) is None
    sched = w.WCFGScalesRange.build(yl_scale=1.0, scales_end={"yl_scale": 3.0}, schedule="sine")
    assert isinstance(sched, w.WCFGScalesRange) and sched.scheduler.schedule == w.WCFGSchedule.SINE
    assert w.WCFGSchedule.HALF_COSINE.interp(0.5) == pytest.approx(0.5)
    with pytest.raises(TypeError):
        w.WCFGRule.build(blend_strength="x")


def _gloo_worker(rank: int, world: int, port: int, q) -> None:
    import os

    import torch.distributed as dist

    sys.path.insert(0, str(REPO))
    import sonar_b200 as sb

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        full = torch.arange(5 * 3 * 4, dtype=torch.float64).reshape(5, 3, 4)
        with sb.parallel.sharded(5) as ctx:
            local = sb.parallel.shard(full)
            sums = torch.stack((local.sum(), (local * local).sum()))
            count = sb.parallel.global_count(local.numel(), sums)
            total, begin = sb.parallel.global_draw_geometry(local.shape)
            gathered = sb.parallel.gather(local)
            to_zero = sb.parallel.gather(local, dst=0)
            # no peer memory under gloo: the device rendezvous declines instead of hanging or touching CUDA
            assert ctx.peers is None and sb.parallel.device_barrier() is False
            q.put((rank, ctx.batch_sizes, count, sums.tolist(), total, begin, torch.equal(gathered, full),
                   None if to_zero is None else torch.equal(to_zero, full), ctx.collectives))
    finally:
        dist.destroy_process_group()


def test_sharding_world_size_2_gloo():
    """N > 1 host path on CPU: ragged batch split, 2-double all-reduce, gather."""
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (hash(str(REPO)) % 500)
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    full = torch.arange(60, dtype=torch.float64)
    for rank, sizes, count, sums, total, begin, gathered_ok, dst_ok, collectives in results:
        assert sizes == [3, 2] and count == 60 and total == 60
        assert begin == (0 if rank == 0 else 36)
        assert sums == pytest.approx([float(full.sum()), float((full * full).sum())])
        assert gathered_ok and (dst_ok is True if rank == 0 else dst_ok is None)
        assert collectives == 3


def test_host_schedule_copy_is_cached_per_tensor_object(sb):
    """The sampler's one device->host copy of the sigma schedule is reused only for the very same, unmodified
    tensor object (weak reference + version counter), never for another tensor with equal values."""
    from sonar_b200 import samplers

    a = torch.linspace(14.6, 0.0, 8)
    h1 = samplers._host_schedule(a)
    assert h1 is samplers._host_schedule(a) and h1.data_ptr() != a.data_ptr()
    a.mul_(0.5)  # in-place edit bumps the version: fresh copy
    h2 = samplers._host_schedule(a)
    assert h2 is not h1 and torch.equal(h2, a)
    b = a.clone()
    h3 = samplers._host_schedule(b)
    assert h3 is not h2 and torch.equal(h3, b)
    del b
    assert torch.equal(samplers._host_schedule(a), a)  # cache slot now points at a dead tensor: recopied
    assert samplers._host_schedule(a.double()).dtype == torch.float32


def test_sharding_rejects_a_batch_smaller_than_the_world(sb):
    """a rank with an empty shard would skip the exchanges its peers wait for."""
    import pytest

    with pytest.raises(ValueError, match="at least one item"):
        with sb.parallel.sharded(1, rank=0, world_size=2):
            pass
    with sb.parallel.sharded(5, rank=1, world_size=2) as ctx:
        assert ctx.batch_sizes == [3, 2] and ctx.batch_begin == 3


def test_channel_mixer_is_memoised_and_identity_stays_on_the_host():
    """ChannelMixer.build is a pure function of (channels, common_mode, correlation): built once, cloned per sampler;
    the default (common_mode == 0) mixer is the identity and is never moved to the device."""
    sn = importlib.import_module("sonar_b200.spectral_noise")
    corr = torch.tensor([1.0, 0.5, 0.25])
    a = sn.ChannelMixer(6, 0.3, corr)
    b = sn.ChannelMixer(6, 0.3, corr.clone())
    assert torch.equal(a.mixer, b.mixer) and a.mixer.data_ptr() != b.mixer.data_ptr()
    assert not a.is_identity
    fresh = a._build()  # noqa: SLF001
    assert torch.equal(a.mixer, fresh)
    # rows have unit norm (:88-89)
    assert torch.allclose(a.mixer.norm(dim=1), torch.ones(6), atol=1e-6)
    ident = sn.ChannelMixer(528, 0.0, torch.ones(6))
    assert ident.is_identity and torch.equal(ident.mixer, torch.eye(528))
    moved = ident.to("meta")  # would fail loudly if the identity were copied anywhere
    assert moved.mixer.device.type == "cpu"


def test_peer_exchange_cache_is_keyed_on_the_group_object(monkeypatch):
    """A process group that was destroyed and re-created must not be handed the mailboxes of the old one: the cache
    compares the group OBJECT (held strongly), not its id()."""
    par = importlib.import_module("sonar_b200.parallel")
    monkeypatch.setattr(par, "_PEERS", [])

    class FakeGroup:
        pass

    a, b = FakeGroup(), FakeGroup()
    monkeypatch.setattr(par.dist, "is_initialized", lambda: False)  # no NCCL here: entries are None, the keying is what counts
    assert par.peer_exchange(0, 2, a) is None and len(par._PEERS) == 1  # noqa: SLF001
    assert par.peer_exchange(0, 2, a) is None and len(par._PEERS) == 1  # noqa: SLF001  (hit)
    assert par.peer_exchange(0, 2, b) is None and len(par._PEERS) == 2  # noqa: SLF001  (another object: miss)
    assert par._PEERS[0][0] is a and par._PEERS[1][0] is b  # noqa: SLF001
    assert par.peer_exchange(1, 2, a) is None and len(par._PEERS) == 3  # noqa: SLF001  (another rank)
