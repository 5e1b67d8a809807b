"""Static checks on the built library (cuobjdump, no GPU): register / stack budgets of the hot kernels and the
design claim that tensor cores are deliberately unused. Guards against silent regressions such as a kernel
tipping over its launch-bounds register budget and spilling (the 2-item unroll of the wavelet-CFG kernel
cost 37 -> 43 us that way)."""
from __future__ import annotations

import re
import shutil
import subprocess

import pytest

from conftest import REPO

LIB = REPO / "comfyui-sonar_b200" / "libsonar_b200.so"
pytestmark = pytest.mark.skipif(shutil.which("cuobjdump") is None or shutil.which("c++filt") is None,
                                reason="needs the CUDA toolkit's cuobjdump")


@pytest.fixture(scope="module")
def usage(sb):
    sb._native.load()  # builds the library if it is missing
    out = subprocess.run(["cuobjdump", "-res-usage", str(LIB)], capture_output=True, text=True, check=True).stdout
    rows = re.findall(r"Function (\S+):\s*\n?\s*REG:(\d+) STACK:(\d+) SHARED:(\d+) LOCAL:(\d+)", out)
    assert len(rows) > 100, "no kernels parsed from cuobjdump -res-usage"
    names = subprocess.run(["c++filt", *[r[0] for r in rows]], capture_output=True, text=True, check=True).stdout.splitlines()
    return [(name, int(reg), int(stack), int(shared), int(local)) for name, (_, reg, stack, shared, local) in zip(names, rows)]


def _select(usage, pattern):
    hit = [u for u in usage if re.search(pattern, u[0])]
    assert hit, f"no kernel matches {pattern}"
    return hit


def test_library_is_sm_100a_only():
    out = subprocess.run(["cuobjdump", "-lelf", str(LIB)], capture_output=True, text=True, check=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_register_and_stack_budgets(usage):
    # single-wave step kernel for SDXL-sized draws: 32 registers so that 8 CTAs of 256 threads fit an SM
    for name, reg, stack, _shared, _local in _select(usage, r"sonar_step_fast_philox2_kernel"):
        assert reg <= 32 and stack == 0, (name, reg, stack)
    for name, reg, stack, _shared, _local in _select(usage, r"sonar_step"):  # every specialised and generic variant
        # (philox4 keeps the 16 normals of four Philox calls in registers: 64 = four CTAs of 256 per SM)
        assert reg <= (64 if "philox4" in name else 48) and stack <= 8, (name, reg, stack)
    # 1024 threads per SM at 64 registers: the shared-memory-resident kernels may spill a few words, not more
    # (<2> regenerates its input from the Philox stream: the small-draw variant carries the generator state as well)
    for name, reg, stack, _shared, _local in _select(usage, r"spectral_batched_kernel"):
        assert reg <= 64 and stack <= (192 if "<2>" in name else 128), (name, reg, stack)
    for name, reg, stack, _shared, _local in _select(usage, r"wcfg_fused_kernel<double, 4"):
        assert reg <= 64 and stack <= 32, (name, reg, stack)
    for name, reg, stack, _shared, _local in _select(usage, r"wcfg_fused_kernel"):
        assert reg <= 64 and stack <= 192, (name, reg, stack)
    # streaming kernels never touch local memory
    for name, _reg, stack, _shared, _local in _select(usage, r"(blend_kernel|scale_noise_kernel|moments_kernel|freeu_apply_kernel|philox_fill)"):
        assert stack == 0, (name, stack)
    assert all(u[4] == 0 for u in usage), "no kernel declares static local memory"


def test_co_scheduling_register_budget(usage):
    """The noise pipeline rests on register arithmetic: 4 fill CTAs + 4 step CTAs of 256 threads
    fit one SM only at <= 32 registers each (2 x 32 K), and the 3 x 256-thread FFT CTAs at 64 registers (48 K) leave
    exactly the 16 K two more step CTAs need."""
    for name, reg, _stack, _shared, _local in _select(usage, r"philox_fill(_batch)?_kernel"):
        assert reg <= 32, (name, reg)
    # the C5 step: DPM++ half step, NEW momentum mode, history present, noise normalised on load
    for name, reg, _stack, _shared, _local in _select(usage, r"sonar_step_fast_vec_kernel<1, true, true, 2>|sonar_step_fast_vec_kernel<\(int\)1, \(bool\)1, \(bool\)1, \(int\)2>"):
        assert reg <= 32, (name, reg)
    for name, reg, _stack, _shared, _local in _select(usage, r"spectral_batched_kernel<\(?(int\))?0>"):
        assert 3 * 256 * reg + 2 * 256 * 32 <= 65536, (name, reg)
    for name, reg, _stack, _shared, _local in _select(usage, r"wcfg_strip_kernel"):
        assert reg <= 64, (name, reg)


def test_tensor_cores_are_deliberately_unused():
    """Nothing on this path is a dense contraction: no MMA instruction of any generation."""
    sass = subprocess.run(["cuobjdump", "-sass", str(LIB)], capture_output=True, text=True, check=True).stdout
    assert not re.search(r"\b(HMMA|IMMA|DMMA|QMMA|UTCHMMA|UTCQMMA|UTCMMA|HGMMA)\b", sass)
    # ... while the kernels do use what the path needs: Philox wide multiplies, MUFU for Box-Muller, fp64 FMAs
    for needle in ("IMAD.WIDE", "MUFU.RSQ", "DFMA"):
        assert needle in sass, needle
