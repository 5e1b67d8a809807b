"""ctypes binding of libsonar_b200.so (the C ABI declared in include/sonar_b200.h).

There is deliberately no fallback: if the library cannot be loaded (or built with nvcc), every
product entry point raises. PyTorch is only used for device memory and streams; the arithmetic
runs in the hand-written sm_100a kernels behind these symbols.
"""

from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, c_double, c_float, c_int, c_int32, c_int64, c_uint32, c_uint64, c_void_p
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
LIB_NAME = "libsonar_b200.so"


class NativeLibraryError(RuntimeError):
    pass


class SonarStepParams(ctypes.Structure):
    """Mirror of `struct SonarStepParams` (include/sonar_b200.h)."""

    _fields_ = [
        ("x", c_void_p),
        ("denoised", c_void_p),
        ("hist_in", c_void_p),
        ("noise", c_void_p),
        ("x_out", c_void_p),
        ("hist_out", c_void_p),
        ("n", c_int64),
        ("kind", c_int32),
        ("mode", c_int32),
        ("momentum_blend", c_int32),
        ("history_blend", c_int32),
        ("hist_state", c_int32),
        ("momentum_active", c_int32),
        ("history_active", c_int32),
        ("noise_kind", c_int32),
        ("momentum", c_float),
        ("sigma", c_float),
        ("c0", c_float),
        ("c1", c_float),
        ("hd_ratio", c_float),
        ("hd_scale", c_float),
        ("md_scale", c_float),
        ("hist_in_div", c_float),
        ("noise_scale", c_float),
        ("noise_factor", c_float),
        ("noise_threshold_std_devs", c_float),
        ("philox_seed", c_uint64),
        ("philox_offset", c_uint64),
        ("philox_grid_blocks", c_uint32),
        ("noise_begin", c_int64),
        ("noise_numel_total", c_int64),
        ("noise_sums", c_void_p),
        ("noise_count", c_int64),
        ("noise_decision", c_void_p),
        ("peer_world", c_int32),
        ("peer_mailbox", c_void_p),
        ("peer_epoch", c_double),
    ]


ABI_VERSION = 9  # SONAR_B200_ABI_VERSION of include/sonar_b200.h this binding was written against
PEER_MAX_RANKS = 8
PEER_TABLE_MAX = 512
PYRAMID_MAX_LEVELS = 16
PERLIN_MAX_ITERS = 8
FFT_MAX_FACTORS = 24
DWT_MAX_TAPS = 40


FILL_BATCH_MAX = 32
MIXER_SMALL_MAX = 8  # SONAR_MIXER_SMALL_MAX


class SonarFillDesc(ctypes.Structure):
    _fields_ = [
        ("out", c_void_p),
        ("begin", c_int64),
        ("count", c_int64),
        ("numel_total", c_int64),
        ("offset", c_uint64),
        ("grid_blocks", c_uint32),
        ("kind", c_int32),
        ("p0", c_float),
        ("p1", c_float),
    ]


class SonarFillBatch(ctypes.Structure):
    _fields_ = [("n", c_int32), ("seed", c_uint64), ("draws", SonarFillDesc * FILL_BATCH_MAX)]


class SonarGuidanceParams(ctypes.Structure):
    _fields_ = [
        ("x", c_void_p),
        ("ref", c_void_p),
        ("item_sums", c_void_p),
        ("out", c_void_p),
        ("items", c_int64),
        ("per_item", c_int64),
        ("ref_items", c_int32),
        ("kind", c_int32),
        ("blend_mode", c_int32),
        ("factor", c_float),
        ("sigma", c_float),
        ("dt", c_float),
    ]


class SonarPyramidParams(ctypes.Structure):
    _fields_ = [
        ("out", c_void_p),
        ("base", c_void_p),
        ("levels", c_void_p * PYRAMID_MAX_LEVELS),
        ("level_h", c_int32 * PYRAMID_MAX_LEVELS),
        ("level_w", c_int32 * PYRAMID_MAX_LEVELS),
        ("weights", c_float * PYRAMID_MAX_LEVELS),
        ("planes", c_int64),
        ("H", c_int32),
        ("W", c_int32),
        ("n_levels", c_int32),
        ("mode", c_int32),
        ("base_scale", c_float),
        ("sums", c_void_p),
        ("sums_clear", c_void_p),
    ]


MIX_MAX_TABLES = 4  # SONAR_MIX_MAX_TABLES


class SonarMixTerm(ctypes.Structure):
    _fields_ = [
        ("kind", c_int32),
        ("n_levels", c_int32),
        ("mode", c_int32),
        ("full_level", c_int32),
        ("iterations", c_int32),
        ("base_scale", c_float),
        ("div_fac", c_float),
        ("uniform_from", c_float),
        ("uniform_to", c_float),
        ("base_offset", c_uint64),
        ("level_offset", c_uint64 * PYRAMID_MAX_LEVELS),
        ("levels", c_void_p * PYRAMID_MAX_LEVELS),
        ("level_h", c_int32 * PYRAMID_MAX_LEVELS),
        ("level_w", c_int32 * PYRAMID_MAX_LEVELS),
        ("weights", c_float * PYRAMID_MAX_LEVELS),
        ("tables", c_void_p * MIX_MAX_TABLES),
    ]


class SonarNoiseMixParams(ctypes.Structure):
    _fields_ = [
        ("out", c_void_p),
        ("n", c_int64),
        ("begin", c_int64),
        ("numel_total", c_int64),
        ("C", c_int32),
        ("H", c_int32),
        ("W", c_int32),
        ("blend_mode", c_int32),
        ("blend_t", c_float),
        ("grid_blocks", c_uint32),
        ("seed", c_uint64),
        ("a", SonarMixTerm),
        ("b", SonarMixTerm),
        ("sums", c_void_p),
        ("sums_clear", c_void_p),
    ]


class SonarPerlinParams(ctypes.Structure):
    _fields_ = [
        ("out", c_void_p),
        ("base", c_void_p),
        ("angles", c_void_p * PERLIN_MAX_ITERS),
        ("B", c_int32),
        ("C", c_int32),
        ("H", c_int32),
        ("W", c_int32),
        ("iterations", c_int32),
        ("blend_mode", c_int32),
        ("div_fac", c_float),
        ("sums", c_void_p),
        ("sums_clear", c_void_p),
    ]


class SonarSpectralParams(ctypes.Structure):
    _fields_ = [
        ("out", c_void_p),
        ("in_real", c_void_p),
        ("in_spec", c_void_p),
        ("mask", c_void_p),
        ("scratch", c_void_p),
        ("planes", c_int64),
        ("H", c_int32),
        ("W", c_int32),
        ("out_scale", c_float),
        ("sums", c_void_p),
        ("sums_clear", c_void_p),
        ("sums_segment_planes", c_int64),
        ("philox_seed", c_uint64),
        ("philox_offset", c_uint64),
        ("philox_grid_blocks", c_uint32),
        ("philox_std", c_float),
        ("philox_begin", c_int64),
        ("philox_numel_total", c_int64),
    ]


SPECTRAL_PLAN_MAX_STAGES = 16


class SonarSpectralPlanInfo(ctypes.Structure):
    _fields_ = [
        ("batched", c_int32),
        ("group", c_int32),
        ("threads", c_int32),
        ("ctas_per_sm", c_int32),
        ("grid", c_int64),
        ("smem_bytes", c_int64),
        ("n_col_stages", c_int32),
        ("n_row_stages", c_int32),
        ("col_radix", c_int32 * SPECTRAL_PLAN_MAX_STAGES),
        ("row_radix", c_int32 * SPECTRAL_PLAN_MAX_STAGES),
        ("cluster", c_int32),
        ("pad_", c_int32),
    ]


class SonarWaveletFilters(ctypes.Structure):
    _fields_ = [
        ("length", c_int32),
        ("dec_lo", c_double * DWT_MAX_TAPS),
        ("dec_hi", c_double * DWT_MAX_TAPS),
        ("rec_lo", c_double * DWT_MAX_TAPS),
        ("rec_hi", c_double * DWT_MAX_TAPS),
    ]


class SonarDwtAnalysisParams(ctypes.Structure):
    _fields_ = [
        ("in_a", c_void_p),
        ("in_b", c_void_p),
        ("ll", c_void_p),
        ("hi", c_void_p),
        ("planes", c_int64),
        ("H", c_int32),
        ("W", c_int32),
        ("in_stride_h", c_int32),
        ("h", c_int32),
        ("w", c_int32),
        ("mode", c_int32),
        ("in_is_f32", c_int32),
        ("use_f64", c_int32),
        ("filters", SonarWaveletFilters),
    ]


class SonarDwtSynthesisParams(ctypes.Structure):
    _fields_ = [
        ("ll", c_void_p * 2),
        ("hi", c_void_p * 2),
        ("ll_rows", c_int32 * 2),
        ("ll_cols", c_int32 * 2),
        ("scales", (c_double * 4) * 2),
        ("n_sets", c_int32),
        ("planes", c_int64),
        ("h", c_int32),
        ("w", c_int32),
        ("out", c_void_p),
        ("out_f32", c_void_p),
        ("crop_h", c_int32),
        ("crop_w", c_int32),
        ("addend", c_void_p),
        ("addend_scale", c_float),
        ("x", c_void_p),
        ("x_scale", c_float),
        ("recon_sign", c_float),
        ("use_f64", c_int32),
        ("filters", SonarWaveletFilters),
    ]


WCFG_MAX_LEVELS = 8


class SonarWcfgFusedParams(ctypes.Structure):
    _fields_ = [
        ("in_a", c_void_p),
        ("in_b", c_void_p),
        ("out", c_void_p),
        ("addend", c_void_p),
        ("addend_scale", c_float),
        ("x", c_void_p),
        ("x_scale", c_float),
        ("recon_sign", c_float),
        ("planes", c_int64),
        ("H", c_int32),
        ("W", c_int32),
        ("levels", c_int32),
        ("mode", c_int32),
        ("use_f64", c_int32),
        ("scale_ll", c_double),
        ("scale_hi", (c_double * 3) * WCFG_MAX_LEVELS),
        ("filters", SonarWaveletFilters),
    ]


class SonarFreeuParams(ctypes.Structure):
    _fields_ = [
        ("x", c_void_p),
        ("filtered", c_void_p),
        ("hidden", c_void_p),
        ("hidden_range", c_void_p),
        ("batch", c_int64),
        ("channels", c_int64),
        ("hw", c_int64),
        ("slice_offset", c_int64),
        ("slice_channels", c_int64),
        ("scale", c_float),
        ("scale_minus_one", c_float),
        ("blend", c_float),
        ("blend_mode", c_int32),
        ("use_blend", c_int32),
    ]


# name -> argtypes; every function returns int. Kept in one table so tests can check that the
# library exports exactly what include/sonar_b200.h declares.
SIGNATURES: dict[str, list] = {
    "sonar_abi_version": [],
    "sonar_set_device": [c_int],
    "sonar_set_grid_limit": [c_int],
    "sonar_philox_policy": [c_int64, POINTER(c_uint32), POINTER(c_uint64)],
    "sonar_philox_normal_f32": [
        c_void_p, c_int64, c_int64, c_int64, c_uint64, c_uint64, c_uint32, c_float, c_float, c_void_p,
    ],
    "sonar_philox_uniform_f32": [
        c_void_p, c_int64, c_int64, c_int64, c_uint64, c_uint64, c_uint32, c_float, c_float, c_void_p,
    ],
    "sonar_philox_fill_batch": [POINTER(SonarFillBatch), c_void_p],
    "sonar_moments_f32": [c_void_p, c_int64, c_void_p, c_void_p],
    "sonar_philox_normal_moments": [c_int64, c_int64, c_int64, c_uint64, c_uint64, c_uint32, c_void_p, c_void_p],
    "sonar_philox_normal_moments_batch": [
        POINTER(c_uint64), c_int, c_int64, c_int64, c_int64, c_uint64, c_uint32, c_void_p, c_void_p,
    ],
    "sonar_norm_decisions": [c_void_p, c_int, c_int64, c_float, c_void_p, c_void_p],
    "sonar_scale_noise_f32": [c_void_p, c_void_p, c_int64, c_void_p, c_int64, c_float, c_float, c_void_p],
    "sonar_scale_by_std_f32": [c_void_p, c_void_p, c_int64, c_void_p, c_int64, c_float, c_void_p],
    "sonar_affine_f32": [c_void_p, c_void_p, c_int64, c_float, c_float, c_float, c_void_p],
    "sonar_div_scalar_f32": [c_void_p, c_void_p, c_int64, c_float, c_void_p],
    "sonar_step_f32": [POINTER(SonarStepParams), c_void_p],
    "sonar_philox_normal_fill_moments_f32": [
        c_void_p, c_int64, c_int64, c_int64, c_uint64, c_uint64, c_uint32, c_void_p, c_void_p,
    ],
    "sonar_peer_alloc": [POINTER(c_void_p)],
    "sonar_peer_free": [c_void_p],
    "sonar_peer_get_handle": [c_void_p, ctypes.c_char_p],
    "sonar_peer_open_handle": [ctypes.c_char_p, POINTER(c_void_p)],
    "sonar_peer_close_handle": [c_void_p],
    "sonar_scale_noise_peers_f32": [
        c_void_p, c_void_p, c_int64, c_void_p, c_int, c_double, c_int64, c_float, c_float, c_void_p,
    ],
    "sonar_peer_publish_sums": [POINTER(c_void_p), c_int, c_int, c_void_p, c_double, c_void_p],
    "sonar_peer_allreduce_table": [POINTER(c_void_p), c_int, c_int, c_void_p, c_int, c_double, c_void_p],
    "sonar_item_moments_f32": [c_void_p, c_int64, c_int64, c_void_p, c_void_p],
    "sonar_guidance_f32": [POINTER(SonarGuidanceParams), c_void_p],
    "sonar_pyramid_accum_f32": [POINTER(SonarPyramidParams), c_void_p],
    "sonar_perlin_accum_f32": [POINTER(SonarPerlinParams), c_void_p],
    "sonar_noise_mix_f32": [POINTER(SonarNoiseMixParams), c_void_p],
    "sonar_perlin_tables_f32": [POINTER(c_void_p), POINTER(c_uint64), POINTER(c_uint32), c_int32, c_uint64, c_int32, c_int32, c_int32, c_int32, c_void_p],
    "sonar_blend_f32": [c_void_p, c_void_p, c_void_p, c_float, c_void_p, c_int64, c_int, c_void_p, c_void_p, c_void_p],
    "sonar_axpby_f32": [c_void_p, c_float, c_void_p, c_float, c_void_p, c_int64, c_void_p, c_void_p, c_void_p],
    "sonar_composite_f32": [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int64, c_void_p],
    "sonar_powerlaw_f32": [c_void_p, c_void_p, c_int64, c_float, c_int, c_void_p],
    "sonar_item_range_scratch_bytes": [c_int64],
    "sonar_item_div_max_f32": [c_void_p, c_void_p, c_int64, c_int64, c_int, c_void_p, c_void_p],
    "sonar_item_minmax_rescale_f32": [
        c_void_p, c_void_p, c_int64, c_int64, c_float, c_float, c_float, c_void_p, c_void_p,
    ],
    "sonar_spectral_scratch_bytes": [c_int, c_int],
    "sonar_spectral_filter_f32": [POINTER(SonarSpectralParams), c_void_p],
    "sonar_channel_mix_f32": [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int32, c_int64, c_void_p, c_void_p, c_void_p],
    "sonar_channel_mix_packed_floats": [c_int32],
    "sonar_spectral_plan": [c_int, c_int, c_int64, c_int, POINTER(SonarSpectralPlanInfo)],
    "sonar_dwt_coeff_len": [c_int, c_int],
    "sonar_dwt2_analysis": [POINTER(SonarDwtAnalysisParams), c_void_p],
    "sonar_dwt2_synthesis": [POINTER(SonarDwtSynthesisParams), c_void_p],
    "sonar_dwt2_synthesis_per": [POINTER(SonarDwtSynthesisParams), c_void_p],
    "sonar_wcfg_fused_smem_bytes": [c_int, c_int, c_int, c_int, c_int],
    "sonar_wcfg_fused": [POINTER(SonarWcfgFusedParams), c_void_p],
    "sonar_freeu_range_bytes": [c_int64],
    "sonar_freeu_hidden_mean_f32": [c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int64, c_void_p],
    "sonar_freeu_apply_f32": [POINTER(SonarFreeuParams), c_void_p],
}

# functions whose return value is not an error code
RESTYPES = {
    "sonar_channel_mix_packed_floats": c_int64,
    "sonar_spectral_scratch_bytes": c_int64,
    "sonar_wcfg_fused_smem_bytes": c_int64,
    "sonar_freeu_range_bytes": c_int64,
}

_LIB: ctypes.CDLL | None = None


def library_path() -> Path:
    override = os.environ.get("SONAR_B200_LIB")
    return Path(override) if override else PKG_DIR / LIB_NAME


def _build_if_possible() -> None:
    import importlib.util

    spec = importlib.util.spec_from_file_location("_sonar_b200_build", PKG_DIR / "build.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    mod.build_library()


def _warn_if_stale(path: Path) -> None:
    """A library older than its sources runs old kernels silently: say so (a rebuild is `python build.py`)."""
    try:
        built = path.stat().st_mtime
        deps = [*(PKG_DIR / "csrc").glob("*.cu"), *(PKG_DIR / "csrc").glob("*.cuh"), *(PKG_DIR.parent / "include").glob("*.h")]
        newer = [d.name for d in deps if d.stat().st_mtime > built + 1.0]
    except OSError:
        return
    if newer and path.parent == PKG_DIR:
        import warnings

        warnings.warn(
            f"{path.name} is older than {', '.join(sorted(newer)[:4])}: rebuild with `python comfyui-sonar_b200/build.py`",
            RuntimeWarning,
            stacklevel=3,
        )


def load(*, build_if_missing: bool = True) -> ctypes.CDLL:
    """Loads the shared library (building it with nvcc first if it is absent). Raises loudly."""
    global _LIB  # noqa: PLW0603
    if _LIB is not None:
        return _LIB
    path = library_path()
    if not path.exists() and build_if_missing:
        try:
            _build_if_possible()
        except Exception as exc:  # noqa: BLE001
            raise NativeLibraryError(
                f"{LIB_NAME} is missing and could not be built ({exc}). sonar_b200 has no CPU or "
                "eager-PyTorch fallback: run `python comfyui-sonar_b200/build.py`.",
            ) from exc
    if not path.exists():
        raise NativeLibraryError(f"{path} not found; sonar_b200 has no fallback path")
    _warn_if_stale(path)
    try:
        lib = ctypes.CDLL(str(path))
    except OSError as exc:
        raise NativeLibraryError(f"failed to load {path}: {exc}") from exc
    for name, argtypes in SIGNATURES.items():
        try:
            fn = getattr(lib, name)
        except AttributeError as exc:
            raise NativeLibraryError(f"{path} does not export {name}; rebuild the library") from exc
        fn.argtypes = argtypes
        fn.restype = RESTYPES.get(name, c_int)
    if lib.sonar_abi_version() != ABI_VERSION:
        raise NativeLibraryError(
            f"{path} has ABI version {lib.sonar_abi_version()}, this package needs {ABI_VERSION}: rebuild the library",
        )
    _LIB = lib
    return lib


def check(code: int, what: str) -> None:
    if code != 0:
        raise NativeLibraryError(f"{what} failed with CUDA error {code}")


__all__ = [
    "SIGNATURES",
    "SonarPyramidParams",
    "SonarPerlinParams",
    "SonarSpectralParams",
    "SonarWaveletFilters",
    "SonarDwtAnalysisParams",
    "SonarDwtSynthesisParams",
    "SonarWcfgFusedParams",
    "NativeLibraryError",
    "SonarStepParams",
    "check",
    "library_path",
    "load",
    "c_double",
]
