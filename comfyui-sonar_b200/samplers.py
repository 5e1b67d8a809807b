"""Sonar momentum samplers on the fused CUDA step.

Host-side mirror of the reference's py/sonar.py: same config objects (`SonarConfig`,
`GuidanceConfig`, enums), same `sonar_params` merging and error messages (:98-131), same sampler
function signatures and callback payloads (:483-526, :576-623, :773-820), same registration
(`add_samplers`, :823-847). The per-step tensor arithmetic -- ~15 full-tensor ATen passes and 4 host
syncs per Euler-ancestral step upstream -- is ONE launch of `sonar_step_f32` per model evaluation.

Default (Gaussian) ancestral noise never exists in HBM: the step kernel regenerates the values
torch.randn(device='cuda') would have drawn from the Philox stream, in registers. The global
mean / std that scale_noise needs (py/utils.py:100-106) depend only on (seed, generator offset), and
the offsets of all remaining draws of a run are known at the first one, so their moments are
reduced in ONE batched launch up front (`SonarBase._lookahead_sums`); each draw re-checks the
generator offset and re-plans if somebody else consumed random numbers in between.

Scalar schedule math (ancestral split, DPM-Solver++ log-sigma arithmetic) is evaluated once per step
on a CPU float32 mirror of `sigmas` with the reference's own op sequence, so the coefficients handed
to the kernel are the float32 values the reference computes and no step waits on the device.
"""

from __future__ import annotations

import ctypes
import importlib
import os
import weakref
from enum import Enum, auto
from sys import stderr
from typing import Any, Callable, NamedTuple

import torch
from torch import Tensor
from tqdm.auto import trange

from . import hostutil, noise_graph as noise, ops, parallel, rng
from ._native import SonarStepParams


# Pipelined production of look-ahead noise (SonarBase._lookahead_noise). SONAR_B200_NOISE_PIPELINE=0 restores the
# batched schedule; the *_CTAS knobs are the CTAs per SM each side may occupy while both run.
NOISE_PIPELINE = os.environ.get("SONAR_B200_NOISE_PIPELINE", "1") != "0"
NOISE_PIPELINE_CHUNK = int(os.environ.get("SONAR_B200_PIPELINE_CHUNK", "1"))
PIPELINE_STEP_CTAS = int(os.environ.get("SONAR_B200_PIPELINE_STEP_CTAS", "4"))
PIPELINE_FILL_CTAS = int(os.environ.get("SONAR_B200_PIPELINE_FILL_CTAS", "4"))
PIPELINE_FFT_CTAS = int(os.environ.get("SONAR_B200_PIPELINE_FFT_CTAS", "3"))  # >= 3: 256-thread FFT CTAs, 16 K registers stay free
# The step launch is made in two parts: the first PIPELINE_STEP_SPLIT of the elements beside the Philox fill (4 + 4
# CTAs per SM), the rest beside the FFT (3 x 256 threads x 64 registers) with the 2 step CTAs per SM the register file
# still holds. tools/sweep_pipeline.sh on B200, interval between model calls at 8 video latents: 466 us unsplit;
# 406 / 405 / 413 / 420 us at split 0.65 / 0.7 / 0.75 / 0.8 (with the c2r fold fused into the FFT's first row stage; before
# that 417 / 411 / 408 / 405 / 415 us at 0.5 / 0.55 / 0.6 / 0.65 / 0.7). At ~405 us the interval moves its 2.2 GB of actual
# DRAM traffic at 5.4 TB/s (0.84 of the measured peak): a faster FFT alone no longer shortens it.
PIPELINE_STEP_SPLIT = float(os.environ.get("SONAR_B200_PIPELINE_STEP_SPLIT", "0.7"))
PIPELINE_STEP_B_CTAS = int(os.environ.get("SONAR_B200_PIPELINE_STEP_B_CTAS", "2"))
# Below this many elements per sample the batched schedule wins: a one-sample producer launch is far from the throughput
# of a batch of 9-18 (tail of the persistent FFT grid), and a 30 us step hides little of it. Measured on B200 with the
# C5 job (tools/sweep_pipeline.sh, profiles/r02b_noise_pipeline_sweeps.txt): 1 video latent per GPU 1.27 ms batched vs
# 1.57 ms pipelined, 2 / 4 / 8 latents 2.53 / 5.04 / 9.93 ms batched vs 2.21 / 3.97 / 7.65 ms pipelined.
PIPELINE_MIN_NUMEL = int(os.environ.get("SONAR_B200_PIPELINE_MIN_NUMEL", str(12_000_000)))
# Batch-sharded runs exchange the statistics of every pipelined sample separately (18 exchanges per C5 run instead of
# 1-2 per look-ahead batch), each of which also absorbs the skew between the ranks: on 4 GPUs x 2 latents the pipelined
# job took 2.93 ms against 2.49 ms batched (2.38 vs 2.53 ms on one GPU with the same 2 latents); on 2 GPUs x 4 latents
# 4.17 vs 4.67 ms. Sharded runs therefore pipeline only from twice the size.
PIPELINE_MIN_NUMEL_SHARDED = int(os.environ.get("SONAR_B200_PIPELINE_MIN_NUMEL_SHARDED", str(24_000_000)))
_PRODUCER_STREAMS: dict = {}


def _producer_stream(device_index: int) -> "torch.cuda.Stream":
    """The second stream of a device, on which look-ahead noise is produced (high priority: its CTAs take the slots
    that free up first)."""
    st = _PRODUCER_STREAMS.get(device_index)
    if st is None:
        st = _PRODUCER_STREAMS[device_index] = torch.cuda.Stream(device=device_index, priority=-1)
    return st


class HistoryType(Enum):
    ZERO = auto()
    RAND = auto()
    SAMPLE = auto()
    SAMPLE_NORM = auto()


class GuidanceType(Enum):
    LINEAR = auto()
    EULER = auto()


class GuidanceConfig(NamedTuple):
    guidance_type: GuidanceType = GuidanceType.LINEAR
    factor: float = 0.01
    start_step: int = 1
    end_step: int = 9999
    latent: Tensor | None = None


class MomentumMode(Enum):
    CLASSIC = auto()
    NEW = auto()
    DENOISED = auto()


_MODE_IDS = {MomentumMode.CLASSIC: ops.MODE_CLASSIC, MomentumMode.NEW: ops.MODE_NEW, MomentumMode.DENOISED: ops.MODE_DENOISED}


class SonarConfig(NamedTuple):
    momentum: float = 0.95
    momentum_hist: float = 0.75
    direction: float = 1.0
    momentum_start_step: int = 0
    momentum_end_step: int = 9999
    always_update_history: bool = True
    momentum_mode: MomentumMode = MomentumMode.NEW
    init: HistoryType = HistoryType.ZERO
    noise_type: noise.NoiseType | None = None
    custom_noise: noise.CustomNoise | None = None
    rand_init_noise_type: noise.NoiseType | None = None
    rand_init_noise_multiplier: float | int = 1.0
    guidance: GuidanceConfig | None = None
    blend_mode: str = "lerp"
    momentum_blend_mode: str | None = None
    history_blend_mode: str | None = None
    guidance_blend_mode: str | None = None

    def get_with_default(self, k: str, default: Any) -> Any:  # noqa: ANN401
        val = getattr(self, k)
        return val if val is not None else default


LOOKAHEAD_MAX_DRAWS = 256  # statistics of at most this many future noise draws per batched launch


class SonarBase:
    """Momentum state + the fused step launcher (reference SonarBase, py/sonar.py:70-320)."""

    DEFAULT_NOISE_TYPE = noise.NoiseType.GAUSSIAN

    def __init__(self, cfg: SonarConfig) -> None:
        self.history_d: Tensor | None = None
        self.cfg = cfg
        self.noise_sampler = None
        base = cfg.blend_mode
        # BLENDING_MODES[...] raises KeyError for unknown names, like the reference
        self.blend = hostutil.BLENDING_MODES[base]
        self.momentum_blend = hostutil.BLENDING_MODES[cfg.get_with_default("momentum_blend_mode", base)]
        self.history_blend = hostutil.BLENDING_MODES[cfg.get_with_default("history_blend_mode", base)]
        self.guidance_blend = hostutil.BLENDING_MODES[cfg.get_with_default("guidance_blend_mode", base)]
        self._hist_pending_init: tuple[Tensor, float] | None = None
        # look-ahead statistics of the fused Gaussian noise draws (see _lookahead_sums)
        self._lookahead: dict | None = None
        self._lookahead_cap = LOOKAHEAD_MAX_DRAWS
        self.noise_draws_left = 1  # samplers set this to the number of ancestral draws of the run

    _cfg_fixups = (
        ("momentum_mode", MomentumMode),
        ("init", HistoryType),
        ("noise_type", noise.NoiseType),
    )

    @classmethod
    def get_config(cls, cfg: SonarConfig | None = None, ext: dict | None = None) -> SonarConfig:
        merged = dict(ext) if ext is not None else {}
        missing = object()
        for key, enum_class in cls._cfg_fixups:
            val = merged.get(key, missing)
            if val is missing:
                continue
            if isinstance(val, str):
                member = getattr(enum_class, val.strip().upper(), missing)
                if member is missing:
                    valid = ", ".join(enum_class.__members__.keys())
                    raise ValueError(
                        f"Bad value for {key} of type enum {enum_class.__name__}, must be one of the following: {valid}",
                    )
                merged[key] = member
            elif not isinstance(val, enum_class):
                raise TypeError(
                    f"Bad parameter type for {key}: Must be valid string or instance of {enum_class.__name__}",
                )
        if cfg is None:
            return SonarConfig(**merged)
        return SonarConfig(**(cfg._asdict() | merged))

    def set_noise_sampler(self, x: Tensor, sigmas: Tensor, noise_sampler: Callable | None, seed: int | None = None):
        sigmas_host = getattr(self, "sigmas_host", None)  # the run's one device->host copy of the schedule
        if sigmas_host is None or sigmas_host.shape != sigmas.shape:
            sigmas_host = sigmas.detach().float().cpu()
        sigma_min, sigma_max = sigmas_host[sigmas_host > 0].min(), sigmas_host.max()
        if noise_sampler is not None and self.cfg.noise_type not in {None, self.DEFAULT_NOISE_TYPE}:
            print("Sonar: Warning: Noise sampler supplied, overriding noise type from settings", file=stderr)
        if self.cfg.custom_noise:
            noise_sampler = self.cfg.custom_noise.make_noise_sampler(x, sigma_min, sigma_max, seed=seed)
        elif noise_sampler is None:
            noise_sampler = noise.get_noise_sampler(
                self.cfg.noise_type or self.DEFAULT_NOISE_TYPE,
                x,
                sigma_min,
                sigma_max,
                seed=seed,
                cpu=True,
                normalized=True,
            )
        self.noise_sampler = noise_sampler
        return noise_sampler

    @property
    def history_ratios(self) -> tuple[float, float, float]:
        direction, momentum_hist = self.cfg.direction, self.cfg.momentum_hist
        hd_scale = 1.0 + abs(direction) * (1 - momentum_hist) if direction < 0 else 2.0 - direction
        return (momentum_hist, hd_scale, direction)

    def check_step(self, step: int, *, is_history: bool = False) -> bool:
        cfg = self.cfg
        if is_history and cfg.always_update_history:
            return True
        return cfg.momentum_start_step <= step <= cfg.momentum_end_step

    # ------------------------------------------------------------------------------------
    # history initialisation (reference init_hist_d, :169-206). It runs once, after the first
    # momentum mix of the call, so the kernel gets HIST_INIT: used by updates, not by the first mix.
    # ------------------------------------------------------------------------------------
    def _initial_history(self, x: Tensor, denoised: Tensor, sigma: float, *, step: int):
        """Returns (tensor, divisor) for a history born in this call, or None."""
        if self.history_d is not None or not self.check_step(step, is_history=True):
            return None
        cfg, init = self.cfg, self.cfg.init
        source = denoised if cfg.momentum_mode == MomentumMode.DENOISED else x
        if init == HistoryType.ZERO:
            return None
        if init == HistoryType.SAMPLE:
            return (source, 1.0)
        if init == HistoryType.SAMPLE_NORM:
            return (source, sigma)
        if init == HistoryType.RAND:
            ns = noise.get_noise_sampler(
                cfg.rand_init_noise_type,
                x,
                None,
                None,
                seed=self.extra_args.get("seed"),
                cpu=True,
                normalized=True,
            )
            hist = ns(None, None)
            if cfg.rand_init_noise_multiplier != 1:
                ops.scale(hist, cfg.rand_init_noise_multiplier)
            return (hist, 1.0)
        raise ValueError("Sonar sampler: bad history type")

    # ------------------------------------------------------------------------------------
    # the fused launch
    # ------------------------------------------------------------------------------------
    def _step_params(self) -> SonarStepParams:
        """Per-sampler parameter block; the fields that never change are filled once."""
        p = getattr(self, "_params", None)
        if p is None:
            cfg = self.cfg
            p = self._params = SonarStepParams()
            p.mode = _MODE_IDS[cfg.momentum_mode]
            p.momentum_blend = hostutil.blend_mode_id(self.momentum_blend)
            p.history_blend = hostutil.blend_mode_id(self.history_blend)
            p.momentum = cfg.momentum
            p.hd_ratio, p.hd_scale, p.md_scale = self.history_ratios
            p.noise_threshold_std_devs = 2.5
            p.hist_in_div = 1.0
            self._params_ref = ctypes.byref(p)
            # step gating (check_step, :221-225) as plain attributes: this runs every step
            self._gate = (cfg.momentum_start_step, cfg.momentum_end_step, bool(cfg.always_update_history), cfg.momentum_hist != 1)
        return p

    def fused_step(
        self,
        step: int,
        x: Tensor,
        denoised: Tensor,
        sigma: float,
        *,
        kind: int,
        c0: float,
        c1: float = 0.0,
        noise_tensor: Tensor | None = None,
        noise_scale: float = 0.0,
        noise_philox: dict | None = None,
        noise_deferred: tuple | None = None,
    ) -> Tensor:
        """One C-ABI call: momentum mix, both history updates, Euler / DPM++ update, noise injection.

        This is the per-step host path (the end-to-end number at SDXL-sized latents is bound by it, not
        by the ~6 us kernel), hence the flat code: one parameter block reused across steps, no helper
        layers, no per-step allocation besides the two output tensors."""
        if x.dtype != torch.float32 or not x.is_cuda:
            raise TypeError(f"sonar_b200 samplers run on float32 CUDA latents (got {x.dtype} on {x.device})")
        if not x.is_contiguous():
            x = x.contiguous()
        if denoised.dtype != torch.float32 or not denoised.is_contiguous():
            denoised = denoised.to(torch.float32).contiguous()
        if denoised.device != x.device:
            raise RuntimeError(f"tensors on different devices: {x.device} vs {denoised.device}")
        p = self._step_params()
        self._stock_static = None  # this path rewrites fields the stock lane treats as per-run constants
        start, end, always, hist_on = self._gate
        in_window = start <= step <= end
        history_active = hist_on and (always or in_window)

        hist_in = self.history_d
        x_out = torch.empty_like(x)
        if hist_in is not None:  # steady state: history updated in place
            hist_state, hist_out = ops.HIST_PRESENT, hist_in
            p.hist_in = p.hist_out = hist_in.data_ptr()
            p.hist_in_div = 1.0
        else:
            born, self._hist_pending_init = self._hist_pending_init, None
            if born is None and self.cfg.init != HistoryType.ZERO:
                born = self._initial_history(x, denoised, sigma, step=step)
            if born is None:
                hist_state, p.hist_in, p.hist_in_div = ops.HIST_NONE, 0, 1.0
            else:
                hist_in, p.hist_in_div = born
                hist_in = hist_in.contiguous()
                hist_state, p.hist_in = ops.HIST_INIT, hist_in.data_ptr()
            # a history survives this call if one was born or an update creates it; never in place
            # here (a born history may alias x / denoised)
            hist_out = torch.empty_like(x) if (born is not None or history_active) else None
            p.hist_out = 0 if hist_out is None else hist_out.data_ptr()
        p.x, p.denoised, p.x_out = x.data_ptr(), denoised.data_ptr(), x_out.data_ptr()
        p.n = x.numel()
        p.kind = kind
        p.hist_state = hist_state
        p.momentum_active = in_window
        p.history_active = history_active
        p.sigma, p.c0, p.c1 = sigma, c0, c1
        p.noise_scale = noise_scale
        keep = None
        if noise_philox is not None:
            draw = noise_philox["draw"]
            p.noise_factor = noise_philox["factor"]
            p.philox_seed, p.philox_offset, p.philox_grid_blocks = draw.seed, draw.offset, draw.grid_blocks
            p.noise_begin, p.noise_numel_total = noise_philox["begin"], draw.numel
            p.peer_world = 0
            if noise_philox["normalized"]:
                raw = noise_philox.get("tensor")
                if raw is not None:  # materialised raw normals, normalised on load from device sums
                    keep = (noise_philox["sums"], raw)
                    p.noise_kind, p.noise = ops.NOISE_TENSOR_NORMALIZED, raw.data_ptr()
                    p.noise_sums = keep[0].data_ptr()
                else:  # regenerated in registers; sums reduced ahead of time (look-ahead batch)
                    keep = noise_philox["sums"]
                    p.noise_kind, p.noise_sums = ops.NOISE_PHILOX_NORMALIZED, noise_philox["sums_ptr"]
                    p.noise_decision = noise_philox.get("decision_ptr", 0)
                p.noise_count = noise_philox["count"]
            else:
                p.noise_kind = ops.NOISE_PHILOX
        elif noise_deferred is not None:  # un-normalised tensor + its global {sum, sum^2}: scale_noise on load
            raw, sums, count, factor = keep = noise_deferred
            if raw.device != x.device:
                raise RuntimeError(f"tensors on different devices: {x.device} vs {raw.device}")
            p.noise_kind, p.noise, p.peer_world = ops.NOISE_TENSOR_NORMALIZED, raw.data_ptr(), 0
            p.noise_sums, p.noise_count, p.noise_factor = sums.data_ptr(), count, factor
        elif noise_tensor is not None:
            if noise_tensor.dtype != torch.float32 or not noise_tensor.is_contiguous():
                noise_tensor = noise_tensor.to(torch.float32).contiguous()
            if noise_tensor.device != x.device:
                raise RuntimeError(f"tensors on different devices: {x.device} vs {noise_tensor.device}")
            p.noise_kind, p.noise, p.peer_world, keep = ops.NOISE_TENSOR, noise_tensor.data_ptr(), 0, noise_tensor
        else:
            p.noise_kind = ops.NOISE_NONE
        join = getattr(self, "_noise_join", None)
        if join is None:
            ops.launch_step(self._params_ref, x.device.index)
        else:
            # a producer is running on the second stream: leave it thread slots on every SM, then make everything
            # enqueued after this half step wait for it (the next model call starts with the next sample ready)
            self._noise_join = None
            n = p.n
            n_a = (int(n * PIPELINE_STEP_SPLIT) // 4) * 4 if PIPELINE_STEP_SPLIT < 1.0 else n
            try:
                ops.set_grid_limit(PIPELINE_STEP_CTAS)
                if 0 < n_a < n:
                    # two launches over the flat tensors: the first beside the Philox fill, the second (fewer CTAs
                    # per SM) beside the co-scheduled FFT
                    p.n = n_a
                    ops.launch_step(self._params_ref, x.device.index)
                    ops.set_grid_limit(PIPELINE_STEP_B_CTAS)
                    shift = 4 * n_a
                    p.n = n - n_a
                    for name in ("x", "denoised", "x_out", "hist_in", "hist_out", "noise"):
                        ptr = getattr(p, name)
                        if ptr:
                            setattr(p, name, ptr + shift)
                    ops.launch_step(self._params_ref, x.device.index)
                else:
                    ops.launch_step(self._params_ref, x.device.index)
            finally:
                ops.set_grid_limit(0)
            torch.cuda.current_stream(x.device).wait_event(join)  # (the stream of x's device: the one the step was launched on)
        del keep
        if hist_out is not None:
            self.history_d = hist_out
        return x_out

    def _lookahead_sums(self, draw: ops.PhiloxDraw, begin: int, count: int, device, total: int) -> tuple[Tensor, int, int]:
        """Device statistics of the whole (global) normal draw `draw`: (keep-alive, pointer to its
        (sum, sum^2), pointer to its precomputed scale_noise decision).

        First request of a run: ONE batched launch reduces the moments of this rank's slice of this
        draw and of the next `noise_draws_left - 1` draws (their offsets follow from the generator's
        fixed increment per draw), plus one all-reduce of the (K, 2) table when the batch is sharded.
        Later requests look their offset up. A miss means another consumer advanced the generator
        in between: re-plan from the current offset and stop looking more than one draw ahead."""
        key = (draw.seed, draw.grid_blocks, draw.numel, begin, count, device)
        la = self._lookahead
        idx = None
        if la is not None and la["key"] == key:
            idx = la["index"].get(draw.offset)
            if idx is None:
                self._lookahead_cap = 1
        if idx is None:
            k = max(1, min(self.noise_draws_left, self._lookahead_cap))
            offsets = [draw.offset + j * draw.counter_offset for j in range(k)]
            sums = ops.philox_normal_moments_batch(draw, offsets, begin=begin, count=count, device=device)
            parallel.allreduce_table(sums)  # sharded: sum the partial sums over ranks, once per table
            # the conditional normalisation of every draw, decided once (fp64) instead of in every CTA of
            # every step launch
            decisions = ops.norm_decisions(sums, total)
            la = self._lookahead = {
                "key": key, "index": {o: j for j, o in enumerate(offsets)}, "sums": (sums, decisions),
                "ptr": sums.data_ptr(), "dec_ptr": decisions.data_ptr(), "inc": draw.counter_offset,
            }  # fmt: skip
            idx = 0
        return la["sums"], la["ptr"] + 16 * idx, la["dec_ptr"] + 16 * idx

    # ------------------------------------------------------------------------------------
    # stock lane: ZERO-init history + (optional) fused Gaussian noise, nothing else configured
    # ------------------------------------------------------------------------------------
    def stock_lane(self, x: Tensor) -> bool:
        """True when every step of this run can take `stock_step`: float32 CUDA latent, ZERO history
        init, no guidance, and the ancestral noise (if any) is the fusable plain Gaussian. Decided once."""
        lane = getattr(self, "_stock_lane", None)
        if lane is None:
            guided = getattr(self, "guidance", None) is not None and self.guidance.factor != 0.0
            lane = self._stock_lane = (
                not guided
                and self.cfg.init == HistoryType.ZERO
                and x.dtype == torch.float32
                and x.is_cuda
                and (self.noise_draws_left == 0 or self._fused_noise_spec(x) is not None)
            )
            if lane:
                self._step_params()
                spec = self._fused_noise_spec(x) if self.noise_draws_left else None
                self._stock_noise = spec  # (factor, normalized) or None
                self._stock_static = None
                idx = x.device.index if x.device.index is not None else torch.cuda.current_device()
                self._stock_gen, self._stock_dev = torch.cuda.default_generators[idx], idx
        return lane

    def stock_step(self, step: int, x: Tensor, denoised: Tensor, sigma: float, kind: int, c0: float, c1: float,
                   noise_scale: float | None) -> Tensor | None:
        """fused_step + ancestral_noise for the stock lane with every helper layer flattened out (this is
        the per-step host cost the end-to-end number is made of). Returns None when the step must take
        the general path (unexpected tensor layout, injected RNG, look-ahead miss)."""
        if (
            denoised.dtype != torch.float32
            or denoised.device != x.device
            or not denoised.is_contiguous()
            or not x.is_contiguous()
            or rng._INJECT is not None  # noqa: SLF001
        ):
            return None
        p = self._params
        n = x.numel()
        static = self._stock_static
        if static != (n, kind):
            # fields that do not change from step to step are written once per run (a ctypes field store
            # costs ~0.2 us; the general path invalidates this by resetting _stock_static)
            start, end, always, hist_on = self._gate
            self._stock_window = (start, end, hist_on and always, hist_on)
            # batch-sharded runs regenerate this rank's slice [begin, begin + n) of the GLOBAL draw of `total` values
            ctx = parallel.active()
            total, begin = parallel.global_draw_geometry(x.shape) if ctx is not None and ctx.world_size > 1 else (n, 0)
            self._stock_total = total
            p.n, p.kind, p.hist_in_div, p.peer_world, p.noise_begin, p.noise_numel_total, p.noise_count = n, kind, 1.0, 0, begin, total, n
            spec = self._stock_noise
            if spec is not None:
                p.noise_factor = spec[0]
            self._stock_static = (n, kind)
        if noise_scale is not None:
            spec = self._stock_noise
            if spec is None:
                return None
            gen = self._stock_gen
            offset = gen.get_offset()
            total = self._stock_total
            if spec[1]:  # normalised: statistics from the look-ahead table
                la = self._lookahead
                idx = None if la is None else la["index"].get(offset)
                if idx is None or la["key"][0] != gen.initial_seed() or la["key"][2] != total or la["key"][4] != n:
                    return None  # first draw of the run, or somebody else advanced the generator: re-plan
                p.noise_kind, p.noise_sums, p.noise_decision = ops.NOISE_PHILOX_NORMALIZED, la["ptr"] + 16 * idx, la["dec_ptr"] + 16 * idx
                grid, inc = la["key"][1], la["inc"]
            else:
                p.noise_kind = ops.NOISE_PHILOX
                grid, inc = ops.philox_policy_cached(self._stock_dev, total)
            gen.set_offset(offset + inc)
            self.noise_draws_left -= 1
            p.noise_scale, p.philox_seed, p.philox_offset, p.philox_grid_blocks = noise_scale, gen.initial_seed(), offset, grid
        else:
            p.noise_kind = ops.NOISE_NONE
        start, end, hist_always, hist_on = self._stock_window
        in_window = start <= step <= end
        history_active = hist_always or (hist_on and in_window)
        x_out = torch.empty_like(x)
        hist = self.history_d
        if hist is not None:
            p.hist_state = ops.HIST_PRESENT
            p.hist_in = p.hist_out = hist.data_ptr()
        else:
            p.hist_state, p.hist_in = ops.HIST_NONE, 0
            if history_active:
                hist = torch.empty_like(x)
                p.hist_out = hist.data_ptr()
            else:
                p.hist_out = 0
        p.x, p.denoised, p.x_out = x.data_ptr(), denoised.data_ptr(), x_out.data_ptr()
        p.momentum_active, p.history_active = in_window, history_active
        p.sigma, p.c0, p.c1 = sigma, c0, c1
        ops.launch_step(self._params_ref, self._stock_dev)
        self.history_d = hist
        return x_out

    def _momentum_denoised_preview(self, step: int, denoised: Tensor, sigma: float) -> Tensor:
        """What get_momentum_denoised (:262-283) would return for `denoised` right now, without touching
        the history: only DENOISED mode mixes the prediction with history * sigma."""
        cfg = self.cfg
        hist = self.history_d
        if cfg.momentum_mode != MomentumMode.DENOISED or cfg.momentum == 1 or hist is None or not self.check_step(step):
            return denoised
        scaled = ops.axpby(hist, sigma, None)
        return self.momentum_blend(scaled, denoised.to(torch.float32).contiguous(), cfg.momentum)

    def prime_history(self, step: int, x: Tensor, denoised: Tensor, sigma: float) -> None:
        """Performs the (possibly random) history initialisation of this step NOW. The reference draws
        the RAND history inside momentum_step, i.e. before the step's ancestral noise; callers that
        sample noise ahead of the fused launch call this first so the draw order is preserved."""
        if self.history_d is None and self._hist_pending_init is None and self.cfg.init != HistoryType.ZERO:
            self._hist_pending_init = self._initial_history(x, denoised, sigma, step=step)

    def momentum_step(self, step: int, x: Tensor, denoised: Tensor, sigma: float, sigma_down: float, *, dt=None, **noise_kw):
        """x + momentum_d * (sigma_down - sigma) (:309-320), optionally with the ancestral noise fused in."""
        if dt is None:  # float32 subtraction, like the reference's 0-d tensors
            dt = float(torch.tensor(sigma_down, dtype=torch.float32) - torch.tensor(sigma, dtype=torch.float32))
        return self.fused_step(step, x, denoised, sigma, kind=ops.STEP_EULER, c0=dt, **noise_kw)

    # ------------------------------------------------------------------------------------
    # ancestral noise: fuse the default Gaussian, otherwise sample a tensor
    # ------------------------------------------------------------------------------------
    def ancestral_noise(self, x: Tensor, sigma: Tensor, sigma_next: Tensor, scale: float) -> dict:
        """kwargs for fused_step that add noise_sampler(sigma, sigma_next) * scale."""
        spec = self._fused_noise_spec(x)
        if spec is None:
            ahead = self._lookahead_noise(x)
            if ahead is not None:
                return {"noise_deferred": ahead, "noise_scale": scale}
            deferred = getattr(self.noise_sampler, "deferred", None)
            if deferred is not None:
                # chain noise: skip its final scale_noise pass (a read + a write of the whole tensor) and
                # let the step kernel normalise while it loads, from the statistics the producer kernel
                # reduced (or one moments pass); sharded runs sum the two doubles over ranks first
                raw, factor, normalized = deferred(sigma, sigma_next)
                if normalized and raw.dtype == torch.float32 and raw.shape == x.shape and raw.numel() and raw.is_contiguous():
                    sums = ops.attached_sums(raw)
                    if sums is None:
                        sums = ops.moments(raw)
                    parallel.allreduce_table(sums)
                    count = parallel.global_numel(raw.numel())
                    return {"noise_deferred": (raw, sums, count, factor), "noise_scale": scale}
                return {"noise_tensor": hostutil.scale_noise(raw, factor, normalized=normalized), "noise_scale": scale}
            return {"noise_tensor": self.noise_sampler(sigma, sigma_next), "noise_scale": scale}
        factor, normalized = spec
        ctx = parallel.active()
        if ctx is not None and ctx.world_size > 1:
            total, begin = parallel.global_draw_geometry(x.shape)
        else:
            total, begin = x.numel(), 0
        draw = ops.reserve_draw(total, x.device)
        kw = {"draw": draw, "factor": factor, "normalized": normalized, "begin": begin}
        if normalized:
            kw["sums"], kw["sums_ptr"], kw["decision_ptr"] = self._lookahead_sums(draw, begin, x.numel(), x.device, total)
            kw["count"] = total
        self.noise_draws_left -= 1
        return {"noise_philox": kw, "noise_scale": scale}

    def _lookahead_noise(self, x: Tensor):
        """Custom noise whose samples depend on the generator state only (power noise): samples are made ahead of the
        request and handed out one per request. Every hand-out checks that torch's generator is where the producer
        assumed it would be (nobody else drew in between) and advances it exactly as the draw itself would have;
        otherwise what was made ahead is dropped and made again from the current state. Returns the `noise_deferred`
        tuple.

        Two schedules. Batched (NOISE_PIPELINE off): the samples of the next few draws in one Philox launch + one FFT
        launch + one statistics exchange. Pipelined (default): the producers are instruction-issue-bound (~70 % of
        the issue slots, a third of the HBM bandwidth) and the fused step is HBM-bound (~20 % of the issue slots), so the
        NEXT sample is produced on a second stream while the step kernel that consumes THIS one runs: both kernels
        are launched with grids that leave each other thread slots on every SM (sonar_set_grid_limit) and the main
        stream joins the producer stream right after the step launch, so everything still happens between the
        two model calls that bracket the half step."""
        ns = self.noise_sampler
        make = getattr(ns, "lookahead", None)
        if make is None or self.noise_draws_left < 1:
            return None
        factor, normalized = ns.lookahead_pending
        if not normalized:
            return None
        idx = x.device.index if x.device.index is not None else torch.cuda.current_device()
        gen = torch.cuda.default_generators[idx]
        queue = getattr(self, "_noise_queue", None)
        pending = getattr(self, "_noise_pending", None)
        if not queue and pending is not None:
            batch, ready = pending
            self._noise_pending = None
            torch.cuda.current_stream(x.device).wait_event(ready)  # (already joined after the previous step launch)
            queue = self._noise_queue = batch
        if queue and (queue[0][2].offset != gen.get_offset() or queue[0][2].seed != gen.initial_seed() or queue[0][0].shape != x.shape):
            queue.clear()
        ctx = parallel.active()
        sharded = ctx is not None and ctx.world_size > 1
        pipelined = NOISE_PIPELINE and x.numel() >= (PIPELINE_MIN_NUMEL_SHARDED if sharded else PIPELINE_MIN_NUMEL)
        if not queue:
            if pipelined:
                made = self._produce_noise(make, x, overlapped=False)
                if made is None:
                    return None
                batch, ready = made
                torch.cuda.current_stream(x.device).wait_event(ready)
            else:
                batch = make(self.noise_draws_left)
                if batch is None:
                    return None
                parallel.allreduce_table(batch.table)  # sharded: ONE exchange for the statistics of the whole batch
            queue = self._noise_queue = batch
        raw, sums, draw = queue.pop(0)
        gen.set_offset(draw.offset + draw.counter_offset)
        self.noise_draws_left -= 1
        if pipelined and not queue and self.noise_draws_left >= 1:
            # speculative: the next request normally finds the generator exactly here
            self._noise_pending = self._produce_noise(make, x, overlapped=True)
        return (raw, sums, parallel.global_numel(raw.numel()), factor)

    def _produce_noise(self, make, x: Tensor, *, overlapped: bool):
        """Enqueues the production of the next NOISE_PIPELINE_CHUNK samples on the producer stream: (batch, event).
        overlapped: the consumer launches its step kernel next, to run beside the producers (`_noise_join`)."""
        idx = x.device.index if x.device.index is not None else torch.cuda.current_device()
        side = _producer_stream(idx)
        main = torch.cuda.current_stream(idx)
        slot = getattr(self, "_noise_slot", 0)
        # the buffer about to be overwritten was last read by a step kernel already enqueued on the consumer stream
        side.wait_stream(main)
        with torch.cuda.stream(side):
            batch = make(self.noise_draws_left, chunk=NOISE_PIPELINE_CHUNK, slot=slot,
                         grid_limits=(PIPELINE_FILL_CTAS, PIPELINE_FFT_CTAS) if overlapped else None)
            if batch is None:
                return None
            parallel.allreduce_table(batch.table)  # sharded: the statistics of the chunk, on the producer stream
            ready = torch.cuda.Event()
            ready.record(side)
        self._noise_slot = slot ^ 1
        if overlapped:
            self._noise_join = ready
        return batch, ready

    def _fused_noise_spec(self, x: Tensor):
        """(factor, normalized) when the noise sampler is plain Gaussian noise of x's shape that the
        step kernel can regenerate from the Philox stream; None otherwise. Cached per sampler."""
        if rng._INJECT is not None:  # noqa: SLF001  (parity harness feeds recorded draws as tensors)
            return None
        cached = getattr(self, "_fused_spec", False)
        if cached is False:
            ns = self.noise_sampler
            spec = ns.fused_gaussian() if hasattr(ns, "fused_gaussian") else None
            cached = self._fused_spec = None if spec is None else ((spec[0], spec[1]), tuple(spec[2]))
        if cached is None or cached[1] != x.shape:
            return None
        return cached[0]


class SonarGuidanceMixin:
    """Reference-latent guidance (:323-411): after the fused step, x is pulled towards a reference
    latent that has been given the per-batch-item mean / std of x (LINEAR) or of the denoised prediction
    (EULER). Two launches per guided step: per-item moments, then the shift + blend / Euler update."""

    def __init__(self, cfg: GuidanceConfig | None = None) -> None:
        self.guidance = cfg
        self.ref_latent = self.prepare_ref_latent(cfg.latent) if cfg and cfg.latent is not None else None

    @staticmethod
    def prepare_ref_latent(latent: Tensor | None) -> Tensor | None:
        """Per-plane standardisation of the reference latent (:335-341); setup, once per sampler."""
        if latent is None:
            return None
        avg = latent.mean(dim=(-2, -1), keepdim=True)
        std = latent.std(dim=(-2, -1), keepdim=True)
        return (latent - avg).div_(std).to(latent.dtype)

    def _guidance_ref(self, x: Tensor) -> Tensor:
        ref = self.ref_latent
        if ref.device != x.device or ref.dtype != torch.float32 or not ref.is_contiguous():
            ref = self.ref_latent = ref.to(device=x.device, dtype=torch.float32).contiguous()
        return ref

    def guidance_step(self, step_index: int, x: Tensor, denoised: Tensor) -> Tensor:
        g = self.guidance
        if g is None or g.factor == 0.0 or not g.start_step <= step_index <= g.end_step:
            return x
        ref = self._guidance_ref(x)
        x = x.contiguous()
        if g.guidance_type == GuidanceType.LINEAR:
            return self.guidance_linear(x, ref, g.factor, blend=self.guidance_blend)
        if g.guidance_type == GuidanceType.EULER:
            sh = self.sigma_host_views
            return self.guidance_euler(sh[step_index], sh[step_index + 1], x, denoised, ref, g.factor)
        raise ValueError("Sonar: Guidance: Unknown guidance type")

    @classmethod
    def guidance_euler(cls, sigma, sigma_next, x, denoised, ref_latent, factor: float = 0.2, *, do_shift: bool = True):
        sigma, sigma_next = torch.as_tensor(sigma, dtype=torch.float32).cpu(), torch.as_tensor(sigma_next, dtype=torch.float32).cpu()
        if torch.equal(sigma, sigma_next):
            return cls.guidance_linear(x, ref_latent, factor=factor, do_shift=do_shift)
        sums = ops.item_moments(denoised.to(torch.float32).contiguous()) if do_shift else None
        dt = float((sigma_next - sigma) * factor)  # float32, like the reference's 0-d tensor arithmetic
        return ops.guidance(x, ref_latent, sums, kind=ops.GUIDANCE_EULER, sigma=float(sigma), dt=dt)

    @classmethod
    def guidance_linear(cls, x, ref_latent, factor: float = 0.2, *, blend=None, do_shift: bool = True):
        sums = ops.item_moments(x) if do_shift else None
        mode = hostutil.blend_mode_id("lerp" if blend is None else blend)
        return ops.guidance(x, ref_latent, sums, kind=ops.GUIDANCE_LINEAR, blend_mode=mode, factor=factor)


class SonarWithGuidance(SonarBase, SonarGuidanceMixin):
    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        SonarGuidanceMixin.__init__(self, self.cfg.guidance)


class SonarSampler(SonarWithGuidance):
    def __init__(self, model, sigmas: Tensor, s_in: Tensor, extra_args: dict[str, Any], *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.model = model
        self.sigmas = sigmas
        self.s_in = s_in
        self.extra_args = extra_args
        # one device->host copy of the schedule for the whole run; per-step scalars come from here
        self.sigmas_host = _host_schedule(sigmas)
        self.sigma_views = sigmas.unbind(0)  # 0-d views made once: no per-step indexing op
        self.sigma_host_views = self.sigmas_host.unbind(0)
        # model(x, sigma * s_in): the products of the whole schedule in one op, then views
        self.sigma_in = (sigmas.to(device=s_in.device, dtype=s_in.dtype).unsqueeze(1) * s_in.unsqueeze(0)).unbind(0)
        self.noise_draws_left = self.count_noise_draws()

    NOISE_DRAWS_PER_STEP = 0

    def count_noise_draws(self) -> int:
        """Ancestral noise draws of the whole run (sizes the look-ahead statistics batch)."""
        return self.NOISE_DRAWS_PER_STEP * int((self.sigmas_host[1:] > 0).sum())

    def call_model(self, x: Tensor, sigma: Tensor, *args, s_in=None, extra_args=None) -> Tensor:
        s_in = self.s_in if s_in is None else s_in
        extra_args = self.extra_args if extra_args is None else self.extra_args | extra_args
        return self.model(x, sigma * s_in, *args, **extra_args)

    def run(self, x: Tensor, callback, disable) -> Tensor:
        n_steps = len(self.sigmas) - 1
        for i in range(n_steps) if disable else trange(n_steps, disable=disable):  # (a disabled tqdm still costs ~40 us to build)
            x, sigma, sigma_hat, denoised = self.step(i, x)
            if callback is not None:
                callback({"x": x, "i": i, "sigma": self.sigma_views[i], "sigma_hat": sigma_hat, "denoised": denoised})
        return x

    @classmethod
    def _sample(cls, ctor_args: tuple, model, x, sigmas, extra_args, callback, disable, noise_sampler, sonar_config, sonar_params):
        """Body of the public sampler functions. Latents in another float format (fp16 / bf16 / fp64 models) are
        sampled in float32 -- the kernels' arithmetic type -- and the result is cast back."""
        latent_dtype = x.dtype
        if x.is_cuda and latent_dtype in (torch.float16, torch.bfloat16, torch.float64):
            x = x.to(torch.float32)
        sonar = cls._build(ctor_args, model, x, sigmas, extra_args, noise_sampler, sonar_config, sonar_params)
        out = sonar.run(x, callback, disable)
        return out if out.dtype == latent_dtype else out.to(latent_dtype)

    @classmethod
    def _build(cls, ctor_args: tuple, model, x, sigmas, extra_args, noise_sampler, sonar_config, sonar_params):
        sonar_config = cls.get_config(sonar_config, sonar_params)
        s_in = x.new_ones((x.shape[0],))
        sonar = cls(*ctor_args, model, sigmas, s_in, {} if extra_args is None else extra_args, sonar_config)
        sonar.set_noise_sampler(x, sigmas, noise_sampler, seed=(extra_args or {}).get("seed"))
        return sonar


_SCHEDULE_CACHE: list = [None]  # (weak reference to the device tensor, its version counter, float32 host copy)


def _host_schedule(sigmas: Tensor) -> Tensor:
    """float32 host copy of the sigma schedule. The copy is the only host synchronisation of a sampler run
    (the reference synchronises every step through `.item()`); a run that is handed the very same, unmodified
    tensor object again (serving loops, benchmarks) reuses the previous copy and never blocks on the stream."""
    cached = _SCHEDULE_CACHE[0]
    if cached is not None and cached[0]() is sigmas and cached[1] == sigmas._version:  # noqa: SLF001
        return cached[2]
    host = sigmas.detach().to(dtype=torch.float32, device="cpu")
    if host.data_ptr() == sigmas.data_ptr():  # float32 CPU schedule: `to` returned the same storage
        host = host.clone()
    _SCHEDULE_CACHE[0] = (weakref.ref(sigmas), sigmas._version, host)  # noqa: SLF001
    return host


def ancestral_steps(sigma_from: Tensor, sigma_to: Tensor, eta: float) -> tuple[Tensor, Tensor]:
    """get_ancestral_step for a whole schedule at once: the same float32 element-wise operations
    (mul, sub, div, sqrt, min) the reference applies to 0-d tensors, so identical values."""
    if not eta:
        return sigma_to.clone(), torch.zeros_like(sigma_to)
    sigma_up = torch.minimum(sigma_to, eta * (sigma_to**2 * (sigma_from**2 - sigma_to**2) / sigma_from**2) ** 0.5)
    sigma_down = (sigma_to**2 - sigma_up**2) ** 0.5
    return sigma_down, sigma_up


class SonarEuler(SonarSampler):
    def schedule(self) -> list:
        """Per-step host scalars, computed once with float32 tensor arithmetic like the reference."""
        sched = getattr(self, "_schedule", None)
        if sched is None:
            sh = self.sigmas_host
            dt = (sh[1:] - sh[:-1]).tolist()
            sched = self._schedule = list(zip(sh[:-1].tolist(), sh[1:].tolist(), dt))
        return sched

    def step(self, step_index: int, sample: Tensor):
        sigma = self.sigma_views[step_index]
        sigma_f, sigma_next_f, dt = self.schedule()[step_index]
        denoised = self.model(sample, self.sigma_in[step_index], **self.extra_args)
        if self.stock_lane(sample):
            result = self.stock_step(step_index, sample, denoised, sigma_f, ops.STEP_EULER, dt, 0.0, None)
            if result is not None:
                return (result, sigma, sigma, denoised)
        result = self.momentum_step(step_index, sample, denoised, sigma_f, sigma_next_f, dt=dt)
        if sigma_next_f > 0:
            result = self.guidance_step(step_index, result, denoised)
        return (result, sigma, sigma, denoised)

    @classmethod
    def sampler(
        cls,
        model,
        x: Tensor,
        sigmas: Tensor,
        extra_args: dict | None = None,
        callback=None,
        disable: bool | None = None,  # noqa: FBT001
        noise_sampler: Callable | None = None,
        sonar_config: SonarConfig | None = None,
        sonar_params: dict | None = None,
    ) -> Tensor:
        return cls._sample((), model, x, sigmas, extra_args, callback, disable, noise_sampler, sonar_config, sonar_params)


class SonarEulerAncestral(SonarSampler):
    NOISE_DRAWS_PER_STEP = 1

    def __init__(self, eta: float = 1.0, s_noise: float = 1.0, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.eta = eta
        self.s_noise = s_noise

    def schedule(self) -> list:
        """(sigma, sigma_next, sigma_down, dt, s_noise*sigma_up) per step; get_ancestral_step and the
        subtraction run on float32 0-d host tensors exactly as upstream (:546-551, :317)."""
        sched = getattr(self, "_schedule", None)
        if sched is None:
            sh = self.sigmas_host
            sigma_down, sigma_up = ancestral_steps(sh[:-1], sh[1:], self.eta)
            dt = sigma_down - sh[:-1]
            noise_scale = self.s_noise * sigma_up
            sched = self._schedule = list(
                zip(sh[:-1].tolist(), sh[1:].tolist(), sigma_down.tolist(), dt.tolist(), noise_scale.tolist()),
            )
        return sched

    def step(self, step_index: int, sample: Tensor):
        sigma = self.sigma_views[step_index]
        sigma_f, sigma_next_f, sigma_down_f, dt, noise_scale = self.schedule()[step_index]
        denoised = self.model(sample, self.sigma_in[step_index], **self.extra_args)
        add_noise = sigma_next_f > 0
        if self.stock_lane(sample):
            result = self.stock_step(
                step_index, sample, denoised, sigma_f, ops.STEP_EULER, dt, 0.0, noise_scale if add_noise else None,
            )
            if result is not None:
                return (result, sigma, sigma, denoised)
        noise_kw = {}
        guided = self.guidance is not None and self.guidance.factor != 0.0
        if add_noise and not guided:
            # x' = momentum_step(...) + noise * (s_noise * sigma_up): one launch
            self.prime_history(step_index, sample, denoised, sigma_f)
            sh = self.sigma_host_views
            noise_kw = self.ancestral_noise(sample, sh[step_index], sh[step_index + 1], noise_scale)
        result = self.momentum_step(step_index, sample, denoised, sigma_f, sigma_down_f, dt=dt, **noise_kw)
        if add_noise and guided:
            result = self.guidance_step(step_index, result, denoised)
            drawn = self.noise_sampler(self.sigma_host_views[step_index], self.sigma_host_views[step_index + 1])
            result = ops.axpby(result.contiguous(), 1.0, drawn.contiguous(), noise_scale)
        return (result, sigma, sigma, denoised)

    @classmethod
    def sampler(
        cls,
        model,
        x,
        sigmas,
        extra_args=None,
        callback=None,
        disable=None,
        sonar_config: SonarConfig | None = None,
        sonar_params: dict | None = None,
        eta=1.0,
        s_noise=1.0,
        noise_sampler: Callable | None = None,
    ):
        return cls._sample((eta, s_noise), model, x, sigmas, extra_args, callback, disable, noise_sampler, sonar_config, sonar_params)


class SonarDPMPPSDE(SonarSampler):
    """DPM-Solver++(SDE), r = 1/2, with momentum on both half steps (:626-770): two model
    evaluations and two fused launches per step, four history updates."""

    DEFAULT_NOISE_TYPE = noise.NoiseType.BROWNIAN
    NOISE_DRAWS_PER_STEP = 2

    def __init__(self, eta: float = 1.0, s_noise: float = 1.0, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.eta = eta
        self.s_noise = s_noise

    @staticmethod
    def sigma_fn(t: Tensor) -> Tensor:
        return t.neg().exp()

    @staticmethod
    def t_fn(sigma: Tensor) -> Tensor:
        return sigma.log().neg()

    def schedule(self) -> list:
        """Per-step coefficients of both half steps, host float32 tensor math in the reference's op
        order (:669-719), plus the mid-point sigmas on the device for the second model call."""
        sched = getattr(self, "_schedule", None)
        if sched is not None:
            return sched
        sh = self.sigmas_host
        sigma, sigma_next = sh[:-1], sh[1:]
        last = sigma_next == 0
        safe_next = torch.where(last, sigma, sigma_next)  # keep log() finite on the final step
        r = 1 / 2
        t, t_next = self.t_fn(sigma), self.t_fn(safe_next)
        h = t_next - t
        s = t + h * r
        s_t, s_s = self.sigma_fn(t), self.sigma_fn(s)
        sd, su = ancestral_steps(s_t, s_s, self.eta)
        s_ = self.t_fn(sd)
        s_t_next = self.sigma_fn(t_next)
        sd2, su2 = ancestral_steps(s_t, s_t_next, self.eta)
        t_down = self.t_fn(sd2)
        cols = {
            "sigma": sigma, "sigma_2": s_s,
            "c0_1": (t - s_).expm1(), "c1_1": self.sigma_fn(s_) / s_t, "ns_1": self.s_noise * su,
            "c0_2": (t - t_down).expm1(), "c1_2": self.sigma_fn(t_down) / s_t, "ns_2": self.s_noise * su2,
        }  # fmt: skip
        cols = {k: v.tolist() for k, v in cols.items()}
        down_last, _ = ancestral_steps(sigma, sigma_next, self.eta)
        dt_last = (down_last - sigma).tolist()
        down_last = down_last.tolist()
        sched = self._schedule = []
        for i in range(len(sigma)):
            if bool(last[i]):
                sched.append({"last": True, "sigma": cols["sigma"][i], "sigma_down": down_last[i], "dt": dt_last[i]})
            else:
                row = {k: v[i] for k, v in cols.items()}
                row |= {"last": False, "s_t": s_t[i], "s_s": s_s[i], "s_t_next": s_t_next[i]}
                sched.append(row)
        mid = s_s.to(device=self.s_in.device, dtype=self.s_in.dtype)
        self._sigma_mid_in = (mid.unsqueeze(1) * self.s_in.unsqueeze(0)).unbind(0)  # sigma_2 * s_in, all steps
        return sched

    def dpm_step(self, step_index: int, x: Tensor, denoised: Tensor, sc: dict) -> Tensor:
        guided = self.guidance is not None and self.guidance.factor != 0.0
        # ---- stage 1: x_2 = (sigma_fn(s_)/s_t) * x - momentum(expm1(t - s_) * denoised) + noise ----
        self.prime_history(step_index, x, denoised, sc["sigma"])
        noise_kw = self.ancestral_noise(x, sc["s_t"], sc["s_s"], sc["ns_1"])
        x_2 = self.fused_step(step_index, x, denoised, sc["sigma"], kind=ops.STEP_DPMPP, c0=sc["c0_1"], c1=sc["c1_1"], **noise_kw)
        denoised_2 = self.model(x_2, self._sigma_mid_in[step_index], **self.extra_args)
        # ---- stage 2 (fac = 1/(2r) = 1: denoised_d = 0*md1 + 1*md2) ----
        if guided:
            # the reference guides with denoised_d = get_momentum_denoised(denoised_2) (:720, :731): in
            # DENOISED mode that is the history-mixed prediction, which the fused launch keeps in registers
            denoised_d = self._momentum_denoised_preview(step_index, denoised_2, sc["sigma_2"])
            out = self.fused_step(step_index, x, denoised_2, sc["sigma_2"], kind=ops.STEP_DPMPP, c0=sc["c0_2"], c1=sc["c1_2"])
            out = self.guidance_step(step_index, out, denoised_d)
            drawn = self.noise_sampler(sc["s_t"], sc["s_t_next"])
            return ops.axpby(out.contiguous(), 1.0, drawn.contiguous(), sc["ns_2"])
        noise_kw = self.ancestral_noise(x, sc["s_t"], sc["s_t_next"], sc["ns_2"])
        return self.fused_step(step_index, x, denoised_2, sc["sigma_2"], kind=ops.STEP_DPMPP, c0=sc["c0_2"], c1=sc["c1_2"], **noise_kw)

    def step(self, step_index: int, sample: Tensor):
        sigma = self.sigma_views[step_index]
        sc = self.schedule()[step_index]
        denoised = self.model(sample, self.sigma_in[step_index], **self.extra_args)
        if sc["last"]:
            result = self.momentum_step(step_index, sample, denoised, sc["sigma"], sc["sigma_down"], dt=sc["dt"])
        else:
            result = self.dpm_step(step_index, sample, denoised, sc)
        return (result, sigma, sigma, denoised)

    @classmethod
    def sampler(
        cls,
        model,
        x: Tensor,
        sigmas: Tensor,
        extra_args: dict | None = None,
        callback=None,
        disable: bool | None = None,  # noqa: FBT001
        sonar_config: SonarConfig | None = None,
        sonar_params: dict | None = None,
        eta=1.0,
        s_noise=1.0,
        noise_sampler=None,
    ) -> Tensor:
        return cls._sample((eta, s_noise), model, x, sigmas, extra_args, callback, disable, noise_sampler, sonar_config, sonar_params)


EXTRA_SAMPLERS = {
    "sonar_euler": SonarEuler.sampler,
    "sonar_euler_ancestral": SonarEulerAncestral.sampler,
    "sonar_dpmpp_sde": SonarDPMPPSDE.sampler,
}


def add_samplers() -> None:
    """Registers the three samplers with ComfyUI (reference :823-847). Needs `comfy` importable."""
    from comfy.samplers import KSampler, k_diffusion_sampling

    added = 0
    for name, fn in EXTRA_SAMPLERS.items():
        if name in KSampler.SAMPLERS:
            continue
        try:
            KSampler.SAMPLERS.append(name)
            setattr(k_diffusion_sampling, f"sample_{name}", fn)
            added += 1
        except ValueError as exc:
            print(f"Sonar: Failed to add {name} to built in samplers list: {exc}")
    if added > 0:
        importlib.reload(k_diffusion_sampling)
