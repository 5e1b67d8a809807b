"""FreeU-Extreme on the spectral kernel.

Mirror of the reference's py/nodes/freeu_extreme.py: `ffilter` (:10-29), `FreeUExtremeConfig`
(:112-253, incl. `get_config_list`, `get_scale`, `check_match`, `apply`, `apply_filter`) and the two nodes
(`FreeUExtremeConfig` :32-109, `FreeUExtreme` :256-334). The arithmetic runs on the CUDA kernels:

* the filter is `sonar_spectral_filter_f32` with real input (rfft2 -> PowerFilter gain -> irfft2 per
  (H, W) plane in one launch, several small UNet planes per CTA);
* the "hidden mean" scale map and the write-back over the channel slice are `csrc/freeu.cu`
  (one reduction launch, one in-place streaming launch) instead of ~12 eager passes.

Differences from the reference, on purpose: `cpu_fft` is accepted and ignored (there is no CPU path), and
`ffilter` also works without a filter cache (upstream raises UnboundLocalError there, :12-15).
"""

from __future__ import annotations

import torch

from . import hostutil, ops
from .spectral_noise import PowerFilter


def _work_tensor(x: torch.Tensor) -> torch.Tensor:
    if not x.is_cuda:
        raise RuntimeError("sonar_b200 FreeU-Extreme needs CUDA activations (there is no CPU path)")
    return x if (x.dtype == torch.float32 and x.is_contiguous()) else x.to(torch.float32).contiguous()


def ffilter(x, pfilter, normalization_factor=1.0, cfg_idx=None, filter_cache=None):
    """irfft2(rfft2(x, ortho) * filter, ortho) per (H, W) plane (reference :10-29)."""
    cache_key = None
    filter_rfft = None
    if filter_cache is not None and cfg_idx is not None:
        cache_key = (cfg_idx, x.shape[-2:])
        filter_rfft = filter_cache.get(cache_key)
    if filter_rfft is None:
        filter_rfft = (
            PowerFilter.normalize(pfilter.build(x.shape), x.shape, normalization_factor=normalization_factor)
            .to(device=x.device, dtype=torch.float32, non_blocking=True)
            .contiguous()
        )
    if cache_key:
        filter_cache[cache_key] = filter_rfft
    height, width = x.shape[-2:]
    out = ops.spectral_filter(real=_work_tensor(x), mask=filter_rfft, hw=(height, width), out_scale=1.0 / (height * width))
    ops.drop_sums(out)
    return out.to(x.dtype, non_blocking=True)


class FreeUExtremeConfig:
    _keys = (
        "target", "stage_1", "stage_2", "stage_3", "start", "end", "slice", "slice_offset", "filter_norm", "scale",
        "blend", "blend_mode", "hidden_mean", "final", "sonar_power_filter", "frux_config",
    )  # fmt: skip

    def __init__(
        self,
        *,
        target,
        stage_1=False,
        stage_2=False,
        stage_3=False,
        start=0.0,
        end=1.0,
        slice=1.0,  # noqa: A002
        slice_offset=0.0,
        filter_norm=1.0,
        scale=1.0,
        blend=1.0,
        blend_mode=None,
        hidden_mean=True,
        final=True,
        sonar_power_filter_opt=None,
        frux_config_opt=None,
    ):
        self.target = target
        self.stage_1 = stage_1
        self.stage_2 = stage_2
        self.stage_3 = stage_3
        self.start = start
        self.end = end
        self.slice = slice
        self.slice_offset = slice_offset
        self.filter_norm = filter_norm
        self.scale = scale
        self.blend = blend
        self.blend_mode = blend_mode
        self.hidden_mean = hidden_mean
        self.final = final
        self.sonar_power_filter = sonar_power_filter_opt
        self.frux_config = frux_config_opt

    def get_config_list(self):
        result = [self]
        curr = self
        while cfg := curr.frux_config:
            curr = cfg
            if cfg.start >= 1 or cfg.end <= 0 or cfg.blend == 0 or not (cfg.stage_1 or cfg.stage_2 or cfg.stage_3):
                continue
            result.append(cfg)
        result.reverse()
        return result

    def get_scale(self, h: torch.Tensor):
        """Scalar scale, or the (B, 1, H, W) FreeU-V2 "hidden mean" scale map (reference :183-194)."""
        if not self.hidden_mean:
            return self.scale
        hmean, _rng = ops.freeu_hidden_mean(_work_tensor(h))
        hmean = ops.minmax_rescale(hmean, 0.0, 1.0, eps=0.0)
        return ops.affine(hmean, 0.0, self.scale - 1.0, 1.0, out=hmean)

    def check_match(self, pct, stage, is_skip=False):
        if pct < self.start or pct > self.end:
            return False
        if not getattr(self, f"stage_{stage}"):
            return False
        return not self.target not in {"skip" if is_skip else "backbone", "both"}

    def apply(self, idx, x, filter_cache, cpu_fft=False):
        """Filters / scales the channel slice of `x` in place and returns `x` (reference :203-227)."""
        _batch, features, _height, _width = x.shape
        slice_size = int(features * self.slice)
        slice_offs = int(features * self.slice_offset)
        lo, hi, _ = slice(slice_offs, slice_offs + slice_size).indices(features)
        count = max(hi - lo, 0)
        if self.blend != 1.0 and self.blend_mode not in ops.BLEND_IDS:
            raise KeyError(self.blend_mode)
        work = _work_tensor(x)
        hidden = ops.freeu_hidden_mean(work) if self.hidden_mean else None
        if count > 0:
            filtered = None
            if self.sonar_power_filter is not None:
                filtered = self.apply_filter(idx, work[:, lo:hi], filter_cache, cpu_fft=cpu_fft)
            ops.freeu_apply(
                work, filtered, hidden, slice_offset=lo, slice_channels=count, scale=self.scale, blend=self.blend,
                blend_mode=self.blend_mode,
            )  # fmt: skip
        if work is not x:
            x.copy_(work)
        return x

    def apply_filter(self, idx, xslice, filter_cache, cpu_fft=False):
        _ = cpu_fft  # no CPU FFT path: the spectral kernel is the filter
        filt = self.sonar_power_filter
        if filt is None:
            return xslice
        return ffilter(
            xslice.contiguous(), filt, normalization_factor=self.filter_norm, cfg_idx=idx, filter_cache=filter_cache,
        )

    def clone(self):
        kwargs = {k: getattr(self, k) for k in self._keys}
        kwargs["sonar_power_filter_opt"] = kwargs.pop("sonar_power_filter")
        kwargs["frux_config_opt"] = kwargs.pop("frux_config")
        return self.__class__(**kwargs)

    def __repr__(self):
        return f"<FRUXConfig: { {k: getattr(self, k) for k in self._keys} }>"


def _pct(default, tooltip):
    return ("FLOAT", {"step": 0.001, "min": 0.0, "max": 1.0, "round": False, "default": default, "tooltip": tooltip})


class FreeUExtremeConfigNode:
    DESCRIPTION = "Allows setting configuration for FreeU Extreme."
    RETURN_TYPES = ("FRUX_CONFIG",)
    FUNCTION = "go"
    CATEGORY = "model_patches"

    @classmethod
    def INPUT_TYPES(cls) -> dict:  # noqa: N802
        return {
            "required": {
                "stage_1": ("BOOLEAN", {"default": True, "tooltip": "Controls whether this configuration applies to stage 1."}),
                "stage_2": ("BOOLEAN", {"default": False, "tooltip": "Controls whether this configuration applies to stage 2."}),
                "stage_3": ("BOOLEAN", {"default": False, "tooltip": "Controls whether this configuration applies to stage 3."}),
                "target": (
                    ("backbone", "skip", "both"),
                    {"default": "backbone", "tooltip": "Controls whether this filter applies to backbone or skip layers (or both)."},
                ),
                "start": _pct(0.0, "Start time as percentage of sampling this configuration applies to. Inclusive."),
                "end": _pct(1.0, "End time as percentage of sampling this configuration applies to. Inclusive."),
                "slice": _pct(1.0, "Percentage of the layer the FreeU effect is applied to."),
                "slice_offset": _pct(0.0, "Offset as a percentage the layer is applied to."),
                "filter_norm": (
                    "FLOAT",
                    {"step": 0.001, "min": -10.0, "max": 10.0, "round": False, "default": 0.0,
                     "tooltip": "Normalization factor applied to the filter. 1.0 means 100% normalized."},
                ),  # fmt: skip
                "scale": (
                    "FLOAT",
                    {"step": 0.001, "min": -10000.0, "max": 10000.0, "round": False, "default": 1.0,
                     "tooltip": "Strength of the effects applied by this configuration."},
                ),  # fmt: skip
                "blend": (
                    "FLOAT",
                    {"step": 0.001, "min": -10000.0, "max": 10000.0, "round": False, "default": 1.0,
                     "tooltip": "Blends the filtered result based on the specified strength where 1.0 means 100% filtered."},
                ),  # fmt: skip
                "blend_mode": (tuple(hostutil.BLENDING_MODES), {"default": "lerp", "tooltip": "Mode used when blending."}),
                "hidden_mean": ("BOOLEAN", {"default": True, "tooltip": "You can think of this as FreeU V2 mode."}),
                "final": (
                    "BOOLEAN",
                    {"default": True, "tooltip": "When enabled, other configurations won't be considered if this one matched."},
                ),
            },
            "optional": {
                "sonar_power_filter_opt": (
                    "SONAR_POWER_FILTER",
                    {"tooltip": "Optionally attach a Power Filter here to set filtering parameters."},
                ),
                "frux_config_opt": ("FRUX_CONFIG", {"tooltip": "Optionally attach another configuration node here."}),
            },
        }

    @classmethod
    def go(cls, **kwargs: dict):
        return (FreeUExtremeConfig(**kwargs),)


class FreeUExtremeNode:
    DESCRIPTION = "Main FreeU Extreme node. Allows patching a model with the FreeU (V2) effect with more control."
    RETURN_TYPES = ("MODEL",)
    FUNCTION = "go"
    CATEGORY = "model_patches"

    @classmethod
    def INPUT_TYPES(cls) -> dict:  # noqa: N802
        return {
            "required": {
                "model": ("MODEL", {"tooltip": "Model to patch."}),
                "cpu_fft": ("BOOLEAN", {"default": False, "tooltip": "Ignored: the filter always runs on the CUDA spectral kernel."}),
            },
            "optional": {
                "input_config": ("FRUX_CONFIG", {"tooltip": "Allows specifying configuration for input blocks."}),
                "middle_config": ("FRUX_CONFIG", {"tooltip": "Allows specifying configuration for middle blocks."}),
                "output_config": ("FRUX_CONFIG", {"tooltip": "Allows specifying configuration for output blocks."}),
            },
        }

    @classmethod
    def go(cls, model, cpu_fft, input_config=None, middle_config=None, output_config=None):
        model_channels = model.model.model_config.unet_config["model_channels"]
        stages = {model_channels * 4: 1, model_channels * 2: 2, model_channels: 3}
        icfg, mcfg, ocfg = (() if cfg is None else cfg.get_config_list() for cfg in (input_config, middle_config, output_config))
        m = model.clone()
        ms = m.get_model_object("model_sampling")
        filter_cache = {}

        def handler(_typ, h_shape, cfg, x, toptions, is_skip=False):
            stage = stages.get(h_shape[1])
            if stage is None:
                return x
            sigma = toptions["sigmas"].max().detach().cpu()
            pct = 1.0 - (ms.timestep(sigma) / 999.0)
            for idx, ci in enumerate(cfg):
                if not ci.check_match(pct, stage, is_skip):
                    continue
                x = ci.apply(idx, x, filter_cache, cpu_fft=cpu_fft)
                if ci.final:
                    break
            return x

        def in_patch(h, toptions):
            return handler("input", h.shape, icfg, h, toptions)

        def mid_patch(h, toptions):
            return handler("middle", h.shape, mcfg, h, toptions)

        def out_patch(h, hsp, toptions):
            h = handler("output", h.shape, ocfg, h, toptions)
            hsp = handler("output", h.shape, ocfg, hsp, toptions, is_skip=True)
            return h, hsp

        if icfg:
            m.set_model_input_block_patch(in_patch)
        if mcfg:
            m.set_model_patch(mid_patch, "middle_block_patch")
        if ocfg:
            m.set_model_output_block_patch(out_patch)
        return (m,)


NODE_CLASS_MAPPINGS = {
    "FreeUExtremeConfig": FreeUExtremeConfigNode,
    "FreeUExtreme": FreeUExtremeNode,
}
