"""ComfyUI node surface of the B200 path.

Keeps the reference's node names, input schemas (names, kinds, defaults, ranges), RETURN_TYPES,
FUNCTION names and argument marshalling for every node that fronts the hot path (reference
py/nodes/*.py), so workflows built against blepping/ComfyUI-sonar load
unchanged. Node bodies only marshal arguments into the host objects of this package; nodes of
the reference that front out-of-scope components are not registered.

tests/test_nodes.py checks these schemas against tests/golden/node_schemas.json, which was dumped
from the reference's own INPUT_TYPES().
"""

from __future__ import annotations

import functools
import inspect
import math
import random
from typing import Any, Callable

import numpy as np
import torch
import yaml
from comfy import model_management, samplers as comfy_samplers

from . import hostutil, noise_graph as noise, spectral_noise
from .generators import NoiseType
from .samplers import (
    GuidanceConfig,
    GuidanceType,
    HistoryType,
    SonarConfig,
    SonarDPMPPSDE,
    SonarEuler,
    SonarEulerAncestral,
)
from .wcfg import WaveletCFG, WCFGRules


# ---------------------------------------------------------------------------------------------
# schema helpers
# ---------------------------------------------------------------------------------------------
class Wildcard(str):
    """ComfyUI "any type" marker: never unequal to another type string."""

    __slots__ = ()

    def __ne__(self, _other):
        return False


WILDCARD_NOISE = Wildcard("*")
TRISTATE = ("default", "forced", "disabled")
YAML_OPTS = {"placeholder": "# YAML or JSON here", "dynamicPrompts": False, "multiline": True}


def f_float(default, *, min=-10000.0, max=10000.0, step=0.001, tooltip=None):  # noqa: A002
    opts = {"step": step, "min": min, "max": max, "round": False, "default": default}
    return ("FLOAT", opts | ({"tooltip": tooltip} if tooltip else {}))


def f_int(default, *, min=-10000, max=10000, tooltip=None):  # noqa: A002
    opts = {"min": min, "max": max, "default": default}
    return ("INT", opts | ({"tooltip": tooltip} if tooltip else {}))


def f_bool(default=False, tooltip=None):
    return ("BOOLEAN", {"default": default} | ({"tooltip": tooltip} if tooltip else {}))


def f_choice(options, default=None, tooltip=None):
    opts = ({} if default is None else {"default": default}) | ({"tooltip": tooltip} if tooltip else {})
    return (tuple(options), opts) if opts else (tuple(options),)


def f_noise(tooltip=None):
    return (WILDCARD_NOISE, {"tooltip": tooltip}) if tooltip else (WILDCARD_NOISE,)


def noise_names(*, skip=None, first=None) -> tuple:
    names = tuple(NoiseType.get_names(skip=skip))
    return names if first is None else (first, *names)


def tristate(val: str) -> bool | None:
    """default / forced / disabled -> None / True / False (reference nodes/base.py:287-290)."""
    return None if val == "default" else val == "forced"


FACTOR = f_float(1.0, tooltip="Scaling factor for the generated noise of this type.")
RESCALE = f_float(
    0.0,
    min=0.0,
    tooltip="When non-zero, the factors of this item and the items chained to it are scaled to add up to this value.",
)
CHAIN_OPT = {"sonar_custom_noise_opt": f_noise("Optional input for more custom noise items.")}


# ---------------------------------------------------------------------------------------------
# custom noise chain nodes
# ---------------------------------------------------------------------------------------------
class SonarCustomNoiseNodeBase:
    DESCRIPTION = "A custom noise item."
    RETURN_TYPES = ("SONAR_CUSTOM_NOISE",)
    OUTPUT_TOOLTIPS = ("A custom noise chain.",)
    CATEGORY = "advanced/noise"
    FUNCTION = "go"

    REQUIRED: dict = {}
    OPTIONAL: dict = {}
    WITH_RESCALE = True
    WITH_CHAIN = True

    @classmethod
    def INPUT_TYPES(cls) -> dict:  # noqa: N802
        required = {"factor": FACTOR} | ({"rescale": RESCALE} if cls.WITH_RESCALE else {}) | cls.REQUIRED
        optional = (CHAIN_OPT if cls.WITH_CHAIN else {}) | cls.OPTIONAL
        return {"required": dict(required), "optional": dict(optional)}

    @classmethod
    def get_item_class(cls):
        raise NotImplementedError

    def go(self, factor=1.0, rescale=0.0, sonar_custom_noise_opt=None, **kwargs: Any):
        chain = sonar_custom_noise_opt.clone() if sonar_custom_noise_opt else noise.CustomNoiseChain()
        if factor != 0:
            chain.add(self.get_item_class()(factor, **kwargs))
        return (chain if rescale == 0 else chain.rescaled(rescale),)

    @staticmethod
    def get_normalize(val: str) -> bool | None:
        return tristate(val)


class SonarCustomNoiseNode(SonarCustomNoiseNodeBase):
    REQUIRED = {"noise_type": f_choice(noise_names(), "gaussian", "Sets the type of noise to generate.")}

    @classmethod
    def get_item_class(cls):
        return noise.CustomNoiseItem


class SonarCustomNoiseAdvNode(SonarCustomNoiseNode):
    DESCRIPTION = "A custom noise item allowing advanced YAML parameter input."
    OPTIONAL = {"yaml_parameters": ("STRING", dict(YAML_OPTS))}


class SonarAdvancedPyramidNoiseNode(SonarCustomNoiseNodeBase):
    DESCRIPTION = "Custom noise type that allows specifying parameters for Pyramid variants."
    REQUIRED = {
        "variant": f_choice(("highres_pyramid", "pyramid", "pyramid_old"), "highres_pyramid"),
        "iterations": f_int(-1, min=-1, max=8, tooltip="-1 uses the variant's default."),
        "discount": f_float(0.0, tooltip="0 uses the variant's default."),
        "upscale_mode": f_choice(("default", *hostutil.UPSCALE_METHODS), "default"),
    }

    @classmethod
    def get_item_class(cls):
        return noise.AdvancedPyramidNoise

    def go(self, *, factor, rescale, variant, iterations, discount, upscale_mode, sonar_custom_noise_opt=None):
        return super().go(
            factor,
            rescale=rescale,
            sonar_custom_noise_opt=sonar_custom_noise_opt,
            variant=variant,
            iterations=None if iterations == -1 else iterations,
            discount=None if discount == 0 else discount,
            upscale_mode=None if upscale_mode == "default" else upscale_mode,
        )


class SonarAdvanced1fNoiseNode(SonarCustomNoiseNodeBase):
    DESCRIPTION = "Custom noise type that allows specifying parameters for 1f (pink, green, etc) variants."
    REQUIRED = {
        "alpha": f_float(0.25),
        "k": f_float(1.0),
        "vertical_factor": f_float(1.0),
        "horizontal_factor": f_float(1.0),
        "use_sqrt": f_bool(True),
    }

    @classmethod
    def get_item_class(cls):
        return noise.Advanced1fNoise

    def go(self, *, factor, rescale, alpha, k, vertical_factor, horizontal_factor, use_sqrt, sonar_custom_noise_opt=None):
        return super().go(
            factor,
            rescale=rescale,
            sonar_custom_noise_opt=sonar_custom_noise_opt,
            alpha=alpha,
            k=k,
            hfac=vertical_factor,
            wfac=horizontal_factor,
            use_sqrt=use_sqrt,
        )


class SonarAdvancedPowerLawNoiseNode(SonarCustomNoiseNodeBase):
    DESCRIPTION = "Custom noise type that allows specifying parameters for power law (grey, violet, etc) variants."
    MAX_DIMS_MAP = {  # noqa: RUF012
        "none": None,
        "non-batch": (-3, -2, -1),
        "spatial": (-2, -1),
        "all": (),
        "batch": 0,
        "channel": 1,
        "height": 2,
        "width": 3,
    }
    REQUIRED = {
        "alpha": f_float(0.5),
        "div_max_dims": f_choice(tuple(MAX_DIMS_MAP), "non-batch"),
        "use_div_max_abs": f_bool(True),
        "use_sign": f_bool(False),
    }

    @classmethod
    def get_item_class(cls):
        return noise.AdvancedPowerLawNoise

    def go(self, *, factor, rescale, alpha, div_max_dims, use_sign, use_div_max_abs, sonar_custom_noise_opt=None):
        return super().go(
            factor,
            rescale=rescale,
            sonar_custom_noise_opt=sonar_custom_noise_opt,
            alpha=alpha,
            div_max_dims=self.MAX_DIMS_MAP.get(div_max_dims),
            use_sign=use_sign,
            use_div_max_abs=use_div_max_abs,
        )


class SonarRepeatedNoiseNode(SonarCustomNoiseNodeBase):
    DESCRIPTION = "Custom noise type that caches noise samples and replays them (optionally permuted)."
    WITH_RESCALE = WITH_CHAIN = False
    REQUIRED = {
        "sonar_custom_noise": f_noise("Custom noise input for items to repeat."),
        "repeat_length": f_int(8, min=1, max=100),
        "max_recycle": f_int(1000, min=1, max=1000),
        "normalize": f_choice(TRISTATE, "default"),
        "permute": f_choice(("enabled", "disabled", "always"), "enabled"),
    }

    @classmethod
    def get_item_class(cls):
        return noise.RepeatedNoise

    def go(self, *, factor, sonar_custom_noise, repeat_length, max_recycle, normalize, permute=True):
        return super().go(
            factor,
            noise=sonar_custom_noise,
            repeat_length=repeat_length,
            max_recycle=max_recycle,
            normalize=tristate(normalize),
            permute=permute,
        )


class SonarScheduledNoiseNode(SonarCustomNoiseNodeBase):
    DESCRIPTION = "Custom noise type that allows scheduling custom noise types by sampling percentage."
    WITH_RESCALE = WITH_CHAIN = False
    REQUIRED = {
        "model": ("MODEL",),
        "sonar_custom_noise": f_noise("Custom noise to use inside the range."),
        "start_percent": f_float(0.0, min=0.0, max=1.0),
        "end_percent": f_float(1.0, min=0.0, max=1.0),
        "normalize": f_choice(TRISTATE, "default"),
    }
    OPTIONAL = {"fallback_sonar_custom_noise": f_noise("Custom noise to use outside the range (default: zeros).")}

    @classmethod
    def get_item_class(cls):
        return noise.ScheduledNoise

    def go(self, *, model, factor, sonar_custom_noise, start_percent, end_percent, normalize, fallback_sonar_custom_noise=None):
        ms = model.get_model_object("model_sampling")
        return super().go(
            factor,
            noise=sonar_custom_noise,
            start_sigma=ms.percent_to_sigma(start_percent),
            end_sigma=ms.percent_to_sigma(end_percent),
            normalize=tristate(normalize),
            fallback_noise=fallback_sonar_custom_noise,
        )


class SonarCompositeNoiseNode(SonarCustomNoiseNodeBase):
    DESCRIPTION = "Custom noise type that composites two other custom noise generators based on a mask."
    WITH_RESCALE = WITH_CHAIN = False
    REQUIRED = {
        "sonar_custom_noise_dst": f_noise("Noise where the mask is not set."),
        "sonar_custom_noise_src": f_noise("Noise where the mask is set."),
        "normalize_dst": f_choice(TRISTATE, "default"),
        "normalize_src": f_choice(TRISTATE, "default"),
        "normalize_result": f_choice(TRISTATE, "default"),
        "mask": ("MASK",),
    }

    @classmethod
    def get_item_class(cls):
        return noise.CompositeNoise

    def go(self, *, factor, sonar_custom_noise_dst, sonar_custom_noise_src, normalize_src, normalize_dst, normalize_result, mask):
        # The reference node hands normalize_src to the dst slot and vice versa
        # (nodes/noise_filters.py:246-247); preserved so existing workflows behave identically.
        return super().go(
            factor,
            dst_noise=sonar_custom_noise_dst,
            src_noise=sonar_custom_noise_src,
            normalize_dst=tristate(normalize_src),
            normalize_src=tristate(normalize_dst),
            normalize_result=tristate(normalize_result),
            mask=mask,
        )


class SonarGuidedNoiseNode(SonarCustomNoiseNodeBase):
    DESCRIPTION = "Custom noise type that mixes a references with another custom noise generator to guide the generation."
    WITH_RESCALE = WITH_CHAIN = False
    REQUIRED = {
        "latent": ("LATENT", {"tooltip": "Latent to use for guidance."}),
        "method": f_choice(("euler", "linear"), "euler"),
        "guidance_factor": f_float(0.0125, min=-100.0, max=100.0),
        "normalize_noise": f_choice(TRISTATE, "default"),
        "normalize_result": f_choice(TRISTATE, "default"),
        "normalize_ref": f_bool(True),
    }
    OPTIONAL = {"sonar_custom_noise": f_noise("Optional custom noise input to combine with the guidance.")}

    @classmethod
    def get_item_class(cls):
        return noise.GuidedNoise

    def go(self, *, factor, latent, normalize_noise, normalize_result, normalize_ref=True, method="euler",
           guidance_factor=0.5, sonar_custom_noise=None):  # fmt: skip
        from .samplers import SonarGuidanceMixin

        # the reference standardises the latent where it lives (CPU); this package has no CPU path, so the one-off
        # setup runs on the compute device (nodes/noise_filters.py:302-308)
        samples = latent["samples"].to(device=model_management.get_torch_device(), dtype=torch.float32, copy=True)
        ref_latent = hostutil.scale_noise(SonarGuidanceMixin.prepare_ref_latent(samples), normalized=normalize_ref)
        return super().go(
            factor,
            ref_latent=ref_latent,
            guidance_factor=guidance_factor,
            noise=sonar_custom_noise.clone() if sonar_custom_noise is not None else None,
            method=method,
            normalize_noise=tristate(normalize_noise),
            normalize_result=tristate(normalize_result),
        )


class SonarBlendedNoiseNode(SonarCustomNoiseNodeBase):
    DESCRIPTION = "Custom noise type that allows blending two other noise items."
    REQUIRED = {
        "noise_2_percent": f_float(0.5),
        "blend_mode": f_choice(tuple(hostutil.BLENDING_MODES), "lerp"),
        "normalize": f_choice(TRISTATE, "default"),
    }
    OPTIONAL = {"custom_noise_1": f_noise(), "custom_noise_2": f_noise(), "custom_noise_mask": f_noise()}

    @classmethod
    def get_item_class(cls):
        return noise.BlendedNoise

    def go(
        self,
        *,
        factor,
        rescale,
        sonar_custom_noise_opt=None,
        normalize,
        noise_2_percent,
        custom_noise_1=None,
        custom_noise_2=None,
        custom_noise_mask=None,
        blend_mode="lerp",
    ):
        blend_function = hostutil.BLENDING_MODES.get(blend_mode)
        if blend_function is None:
            raise ValueError("Unknown blend mode")
        return super().go(
            factor,
            rescale=rescale,
            sonar_custom_noise_opt=sonar_custom_noise_opt,
            blend_function=blend_function,
            normalize=tristate(normalize),
            custom_noise_1=custom_noise_1,
            custom_noise_2=custom_noise_2,
            custom_noise_mask=custom_noise_mask,
            noise_2_percent=noise_2_percent,
        )


_PARAM_DTYPES = (
    "default", "float64", "float32", "float16", "bfloat16", "float8_e4m3fn", "float8_e4m3fnuz", "float8_e5m2",
    "float8_e5m2fnuz", "float8_e8m0fnu", "int64", "int32", "int16", "int8",
)  # fmt: skip


WAVELET_FILTER_YAML = """# YAML or JSON. Every key is optional; the values below are the defaults.
wave: haar              # haar or db1..db12
level: 3                # decomposition levels
mode: periodization     # padding: periodization, symmetric, zero, reflect, periodic
use_dtcwt: false        # DTCWT is not built in sonar_b200 (2-D DWT only)
biort: near_sym_a       # DTCWT only
qshift: qshift_a        # DTCWT only
two_step_inverse: false # invert the low and high parts separately and add them (same result: the transform is linear)
# inv_wave / inv_mode / inv_biort / inv_qshift override the settings of the inverse transform
yl_scale: 1.0           # scale of the approximation (low-frequency) band
yh_scales: 1.0          # scale(s) of the detail bands: a number, a list per level, or a list of [h, v, d] lists
"""


class SonarWaveletFilteredNoiseNode(SonarCustomNoiseNodeBase):
    DESCRIPTION = "Custom noise type that allows filtering another custom noise source with wavelets."
    REQUIRED = {  # noqa: RUF012
        "normalize_noise": f_bool(False, "Controls whether the noise source is normalized before wavelet filtering occurs."),
        "normalize": f_choice(TRISTATE, "default", "Controls whether the generated noise is normalized to 1.0 strength."),
    }
    OPTIONAL = {  # noqa: RUF012
        "custom_noise": f_noise("Optional: Custom noise input. If unconnected will default to Gaussian noise."),
        "custom_noise_high": f_noise(
            "Optional: noise for the high-frequency side of the wavelet. If unconnected the same generator as custom_noise is used.",
        ),
        "yaml_parameters": ("STRING", YAML_OPTS | {"placeholder": WAVELET_FILTER_YAML}),
    }

    @classmethod
    def get_item_class(cls):
        return noise.WaveletFilteredNoise

    def go(self, *, factor, rescale, normalize, normalize_noise, custom_noise=None, custom_noise_high=None,
           yaml_parameters=None, sonar_custom_noise_opt=None):  # fmt: skip
        return super().go(
            factor,
            rescale=rescale,
            sonar_custom_noise_opt=sonar_custom_noise_opt,
            normalize=self.get_normalize(normalize),
            normalize_noise=normalize_noise,
            noise=custom_noise,
            noise_high=custom_noise_high if custom_noise_high is not None else custom_noise,
            yaml_parameters=yaml_parameters,
        )


class SonarCustomNoiseParametersNode(SonarCustomNoiseNodeBase):
    DESCRIPTION = "Allows overriding shape / dtype / RNG parameters for the attached custom noise (e.g. video latents)."
    WITH_RESCALE = WITH_CHAIN = False
    REQUIRED = {
        "custom_noise": f_noise(),
        "rng_state_offset": f_int(0, min=0),
        "rng_offset_mode": f_choice(("disabled", "override", "add"), "disabled"),
        "rng_mode": f_choice(("default", "separate", "fork"), "default"),
        "frames_to_channels": f_bool(False),
        "ensure_square_aspect_ratio": f_bool(False),
        "fix_invalid": f_bool(False),
        "override_dtype": f_choice(_PARAM_DTYPES, "default"),
        "override_device": f_choice(("default", "cpu", "gpu"), "default"),
        "normalize": f_choice(TRISTATE, "default"),
    }

    @classmethod
    def get_item_class(cls):
        return noise.CustomNoiseParametersNoise

    def go(
        self,
        *,
        factor,
        rng_state_offset: int,
        rng_offset_mode: str,
        rng_mode: str,
        frames_to_channels: bool,
        ensure_square_aspect_ratio: bool,
        fix_invalid: bool,
        override_dtype: str,
        override_device: str,
        normalize: str,
        custom_noise: object,
    ):
        dt = getattr(torch, override_dtype, None)
        if override_dtype not in _PARAM_DTYPES or (override_dtype != "default" and dt is None):
            raise ValueError("Bad dtype, may not be supported by your PyTorch version")
        device = {"default": None, "cpu": "cpu", "gpu": model_management.get_torch_device()}.get(override_device)
        return super().go(
            factor,
            rng_state_offset=rng_state_offset,
            rng_offset_mode=rng_offset_mode,
            rng_mode=rng_mode,
            frames_to_channels=frames_to_channels,
            ensure_square_aspect_ratio=ensure_square_aspect_ratio,
            fix_invalid=fix_invalid,
            override_dtype=dt,
            override_device=device,
            normalize=normalize,
            noise=custom_noise,
        )


# ---------------------------------------------------------------------------------------------
# power noise nodes
# ---------------------------------------------------------------------------------------------
_FILTER_FIELDS = {
    "alpha": f_float(0.0, min=-5.0, max=5.0, tooltip="Above 0 amplifies low frequencies, below 0 high frequencies."),
    "max_freq": f_float(0.7071, min=0.0, max=0.7071, tooltip="Maximum frequency to pass through the filter."),
    "min_freq": f_float(0.0, min=0.0, max=0.7071, tooltip="Minimum frequency to pass through the filter."),
    "stretch": f_float(1.0, min=0.01, max=100.0, tooltip="Stretches the filter's shape by the specified factor."),
    "rotate": f_float(0.0, min=-90.0, max=90.0, step=5.0, tooltip="Rotates the filter."),
    "pnorm": f_float(2.0, min=0.125, max=100.0, step=0.1, tooltip="Factor used for cushioning the band-pass region."),
}


class SonarPowerNoiseNode(SonarCustomNoiseNodeBase):
    DESCRIPTION = "Custom noise type that applies a filter to generated noise."
    REQUIRED = {
        "time_brownian": f_bool(False, "Controls whether brownian noise is used when mix isn't 1.0."),
        **_FILTER_FIELDS,
        "mix": f_float(1.0, min=0.0, max=1.0, tooltip="Ratio of filtered noise; 0.75 means 25% raw noise."),
        "common_mode": f_float(0.0, min=-100.0, max=100.0, tooltip="Injects the average across channels."),
        "channel_correlation": ("STRING", {"default": "1, 1, 1, 1, 1, 1"}),
        "preview": f_choice(("none", "no_mix", "mix"), "none"),
    }

    @classmethod
    def get_item_class(cls):
        return spectral_noise.PowerNoiseItem

    def go(self, preview="none", **kwargs: Any):
        if preview != "none":
            raise NotImplementedError("sonar_b200: filter previews (PIL image output) are UI glue outside the hot path")
        return super().go(**kwargs)


class SonarPowerFilterNoiseNode(SonarPowerNoiseNode):
    DESCRIPTION = "Custom noise type that applies a Power Filter to another custom noise generator."

    @classmethod
    def INPUT_TYPES(cls) -> dict:  # noqa: N802
        result = super().INPUT_TYPES()
        for key in (*_FILTER_FIELDS, "time_brownian"):
            del result["required"][key]
        result["required"] |= {
            "sonar_custom_noise": f_noise("Custom noise type to filter."),
            "sonar_power_filter": ("SONAR_POWER_FILTER", {"tooltip": "Filter to use."}),
            "filter_norm_factor": ("FLOAT", {"default": 1.0, "min": 0.0, "max": 1.0, "step": 0.1, "round": False}),
            "normalize_result": (TRISTATE,),
            "normalize_noise": (TRISTATE,),
        }
        result["required"]["preview"] = (("none", "no_mix", "mix", "custom"),)
        return result

    @classmethod
    def get_item_class(cls):
        return spectral_noise.PowerFilterNoiseItem

    def go(self, factor, sonar_custom_noise, sonar_power_filter, filter_norm_factor, normalize_noise, normalize_result, preview="none", **kwargs: Any):
        return super().go(
            factor=factor,
            noise=sonar_custom_noise,
            normalize_noise=tristate(normalize_noise),
            normalize_result=tristate(normalize_result),
            preview=preview,
            time_brownian=True,
            power_filter=sonar_power_filter,
            filter_norm_factor=filter_norm_factor,
            **kwargs,
        )


class SonarPowerFilterNode:
    RETURN_TYPES = ("SONAR_POWER_FILTER",)
    CATEGORY = "advanced/noise"
    FUNCTION = "go"

    @classmethod
    def INPUT_TYPES(cls) -> dict:  # noqa: N802
        return {
            "required": dict(_FILTER_FIELDS)
            | {
                "oversample": ("INT", {"default": 4, "min": 1, "max": 128}),
                "blur": ("FLOAT", {"default": 0.125, "min": -10.0, "max": 10.0, "step": 0.01, "round": False}),
                "scale": ("FLOAT", {"default": 1, "min": -100.0, "max": 100.0, "step": 0.1, "round": False}),
                "compose_mode": (("max", "min", "add", "sub", "mul"),),
            },
            "optional": {"power_filter_opt": ("SONAR_POWER_FILTER",)},
        }

    @classmethod
    def go(
        cls,
        min_freq=0.0,
        max_freq=0.7071,
        stretch=1.0,
        rotate=0.0,
        pnorm=2.0,
        alpha=0.0,
        blur=0.125,
        oversample=4,
        scale=1.0,
        compose_mode="max",
        power_filter_opt=None,
    ):
        return (
            spectral_noise.PowerFilter(
                min_freq=min_freq,
                max_freq=max_freq,
                stretch=stretch,
                rotate=rotate,
                pnorm=pnorm,
                alpha=alpha,
                scale=scale,
                rel_bw=blur,
                oversample=oversample,
                compose_mode=compose_mode,
                compose_with=power_filter_opt,
            ),
        )


# ---------------------------------------------------------------------------------------------
# NOISE object, NoisyLatentLike
# ---------------------------------------------------------------------------------------------
def compute_device() -> torch.device:
    """Device the kernels run on when ComfyUI hands us a CPU latent (its NOISE contract is CPU)."""
    dev = model_management.get_torch_device()
    if torch.device(dev).type != "cuda":
        raise RuntimeError("sonar_b200 needs a CUDA device; there is no CPU generation path")
    return torch.device(dev)


class CustomNOISE:
    """ComfyUI NOISE object (reference nodes/misc.py:360-419): `.seed` + `.generate_noise(latent)`
    returning CPU noise in the latent's dtype, honouring `batch_index`. The latent is moved to the
    GPU, noise is generated there, and the result is copied back (the contract forces the D2H)."""

    def __init__(self, custom_noise, seed, *, cpu_noise=True, normalize=True, multiplier=1.0):
        self.custom_noise = custom_noise
        self.seed = seed
        self.cpu_noise = cpu_noise
        self.normalize = normalize
        self.multiplier = multiplier

    def _sample_noise(self, latent_image: torch.Tensor, seed: int) -> torch.Tensor:
        work = latent_image if latent_image.is_cuda else latent_image.to(compute_device(), non_blocking=True)
        ns = self.custom_noise.make_noise_sampler(work, None, None, seed=seed, cpu=self.cpu_noise, normalized=self.normalize)
        result = ns(None, None)
        if self.multiplier != 1.0:
            result = result.mul_(self.multiplier)
        result = result.to(device="cpu", dtype=latent_image.dtype)
        if latent_image.layout != torch.strided:
            if latent_image.layout == torch.sparse_coo:
                return result.to_sparse()
            raise NotImplementedError(f"Cannot handle latent layout {type(latent_image.layout).__name__}")
        return result

    def generate_noise(self, input_latent: dict) -> torch.Tensor:
        latent_image = input_latent["samples"]
        batch_inds = input_latent.get("batch_index")
        torch.manual_seed(self.seed)
        random.seed(self.seed)
        if self.multiplier == 0.0:
            return torch.zeros(latent_image.shape, dtype=latent_image.dtype, layout=latent_image.layout, device="cpu")
        if batch_inds is None:
            return self._sample_noise(latent_image, self.seed)
        unique_inds, inverse_inds = np.unique(batch_inds, return_inverse=True)
        batch_size = latent_image.shape[0]
        kept = []
        for idx in range(unique_inds[-1] + 1):
            drawn = self._sample_noise(latent_image[idx % batch_size].unsqueeze(0), self.seed + idx)
            if idx in unique_inds:
                kept.append(drawn)
        return torch.cat(tuple(kept[i] for i in inverse_inds), axis=0)


class SonarToComfyNOISENode:
    DESCRIPTION = "Converts SONAR_CUSTOM_NOISE to NOISE (used by SamplerCustomAdvanced and other custom samplers)."
    RETURN_TYPES = ("NOISE",)
    CATEGORY = "sampling/custom_sampling/noise"
    FUNCTION = "go"

    @classmethod
    def INPUT_TYPES(cls) -> dict:  # noqa: N802
        return {
            "required": {
                "custom_noise": f_noise("Custom noise type to convert."),
                "seed": f_int(0, min=0, max=0xFFFFFFFFFFFFFFFF, tooltip="Seed to use for generated noise"),
                "cpu_noise": f_bool(True),
                "normalize": f_bool(True),
                "multiplier": f_float(1.0),
            },
            "optional": {},
        }

    @classmethod
    def go(cls, *, custom_noise, seed, cpu_noise=True, normalize=True, multiplier=1.0):
        return (CustomNOISE(custom_noise, seed, cpu_noise=cpu_noise, normalize=normalize, multiplier=multiplier),)


class NoisyLatentLikeNode:
    DESCRIPTION = "Generates noise (and optionally adds it) based on a reference latent."
    RETURN_TYPES = ("LATENT",)
    OUTPUT_TOOLTIPS = ("The noisy latent image.",)
    CATEGORY = "latent/noise"
    FUNCTION = "go"

    @classmethod
    def INPUT_TYPES(cls) -> dict:  # noqa: N802
        return {
            "required": {
                "noise_type": f_choice(noise_names(), "gaussian"),
                "seed": f_int(0, min=0, max=0xFFFFFFFFFFFFFFFF, tooltip="Seed to use for generated noise"),
                "latent": ("LATENT",),
                "multiplier": f_float(1.0),
                "add_to_latent": f_bool(False),
                "repeat_batch": f_int(1, min=1),
                "cpu_noise": f_bool(True),
                "normalize": f_bool(True),
            },
            "optional": {"custom_noise_opt": f_noise(), "mul_by_sigmas_opt": ("SIGMAS",), "model_opt": ("MODEL",)},
        }

    @classmethod
    def go(
        cls,
        *,
        noise_type: str,
        seed: int | None,
        latent: dict,
        multiplier: float = 1.0,
        add_to_latent=False,
        repeat_batch=1,
        cpu_noise=True,
        normalize=True,
        custom_noise_opt: object | None = None,
        mul_by_sigmas_opt: torch.Tensor | None = None,
        model_opt: object | None = None,
    ):
        model, sigmas = model_opt, mul_by_sigmas_opt
        if sigmas is not None and len(sigmas) > 0:
            if model is None:
                raise ValueError("NoisyLatentLike requires a model when sigmas are connected!")
            while hasattr(model, "model"):
                model = model.model
            model_sigma_max = float(model.model_sampling.sigma_max)
            first_sigma = float(sigmas[0])
            max_denoise = math.isclose(model_sigma_max, first_sigma, rel_tol=1e-05) or first_sigma > model_sigma_max
            multiplier *= float(torch.sqrt(1.0 + sigmas[0] ** 2.0) if max_denoise else sigmas[0]) / model.latent_format.scale_factor
        if sigmas is not None and sigmas.numel() > 1:
            host = sigmas.detach().float().cpu()
            sigma_min, sigma_max, sigma, sigma_next = host[host > 0].min(), host.max(), host[0], host[1]
        else:
            sigma_min = sigma_max = sigma = sigma_next = None
        samples = latent["samples"]
        orig_device = samples.device
        work = samples if samples.is_cuda else samples.detach().clone().to(compute_device())
        if custom_noise_opt is not None:
            ns = custom_noise_opt.make_noise_sampler(
                work, sigma_min=sigma_min, sigma_max=sigma_max, seed=seed, cpu=cpu_noise, normalized=normalize,
            )
        else:
            ns = noise.get_noise_sampler(
                NoiseType[noise_type.upper()], work, sigma_min, sigma_max, seed=seed, cpu=cpu_noise, normalized=normalize,
            )
        host_state = torch.random.get_rng_state()
        dev_state = torch.cuda.get_rng_state(work.device)
        try:
            torch.random.manual_seed(seed)
            result = torch.cat(tuple(ns(sigma, sigma_next) for _ in range(repeat_batch)), dim=0)
        finally:
            torch.random.set_rng_state(host_state)
            torch.cuda.set_rng_state(dev_state, work.device)
        result = hostutil.scale_noise(result.contiguous(), multiplier, normalized=True)
        if add_to_latent:
            result += work.repeat(*(repeat_batch if i == 0 else 1 for i in range(work.ndim))).to(result)
        return ({"samples": result.to(orig_device)},)


# ---------------------------------------------------------------------------------------------
# sampler nodes
# ---------------------------------------------------------------------------------------------
class GuidanceConfigNode:
    DESCRIPTION = "Allows specifying extended guidance parameters for Sonar samplers."
    RETURN_TYPES = ("SONAR_GUIDANCE_CFG",)
    CATEGORY = "sampling/custom_sampling/samplers"
    FUNCTION = "make_guidance_cfg"

    @classmethod
    def INPUT_TYPES(cls) -> dict:  # noqa: N802
        return {
            "required": {
                "factor": f_float(0.01, min=-2.0, max=2.0),
                "guidance_type": f_choice(tuple(t.name.lower() for t in GuidanceType), "linear"),
                "start_step": f_int(0, min=0),
                "end_step": f_int(9999, min=0),
                "latent": ("LATENT",),
            },
            "optional": {},
        }

    @classmethod
    def make_guidance_cfg(cls, guidance_type, factor, start_step, end_step, latent):
        return (
            GuidanceConfig(
                guidance_type=GuidanceType[guidance_type.upper()],
                factor=factor,
                start_step=start_step,
                end_step=end_step,
                latent=latent.get("samples"),
            ),
        )


_MOMENTUM_FIELDS = {
    "momentum": f_float(0.95, min=-0.5, max=2.5, tooltip="How much of the normal result is used; 1.0 disables momentum."),
    "momentum_hist": f_float(0.75, min=-1.5, max=1.5, tooltip="How much of the history is kept when it is updated."),
    "momentum_init": f_choice(tuple(t.name for t in HistoryType), "ZERO"),
    "direction": f_float(1.0, min=-30.0, max=15.0),
    "rand_init_noise_type": f_choice(noise_names(skip=(NoiseType.BROWNIAN,)), "gaussian"),
}


class SamplerNodeSonarEuler:
    DESCRIPTION = "Sonar - momentum based sampler node."
    RETURN_TYPES = ("SAMPLER",)
    CATEGORY = "sampling/custom_sampling/samplers"
    FUNCTION = "get_sampler"

    @classmethod
    def INPUT_TYPES(cls) -> dict:  # noqa: N802
        return {"required": dict(_MOMENTUM_FIELDS), "optional": {"guidance_cfg_opt": ("SONAR_GUIDANCE_CFG",)}}

    @classmethod
    def get_sampler(cls, *, momentum, momentum_hist, momentum_init, direction, rand_init_noise_type, guidance_cfg_opt=None):
        cfg = SonarConfig(
            momentum=momentum,
            init=HistoryType[momentum_init.upper()],
            momentum_hist=momentum_hist,
            direction=direction,
            rand_init_noise_type=NoiseType[rand_init_noise_type.upper()],
            guidance=guidance_cfg_opt,
        )
        return (comfy_samplers.KSAMPLER(SonarEuler.sampler, {"sonar_config": cfg}),)


class SamplerNodeSonarEulerAncestral(SamplerNodeSonarEuler):
    SAMPLER_FN = staticmethod(SonarEulerAncestral.sampler)
    DEFAULT_NOISE = "gaussian"

    @classmethod
    def INPUT_TYPES(cls) -> dict:  # noqa: N802
        base = super().INPUT_TYPES()
        base["required"] |= {
            "s_noise": f_float(1.0, tooltip="Multiplier for noise added during ancestral or SDE sampling."),
            "eta": f_float(1.0, tooltip="Controls the ancestralness of the sampler; 0 gives a non-ancestral sampler."),
            "noise_type": f_choice(noise_names(), cls.DEFAULT_NOISE),
        }
        base["optional"] |= {"custom_noise_opt": f_noise("Custom noise used during ancestral or SDE sampling.")}
        return base

    @classmethod
    def get_sampler(
        cls,
        *,
        momentum,
        momentum_hist,
        momentum_init,
        direction,
        rand_init_noise_type,
        noise_type,
        eta,
        s_noise,
        guidance_cfg_opt=None,
        custom_noise_opt=None,
    ):
        cfg = SonarConfig(
            momentum=momentum,
            init=HistoryType[momentum_init.upper()],
            momentum_hist=momentum_hist,
            direction=direction,
            rand_init_noise_type=NoiseType[rand_init_noise_type.upper()],
            noise_type=NoiseType[noise_type.upper()],
            custom_noise=custom_noise_opt.clone() if custom_noise_opt else None,
            guidance=guidance_cfg_opt,
        )
        return (comfy_samplers.KSAMPLER(cls.SAMPLER_FN, {"sonar_config": cfg, "eta": eta, "s_noise": s_noise}),)


class SamplerNodeSonarDPMPPSDE(SamplerNodeSonarEulerAncestral):
    SAMPLER_FN = staticmethod(SonarDPMPPSDE.sampler)
    DEFAULT_NOISE = "brownian"


class SamplerNodeConfigOverride:
    DESCRIPTION = "Overrides configuration settings (noise type, eta, ...) for other samplers."
    RETURN_TYPES = ("SAMPLER",)
    CATEGORY = "sampling/custom_sampling/samplers"
    FUNCTION = "get_sampler"

    @classmethod
    def INPUT_TYPES(cls) -> dict:  # noqa: N802
        return {
            "required": {
                "sampler": ("SAMPLER",),
                "eta": f_float(1.0),
                "s_noise": f_float(1.0),
                "s_churn": f_float(0.0),
                "r": f_float(0.5),
                "sde_solver": (("midpoint", "heun"),),
                "cpu_noise": f_bool(True),
                "normalize": f_bool(True),
            },
            "optional": {
                "noise_type": f_choice(noise_names(first="DEFAULT"), "DEFAULT"),
                "custom_noise_opt": f_noise(),
                "yaml_parameters": ("STRING", dict(YAML_OPTS)),
            },
        }

    def get_sampler(
        self,
        *,
        sampler,
        eta,
        s_noise,
        s_churn,
        r,
        sde_solver,
        cpu_noise=True,
        noise_type=None,
        custom_noise_opt=None,
        normalize=True,
        yaml_parameters="",
    ):
        sampler_kwargs = {"s_noise": s_noise, "eta": eta, "s_churn": s_churn, "r": r, "solver_type": sde_solver}
        if yaml_parameters:
            extra = yaml.safe_load(yaml_parameters)
            if extra is not None:
                if not isinstance(extra, dict):
                    raise ValueError("SamplerConfigOverride: yaml_parameters must either be null or an object")
                sampler_kwargs |= extra
        override = {
            "sampler": sampler,
            "noise_type": NoiseType[noise_type.upper()] if noise_type not in {None, "DEFAULT"} else None,
            "custom_noise": custom_noise_opt,
            "sampler_kwargs": sampler_kwargs,
            "cpu_noise": cpu_noise,
            "normalize": normalize,
        }
        fn = functools.update_wrapper(
            functools.partial(self.sampler_function, override_sampler_cfg=override),
            sampler.sampler_function,
        )
        return (
            comfy_samplers.KSAMPLER(
                fn,
                extra_options=sampler.extra_options.copy(),
                inpaint_options=sampler.inpaint_options.copy(),
            ),
        )

    @staticmethod
    def sampler_function(
        model,
        x,
        sigmas,
        *args: Any,
        override_sampler_cfg: dict[str, Any] | None = None,
        noise_sampler: Callable | None = None,
        extra_args: dict[str, Any] | None = None,
        **kwargs: Any,
    ) -> torch.Tensor:
        if not override_sampler_cfg:
            raise ValueError("Override sampler config missing!")
        extra_args = {} if extra_args is None else extra_args
        cfg = override_sampler_cfg
        target = cfg["sampler"]
        params = inspect.signature(target.sampler_function).parameters
        if "noise_sampler" in params:
            host = sigmas.detach().float().cpu()
            sigma_min, sigma_max = host[host > 0].min(), host.max()
            common = {"seed": extra_args.get("seed"), "cpu": cfg.get("cpu_noise", True), "normalized": cfg.get("normalize", True)}
            if cfg.get("custom_noise") is not None:
                noise_sampler = cfg["custom_noise"].make_noise_sampler(x, sigma_min, sigma_max, **common)
            elif cfg.get("noise_type") is not None:
                noise_sampler = noise.get_noise_sampler(cfg["noise_type"], x, sigma_min, sigma_max, **common)
            kwargs["noise_sampler"] = noise_sampler
        kwargs |= {k: v for k, v in cfg["sampler_kwargs"].items() if k in params}
        return target.sampler_function(model, x, sigmas, *args, extra_args=extra_args, **kwargs)


# ---------------------------------------------------------------------------------------------
# wavelet CFG node
# ---------------------------------------------------------------------------------------------
WCFG_DEFAULT_YAML = """\
# YAML or JSON here. Same keys as ComfyUI-sonar's SonarWaveletCFG (docs/waveletcfg.md upstream).
# Fields of the node (start_sigma, ...) may be overridden here.

# The CFG scale, per frequency band. All scales sections (difference, cond, uncond, final) share
# this format.
difference:
    # low-frequency (approximation) band
    yl_scale: 5.0
    # high-frequency bands: a scalar, a list (one entry per level, fine to coarse; "fill" repeats the
    # previous entry) or a list of [horizontal, vertical, diagonal] lists
    yh_scales: 3.0
    # optional scales_end block + schedule / schedule_mode / schedule_offset / ... to interpolate
    schedule: linear
    schedule_mode: sampling
    reverse_schedule: false
    schedule_offset: 0.0
    schedule_multiplier: 1.0
    schedule_offset_after: 0.0
    schedule_multiplier_after: 1.0
    schedule_min: 0.0
    schedule_max: 1.0
    blend_mode: lerp

# Daubechies wavelets db1..db12 and haar are built in.
wave: db4
level: 5

### advanced options
padding_mode: symmetric
# only the 2-D DWT has kernels: the two switches below must stay off
use_1d_dwt: false
use_dtcwt: false
biort: near_sym_a
qshift: qshift_a

# denoised, noise or noise_norm: what the wavelet CFG is computed on
target_mode: denoised

# scales applied to cond / uncond before the difference, and to the final result
cond:
    yl_scale: 1.0
    yh_scales: 1.0
uncond:
    yl_scale: 1.0
    yh_scales: 1.0
final:
    yl_scale: 1.0
    yh_scales: 1.0

# float64 coefficients (true) or the latent's dtype (false)
high_precision_mode: true

# how the scaled difference is combined with uncond: inject (uncond + diff * strength), lerp, subtract_b
difference_blend_mode: inject
difference_blend_strength: 1.0

verbose: false
"""


class SonarWaveletCFGNode:
    DESCRIPTION = "Wavelet CFG: classifier-free guidance with a separate scale per wavelet band."
    RETURN_TYPES = ("MODEL",)
    CATEGORY = "model_patches"
    FUNCTION = "go"

    @classmethod
    def INPUT_TYPES(cls) -> dict:  # noqa: N802
        return {
            "required": {
                "model": ("MODEL",),
                "start_sigma": f_float(-1.0, min=-1.0, tooltip="First sigma the rule is active at; negative = infinity."),
                "end_sigma": f_float(0.0, min=0.0),
                "fallback_mode": f_choice(("existing", "own"), "existing"),
                "blend_mode": f_choice(tuple(hostutil.BLENDING_MODES), "lerp"),
                "blend_strength": f_float(1.0),
                "yaml_parameters": ("STRING", YAML_OPTS | {"default": WCFG_DEFAULT_YAML}),
            },
            "optional": {
                name: ("LATENT_OPERATION",)
                for name in (
                    "operation_cond", "operation_uncond", "operation_fallback_cfg", "operation_wavelet_cfg", "operation_result",
                )  # fmt: skip
            },
        }

    @classmethod
    def go(
        cls,
        *,
        model: object,
        start_sigma: float,
        end_sigma: float,
        fallback_mode: str,
        blend_mode: str,
        blend_strength: float,
        yaml_parameters: str,
        operation_cond: Callable | None = None,
        operation_uncond: Callable | None = None,
        operation_fallback_cfg: Callable | None = None,
        operation_wavelet_cfg: Callable | None = None,
        operation_result: Callable | None = None,
        _override_rules_dict: dict | None = None,
    ) -> tuple[object]:
        if start_sigma < 0:
            start_sigma = math.inf
        params = dict(_override_rules_dict) if _override_rules_dict is not None else (yaml.safe_load(yaml_parameters) or {})
        rules = WCFGRules.build(
            **(
                {
                    "start_sigma": start_sigma,
                    "end_sigma": end_sigma,
                    "fallback_existing": fallback_mode == "existing",
                    "blend_mode": blend_mode,
                    "blend_strength": blend_strength,
                }
                | params
            ),
        )
        model = model.clone()
        model.set_model_sampler_cfg_function(
            WaveletCFG(
                existing_cfg=model.model_options.get("sampler_cfg_function"),
                rules=rules,
                operation_cond=operation_cond,
                operation_uncond=operation_uncond,
                operation_fallback_cfg=operation_fallback_cfg,
                operation_wavelet_cfg=operation_wavelet_cfg,
                operation_result=operation_result,
            ),
        )
        return (model,)


NODE_CLASS_MAPPINGS = {
    "SonarCustomNoise": SonarCustomNoiseNode,
    "SonarCustomNoiseAdv": SonarCustomNoiseAdvNode,
    "SonarAdvancedPyramidNoise": SonarAdvancedPyramidNoiseNode,
    "SonarAdvanced1fNoise": SonarAdvanced1fNoiseNode,
    "SonarAdvancedPowerLawNoise": SonarAdvancedPowerLawNoiseNode,
    "SonarRepeatedNoise": SonarRepeatedNoiseNode,
    "SonarScheduledNoise": SonarScheduledNoiseNode,
    "SonarCompositeNoise": SonarCompositeNoiseNode,
    "SonarBlendedNoise": SonarBlendedNoiseNode,
    "SonarGuidedNoise": SonarGuidedNoiseNode,
    "SonarCustomNoiseParameters": SonarCustomNoiseParametersNode,
    "SonarWaveletFilteredNoise": SonarWaveletFilteredNoiseNode,
    "SonarPowerNoise": SonarPowerNoiseNode,
    "SonarPowerFilterNoise": SonarPowerFilterNoiseNode,
    "SonarPowerFilter": SonarPowerFilterNode,
    "SONAR_CUSTOM_NOISE to NOISE": SonarToComfyNOISENode,
    "NoisyLatentLike": NoisyLatentLikeNode,
    "SonarGuidanceConfig": GuidanceConfigNode,
    "SamplerSonarEuler": SamplerNodeSonarEuler,
    "SamplerSonarEulerA": SamplerNodeSonarEulerAncestral,
    "SamplerSonarDPMPPSDE": SamplerNodeSonarDPMPPSDE,
    "SamplerConfigOverride": SamplerNodeConfigOverride,
    "SonarWaveletCFG": SonarWaveletCFGNode,
}
NODE_DISPLAY_NAME_MAPPINGS: dict = {}
