"""comfy.samplers: just enough for sampler registration and KSAMPLER construction."""
from .k_diffusion import sampling as k_diffusion_sampling


class KSampler:
    SAMPLERS = ["euler", "euler_ancestral", "dpmpp_sde"]


class KSAMPLER:
    def __init__(self, sampler_function, extra_options=None, inpaint_options=None):
        self.sampler_function = sampler_function
        self.extra_options = {} if extra_options is None else extra_options
        self.inpaint_options = {} if inpaint_options is None else inpaint_options


def ksampler(sampler_name, extra_options=None, inpaint_options=None):
    return KSAMPLER(getattr(k_diffusion_sampling, f"sample_{sampler_name}"), extra_options, inpaint_options)
