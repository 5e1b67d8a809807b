"""Minimal on-disk stand-in for the parts of ComfyUI the reference node pack imports.

TEST INFRASTRUCTURE ONLY (tests/, tests/golden/make_golden.py, bench.py --impl reference). It lets the
unmodified reference at /root/reference -- and this repo's node surface -- import without ComfyUI.
The helper semantics restate upstream ComfyUI (comfy/k_diffusion/sampling.py, comfy/utils.py); the
reference pins no ComfyUI version.
"""
from . import latent_formats, model_management, samplers, utils  # noqa: F401
from .k_diffusion import sampling  # noqa: F401
