"""sonar_b200 -- B200-native (sm_100a) implementation of ComfyUI-sonar's per-step hot path.

Drop-in ComfyUI custom-node package: exports NODE_CLASS_MAPPINGS / NODE_DISPLAY_NAME_MAPPINGS and
registers the three Sonar samplers when ComfyUI (`comfy`) is importable (reference __init__.py:1-23).
Without ComfyUI (tests, bench) the compute modules are importable on their own:

    ops            tensor-level wrappers over the C ABI (include/sonar_b200.h)
    torch_ops      the same entry points as torch.library custom ops (torch.ops.sonar_b200.*)
    generators     noise generators         (reference py/noise_generation.py)
    noise_graph    chains / items           (reference py/noise.py)
    spectral_noise power-law spectral noise (reference py/nodes/powernoise.py)
    freeu          FreeU-Extreme filter     (reference py/nodes/freeu_extreme.py)
    samplers       Sonar samplers           (reference py/sonar.py)
    wavelets, wcfg wavelet CFG              (reference py/wavelet_functions.py, py/wavelet_cfg.py)
    parallel       batch sharding over GPUs

The arithmetic lives in libsonar_b200.so (hand-written CUDA); there is no CPU or eager fallback.
"""

from __future__ import annotations

import sys

from . import _native, freeu, generators, hostutil, kdiff, noise_graph, ops, parallel, rng, samplers, spectral_noise, torch_ops, wavelets, wcfg

__version__ = "0.1.0"

NODE_CLASS_MAPPINGS: dict = {}
NODE_DISPLAY_NAME_MAPPINGS: dict = {}

try:  # ComfyUI present: expose the node surface and register the samplers
    import comfy.samplers  # noqa: F401
except ImportError:
    HAVE_COMFY = False
else:
    HAVE_COMFY = True
    from . import nodes

    NODE_CLASS_MAPPINGS = nodes.NODE_CLASS_MAPPINGS | freeu.NODE_CLASS_MAPPINGS
    NODE_DISPLAY_NAME_MAPPINGS = nodes.NODE_DISPLAY_NAME_MAPPINGS
    samplers.add_samplers()
    _bi = sys.modules.get("_blepping_integrations", {})
    if "sonar" not in _bi:
        _bi["sonar"] = sys.modules[__name__]
        sys.modules["_blepping_integrations"] = _bi

__all__ = ["NODE_CLASS_MAPPINGS", "NODE_DISPLAY_NAME_MAPPINGS"]
