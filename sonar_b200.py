"""Import alias: `import sonar_b200` loads the package that lives in `comfyui-sonar_b200/`
(a ComfyUI custom-node directory name, not a valid Python identifier)."""

import importlib.util
import sys
from pathlib import Path

_PKG_DIR = Path(__file__).resolve().parent / "comfyui-sonar_b200"
_spec = importlib.util.spec_from_file_location(
    "sonar_b200",
    _PKG_DIR / "__init__.py",
    submodule_search_locations=[str(_PKG_DIR)],
)
_module = importlib.util.module_from_spec(_spec)
sys.modules["sonar_b200"] = _module
_spec.loader.exec_module(_module)
