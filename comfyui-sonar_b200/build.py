"""Builds libsonar_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python comfyui-sonar_b200/build.py [--force] [--verbose]

The shared library is a plain C-ABI object (see include/sonar_b200.h): no torch headers, no
pybind, so a rebuild takes seconds per translation unit. The .so is git-ignored but travels to the
GPU box with the gpurun snapshot.
"""

from __future__ import annotations

import concurrent.futures
import os
import shutil
import subprocess
import sys
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
CSRC = PKG_DIR / "csrc"
INCLUDE = PKG_DIR.parent / "include"
LIB_PATH = PKG_DIR / "libsonar_b200.so"
OBJ_DIR = PKG_DIR / "build"

NVCC_FLAGS = [
    "-O3",
    "-std=c++17",
    "-gencode",
    "arch=compute_100a,code=sm_100a",
    "-lineinfo",
    "-Xcompiler",
    "-fPIC",
    "--expt-relaxed-constexpr",
    "-I",
    str(INCLUDE),
]


def find_nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found: sonar_b200 has no CPU fallback and cannot be built without the CUDA toolkit")


def sources() -> list[Path]:
    return sorted(CSRC.glob("*.cu"))


def needs_rebuild() -> bool:
    if not LIB_PATH.exists():
        return True
    lib_mtime = LIB_PATH.stat().st_mtime
    deps = [*CSRC.glob("*.cu"), *CSRC.glob("*.cuh"), *INCLUDE.glob("*.h"), Path(__file__)]
    return any(d.stat().st_mtime > lib_mtime for d in deps)


def _compile_one(nvcc: str, src: Path, verbose: bool) -> Path:
    obj = OBJ_DIR / (src.stem + ".o")
    hdr_mtime = max(
        [p.stat().st_mtime for p in (*CSRC.glob("*.cuh"), *INCLUDE.glob("*.h"), Path(__file__))],
    )
    if obj.exists() and obj.stat().st_mtime > max(src.stat().st_mtime, hdr_mtime):
        return obj
    cmd = [nvcc, *NVCC_FLAGS, "-c", str(src), "-o", str(obj)]
    if verbose:
        cmd.insert(1, "-Xptxas")
        cmd.insert(2, "-v")
    proc = subprocess.run(cmd, capture_output=True, text=True, check=False)
    if proc.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src.name}:\n{proc.stdout}\n{proc.stderr}")
    if verbose:
        print(proc.stderr, file=sys.stderr)
    return obj


def build_library(*, force: bool = False, verbose: bool = False) -> Path:
    if not force and not needs_rebuild():
        return LIB_PATH
    nvcc = find_nvcc()
    OBJ_DIR.mkdir(exist_ok=True)
    if force:
        for o in OBJ_DIR.glob("*.o"):
            o.unlink()
    srcs = sources()
    with concurrent.futures.ThreadPoolExecutor(max_workers=min(8, len(srcs))) as pool:
        objs = list(pool.map(lambda s: _compile_one(nvcc, s, verbose), srcs))
    tmp = LIB_PATH.with_suffix(".so.tmp")
    link = [
        nvcc,
        "-shared",
        "-gencode",
        "arch=compute_100a,code=sm_100a",
        "-Xcompiler",
        "-fPIC",
        "-o",
        str(tmp),
        *map(str, objs),
    ]
    proc = subprocess.run(link, capture_output=True, text=True, check=False)
    if proc.returncode != 0:
        raise RuntimeError(f"link failed:\n{proc.stdout}\n{proc.stderr}")
    os.replace(tmp, LIB_PATH)
    return LIB_PATH


if __name__ == "__main__":
    path = build_library(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    print(path)
