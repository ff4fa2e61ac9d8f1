"""Builds ``libb200clip.so`` (hand-written sm_100a kernels + C ABI) in-tree with nvcc.

``python -m construction_clip_b200.build`` or ``build()``; nvcc cross-compiles without a GPU.
The shared library is git-ignored but travels with the working tree to the GPU box.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
LIB = PKG / "libb200clip.so"
OBJ = PKG / "_obj"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found (set NVCC=/path/to/nvcc)")


def sources() -> list[Path]:
    return sorted(CSRC.glob("*.cu"))


def _deps_mtime() -> float:
    deps = list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cuh")) + list(CSRC.glob("*.h"))
    deps.append(PKG.parent / "include" / "b200clip.h")
    return max(p.stat().st_mtime for p in deps)


def is_stale() -> bool:
    return not LIB.exists() or LIB.stat().st_mtime < _deps_mtime()


def build(force: bool = False, verbose: bool = False) -> Path:
    if not force and not is_stale():
        return LIB
    nvcc = _nvcc()
    OBJ.mkdir(exist_ok=True)
    hdr_mtime = max(p.stat().st_mtime for p in list(CSRC.glob("*.cuh")) + list(CSRC.glob("*.h")) +
                    [PKG.parent / "include" / "b200clip.h"])

    def compile_one(src: Path) -> Path:
        obj = OBJ / (src.stem + ".o")
        if not force and obj.exists() and obj.stat().st_mtime >= max(src.stat().st_mtime, hdr_mtime):
            return obj
        cmd = [nvcc, *NVCC_FLAGS, "-c", str(src), "-o", str(obj)]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src.name}:\n{r.stdout}\n{r.stderr}")
        if verbose:
            sys.stderr.write(r.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        objs = list(ex.map(compile_one, sources()))
    tmp = LIB.with_suffix(".so.tmp")
    cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", str(tmp), *map(str, objs)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    os.replace(tmp, LIB)
    return LIB


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(path)
