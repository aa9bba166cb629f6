"""Build libbvlm.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

Usage: ``python -m bayesvlm_b200.build [--force] [--verbose]``.  Object files go to ``build/`` at the repo root
(git-ignored); the shared library is written next to this file so that it travels with the source snapshot.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
REPO_ROOT = PKG_DIR.parent
CSRC = PKG_DIR / "csrc"
LIB_PATH = PKG_DIR / "libbvlm.so"
BUILD_DIR = REPO_ROOT / "build" / "bvlm"

SOURCES = ["tmap.cu", "prep.cu", "api.cu", "predictive.cu", "kfac.cu", "epig.cu"]

NVCC_FLAGS = [
    "-O3",
    "-std=c++17",
    "-gencode",
    "arch=compute_100a,code=sm_100a",
    "-lineinfo",
    "--expt-relaxed-constexpr",
    "-Xcompiler",
    "-fPIC",
]


def _nvcc() -> str:
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not Path(exe).exists():
        raise RuntimeError("nvcc not found; libbvlm.so cannot be built")
    return exe


def _signature() -> str:
    h = hashlib.sha256()
    for p in sorted(list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cuh")) + [REPO_ROOT / "include" / "bvlm.h"]):
        h.update(p.name.encode())
        h.update(p.read_bytes())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def needs_build() -> bool:
    stamp = BUILD_DIR / "signature.txt"
    return not (LIB_PATH.exists() and stamp.exists() and stamp.read_text() == _signature())


def build(force: bool = False, verbose: bool = False, diag: bool = False) -> Path:
    """``diag=True`` builds ``libbvlm_diag.so`` with -DBVLM_DIAG (the ablation switches of predictive.cu / epilogues.cuh) next
    to the product library; it is only ever loaded when ``BVLM_LIB`` points at it (scripts/diag_pred_epilogue.sh)."""
    if diag:
        return _build_to(PKG_DIR / "libbvlm_diag.so", REPO_ROOT / "build" / "bvlm_diag", ["-DBVLM_DIAG"], verbose)
    if not force and not needs_build():
        return LIB_PATH
    _build_to(LIB_PATH, BUILD_DIR, [], verbose)
    (BUILD_DIR / "signature.txt").write_text(_signature())
    return LIB_PATH


def _build_to(lib_path: Path, build_dir: Path, defines, verbose: bool) -> Path:
    build_dir.mkdir(parents=True, exist_ok=True)
    nvcc = _nvcc()
    extra = (["-Xptxas", "-v"] if verbose else []) + list(defines)

    def compile_one(name: str) -> Path:
        src = CSRC / name
        obj = build_dir / (src.stem + ".o")
        cmd = [nvcc, *NVCC_FLAGS, *extra, "-c", str(src), "-o", str(obj)]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if verbose or res.returncode != 0:
            sys.stderr.write(f"$ {' '.join(cmd)}\n{res.stdout}{res.stderr}\n")
        if res.returncode != 0:
            raise RuntimeError(f"nvcc failed on {name}")
        return obj

    sources = [s for s in SOURCES if (CSRC / s).exists()]
    with ThreadPoolExecutor(max_workers=min(len(sources), os.cpu_count() or 4)) as ex:
        objs = list(ex.map(compile_one, sources))
    cmd = [nvcc, "-shared", "-o", str(lib_path), *map(str, objs), "-gencode", "arch=compute_100a,code=sm_100a"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError(f"link of {lib_path.name} failed")
    return lib_path


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose="--verbose" in sys.argv, diag="--diag" in sys.argv)
    print(path)
