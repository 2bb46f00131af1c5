"""Builds the C-ABI CUDA library `libc2w_b200.so` in-tree with nvcc for sm_100a.

The library is plain `extern "C"` (see include/c2w_b200.h): no torch headers, no pybind.
nvcc cross-compiles without a GPU, so this also runs in the CPU-only build container.
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
CSRC = PKG_DIR / "csrc"
LIB_PATH = PKG_DIR / "libc2w_b200.so"
OBJ_DIR = PKG_DIR / "build"

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
    raise RuntimeError("nvcc not found: cannot build libc2w_b200.so")


def _sources() -> list[Path]:
    return sorted(CSRC.glob("*.cu"))


def _fingerprint() -> str:
    h = hashlib.sha256()
    for f in sorted(list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cuh")) + [PKG_DIR.parent / "include" / "c2w_b200.h"]):
        if f.exists():
            h.update(f.name.encode())
            h.update(f.read_bytes())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False, diag: bool = False) -> Path:
    """Compile every csrc/*.cu and link libc2w_b200.so.  Returns the library path.
    diag=True builds libc2w_b200_diag.so with -DC2W_DIAG instead: K1 with its per-role cycle counters and the
    load-skipping timing experiments (tools/bringup_conv.py); the shipped library has none of that code."""
    lib_path = PKG_DIR / "libc2w_b200_diag.so" if diag else LIB_PATH
    obj_dir = PKG_DIR / "build" / "diag" if diag else OBJ_DIR
    stamp = obj_dir / "fingerprint.txt"
    fp = _fingerprint()
    if not force and lib_path.exists() and stamp.exists() and stamp.read_text() == fp:
        return lib_path
    nvcc = _nvcc()
    obj_dir.mkdir(parents=True, exist_ok=True)
    inc = ["-I", str(CSRC), "-I", str(PKG_DIR.parent / "include")]
    extra = ["-DC2W_DIAG"] if diag else []

    def compile_one(src: Path) -> Path:
        obj = obj_dir / (src.stem + ".o")
        cmd = [nvcc, *NVCC_FLAGS, *extra, *inc, "-c", str(src), "-o", str(obj)]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src.name}:\n{r.stdout}\n{r.stderr}")
        if verbose:
            sys.stderr.write(r.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        objs = list(ex.map(compile_one, _sources()))
    cmd = [nvcc, "-shared", "-o", str(lib_path), *map(str, objs)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    stamp.write_text(fp)
    return lib_path


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, diag="--diag" in sys.argv))
