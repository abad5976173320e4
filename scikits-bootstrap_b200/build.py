"""In-tree build of libb200boot.so (sm_100a only).

    python scikits-bootstrap_b200/build.py [--force]

nvcc cross-compiles without a GPU; the .so is git-ignored but travels to the GPU
box with the repo snapshot."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, "csrc")
LIB_PATH = os.path.join(PKG_DIR, "libb200boot.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17", "--expt-relaxed-constexpr",
    "--shared", "-Xcompiler", "-fPIC",
]


def _sources():
    out = [os.path.join(PKG_DIR, "..", "include", "b200boot.h")]
    for name in sorted(os.listdir(CSRC)):
        if name.endswith((".cu", ".cuh", ".h")):
            out.append(os.path.join(CSRC, name))
    return out


def needs_build() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    built = os.path.getmtime(LIB_PATH)
    return any(os.path.getmtime(s) > built for s in _sources())


def build(force: bool = False, verbose: bool = False, defines=(), out: str | None = None) -> str:
    """Compiles csrc/b200boot.cu into libb200boot.so; returns the library path.
    `defines` (e.g. ["B2_K1_PIPE=1"]) and `out` build an experimental variant elsewhere."""
    if out is None and not defines and not force and not needs_build():
        return LIB_PATH
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found; cannot build libb200boot.so")
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + \
        ["-D" + d for d in defines] + ["-o", out or LIB_PATH, os.path.join(CSRC, "b200boot.cu")]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + proc.stdout + proc.stderr)
    if verbose:
        print(proc.stderr)
    return out or LIB_PATH


if __name__ == "__main__":
    defs = [a[2:] for a in sys.argv[1:] if a.startswith("-D")]
    outs = [a[2:] for a in sys.argv[1:] if a.startswith("-o")]
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, defines=defs,
                out=outs[0] if outs else None))
