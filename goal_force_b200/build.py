"""Build libgoalforce_b200.so in-tree with nvcc for sm_100a.

The shared library is the C-ABI boundary declared in include/goalforce_b200.h. It is built here (the authoring
container cross-compiles without a GPU) and travels to the GPU box with the repository snapshot.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
LIBDIR = PKG / "_lib"
LIBNAME = "libgoalforce_b200.so"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr",
    "-Xptxas", "-v",
]


def _nvcc() -> str:
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found; cannot build libgoalforce_b200.so")
    return nvcc


def sources() -> list[Path]:
    return sorted(CSRC.glob("*.cu"))


def _digest() -> str:
    h = hashlib.sha256()
    for p in sorted(list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cuh")) + list(CSRC.glob("*.h"))
                    + [PKG.parent / "include" / "goalforce_b200.h"]):
        h.update(p.name.encode())
        h.update(p.read_bytes())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def lib_path() -> Path:
    return LIBDIR / LIBNAME


def build(force: bool = False, verbose: bool = False) -> Path:
    """Compile every .cu under csrc/ into one shared library. Skips the work when sources are unchanged."""
    LIBDIR.mkdir(exist_ok=True)
    stamp = LIBDIR / "build.sha256"
    dig = _digest()
    if not force and lib_path().exists() and stamp.exists() and stamp.read_text().strip() == dig:
        return lib_path()
    nvcc = _nvcc()
    objs = []
    procs = []
    for src in sources():
        obj = LIBDIR / (src.stem + ".o")
        cmd = [nvcc, *NVCC_FLAGS, "-c", str(src), "-o", str(obj)]
        procs.append((src, obj, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    log = []
    for src, obj, pr in procs:
        out, _ = pr.communicate()
        log.append(f"==== {src.name}\n{out}")
        if pr.returncode != 0:
            (LIBDIR / "build.log").write_text("\n".join(log))
            raise RuntimeError(f"nvcc failed on {src.name}:\n{out}")
        objs.append(str(obj))
    link = [nvcc, "-shared", "-o", str(lib_path()), *objs, "-lcudart"]
    r = subprocess.run(link, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    log.append("==== link\n" + r.stdout)
    (LIBDIR / "build.log").write_text("\n".join(log))
    if r.returncode != 0:
        raise RuntimeError("link failed:\n" + r.stdout)
    stamp.write_text(dig)
    if verbose:
        print("\n".join(log))
    return lib_path()


if __name__ == "__main__":
    p = build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(p)
