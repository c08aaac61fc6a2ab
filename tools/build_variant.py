#!/usr/bin/env python
"""Build an experimental copy of libgoalforce_b200.so with extra nvcc flags (kernel A/B work; never the shipped library).

    python tools/build_variant.py trace -DGF_A8_TRACE
    GF_B200_LIB=goal_force_b200/_lib/variants/trace/libgoalforce_b200.so python tools/attn_trace.py

The variant lands in goal_force_b200/_lib/variants/<name>/ (git-ignored, travels to the GPU box with the snapshot).
"""
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from goal_force_b200 import build as B  # noqa: E402


def main():
    name, flags = sys.argv[1], sys.argv[2:]
    out = B.LIBDIR / "variants" / name
    out.mkdir(parents=True, exist_ok=True)
    procs = []
    for src in B.sources():
        obj = out / (src.stem + ".o")
        procs.append((src, obj, subprocess.Popen([B._nvcc(), *B.NVCC_FLAGS, *flags, "-c", str(src), "-o", str(obj)],
                                                 stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    objs = []
    for src, obj, pr in procs:
        log, _ = pr.communicate()
        if pr.returncode:
            raise SystemExit(f"nvcc failed on {src.name}:\n{log}")
        objs.append(str(obj))
    lib = out / B.LIBNAME
    subprocess.run([B._nvcc(), "-shared", "-o", str(lib), *objs, "-lcudart"], check=True)
    for o in objs:
        Path(o).unlink()
    print(lib)


if __name__ == "__main__":
    main()
