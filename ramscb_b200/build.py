"""Build the CUDA library in-tree for sm_100a (nvcc cross-compiles without a GPU).

    python -m ramscb_b200.build [--force]

Produces ramscb_b200/lib/libramscb_gpu.so.  -fmad=false: the EXACT-mode kernels
must round every product separately to stay bit-identical to the reference's
operation order; FAST-mode code requests FMAs explicitly with fma().
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIBDIR, "libramscb_gpu.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-fmad=false", "-shared", "-Xcompiler", "-fPIC,-ffp-contract=off,-O2",
    "-ccbin", "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++",
]


LINK_LIBS = []


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if not f.endswith(".o")] + [os.path.join(HERE, "..", "include", "ramscb_gpu.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not _stale():
        return LIB
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    os.makedirs(LIBDIR, exist_ok=True)
    objdir = os.path.join(LIBDIR, "obj")
    os.makedirs(objdir, exist_ok=True)
    flags = [f for f in NVCC_FLAGS if f != "-shared"] + (["-Xptxas", "-v"] if verbose else [])
    # one nvcc per translation unit, side by side (the two big files take about a minute each)
    procs = []
    for src in sources():
        obj = os.path.join(objdir, os.path.basename(src)[:-3] + ".o")
        procs.append((src, obj, subprocess.Popen([nvcc] + flags + ["-c", "-o", obj, src], stdout=subprocess.PIPE,
                                                 stderr=subprocess.PIPE, text=True)))
    objs, log = [], ""
    failed = False
    for src, obj, pr in procs:
        out, err = pr.communicate()
        log += out + err
        failed = failed or pr.returncode != 0
        objs.append(obj)
    if failed:
        sys.stderr.write(log)
        raise RuntimeError("nvcc failed")
    r = subprocess.run([nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB] + objs + LINK_LIBS,
                       capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("nvcc link failed")
    if verbose:
        sys.stderr.write(log)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
