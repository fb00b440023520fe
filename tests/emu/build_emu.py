"""TEST INFRASTRUCTURE ONLY: compile the device library (RAM and SCB) for the host-CPU CUDA emulator.

    python tests/emu/build_emu.py [--force]

Copies ramscb_b200/csrc/{*.cuh, *.cu} into tests/emu/_gen/, rewriting only what g++
cannot parse -- `kernel<<<grid, block, smem, stream>>>(args);` becomes
`emu::launch(grid, block, smem, stream, kernel, args);`, `extern __shared__ T name[];` becomes a
pointer to the emulated dynamic shared memory, and the six `cp.async` inline-PTX statements become
plain copies -- and compiles the result against tests/emu/cuda_runtime.h into
tests/emu/_gen/libramscb_emu.so.  The kernels' bodies are otherwise untouched, so a parity test of
this library against the oracle exercises the same indexing, barriers, shuffles and arithmetic as
the sm_100a build.  The product (ramscb_b200/) never loads this library.
"""
from __future__ import annotations

import os
import re
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "ramscb_b200", "csrc")
GEN = os.path.join(HERE, "_gen")
LIB = os.path.join(GEN, "libramscb_emu.so")
FILES = ["ram_common.cuh", "ram_kernels.cuh", "ram_fused.cuh", "ram_coulomb.cuh", "ram_diffcoef.cuh", "ram_gpu.cu", "scb_kernels.cuh", "scb_press_front.cuh", "hi_kernels.cuh", "scb_gpu.cu"]


def _match(src, i, open_ch, close_ch):
    """index just after the bracket that closes the one at src[i]"""
    depth = 0
    while True:
        c = src[i]
        if c == open_ch:
            depth += 1
        elif c == close_ch:
            depth -= 1
            if depth == 0:
                return i + 1
        i += 1


def _split(text):
    """split at top-level commas"""
    parts, depth, cur = [], 0, ""
    for c in text:
        if c in "([{":
            depth += 1
        elif c in ")]}":
            depth -= 1
        if c == "," and depth == 0:
            parts.append(cur.strip())
            cur = ""
        else:
            cur += c
    if cur.strip():
        parts.append(cur.strip())
    return parts


def rewrite_launches(src):
    out, pos = [], 0
    while True:
        i = src.find("<<<", pos)
        if i < 0:
            out.append(src[pos:])
            return "".join(out)
        # kernel expression: identifier, optionally followed by a <...> template argument list
        j = i
        if src[j - 1] == ">":
            depth, j = 0, j - 1
            while True:
                if src[j] == ">":
                    depth += 1
                elif src[j] == "<":
                    depth -= 1
                    if depth == 0:
                        break
                j -= 1
        while src[j - 1].isalnum() or src[j - 1] == "_":
            j -= 1
        kernel = src[j:i]
        k = src.index(">>>", i)
        cfg = src[i + 3:k]
        parts = _split(cfg)
        while len(parts) < 4:
            parts.append("0")
        a0 = src.index("(", k)
        a1 = _match(src, a0, "(", ")")
        args = _split(src[a0 + 1:a1 - 1])
        out.append(src[pos:j])
        # arguments are evaluated (and copied) at the launch, like a real launch or a captured graph node
        binds = " ".join(f"auto emu_a{n} = ({a});" for n, a in enumerate(args))
        call = ", ".join(f"emu_a{n}" for n in range(len(args)))
        out.append(f"do {{ {binds} emu::launch({parts[0]}, {parts[1]}, {parts[2]}, {parts[3]}, [=]() {{ {kernel}({call}); }}); }} while (0)")
        pos = a1


def rewrite(src, name):
    src = re.sub(r"extern\s+__shared__\s+(\w+)\s+(\w+)\s*\[\s*\]\s*;", r"\1* \2 = (\1*)emu::dyn_smem();", src)
    src = re.sub(r'asm volatile\("cp\.async\.ca\.shared\.global[^\n]*\n', "memcpy(smem_dst, gsrc, 8);\n", src)
    src = re.sub(r'asm volatile\("cp\.async\.cg\.shared\.global[^\n]*\n', "memcpy(smem_dst, gsrc, 16);\n", src)
    src = re.sub(r'asm volatile\("cp\.async\.(commit_group|wait_group \d+);"\);', ";", src)
    # TMA bulk copies and their mbarrier (ram_fused.cuh): the copy happens at once, the barrier is always passed --
    # the kernels follow the wait with a __syncthreads(), which is what orders the emulated threads
    src = re.sub(r'asm volatile\("cp\.async\.bulk\.shared::cluster\.global[^\n]*\n', "memcpy(smem_dst, gsrc, bytes);\n", src)
    src = re.sub(r'asm volatile\("cp\.async\.bulk\.global\.shared::cta[^\n]*\n', "memcpy(gdst, smem_src, bytes);\n", src)
    src = re.sub(r'asm volatile\("\{ \.reg \.pred p; mbarrier\.try_wait[^\n]*\n', "done = 1;\n", src)
    src = re.sub(r'asm volatile\("(mbarrier\.|fence\.|cp\.async\.bulk\.(commit_group|wait_group))[^\n]*\n', ";\n", src)
    if "asm volatile" in src or "asm(" in src:
        raise SystemExit(f"{name}: inline PTX the emulator does not know")
    return rewrite_launches(src)


def build(force=False):
    os.makedirs(GEN, exist_ok=True)
    deps = [os.path.join(CSRC, f) for f in FILES] + [os.path.join(HERE, "cuda_runtime.h"), os.path.abspath(__file__),
                                                      os.path.join(ROOT, "include", "ramscb_gpu.h")]
    if not force and os.path.exists(LIB) and all(os.path.getmtime(d) <= os.path.getmtime(LIB) for d in deps):
        return LIB
    for f in FILES:
        with open(os.path.join(CSRC, f)) as fh:
            src = fh.read()
        src = rewrite(src, f).replace('"../../include/ramscb_gpu.h"', '"ramscb_gpu.h"')
        with open(os.path.join(GEN, f.replace(".cu", ".cpp") if f.endswith(".cu") else f), "w") as fh:
            fh.write(src)
    with open(os.path.join(GEN, "emu_impl.cpp"), "w") as fh:      # the one definition of the context switch
        fh.write('#define EMU_IMPLEMENT\n#include "cuda_runtime.h"\n')
    cmd = ["g++", "-std=c++17", "-O2", "-g", "-ffp-contract=off", "-fno-fast-math", "-fPIC", "-shared",
           "-fno-stack-protector", "-Wno-unused-result",
           "-I", HERE, "-I", os.path.join(ROOT, "include"), "-I", GEN,
           "-o", LIB, os.path.join(GEN, "ram_gpu.cpp"), os.path.join(GEN, "scb_gpu.cpp"), os.path.join(GEN, "emu_impl.cpp")]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr[-20000:])
        raise RuntimeError("g++ failed on the emulator build")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
