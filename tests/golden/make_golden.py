"""Regenerates the golden fixtures under tests/golden/ from the reference tree.

Run in the build container only (needs /root/reference):
    python tests/golden/make_golden.py

ekev_test1_dsbnd.txt : first column of output/test1/dsbnd.ref = EKEV(2..35) written by
                       write_dsbnd (src/ModRamIO.f90:1293-1320) -- a known-answer vector
                       for the energy ladder of ARRAYS (src/ModRamInit.f90:428-459).
gcoul_kat.txt        : the reference's own unit test test_Gcoul
                       (src/ModRamFunctions.f90:470-511): x, Gcoul(x) pairs, tolerance 1e-8.
lz_mlt_test1.txt     : L and MLT columns of output/test1/pressure.ref (pins LZ(2:NR), MLT).
"""
import os
import re

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    with open(os.path.join(REF, "output/test1/dsbnd.ref")) as f:
        lines = f.read().splitlines()[1:]
    with open(os.path.join(HERE, "ekev_test1_dsbnd.txt"), "w") as f:
        for ln in lines:
            f.write(ln.split()[0] + "\n")
    src = open(os.path.join(REF, "src/ModRamFunctions.f90")).read()
    blk = src[src.index("subroutine test_Gcoul"):src.index("end subroutine test_Gcoul")]
    m = re.search(r"expect\s*=\s*\(/(.*?)/\)", blk, re.S)
    vals = [v.strip() for v in m.group(1).replace("&", " ").replace("\n", " ").split(",")]
    with open(os.path.join(HERE, "gcoul_kat.txt"), "w") as f:
        for x, v in zip((0.5, 1.0, 1.5), vals):      # Gcoul(REAL(ii)/2.d0), ii = 1..3
            f.write(f"{x} {v}\n")
    rows = []
    with open(os.path.join(REF, "output/test1/pressure.ref")) as f:
        for ln in f.read().splitlines()[2:]:
            p = ln.split()
            rows.append((p[0], p[1]))
    with open(os.path.join(HERE, "lz_mlt_test1.txt"), "w") as f:
        for a, b in rows:
            f.write(f"{a} {b}\n")


if __name__ == "__main__":
    main()
