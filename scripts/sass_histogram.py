"""SASS mnemonic histogram of the hot kernels in ramscb_b200/lib/libramscb_gpu.so (cuobjdump -sass), written to
profiles/r2/sass_histogram.txt.  What it is for: showing which memory path each kernel uses (UBLKCP / SYNCS = TMA bulk
copies completing on an mbarrier, LDGSTS = cp.async, LDG/STG = plain), how much of the instruction stream is FP64
arithmetic (DFMA / DADD / DMUL / DSETP), and that nothing spills (LDL / STL).

    python scripts/sass_histogram.py [kernel-regex ...]
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "ramscb_b200", "lib", "libramscb_gpu.so")
WANT = sys.argv[1:] or [r"k_plane_rp", r"k_col_fused", r"k_anisch_pa_fast", r"k_finalize", r"k_wpadif_tables", r"k_coulmu_tables",
                        r"k_scb_sor_cluster_reg", r"k_scb_map_w", r"k_hi_nn9", r"k_hi_lines", r"k_peer_barrier", r"k_diffcoef"]
GROUPS = [("TMA bulk copy (UBLKCP)", r"^UBLKCP"), ("mbarrier (SYNCS)", r"^SYNCS"), ("cp.async (LDGSTS)", r"^LDGSTS"),
          ("global load (LDG)", r"^LDG"), ("global store (STG)", r"^STG"), ("shared load (LDS)", r"^LDS"), ("shared store (STS)", r"^STS"),
          ("local = spill (LDL/STL)", r"^(LDL|STL)"), ("FP64 FMA (DFMA)", r"^DFMA"), ("FP64 add/mul (DADD/DMUL)", r"^(DADD|DMUL)"),
          ("FP64 compare/select (DSETP/FSEL)", r"^(DSETP|FSEL)"), ("MUFU (rcp/sqrt/ex2 seeds)", r"^MUFU"), ("barrier (BAR/UCGABAR)", r"^(BAR|UCGABAR)"),
          ("shuffle (SHFL)", r"^SHFL"), ("atomics / RED", r"^(ATOM|RED|ATOMG)"), ("cluster / DSMEM (ST.*CLUSTER via STS/MAPA)", r"^MAPA"),
          ("system-scope fence / flag (MEMBAR, LD/ST .SYS)", r"^(MEMBAR|ERRBAR|CCTL)")]


def main():
    txt = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    out = ["SASS mnemonic histogram, sm_100a, " + os.path.relpath(LIB, ROOT), ""]
    cur, hist = None, {}
    for line in txt.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            hist[cur] = collections.Counter()
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and cur:
            hist[cur][m.group(1)] += 1
    for fn in sorted(hist):
        dem = subprocess.run(["c++filt", fn], capture_output=True, text=True).stdout.strip().split("(")[0]
        if not any(re.search(w, dem) for w in WANT):
            continue
        h = hist[fn]
        total = sum(h.values())
        out.append(f"{dem}   [{total} instructions]")
        for label, rx in GROUPS:
            n = sum(v for k, v in h.items() if re.match(rx, k))
            if n:
                out.append(f"    {label:52s} {n:6d}  {100.0 * n / total:5.1f} %")
        out.append("")
    p = os.path.join(ROOT, "profiles", "r2", "sass_histogram.txt")
    open(p, "w").write("\n".join(out) + "\n")
    print(p, len(out), "lines")


if __name__ == "__main__":
    main()
