"""Where does FAST mode leave the 1e-12 per-cell band?  Operator-by-operator lock-step of the FAST kernels (one kernel per
operator; the fused step is bit-identical to it) against the CPU oracle over three ram_run steps of the noisy default-grid
input.  After every operator:
  * strict per-cell relative error, binned by |F2_ref| / max|F2_ref| of the species;
  * clamp flips: cells that one side clamped to 1e-15 (src/ModRamDrift.f90:187-190 and the like) and the other did not;
  * the "tainted" set: cells in the domain of dependence of a flip (the flip cell dilated by the 2-cell limiter stencil of
    every later sweep along its direction), and the worst strict error OUTSIDE that set.
Runs on the GPU, or on the host-CPU emulator of the test-suite with RSG_EMU=1 (same kernel bodies).
    RSG_EMU=1 python scripts/fast_parity_study.py [--steps 3] [--out PARITY_data.json]
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
if os.environ.get("RSG_EMU") == "1":
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import conftest
    conftest.use_emulator()

from oracle import oracle  # noqa: E402
from ramscb_b200 import grids, host, synthetic  # noqa: E402

AXIS = {"driftr": 0, "driftp": 1, "drifte": 2, "driftmu": 3}      # axis of F2[s] (NR,NT,NE,NPA) a sweep couples


def dilate(mask, axis, r=2, periodic=False):
    out = mask.copy()
    for sft in range(1, r + 1):
        for sg in (+1, -1):
            m = np.roll(mask, sg * sft, axis=axis)
            if not periodic:
                sl = [slice(None)] * mask.ndim
                sl[axis] = slice(0, sft) if sg > 0 else slice(-sft, None)
                m[tuple(sl)] = False
            out |= m
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--out", default=os.path.join(ROOT, "profiles", "r2", "fast_parity_study.json"))
    ap.add_argument("--small", action="store_true", help="reduced grid (quick look)")
    a = ap.parse_args()
    g = grids.build_grids(NR=12, NT=13, NE=15) if a.small else grids.build_grids()
    inp = synthetic.make_inputs(g, f2_kind="noisy", inductive=True, mgnp=True)
    oracle.build()
    o = oracle.RamOracle(g, inp, DTs=5.0)
    of = oracle.RamOracle(g, inp, DTs=5.0, variant="fma")     # the reference's arithmetic with FMA contraction allowed
    gpu = host.RamGpu(g, mode=host.MODE_FAST)
    gpu.set_inputs(inp)
    tainted = np.zeros(inp.F2.shape, dtype=bool)
    log = []
    edges = [0.0, 1e-60, 1e-40, 1e-30, 1e-25, 1e-20, 1e-15, 1e-10, 1e-5, 1.0001]

    def compare(step, S, name):
        got, ref = gpu.f2_d2h()[S - 1], o.F2[S - 1]
        strict = np.abs(got - ref) / np.maximum(np.abs(ref), 1e-300)
        flip = (got == 1e-15) != (ref == 1e-15)
        if name in AXIS:
            tainted[S - 1] = dilate(tainted[S - 1], AXIS[name], 2, periodic=(name == "driftp"))
        tainted[S - 1] |= flip
        clean = ~tainted[S - 1]
        relmag = np.abs(ref) / np.abs(ref).max()
        hist = []
        for lo, hi in zip(edges[:-1], edges[1:]):
            m = (relmag >= lo) & (relmag < hi)
            hist.append({"rel_magnitude": [lo, hi], "cells": int(m.sum()), "over_1e-12": int((strict[m] > 1e-12).sum()),
                         "over_1e-12_untainted": int((strict[m & clean] > 1e-12).sum()),
                         "max_strict": float(strict[m].max()) if m.any() else 0.0})
        # the reference against itself: strict (-ffp-contract=off) vs contracted (-ffp-contract=fast -mfma) build of the oracle
        sf = np.abs(of.F2[S - 1] - ref) / np.maximum(np.abs(ref), 1e-300)
        both = (strict > 1e-12) & (sf > 1e-12)
        rec_self = {"oracle_fma_cells_over_1e-12": int((sf > 1e-12).sum()), "oracle_fma_max_strict": float(sf.max()),
                    "fast_over_and_oracle_fma_over": int(both.sum()),
                    "oracle_fma_flips": int(((of.F2[S - 1] == 1e-15) != (ref == 1e-15)).sum())}
        rec = {"step": step, "S": S, "op": name, "new_flips": int(flip.sum()), **rec_self, "tainted": int(tainted[S - 1].sum()),
               "cells_over_1e-12": int((strict > 1e-12).sum()), "cells_over_1e-12_untainted": int((strict[clean] > 1e-12).sum()),
               "max_strict": float(strict.max()), "max_strict_untainted": float(strict[clean].max()),
               "max_abs_diff_tainted": float(np.abs(got - ref)[tainted[S - 1]].max()) if tainted[S - 1].any() else 0.0,
               "max_ref_tainted_over_max": float(relmag[tainted[S - 1]].max()) if tainted[S - 1].any() else 0.0, "hist": hist}
        log.append(rec)
        print(f"step {step} S={S} {name:12s} flips {rec['new_flips']:5d} tainted {rec['tainted']:7d}  >1e-12: {rec['cells_over_1e-12']:6d} "
              f"(untainted {rec['cells_over_1e-12_untainted']:4d})  max strict {rec['max_strict']:.2e} untainted {rec['max_strict_untainted']:.2e} "
              f"| oracle-vs-oracle(fma): >1e-12 {rec['oracle_fma_cells_over_1e-12']:6d} max {rec['oracle_fma_max_strict']:.2e} flips {rec['oracle_fma_flips']}", flush=True)

    for step in range(a.steps):
        dts = [5.0, 7.5, 20.0][step % 3]
        o.set_scalar("DTs", dts)
        of.set_scalar("DTs", dts)
        for S in range(1, g.nS + 1):
            el = g.kind[S - 1] == 3
            o.op("cepara", S); o.op("driftpara", S)
            of.op("cepara", S); of.op("driftpara", S)
            gpu.CEPARA(S, dts); gpu.DRIFTPARA(S, dts)
            seq = ["driftr", "driftp", "drifte", "driftmu", "wavelo" if el else "charexchange", "atmol", "atmol",
                   "wavelo" if el else "charexchange", "driftmu", "drifte", "driftp", "driftr"]
            for name in seq:
                o.op(name, S)
                of.op(name, S)
                if name == "wavelo":
                    gpu.WAVELO(S, dts)
                else:
                    getattr(gpu, name.upper())(S)
                compare(step, S, name)
        # epilogue of ram_run: both sides take the library's / oracle's own (F2(J=NT) = F2(J=1), 1e-31 outside the magnetopause,
        # ANISCH's F2(L=1) = F2(L=2)); re-seat both on the same post-epilogue rule by applying it to the two arrays in numpy
        F_o, F_g = o.F2, gpu.f2_d2h()
        for F in (F_o, F_g, of.F2):
            F[:, :, -1] = F[:, :, 0]
            F[:, inp.outsideMGNP == 1] = 1e-31
            F[:, 1:, :, 1:, 0] = F[:, 1:, :, 1:, 1]
        tainted[:, :, -1] = tainted[:, :, 0]
        tainted[:, 1:, :, 1:, 0] = tainted[:, 1:, :, 1:, 1]
        gpu.f2_h2d(F_g)
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    with open(a.out, "w") as f:
        json.dump({"grid": [g.nS, g.NR, g.NT, g.NE, g.NPA], "input": "noisy, inductive, mgnp", "records": log}, f)
    print("wrote", a.out)


if __name__ == "__main__":
    main()
