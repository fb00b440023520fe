"""One coupled RAM <-> SCB cycle on one GPU at the default grids (configs[0] / [4] physics composed from the device entry
points): RAM steps for 300 s of simulated time between two SCB updates (src/ModRamScbRun.f90; DTs <= 5 s following the returned CFL limit), then pressure ->
scb_run -> computehI -> new fields into the RAM state.  Wall clock per phase; informational."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from ramscb_b200 import grids, host, scb_synthetic, synthetic  # noqa: E402

g = grids.build_grids()
inp = synthetic.make_inputs(g, f2_kind="smooth", inductive=True)
sinp = scb_synthetic.build_scb(nthe=101, npsi=45, nzeta=97, warp=0.2)
ram = host.RamGpu(g, mode=host.MODE_FAST)
ram.set_inputs(inp)
sg = host.ScbGpu(sinp)
sg.set_map_targets(sinp.alphaVal, sinp.psiVal, sinp.chiVal)
LZ, PHI = g.LZ[:g.NR + 1], g.PHI[:g.NT]
bnormal = 0.31 / 6.6 ** 3 * 1.0e5               # src/ModScbInit.f90:246-273 (the SCB field is in units of bnormal)
hi = host.HiGpu(101, 45, 97, LZ, g.MLT[:g.NT], g.MU, g.PA, g.PAbn, sinp.chiVal, 51, bnormal)
hi.set_ram_fields({n: getattr(inp, n) for n in host.HI_RAM_NAMES})
flags_scb = np.array([1, 1, 1, 0][:g.nS], dtype=np.int32)
res = []
for cycle in range(3):
    t = {}
    t0 = time.perf_counter()
    sim, dts, nstep = 0.0, 5.0, 0
    while sim < 300.0 and nstep < 400:             # the host's step control: DTs follows the CFL limit the step returns
        out = ram.ram_run(dts, DtsMin=1.0, flags=0)
        sim += dts
        nstep += 1
        dts = min(5.0, out["DtsNext"], 300.0 - sim) if sim < 300.0 else dts
    torch.cuda.synchronize()
    t["ram_steps"], t["ram_last_DtsNext"] = nstep, float(out["DtsNext"])
    t["ram_300s_ms"] = (time.perf_counter() - t0) * 1e3
    scale = 12.0 / float(out["PPERT"].max())          # synthetic F2 is not ring-current sized: one scale (tests/test_zz_late_additions_gpu.py)
    t0 = time.perf_counter()
    sg.set_ram_pressure(out["PPERT"] * scale, out["PPART"] * scale, flags_scb, LZ, PHI)
    t["pressure_handover_ms"] = (time.perf_counter() - t0) * 1e3
    t0 = time.perf_counter()
    r = sg.scb_run(None, ordering=host.SOR_COLOR4)
    t["scb_run_ms"] = (time.perf_counter() - t0) * 1e3
    t["scb_outer_iterations"], t["SORFail"] = int(r["iterations"]), int(r["SORFail"])
    t0 = time.perf_counter()
    err = hi.computehI(sg, 300.0, True)
    hi.push_to_ram(ram)
    t["computehI_and_field_handover_ms"] = (time.perf_counter() - t0) * 1e3
    t["gslerr"] = int(err)
    t["new_fields_finite"] = bool(all(np.all(np.isfinite(hi.get(n)[1:])) for n in ("FNHS", "FNIS", "BOUNHS", "HDNS")))
    t["pressure_finite"] = bool(np.all(np.isfinite(out["PPERT"])))
    t["cycle_ms"] = t["ram_300s_ms"] + t["pressure_handover_ms"] + t["scb_run_ms"] + t["computehI_and_field_handover_ms"]
    res.append(t)
F = ram.f2_d2h()
print(json.dumps({"cycles": res, "F2_finite": bool(np.all(np.isfinite(F))), "F2_min": float(F.min())}))
