"""Single-GPU timings at the grid sizes configs[4] names (8 x the default RAM grid, 4 x the default SCB grid): informational."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from ramscb_b200 import grids, host, scb_synthetic, synthetic  # noqa: E402

out = {}
g = grids.build_grids(NR=80, NT=97, NE=70, energy_refine=2)
inp = synthetic.make_inputs(g, f2_kind="noisy", inductive=True)
D = synthetic.synthetic_daa(g, inp)
gpu = host.RamGpu(g, mode=host.MODE_FAST)
gpu.set_inputs(inp)
gpu.set_diffcoef(1, D)
gpu.set_diffcoef(2, D)
flush = torch.empty(512 * 1024 * 1024 // 4, dtype=torch.float32, device="cuda")
for _ in range(3):
    gpu.ram_run(5.0, DtsMin=1.0, flags=5)
ms = 0.0
for _ in range(10):
    flush.zero_()
    torch.cuda.synchronize()
    gpu.timer_begin()
    gpu.ram_run(5.0, DtsMin=1.0, flags=5)
    ms += gpu.timer_end()
cells = g.nS * g.NR * g.NT * g.NE * g.NPA
out["ram_8x_grid_flags5"] = {"cells": cells, "ms_per_step": ms / 10, "cell_updates_per_s": 13.0 * cells / (ms / 10 * 1e-3)}
gpu.close()
sinp = scb_synthetic.build_scb(nthe=201, npsi=89, nzeta=97, warp=0.2)
sg = host.ScbGpu(sinp)
sg.computeBandJacob(); sg.metrica(); sg.newk()
alfa0 = sg.get_field("alfa")
best = None
for _ in range(3):
    sg.set_field("alfa", alfa0)
    r = sg.iterateAlpha(1e-6, ordering=host.SOR_COLOR4)
    best = r["ms"] if best is None else min(best, r["ms"])
out["scb_4x_grid_iterate_alpha"] = {"ms": best, "max_sweeps": int(r["nisave"]), "cluster": sg.last_cluster(), "SORFail": int(r["SORFail"])}
sg.metric(); sg.newj()
psi0 = sg.get_field("psi")
best = None
for _ in range(3):
    sg.set_field("psi", psi0)
    r = sg.iteratePsi(1e-6, ordering=host.SOR_COLOR4)
    best = r["ms"] if best is None else min(best, r["ms"])
out["scb_4x_grid_iterate_psi"] = {"ms": best, "max_sweeps": int(r["nisave"]), "cluster": sg.last_cluster(), "SORFail": int(r["SORFail"])}
print(json.dumps(out))
