"""Small-grid pass over the hot kernels for compute-sanitizer (memcheck / racecheck): fused FAST steps with every flag
set, the pipelined host step, an in-process 3-rank sharded step, the SCB cluster SOR and the resident computehI.
    compute-sanitizer --tool memcheck  python scripts/sanitize_small.py
    compute-sanitizer --tool racecheck python scripts/sanitize_small.py ram
"""
import os
import sys

os.environ.setdefault("CUDA_MODULE_LOADING", "EAGER")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

from ramscb_b200 import grids, host, scb_synthetic, synthetic  # noqa: E402

what = sys.argv[1] if len(sys.argv) > 1 else "all"
g = grids.build_grids(NR=11, NT=13, NE=35)
inp = synthetic.make_inputs(g, f2_kind="noisy", inductive=True, mgnp=True)
D = synthetic.synthetic_daa(g, inp)
if what != "scb":
    gpu = host.RamGpu(g, mode=host.MODE_FAST)
    gpu.set_inputs(inp)
    gpu.set_diffcoef(1, D)
    gpu.set_diffcoef(2, D)
    for flags in (0, 5, 7):
        gpu.ram_run(5.0, DtsMin=1.0, flags=flags, T=5.0)
    F = inp.F2.copy(order="F")
    gpu.ram_run_host(F, 5.0, DtsMin=1.0, flags=5)
    print("ram fused + host pipeline ok", float(np.nanmax(F)))
    gpu.close()
if what in ("all", "shard"):
    ranks = [host.RamGpu(g, mode=host.MODE_FAST) for _ in range(3)]
    for r, q in enumerate(ranks):
        q.set_inputs(inp)
        q.set_diffcoef(1, D)
        q.set_diffcoef(2, D)
    for r, q in enumerate(ranks):
        q.peer_attach_local(r, ranks, host.SHARD_SLABS)
    for q in ranks:
        q.run_sharded_enqueue(5.0, flags=5)
    outs = [q.run_sharded_collect() for q in ranks]
    print("sharded (3 in-process ranks) ok", outs[0]["DtsNext"])
    for q in ranks:
        q.close()
if what in ("all", "scb"):
    sinp = scb_synthetic.build_scb(nthe=51, npsi=23, nzeta=49, warp=0.3)
    sg = host.ScbGpu(sinp)
    sg.computeBandJacob(); sg.metrica(); sg.newk()
    r = sg.iterateAlpha(1e-6, ordering=host.SOR_COLOR4)
    sg.metric(); sg.newj()
    r2 = sg.iteratePsi(1e-6, ordering=host.SOR_COLOR4)
    print("scb cluster SOR ok", r["nisave"], r2["nisave"], "cluster", sg.last_cluster() if hasattr(sg, "last_cluster") else "")
    Lz = g.LZ[:g.NR + 1]
    rng = np.random.default_rng(1)
    shape3 = (g.NR + 1, g.NT, g.NPA)
    ram = {n: np.asfortranarray(rng.random(shape3)) for n in ("FNHS", "FNIS", "BOUNHS", "BOUNIS", "HDNS")}
    ram["BNES"] = np.asfortranarray(1e-7 * rng.random((g.NR + 1, g.NT)))
    hi = host.HiGpu(51, 23, 49, Lz, g.MLT[:g.NT], g.MU, g.PA, g.PAbn, sinp.chiVal, 26, 1.0)
    hi.set_ram_fields(ram)
    hi.computehI(sg, 300.0, True)
    print("resident computehI ok", float(np.nanmax(hi.get("FNHS"))))
    hi.close(); sg.close()
print("SANITIZE_SCRIPT_DONE")
