"""GPU parity tests of the components written after the round's GPU minutes were spent (developed against the
host-CPU emulator, tests/emu): PARA_FLC on the device, rsg_scb_run (the whole SCB outer iteration in one call)
and iterateAlpha sharded along zeta.  Same bars and helpers as tests/test_ram_parity_gpu.py /
tests/test_scb_parity_gpu.py; kept in a file that sorts last so that `pytest -x` reaches them after every
hardware-proven test."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import test_ram_parity_gpu as TR          # noqa: E402
import test_scb_parity_gpu as TS          # noqa: E402
from ramscb_b200 import scb_synthetic as SCBSYN          # noqa: E402
from ramscb_b200 import synthetic          # noqa: E402

# first hardware run of these entry points: a device-side loop that never ends must end the run loudly, not hang the
# box (pytest-timeout's thread method interrupts a blocked C call by exiting the process)
pytestmark = [pytest.mark.gpu, pytest.mark.timeout(600, method="thread")]

_mk, _pair_ram, _relerr, DTS = TR._mk, TR._pair, TR._relerr, TR.DTS
_pair, _same, SMALL = TS._pair, TS._same, TS.SMALL


def flc_radius_inputs(g, S=1, seed=11):
    """Synthetic outputs of FLC_Radius (src/ModRamLoss.f90:176-340): curvature radius of the order of the gyroradius
    of the hot tail (epsilon spans < 0.1, the fitted range and the 0.584 clamp) and zeta parameters of O(1)."""
    rng = np.random.default_rng(seed)
    rc = np.asfortranarray(6.4e6 * (0.02 + 0.5 * rng.random((g.NR, g.NT))) * (g.LZ[:g.NR, None] / 6.5) ** 2
                           * np.sqrt(g.RMAS[S - 1] / g.RMAS[0]))       # gyroradius at fixed energy ~ sqrt(mass)
    z1 = np.asfortranarray(2.0 + 3.0 * rng.random((g.NR, g.NT)))
    z2 = np.asfortranarray(2.0 + 3.0 * rng.random((g.NR, g.NT)))
    return rc, z1, z2


@pytest.mark.parametrize("S", [1, 2, 4])
def test_para_flc_on_device(default_grids, oracle_built, S):
    """PARA_FLC (src/ModRamLoss.f90:342-455) on the device from the (NR,NT) outputs of FLC_Radius: FLC_coef within
    1e-12 relative of the oracle (exp / pow / sin / acos of the device library vs libm), zero pattern identical
    (epsilon < 0.1, L = NPA); then FLCscatter with the device-built coefficient against the oracle chain."""
    g = default_grids
    inp = _mk(g, f2_kind="noisy")
    o, gpu = _pair_ram(g, inp, oracle_built)
    rc, z1, z2 = flc_radius_inputs(g, S)
    for n, a in (("r_curvEq", rc), ("zeta1Eq", z1), ("zeta2Eq", z2)):
        o.set_array(n, a)
    o.op("para_flc", S)
    gpu.PARA_FLC(S, rc, z1, z2)
    ref, got = o.arr["FLC_coef"], gpu.get_flc_coef(S)
    assert np.isfinite(ref).all() and (ref >= 0).all() and (ref[..., -1] == 0).all()
    assert np.array_equal(got == 0, ref == 0)
    nz = ref != 0
    assert 0.05 < nz.mean() < 0.95                                   # both branches of the epsilon gate are exercised
    assert np.max(np.abs(got[nz] - ref[nz]) / np.abs(ref[nz])) <= 1e-12
    o.set_scalar("T", 900.0)
    nv_ref = o.op("flcscatter", S)
    nv = gpu.FLCscatter(S, DTS, 900.0)
    a, b = gpu.f2_d2h()[S - 1], o.F2[S - 1]
    scale = np.maximum(np.max(np.abs(b), axis=3, keepdims=True), 1e-300)      # per pitch-angle line: the solve couples all L
    assert np.max(np.abs(a - b) / scale) <= 1e-11, _relerr(a, b)
    assert nv == nv_ref


def test_scb_run_outer_iterations_resident(oracle_built, numit=3, color4=True):
    """scb_run (src/ModScbRun.f90:149-440) in ONE C call with every 3-D array resident (rsg_scb_run; only the
    2-D pressure front end is a host callback) against the same loop composed from the oracle's routines:
    with RSG_SOR_LEX the iteration counts, blends, residuals and x, y, z, alfa, psi are BIT-IDENTICAL; the
    norms of Compute_convergence agree to 1e-12 (tree vs serial sums).  A second run with the production
    4-colour ordering must take the same number of outer iterations and land within 1e-6 of the same points
    (each SOR solve stops at InCon = 1e-6; the orderings share the fixed point, not the iterates)."""
    from ramscb_b200 import host
    inp, o, gpu = _pair(oracle_built, **SMALL)
    fn = SCBSYN.equatorial_pressure_fn()
    kw = dict(numit=numit, MinSCBIterations=2, decreaseConvAlpha=1e-30, decreaseConvPsi=1e-30)     # exactly `numit` outer iterations
    gpu.set_map_targets(inp.alphaVal, inp.psiVal, inp.chiVal)
    ro = o.scb_run(fn, **kw)
    rg = gpu.scb_run(fn, ordering=host.SOR_LEX, **kw)
    assert ro["SORFail"] == 0 and rg["SORFail"] == 0
    assert rg["iterations"] == ro["iterations"] == numit
    for k in ("blendAlpha", "blendPsi", "errorAlpha", "errorPsi", "nisaveAlpha", "nisavePsi", "blendRetries"):
        assert rg[k] == ro[k], (k, rg[k], ro[k])
    for k in ("sumdbAlpha", "sumdbPsi"):                       # tree sum vs serial sum
        assert abs(rg[k] - ro[k]) <= 1e-12 * abs(ro[k]), (k, rg[k], ro[k])
    _same(gpu, o, ("x", "y", "z", "alfa", "psi", "jacobian", "bsq", "pper", "sigma"))
    for a, b in zip((rg["normDiff"], rg["normJxB"], rg["normGradP"]), ro["norm"]):
        assert abs(a - b) <= 1e-12 * abs(b)
    for a, b in zip((rg["normDiffStart"], rg["normJxBStart"], rg["normGradPStart"]), ro["normStart"]):
        assert abs(a - b) <= 1e-12 * abs(b)
    # production ordering from the same start
    if color4:
        g2 = host.ScbGpu(inp)
        g2.set_map_targets(inp.alphaVal, inp.psiVal, inp.chiVal)
        r2 = g2.scb_run(fn, ordering=host.SOR_COLOR4, **kw)
        assert r2["SORFail"] == 0 and r2["iterations"] == numit
        for n in ("x", "y", "z"):
            a, b = g2.get_field(n), getattr(o, n)
            assert np.max(np.abs(a - b)) <= 1e-6 * np.max(np.abs(b)), (n, np.max(np.abs(a - b)))
    # failure path: a callback that reports failure aborts the call; one that returns NaN pressures makes the
    # solve fail (SORFail) and x, y, z, alfa, psi come back as they were at entry (:397-413)
    g3 = host.ScbGpu(inp)
    g3.set_map_targets(inp.alphaVal, inp.psiVal, inp.chiVal)
    with pytest.raises(ZeroDivisionError):
        g3.scb_run(lambda xe, ye: 1 / 0, **kw)

    def bad(xe, ye):
        a, b = fn(xe, ye)
        a[3, 5] = np.nan
        return a, b

    r3 = g3.scb_run(bad, ordering=host.SOR_LEX, **kw)
    assert r3["SORFail"] == 1
    for n in ("x", "y", "z", "alfa", "psi"):
        assert np.array_equal(g3.get_field(n), getattr(inp, n)), n


@pytest.mark.parametrize("p0,outcome", [(2000.0, "retry"), (20000.0, "fail")])
def test_scb_run_damped_retries(oracle_built, p0, outcome):
    """The Move_points loops of scb_run (src/ModScbRun.f90:232-262, 418-440) when the re-gridded points give a
    negative Jacobian: revert x, y, z, damp the blend, try again -- and give up (SORFail, everything restored)
    when that does not help.  A pressure 1000x / 10000x the nominal one with blendInitial = 1 drives the psi half
    into that path; device and oracle must take the same decisions and end bit-identical."""
    from ramscb_b200 import host
    inp, o, gpu = _pair(oracle_built, **SMALL)
    fn = SCBSYN.equatorial_pressure_fn(p0_nPa=p0)
    kw = dict(numit=2, MinSCBIterations=2, decreaseConvAlpha=1e-30, decreaseConvPsi=1e-30, blendInitial=1.0, damp=0.5)
    gpu.set_map_targets(inp.alphaVal, inp.psiVal, inp.chiVal)
    ro = o.scb_run(fn, **kw)
    rg = gpu.scb_run(fn, ordering=host.SOR_LEX, **kw)
    assert ro["blendRetries"] >= 1
    assert ro["SORFail"] == (1 if outcome == "fail" else 0)
    for k in ("SORFail", "iterations", "blendRetries", "blendAlpha", "blendPsi", "errorAlpha", "errorPsi"):
        assert rg[k] == ro[k], (k, rg[k], ro[k])
    _same(gpu, o, ("x", "y", "z", "alfa", "psi"))
    if outcome == "fail":
        for n in ("x", "y", "z", "alfa", "psi"):
            assert np.array_equal(gpu.get_field(n), getattr(inp, n)), n


@pytest.mark.parametrize("slabs", [1, 3])
def test_zeta_sharded_alpha_slabs_on_one_device(slabs):
    """SURVEY 8(e), iterateAlpha sharded along zeta (rsg_scb_zsolve_*): the ranks' kernels and the halo /
    all-reduce protocol of parallel.ScbZetaSharded, stepped in lock-step for `slabs` handles that share
    ONE device (plane copies and the MAX of the state vectors done with torch on the library's memory).
    alfa, ni, diffmx, sumb, sumdb bit-identical to the one-GPU RSG_SOR_COLOR4 solve.  The same protocol
    over a process group: tests/test_parallel_cpu.py (gloo), tests/multi_gpu_scb_check.py (NCCL)."""
    import os
    import torch
    from ramscb_b200 import host, parallel
    emu = os.environ.get("RSG_EMU") == "1" or not host.LIB_PATH.endswith("libramscb_gpu.so")
    inp = SCBSYN.build_scb(nthe=51, npsi=23, nzeta=50, warp=0.3)       # 49 relaxed planes: slabs of 17 + 16 + 16
    ref = host.ScbGpu(inp)
    ref.computeBandJacob(); ref.metrica(); ref.newk()
    r1 = ref.iterateAlpha(1e-7, ordering=host.SOR_COLOR4)
    assert r1["SORFail"] == 0 and r1["nisave"] > 20
    import contextlib
    tst = None if emu else torch.cuda.Stream()                        # not the legacy default stream (the library's own stream does not wait for it)
    with (contextlib.nullcontext() if emu else torch.cuda.stream(tst)):
        gs, al, st, rng = [], [], [], []
        for r in range(slabs):
            g = host.ScbGpu(inp)
            if not emu:
                g.set_stream(tst.cuda_stream)                             # torch's copies and the kernels: one stream
            g.computeBandJacob(); g.metrica(); g.newk()
            k0, nk = parallel._split(inp.nzeta - 1, slabs, r)
            g.zsolve_begin(1e-7, k0 + 1, nk)
            ptr, n = g.field_device("alfa")
            al.append(parallel._dev_tensor(ptr, n, not emu).view(inp.nzeta + 1, -1))
            ptr, n = g.zsolve_state_device()
            st.append(parallel._dev_tensor(ptr, n, not emu))
            gs.append(g); rng.append((k0 + 1, k0 + nk))
        for sweep in range(5001):
            for parity in (0, 1):
                for g in gs:
                    g.zsolve_half(parity)
                for r in range(slabs - 1):                                # edge planes of this parity cross the cut
                    hi, lo = rng[r][1], rng[r + 1][0]
                    if hi % 2 == parity:
                        al[r + 1][hi].copy_(al[r][hi])
                    if lo % 2 == parity:
                        al[r][lo].copy_(al[r + 1][lo])
            m = st[0].clone()
            for t in st[1:]:
                m = torch.maximum(m, t)
            for t, g in zip(st, gs):
                t.copy_(m)
                g.zsolve_commit()
            if sweep % 8 == 7 and all(g.zsolve_pending() == 0 for g in gs):
                break
        for r, (a, b) in enumerate(rng):                                  # "all-gather" of the relaxed planes
            for q in range(slabs):
                if q != r:
                    al[q][a:b + 1].copy_(al[r][a:b + 1])
        for g in gs:
            res = g.iterate_finish(True)
            assert np.array_equal(g.get_field("alfa"), ref.get_field("alfa"))
            assert np.array_equal(res["ni"], r1["ni"]) and res["nisave"] == r1["nisave"]
            assert res["diffmx"] == r1["diffmx"] and res["sumb"] == r1["sumb"] and res["sumdb"] == r1["sumdb"]
            assert res["SORFail"] == 0


@pytest.mark.parametrize("nthe,wiggle,outside", [(101, 0.0, 0.0), (101, 0.05, 0.05), (51, 0.15, 0.1)])
def test_hI_integrals_on_device(default_grids, oracle_built, nthe, wiggle, outside):
    """computehI's integral block (src/ModRamScb.f90:372-410: length, r0, the equatorial-B fix-up, bfMirror,
    GSL_Integration_hI + GSL_BounceAverage, I_cart / H_cart / HDens_cart / bZEq_Cart) through rsg_hI_integrals:
    bit-identical to the oracle on dipole lines and on lines with non-monotonic B (mirror-search fall-backs,
    short spans) and skipped (outsideMGNP) lines; HDens_cart keeps its value on skipped lines."""
    from ramscb_b200 import host
    g = default_grids
    d = SCBSYN.ram_field_lines(g.LZ[1:g.NR + 1] if len(g.LZ) > g.NR else g.LZ, g.MLT[:g.NT], nthe=nthe, wiggle=wiggle,
                               outside_fraction=outside, seed=5)
    D0 = np.full((g.NR, g.NT, g.NPA), -7.0, order="F")
    ref = oracle_built.hi_integrals(mu=g.MU, HDens_cart=D0, **d)
    out = host.hI_integrals(mu=g.MU, HDens_cart=D0, **d)
    for name, a, b in zip(("I_cart", "H_cart", "HDens_cart", "bZEq_Cart"), ref, out):
        assert np.all(np.isfinite(b)), name
        assert np.array_equal(a, b), (name, float(np.max(np.abs(a - b))))
    I, H, D, bz = out[:4]
    skipped = d["outsideMGNP"] != 0
    assert np.all(D[skipped] == -7.0) and np.all(I[skipped] == 0.0) and np.all(bz[skipped] == 0.0)
    live = ~skipped
    assert np.all(H[live][:, 1:] > 0.0) and np.all(I[live][:, 1:] > 0.0) and np.all(D[live] > 0.0)
    if wiggle == 0.0:
        # dipole: h runs from 0.74 (90 deg) to 1.38 (0 deg); the pitch angles whose mirror points span <= 4 nodes copy
        # their neighbour (src/RamGSL.c:583-587), so the first values sit near 0.74 within the error of the coarse linear table (h is not monotonic in L: B is a linear table, and the
        # reference repairs that afterwards, src/ModRamScb.f90:506-514); I grows towards the loss cone
        assert np.all(H[:, :, 1:] > 0.70) and np.all(H < 1.39) and np.all(H[:, :, 1] < 0.80)
        assert np.all(np.diff(I[:, :, 8:], axis=2) >= 0.0)       # beyond the pitch angles that copy their neighbour
    assert out[4] >= 0.0


def hi_tail_inputs(g, seed, scale_cols=True, negatives=False, nans=False):
    """Integral-block output on perturbed dipole lines plus the ingredients of the tail: an outer SCB boundary (ScaleAt)
    on part of the MLT sectors with some points outside the magnetopause, previous RAM variables, optionally negative
    and NaN entries (the repairs of src/ModRamScb.f90:489-508 and :625-637)."""
    rng = np.random.default_rng(seed)
    d = SCBSYN.ram_field_lines(g.LZ[1:g.NR + 1] if len(g.LZ) > g.NR else g.LZ, g.MLT[:g.NT], nthe=41, wiggle=0.05, seed=seed)
    ScaleAt = np.zeros(g.NT, dtype=np.int32)
    if scale_cols:
        for j in range(g.NT):
            if rng.random() < 0.5:
                ScaleAt[j] = rng.integers(max(3, g.NR - 6), g.NR + 1)
                for i in range(ScaleAt[j], g.NR):        # the point the scaling is anchored on stays inside
                    if rng.random() < 0.3:
                        d["outsideMGNP"][i, j] = 1
    return d, ScaleAt, rng


@pytest.mark.parametrize("smooth,DthI,variant", [(1, 300.0, "scaled"), (0, 0.0, "plain"), (1, 300.0, "repairs")])
def test_hI_tail_on_device(default_grids, oracle_built, smooth, DthI, variant):
    """The rest of computehI (src/ModRamScb.f90:413-637) through rsg_hI_tail: every output bit-identical to the oracle
    -- with an outer SCB boundary and points outside the magnetopause, without (and DthI = 0), and with negative / NaN /
    non-monotonic entries that trigger the repairs."""
    from ramscb_b200 import host
    g = default_grids
    d, ScaleAt, rng = hi_tail_inputs(g, seed=21, scale_cols=variant != "plain")
    I, H, D, bz, _ = oracle_built.hi_integrals(mu=g.MU, **d)
    D = np.where(np.isfinite(D), D, 1.0)
    if variant == "repairs":
        for a in (I, H, D):
            m = rng.random(a.shape) < 0.01
            a[m] = -a[m]                                     # negatives (:489-508)
            m = rng.random(a.shape) < 0.01
            a[m] = 3.0 * a[m]                                # too large relative to L+1 (:509-517)
        H[rng.random(H.shape) < 0.002] = np.nan              # NaN repair (:625-637)
    shape3 = (g.NR + 1, g.NT, g.NPA)
    ram = {n: np.asfortranarray(rng.random(shape3)) for n in ("FNHS", "FNIS", "BOUNHS", "BOUNIS", "HDNS")}
    ram["BNES"] = np.asfortranarray(1e-7 * rng.random((g.NR + 1, g.NT)))
    Lz = g.LZ[:g.NR + 1] if len(g.LZ) > g.NR else np.append(2 * g.LZ[0] - g.LZ[1], g.LZ)
    args = (I, H, D, bz, ScaleAt, d["outsideMGNP"], Lz, g.PA, g.PAbn, smooth, DthI, ram)
    ref = oracle_built.hi_tail(*args)
    out = host.hI_tail(*args)
    assert ref["gslerr"] == 0 and out["gslerr"] == 0
    for n in ("I_cart", "H_cart", "HDens_cart", "bZEq_cart", "FNHS", "FNIS", "BOUNHS", "BOUNIS", "HDNS", "BNES", "dIdt", "dHdt",
              "dIbndt", "dBdt"):
        assert np.array_equal(ref[n], out[n], equal_nan=True), (n, float(np.nanmax(np.abs(ref[n] - out[n]))))
    assert np.array_equal(out["FNHS"][0], out["FNHS"][1], equal_nan=True) and np.all(out["dIdt"][0] == 0.0)
    if variant != "repairs":
        assert all(np.all(np.isfinite(out[n])) for n in ("FNHS", "FNIS", "BOUNHS", "BOUNIS", "HDNS", "BNES", "dIdt", "dBdt"))
        assert np.all(np.diff(out["FNIS"][1:], axis=2) >= 0.0) if not smooth else True
    if DthI == 0.0:
        assert not out["dIdt"].any() and not out["dHdt"].any() and not out["dIbndt"].any() and not out["dBdt"].any()


@pytest.mark.parametrize("dims", [(41, 25, 33, 10, 12), (101, 45, 97, 20, 25)])
def test_hI_convert_lines_on_device(oracle_built, dims):
    """computehI's field-line conversion (src/ModRamScb.f90:252-300: winding-number test, psiRAM, then x, y, z, bf of every
    node through Interpolation_2D_NN_point = nine MINLOC passes + inverse-distance-squared weights) through
    rsg_hI_convert_lines: bit-identical to the oracle, inside and outside the SCB domain (RAM shells beyond the last SCB
    surface), incl. a RAM point that coincides with an SCB node (the d <= 1e-9 branch of NN_Interpolation_2D)."""
    from ramscb_b200 import host
    nthe, npsi, nzeta, nR, nT = dims
    inp = SCBSYN.build_scb(nthe=nthe, npsi=npsi, nzeta=nzeta, warp=0.2)
    ke = nthe // 2 + 1
    r = np.sqrt(inp.x ** 2 + inp.y ** 2 + inp.z ** 2)
    bf = np.asfortranarray(30574.0 / r ** 3 * np.sqrt(1.0 + 3.0 * (inp.z / r) ** 2))
    Lz = np.linspace(1.75, 8.5, nR + 1)                       # the SCB domain ends at 7.5: the outer shells are outside
    MLT = np.linspace(0.0, 24.0, nT)
    # one RAM point exactly on an SCB equatorial node
    jn, kn = npsi // 2, nzeta // 3
    xe, ye = inp.x[ke - 1, jn, kn], inp.y[ke - 1, jn, kn]
    Lz[3] = np.hypot(xe, ye)
    MLT[2] = ((np.arctan2(ye, xe) + np.pi) * 24.0 / (2.0 * np.pi)) % 24.0
    args = (inp.x, inp.y, inp.z, bf, inp.psi, inp.alfa, Lz, MLT, ke)
    ref = oracle_built.hi_convert_lines(*args)
    out = host.hI_convert_lines(*args)
    assert np.array_equal(ref[4], out[4]) and 0 < out[4].sum() < out[4].size
    for name, a, b in zip(("xRAM", "yRAM", "zRAM", "bRAM"), ref, out):
        assert np.all(np.isfinite(b)), name
        assert np.array_equal(a, b), (name, float(np.max(np.abs(a - b))))
    inside = out[4] == 0
    # the interpolated equatorial foot points land near the RAM points they were asked for
    xo = Lz[1:, None] * np.cos(MLT[None, :] * 2 * np.pi / 24 - np.pi)
    yo = Lz[1:, None] * np.sin(MLT[None, :] * 2 * np.pi / 24 - np.pi)
    err = np.hypot(out[0][ke - 1] - xo, out[1][ke - 1] - yo)[inside]
    # (nine unnormalised nearest neighbours in (psi, alfa): 0.04 RE median at the default SCB grid, coarser grids are worse)
    assert np.all(out[0][:, ~inside] == 0.0) and (nthe < 101 or np.median(err) < 0.1)


@pytest.mark.parametrize("dims", [(41, 25, 33), (101, 45, 97)])
def test_computehI_composed_on_device(default_grids, oracle_built, dims):
    """computehI end to end (conversion -> ScaleAt / outsideMGNP on the host -> integrals -> tail) through
    host.computehI against the same composition of the oracle's three restatements: every RAM variable and time derivative
    bit-identical, with RAM shells beyond the SCB domain (no tracer: flagged outsideMGNP like the 'SWMF' branch)."""
    from ramscb_b200 import host
    g = default_grids
    nthe, npsi, nzeta = dims
    inp = SCBSYN.build_scb(nthe=nthe, npsi=npsi, nzeta=nzeta, warp=0.2)
    r = np.sqrt(inp.x ** 2 + inp.y ** 2 + inp.z ** 2)
    scb = dict(x=inp.x, y=inp.y, z=inp.z, psi=inp.psi, alfa=inp.alfa, chiVal=inp.chiVal, nThetaEquator=nthe // 2 + 1, bnormal=1.0,
               bf=np.asfortranarray(30574.0 / r ** 3 * np.sqrt(1.0 + 3.0 * (inp.z / r) ** 2)))
    Lz = np.linspace(1.75, 8.0, g.NR + 1)                    # the last shells lie outside the SCB domain (7.5)
    rng = np.random.default_rng(4)
    shape3 = (g.NR + 1, g.NT, g.NPA)
    ram = {n: np.asfortranarray(rng.random(shape3)) for n in ("FNHS", "FNIS", "BOUNHS", "BOUNIS", "HDNS")}
    ram["BNES"] = np.asfortranarray(1e-7 * rng.random((g.NR + 1, g.NT)))
    dens = lambda dist: 10 ** (13.326 - 3.6908 * dist + 1.1362 * dist ** 2 - 0.16984 * dist ** 3 + 0.009553 * dist ** 4)   # RAIRDEN
    args = (scb, Lz, g.MLT[:g.NT], g.MU, g.PA, g.PAbn, ram, 300.0)
    out = host.computehI(*args, integral_smooth=True, density_fn=dens)
    ref = host.computehI(*args, integral_smooth=True, density_fn=dens,
                         _impl=(oracle_built.hi_convert_lines, oracle_built.hi_integrals, oracle_built.hi_tail))
    assert out["ScaleAt"].any() and np.array_equal(out["ScaleAt"], ref["ScaleAt"]) and np.array_equal(out["outsideMGNP"], ref["outsideMGNP"])
    for n in ("xRAM", "bRAM", "FNHS", "FNIS", "BOUNHS", "BOUNIS", "HDNS", "BNES", "dIdt", "dHdt", "dIbndt", "dBdt"):
        assert np.array_equal(ref[n], out[n], equal_nan=True), (n, float(np.nanmax(np.abs(ref[n] - out[n]))))
    inside = out["outsideMGNP"] == 0
    assert np.all(np.isfinite(out["FNHS"][1:][inside])) and np.all(out["FNHS"][1:][inside][:, 4:] > 0.0)


@pytest.mark.gpu
@pytest.mark.parametrize("dims", [(51, 25, 41)])
def test_computehI_resident_handle(default_grids, oracle_built, dims):
    """rsg_hi (the resident computehI): same kernels as the three stateless calls with nothing going through the host in
    between.  (a) host-driven middle (ScaleAt / outsideMGNP / density from the host) bit-identical to host.computehI and so
    to the oracle's composition; (b) a second call continues from the persisted HDens_cart and RAM variables exactly as
    the stateless chain fed with its own outputs does; (c) the all-device defaults ('SWMF' boundary branch, RAIRDEN on the
    device): flags identical, fields within the last bits of pow(); (d) the new field arrays device-to-device into a RamGpu
    give the same ram_run as set_fields from the host."""
    from ramscb_b200 import host
    g = default_grids
    nthe, npsi, nzeta = dims
    inp = SCBSYN.build_scb(nthe=nthe, npsi=npsi, nzeta=nzeta, warp=0.2)
    r = np.sqrt(inp.x ** 2 + inp.y ** 2 + inp.z ** 2)
    scb = dict(x=inp.x, y=inp.y, z=inp.z, psi=inp.psi, alfa=inp.alfa, chiVal=inp.chiVal, nThetaEquator=nthe // 2 + 1, bnormal=1.0,
               bf=np.asfortranarray(30574.0 / r ** 3 * np.sqrt(1.0 + 3.0 * (inp.z / r) ** 2)))
    Lz = np.linspace(1.75, 8.0, g.NR + 1)
    rng = np.random.default_rng(4)
    shape3 = (g.NR + 1, g.NT, g.NPA)
    ram = {n: np.asfortranarray(rng.random(shape3)) for n in ("FNHS", "FNIS", "BOUNHS", "BOUNIS", "HDNS")}
    ram["BNES"] = np.asfortranarray(1e-7 * rng.random((g.NR + 1, g.NT)))

    def dens(d):        # RAIRDEN with the Fortran's integer powers (x**3 = x*x*x, x**4 = (x*x)*(x*x))
        d2 = d * d
        return 10.0 ** (13.326 - 3.6908 * d + 1.1362 * d2 - 0.16984 * (d2 * d) + 0.009553 * (d2 * d2))

    MLT = g.MLT[:g.NT]
    args = (scb, Lz, MLT, g.MU, g.PA, g.PAbn)
    ref1 = host.computehI(*args, ram, 300.0, integral_smooth=True, density_fn=dens)
    ram2 = {n: ref1[n] for n in host.HI_RAM_NAMES}
    hi = host.HiGpu(nthe, npsi, nzeta, Lz, MLT, g.MU, g.PA, g.PAbn, inp.chiVal, nthe // 2 + 1, 1.0)
    hi.set_ram_fields(ram)
    # (a) host-driven middle
    outside, nout = hi.convert(scb)
    assert nout > 0 and np.array_equal(outside, ref1["outsideSCB"])
    xR, yR, zR = hi.get("xRAM"), hi.get("yRAM"), hi.get("zRAM")
    assert np.array_equal(xR, ref1["xRAM"]) and np.array_equal(hi.get("bRAM"), ref1["bRAM"])
    err = hi.finish(300.0, True, ScaleAt=ref1["ScaleAt"], outsideMGNP=ref1["outsideMGNP"],
                    density=dens(np.sqrt(xR ** 2 + yR ** 2 + zR ** 2)))
    assert err == ref1["gslerr"]
    got = hi.results()
    for n in ("FNHS", "FNIS", "BOUNHS", "BOUNIS", "HDNS", "BNES", "dIdt", "dHdt", "dIbndt", "dBdt", "I_cart", "H_cart", "HDens_cart", "bZEq_cart"):
        assert np.array_equal(ref1[n], got[n], equal_nan=True), n
    # (b) second call: state persisted on the device
    warped = dict(scb)
    warped["bf"] = np.asfortranarray(scb["bf"] * 1.01)
    hi.convert(warped, want_outside=False)
    xR, yR, zR = hi.get("xRAM"), hi.get("yRAM"), hi.get("zRAM")
    hi.finish(60.0, True, ScaleAt=ref1["ScaleAt"], outsideMGNP=ref1["outsideMGNP"], density=dens(np.sqrt(xR ** 2 + yR ** 2 + zR ** 2)))
    xR2, yR2, zR2, bR2, _o = host.hI_convert_lines(warped["x"], warped["y"], warped["z"], warped["bf"], warped["psi"], warped["alfa"], Lz, MLT,
                                                   nthe // 2 + 1)[:5]
    I, H, D, bz = host.hI_integrals(inp.chiVal, g.MU, xR2, yR2, zR2, bR2, dens(np.sqrt(xR2 ** 2 + yR2 ** 2 + zR2 ** 2)), ref1["outsideMGNP"],
                                    nthe // 2 + 1, 1.0, HDens_cart=ref1["HDens_cart"])[:4]
    ref2 = host.hI_tail(I, H, D, bz, ref1["ScaleAt"], ref1["outsideMGNP"], Lz, g.PA, g.PAbn, True, 60.0, ram2)
    got2 = hi.results()
    for n in ("FNHS", "FNIS", "BOUNHS", "BOUNIS", "HDNS", "BNES", "dIdt", "dHdt", "dIbndt", "dBdt", "HDens_cart"):
        assert np.array_equal(ref2[n], got2[n], equal_nan=True), n
    # (c) everything on the device in one call
    hi2 = host.HiGpu(nthe, npsi, nzeta, Lz, MLT, g.MU, g.PA, g.PAbn, inp.chiVal, nthe // 2 + 1, 1.0)
    hi2.set_ram_fields(ram)
    n0 = hi2.launch_count()
    assert hi2.computehI(scb, 300.0, True) == ref1["gslerr"]
    assert hi2.launch_count() - n0 == 9          # nn9 x2, ScaleAt, RAIRDEN, lines, tail x4
    got3 = hi2.results()
    assert np.array_equal(got3["ScaleAt"], ref1["ScaleAt"]) and np.array_equal(got3["outsideMGNP"], ref1["outsideMGNP"])
    for n in ("FNHS", "FNIS", "BOUNHS", "BOUNIS", "BNES", "dIdt", "dIbndt", "dBdt"):
        assert np.array_equal(ref1[n], got3[n], equal_nan=True), n
    for n in ("HDNS", "dHdt"):                   # the bounce-averaged density sees pow() of the device
        assert np.allclose(ref1[n], got3[n], rtol=1e-13, atol=0, equal_nan=True), n
    # (d) straight into the RAM state
    inp_r = synthetic.make_inputs(g, f2_kind="smooth", inductive=True)
    outs = []
    for dev_path in (False, True):
        gpu = host.RamGpu(g)
        gpu.set_inputs(inp_r)
        if dev_path:
            hi2.push_to_ram(gpu)
        else:
            L = host.lib()
            a = [np.asfortranarray(got3[n]) for n in ("BNES", "dBdt", "FNHS", "FNIS", "BOUNHS", "BOUNIS", "HDNS", "dIdt", "dIbndt")]
            om = np.asfortranarray(got3["outsideMGNP"], dtype=np.int32)
            host._ck(L.rsg_ram_set_fields(gpu.h, *[host._p(x) for x in a], om.ctypes.data))
        gpu.ram_run(5.0, DtsMin=1.0, flags=0)
        outs.append(gpu.f2_d2h())
        gpu.close()
    assert np.array_equal(outs[0], outs[1], equal_nan=True)
    hi.close(); hi2.close()


@pytest.mark.gpu
def test_coupled_ram_scb_cycle_stays_on_the_device(default_grids, oracle_built, dims=(51, 23, 49)):
    """BASELINE configs[4] in miniature -- the coupled cycle ram_run -> pressure -> scb_run -> computehI -> ram_run
    (src/ModRamScbRun.f90:60-130) composed from the device entry points so that no 3-D or 4-D array crosses PCIe between
    the two hot paths: the (nS,NR,NT) pressures of ram_run go to rsg_scb_set_ram_pressure, rsg_scb_run runs with the front
    end on the device, rsg_computehI reads the SCB handle's arrays in place, rsg_ram_set_fields_device takes the new field
    arrays.  Against the oracle's composition of the same cycle: SCB decisions identical and fields <= 1e-9 (the two
    pressures differ in the last bits of the RAM moments), computehI's RAM variables <= 1e-7 of the oracle chain fed with
    the oracle's own SCB state, and the second ram_run -- both sides on the DEVICE-built fields -- to the EXACT-mode bars."""
    from ramscb_b200 import host
    g = default_grids
    nthe, npsi, nzeta = dims
    inp = synthetic.make_inputs(g, f2_kind="smooth", inductive=True)
    sinp = SCBSYN.build_scb(nthe=nthe, npsi=npsi, nzeta=nzeta, warp=0.3)
    ram, o = host.RamGpu(g, mode=host.MODE_EXACT), oracle_built.RamOracle(g, inp, DTs=5.0)
    ram.set_inputs(inp)
    out = ram.ram_run(5.0, DtsMin=1.0, flags=0)
    o.ram_run(flags=0)
    assert np.allclose(out["PPERT"], o.PPERT, rtol=1e-12, atol=0)
    # pressures of the synthetic F2 are not ring-current sized: one scale on both sides (12 keV/cm^3 peak, output/test1)
    scale = 12.0 / float(out["PPERT"].max())
    flags_scb = np.array([1, 1, 1, 0][:g.nS], dtype=np.int32)
    LZ, PHI = g.LZ[:g.NR + 1], g.PHI[:g.NT]
    sg, so = host.ScbGpu(sinp), oracle_built.ScbOracle(sinp)
    sg.set_ram_pressure(out["PPERT"] * scale, out["PPART"] * scale, flags_scb, LZ, PHI)
    so.pressure_raw(np.asfortranarray(o.PPERT * scale), np.asfortranarray(o.PPART * scale), flags_scb, LZ, PHI)
    sg.set_map_targets(sinp.alphaVal, sinp.psiVal, sinp.chiVal)
    kw = dict(numit=2, MinSCBIterations=2, decreaseConvAlpha=1e-30, decreaseConvPsi=1e-30)
    ro = so.scb_run(None, **kw)
    rg = sg.scb_run(None, ordering=host.SOR_LEX, **kw)
    assert ro["SORFail"] == 0 and rg["SORFail"] == 0 and rg["iterations"] == ro["iterations"] == 2
    assert rg["nisaveAlpha"] == ro["nisaveAlpha"] and rg["nisavePsi"] == ro["nisavePsi"]
    for n in ("x", "y", "z", "alfa", "psi", "bf"):
        a, b = sg.get_field(n), getattr(so, n)
        assert np.max(np.abs(a - b)) <= 1e-9 * np.max(np.abs(b)), n
    # computehI on the SCB handle's device arrays, new fields device-to-device into the RAM state
    ieq = nthe // 2 + 1
    bnormal = so.get("bnormal")
    prev = {n: getattr(inp, n) for n in host.HI_RAM_NAMES}
    hi = host.HiGpu(nthe, npsi, nzeta, LZ, g.MLT[:g.NT], g.MU, g.PA, g.PAbn, sinp.chiVal, ieq, bnormal)
    hi.set_ram_fields(prev)
    assert hi.computehI(sg, 300.0, True) == 0
    hi.push_to_ram(ram)
    new = hi.results()
    assert not new["outsideSCB"].any()                     # the RAM domain (L <= 6.5) lies inside this SCB domain

    def dens(d):
        d2 = d * d
        return 10.0 ** (13.326 - 3.6908 * d + 1.1362 * d2 - 0.16984 * (d2 * d) + 0.009553 * (d2 * d2))

    scb_o = dict(x=so.x, y=so.y, z=so.z, bf=so.bf, psi=so.psi, alfa=so.alfa, chiVal=sinp.chiVal, nThetaEquator=ieq, bnormal=bnormal)
    ref = host.computehI(scb_o, LZ, g.MLT[:g.NT], g.MU, g.PA, g.PAbn, prev, 300.0, integral_smooth=True, density_fn=dens,
                         _impl=(oracle_built.hi_convert_lines, oracle_built.hi_integrals, oracle_built.hi_tail))
    for n in ("FNHS", "FNIS", "BOUNHS", "BOUNIS", "HDNS", "BNES"):
        a, b = new[n], ref[n]
        ok = np.isfinite(b)
        assert np.array_equal(np.isfinite(a), ok), n
        assert np.max(np.abs(a[ok] - b[ok])) <= 1e-7 * np.max(np.abs(b[ok])), (n, float(np.max(np.abs(a[ok] - b[ok]))))
    # second RAM step: the device on the arrays it was handed on the device, the oracle on copies of the same arrays
    for n in ("BNES", "dBdt", "FNHS", "FNIS", "BOUNHS", "BOUNIS", "HDNS", "dIdt", "dIbndt"):
        o.set_array(n, new[n])
    o.set_array("outsideMGNP", new["outsideMGNP"])
    out2 = ram.ram_run(5.0, DtsMin=1.0, flags=0)
    dtn = o.ram_run(flags=0)
    assert abs(out2["DtsNext"] - dtn) <= 1e-13 * dtn
    F = ram.f2_d2h()
    assert np.all(np.isfinite(F))
    assert np.max(np.abs(F - o.F2) / np.maximum(np.abs(o.F2), 1e-300)) <= 1e-12
    assert np.allclose(out2["PPERT"], o.PPERT, rtol=1e-12, atol=0)
    hi.close(); sg.close(); ram.close()


def test_anisch_diffcoef_rebuild_feeds_the_wpi_step(default_grids, oracle_built):
    """ANISCH's diffusion-coefficient rebuild (src/ModRamRun.f90:422-605) on the device, then a WPI + EMIC
    ram_run that never sees a host-built coefficient array.  Coefficients <= 1e-13 of the oracle's (device log10 / pow vs
    libm; the kernels keep the reference's operation order and are bit-identical in the emulator), zero pattern and the
    1e-31 floor identical; the step with them within the bars of the WPADIF tests."""
    from ramscb_b200 import host
    g = default_grids
    inp = synthetic.make_inputs(g, f2_kind="noisy", inductive=True)
    t = synthetic.synthetic_wave_tables(g, inp)
    o = oracle_built.RamOracle(g, inp, DTs=5.0)
    gpu = host.RamGpu(g, mode=host.MODE_EXACT)
    gpu.set_inputs(inp)
    gpu.set_wave_tables(t)
    for S, fl in ((4, host.F_WPI), (1, host.F_EMIC)):
        assert o.anisch_diffcoef(S, fl, t, AE=350) == 0
        assert gpu.ANISCH_diffcoef(S, fl, t["XNE"], AE=350) == 0
    for which, name in enumerate(("ATAW", "ATAC", "ATAW_emic_h", "ATAW_emic_he")):
        a, b = gpu.get_diffcoef(which), getattr(o, name)
        assert np.array_equal(a == 0, b == 0) and np.array_equal(a == 1e-31, b == 1e-31), name
        assert np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300)) <= 1e-13, name
        o.set_array(name, a)     # the step below compares the kernels, not the last bit of pow / log10: same coefficients on both sides
    dtn = o.ram_run(flags=5)
    out = gpu.ram_run(5.0, DtsMin=1.0, flags=5)
    got = gpu.f2_d2h()
    strict = np.abs(got - o.F2) / np.maximum(np.abs(o.F2), 1e-300)
    assert strict.max() <= 1e-12, (strict.max(), int((strict > 1e-12).sum()))          # EXACT mode, bar of test_full_ram_run
    assert out["DtsNext"] == dtn
    gpu.close()


def test_flc_radius_and_pressure_front_end_on_hardware(default_grids, oracle_built):
    """The two device paths written in round 2 for the RAM <-> SCB coupling, on the configs[3] grid: FLC_Radius (R12) and the 2-D
    front end of `pressure` (8(f)-2).  Same kernels are bit-identical to the oracle in the emulator (tests/test_emu_cpu.py); on
    hardware asin / sqrt / division of the device library round like libm except in the last bit of asin: <= 1e-13."""
    import test_emu_cpu as TE
    from ramscb_b200 import host
    g = default_grids
    sinp = SCBSYN.build_scb(nthe=101, npsi=45, nzeta=97, warp=0.2)
    o, gpu = oracle_built.ScbOracle(sinp), host.ScbGpu(sinp)
    o.bandjacob(); gpu.computeBandJacob()
    radRaw = 1.75 + (6.75 - 1.75) * np.arange(1, g.NR + 1) / g.NR
    azimRaw = 24.0 * np.arange(g.NT) / (g.NT - 1)
    ref = oracle_built.flc_radius(o.x, o.y, o.z, o.Bx, o.By, o.Bz, radRaw, azimRaw, (sinp.nthe + 1) // 2, o.get("bnormal"))
    for a, b in zip(gpu.FLC_Radius(radRaw, azimRaw), ref):
        assert np.array_equal(a, b)                      # +, -, *, /, sqrt only: correctly rounded on the device too
    PPerT, PParT, scb, LZ, PHI = TE._ram_pressures(g)
    r = o.pressure_raw(PPerT, PParT, scb, LZ, PHI)
    gpu.set_ram_pressure(PPerT, PParT, scb, LZ, PHI)
    for a, b in zip(r, gpu.get_ram_pressure()):
        assert np.array_equal(a, b)
    pe, pa = o.pressure_front()
    ge, ga = gpu.pressure_front()
    assert np.max(np.abs(ge - pe) / pe) <= 1e-13 and np.max(np.abs(ga - pa) / pa) <= 1e-13
    # scb_run without a host callback: same decisions as the oracle's composition, fields to 1e-10
    gpu.set_map_targets(sinp.alphaVal, sinp.psiVal, sinp.chiVal)
    kw = dict(numit=2, MinSCBIterations=2, decreaseConvAlpha=1e-30, decreaseConvPsi=1e-30)
    ro = o.scb_run(None, **kw)
    n0 = gpu.launch_count()
    rg = gpu.scb_run(None, ordering=host.SOR_LEX, **kw)
    assert ro["SORFail"] == 0 and rg["SORFail"] == 0 and rg["iterations"] == ro["iterations"] == 2
    assert rg["nisaveAlpha"] == ro["nisaveAlpha"] and rg["nisavePsi"] == ro["nisavePsi"]
    for n in ("x", "y", "z", "alfa", "psi"):
        a, b = gpu.get_field(n), getattr(o, n)
        assert np.max(np.abs(a - b)) <= 1e-10 * np.max(np.abs(b)), n
    gpu.close()
