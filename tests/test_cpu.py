"""CPU-only tests (run with -m "not gpu"): the oracle against the reference's golden
vectors, the host logic (grids / synthetic inputs / slot bookkeeping), properties of
the restated operators, and that the C-ABI library loads and exports every symbol
include/*.h declares (no compute call without a GPU)."""
import ctypes
import os
import re
import sys

import numpy as np
import pytest

from ramscb_b200 import grids, synthetic

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")


# ---- golden vectors held by the reference's own tests -------------------------------
def test_energy_ladder_matches_dsbnd_ref(default_grids):
    """EKEV(2..35) printed by write_dsbnd into output/test1/dsbnd.ref (17 digits)."""
    gold = np.loadtxt(os.path.join(GOLD, "ekev_test1_dsbnd.txt"))
    assert gold.shape == (34,)
    assert np.array_equal(default_grids.EKEV[1:], gold)


def test_lz_mlt_match_pressure_ref(default_grids):
    g = default_grids
    rows = np.loadtxt(os.path.join(GOLD, "lz_mlt_test1.txt"))
    assert rows.shape == ((g.NR - 1) * g.NT, 2)
    assert np.allclose(np.unique(rows[:, 0]), g.LZ[1:g.NR], rtol=0, atol=1e-12)
    assert np.allclose(np.unique(rows[:, 1]), g.MLT, rtol=0, atol=1e-12)


def test_gcoul_known_answers(oracle_built):
    """The reference's unit test test_Gcoul (src/ModRamFunctions.f90:470-511): tol 1e-8, Gcoul(0)=NaN."""
    lib = oracle_built.ram_lib()
    for x, expect in np.loadtxt(os.path.join(GOLD, "gcoul_kat.txt")):
        assert abs(lib.orc_gcoul(x) - expect) <= 1e-8
        assert abs(grids.gcoul(x) - expect) <= 1e-8
    assert np.isnan(lib.orc_gcoul(0.0))


def test_funt_funi_limits(oracle_built):
    lib = oracle_built.ram_lib()
    # Ejiri: I(90 deg) = 0, h(0 deg) = alpha, I(0 deg) = 2 alpha
    alpha = 1.0 + np.log(2.0 + np.sqrt(3.0)) / 2.0 / np.sqrt(3.0)
    assert lib.orc_funi(0.0) == 0.0
    assert abs(lib.orc_funt(1.0) - alpha) < 1e-15 and abs(lib.orc_funi(1.0) - 2 * alpha) < 1e-15
    for m in (0.1, 0.5, 0.9):
        assert lib.orc_funt(m) == grids.funt(m) and lib.orc_funi(m) == grids.funi(m)


# ---- host logic -------------------------------------------------------------------------
def test_pitch_angle_grid(default_grids):
    g = default_grids
    assert g.PA[0] == 90.0 and g.PA[-1] == 0.0 and g.MU[0] == 0.0 and g.MU[-1] == 1.0
    assert np.all(np.diff(g.MU) > 0) and np.all(g.WMU > 0) and np.all(g.DMU > 0)
    # loss-cone index of SURVEY appendix E
    assert list(g.UPA.astype(int)) == [47, 49, 51] + list(range(52, 69))
    # comment at src/ModRamRun.f90:303: khi = 6,10,25,30,35 <-> 0.4, 1, 39, 129, 325 keV
    assert np.allclose(g.EKEV[[5, 9, 24, 29]], [0.4, 1.0, 39.0, 129.0], rtol=0.08)


def test_scaled_grid_is_consistent():
    g = grids.build_grids(NR=80, NT=49, NE=70, energy_refine=2)
    g0 = grids.build_grids()
    assert np.array_equal(g.MU, g0.MU) and np.array_equal(g.WMU, g0.WMU)
    assert abs(g.EBND[-1] - g0.EBND[-1]) < 1e-9 * g0.EBND[-1]          # same energy span
    assert np.all(np.diff(g.UPA) >= 0) and g.UPA[0] == 47 and g.UPA[-1] == 68
    assert list(g.khi) == [12, 20, 50, 60, 70]


def test_synthetic_inputs_are_periodic_and_positive(default_grids):
    g = default_grids
    inp = synthetic.make_inputs(g, f2_kind="adversarial", inductive=True, efield_ind=True, mgnp=True)
    assert np.all(inp.F2 > 0) and np.all(inp.FNHS > 0) and np.all(inp.BOUNIS > 0) and np.all(inp.BNES > 0)
    assert np.array_equal(inp.F2[:, :, 0], inp.F2[:, :, -1])
    assert np.array_equal(inp.VT[:, 0], inp.VT[:, -1])
    assert inp.outsideMGNP.sum() == 4
    for a in (inp.F2, inp.FGEOS, inp.FNHS, inp.BNES):
        assert a.flags.f_contiguous


# ---- properties of the restated operators (the reference has no per-operator vectors) -----
@pytest.fixture(scope="module")
def small():
    g = grids.build_grids()
    inp = synthetic.make_inputs(g, f2_kind="smooth")
    return g, inp


def test_driftp_conserves_particles(oracle_built, small):
    """Periodic azimuthal advection in flux form: sum over J=2..NT of F2 is conserved."""
    g, inp = small
    o = oracle_built.RamOracle(g, inp, DTs=5.0)
    o.op("driftpara", 1); o.op("driftp", 1)
    t0 = inp.F2[0][1:, 1:].sum(axis=1)
    t1 = o.F2[0][1:, 1:].sum(axis=1)
    assert np.max(np.abs(t1 - t0) / t0) < 1e-12
    assert np.array_equal(o.F2[0][:, 0], o.F2[0][:, -1])


def test_drifts_keep_positivity_and_ghost_shell(oracle_built, small):
    g, inp = small
    o = oracle_built.RamOracle(g, inp, DTs=5.0)
    for S in (1, 4):
        o.op("driftpara", S)
        for op in ("driftr", "driftp", "drifte", "driftmu"):
            o.op(op, S)
    assert np.all(o.F2 > 0)
    assert np.array_equal(o.F2[:, 0], inp.F2[:, 0])          # F2(S,1,...) is never modified by a drift
    assert np.all(o.DtDriftR[[0, 3]] > 1) and np.all(o.DtDriftMu[[0, 3]] < 1e4)


def test_zero_timestep_is_identity(oracle_built, small):
    g, inp = small
    o = oracle_built.RamOracle(g, inp, DTs=0.0)
    o.op("driftpara", 2)
    for op in ("driftr", "driftp", "drifte"):
        o.op(op, 2)
    f = o.F2[1]
    assert np.array_equal(f[:, 1:], inp.F2[1][:, 1:])        # all Courant numbers vanish


def test_wpadif_conserves_and_smooths(oracle_built, small):
    """Implicit pitch-angle diffusion: positivity, and the mu-integral of f*dmu changes only
    through the boundary (here: strictly smaller anisotropy)."""
    g, inp = small
    o = oracle_built.RamOracle(g, inp, DTs=5.0)
    D = synthetic.synthetic_daa(g, inp) * 50.0
    o.set_array("ATAC", D)
    before = o.F2[3].copy()
    nv = o.op("wpadif", 4)
    after = o.F2[3]
    assert nv == 0
    assert np.all(after[1:, :, 1:, 1:] >= 0)
    r0 = before[5, 3, 10, 5:40] / (inp.FNHS[5, 3, 5:40] * g.MU[5:40])
    r1 = after[5, 3, 10, 5:40] / (inp.FNHS[5, 3, 5:40] * g.MU[5:40])
    assert np.ptp(r1) < np.ptp(r0)


def test_loss_operators_only_decay(oracle_built, small):
    g, inp = small
    o = oracle_built.RamOracle(g, inp, DTs=5.0)
    for S in (1, 2, 3):
        o.op("cepara", S); o.op("charexchange", S); o.op("atmol", S)
    o.op("cepara", 4); o.op("wavelo", 4); o.op("atmol", 4)
    assert np.all(o.F2 <= inp.F2) and np.all(o.F2 > 0)
    assert np.all(o.CHARGE[:3, 1:, :, 1:, 1:] < 1.0) and np.all(o.CHARGE[3] == 1.0)


@pytest.mark.parametrize("variant", ["noisy_fields", "adversarial", "carry_over"])
def test_independent_numpy_restatement_of_the_drifts(oracle_built, variant):
    """tests/independent_ram.py restates DRIFTPARA/DRIFTR/P/E/MU a second time, array-at-a-time in
    numpy straight from the Fortran text.  The C++ oracle (scalar loops) must agree with it bit for
    bit -- F2 and the CFL limit -- on inductive fields with an E field and magnetopause flags, on the
    adversarial distribution, and on an input that triggers DRIFTR's ghost-cell carry-over."""
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import independent_ram as ind
    g = grids.build_grids(NR=12, NT=9, NE=16)
    if variant == "noisy_fields":
        inp = synthetic.make_inputs(g, f2_kind="noisy", inductive=True, efield_ind=True, mgnp=True)
    elif variant == "adversarial":
        inp = synthetic.make_inputs(g, f2_kind="adversarial", inductive=True, efield_ind=True, mgnp=True)
    else:
        inp = synthetic.make_inputs(g, f2_kind="noisy", efield_ind=True)
        inp.EIP[g.NR - 1, :] = 6e-4 * np.cos(g.PHI)       # reverse the radial drift between the last two shells
        inp.EIP[g.NR, :] = -6e-4 * np.cos(g.PHI)
    beta = 1.5
    for S in (1, 4):
        for name, fn, dtn in (("driftr", ind.driftr, "DtDriftR"), ("driftp", ind.driftp, "DtDriftP"),
                              ("drifte", ind.drifte, "DtDriftE"), ("driftmu", ind.driftmu, "DtDriftMu")):
            if variant == "carry_over" and name != "driftr":
                continue
            o = oracle_built.RamOracle(g, inp, DTs=5.0)
            o.op("driftpara", S)
            o.op(name, S)
            if variant == "carry_over":
                c = o.cdrift(S, 0)
                assert np.sum((c[g.NR - 1] >= 0) & (c[g.NR - 2] < 0)) > 0, "input does not trigger the carry-over"
            new, dt = fn(g, inp, inp.F2, S, 5.0, beta)
            assert np.array_equal(new, o.F2[S - 1]), (variant, S, name, float(np.abs(new - o.F2[S - 1]).max()))
            assert dt == getattr(o, dtn)[S - 1], (variant, S, name)
            assert not np.array_equal(new, inp.F2[S - 1])


def test_independent_numpy_restatement_of_losses_and_moments(oracle_built):
    """Same cross-check for CEPARA (CHARGE, ATLOS), CHAREXCHANGE, ATMOL, SUMRC (reference summation
    order), the ANISCH pressures (+ the F2(L=1)=F2(L=2) side effect), WPADIF, WAVELO and the Coulomb
    operators (COULPARA tables, COULEN, COULMU)."""
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import independent_ram as ind
    g = grids.build_grids(NR=10, NT=9, NE=35)
    inp = synthetic.make_inputs(g, f2_kind="noisy", inductive=True, mgnp=True)
    for S in (1, 2, 3, 4):
        o = oracle_built.RamOracle(g, inp, DTs=5.0)
        o.op("cepara", S)
        CH, AT = ind.cepara(g, inp, S, 5.0)
        assert np.array_equal(CH, o.CHARGE[S - 1]) and np.array_equal(AT[1:, 1:], o.ATLOS[S - 1][1:, 1:]), S
        o.op("charexchange", S)
        F1 = ind.charexchange(inp.F2[S - 1], CH)
        assert np.array_equal(F1, o.F2[S - 1]), S
        o.op("atmol", S)
        F2 = ind.atmol(g, inp, F1, AT)
        assert np.array_equal(F2, o.F2[S - 1]) and not np.array_equal(F2, F1), S
        o.op("sumrc", S)
        assert ind.sumrc(g, F2) == o.SETRC[S - 1], S
        o.op("anisch", S)
        pe, pa, F3 = ind.anisch_pressures(g, inp, F2, S)
        assert np.array_equal(pe[1:], o.PPERT[S - 1][1:]) and np.array_equal(pa[1:], o.PPART[S - 1][1:]), S
        assert np.array_equal(F3, o.F2[S - 1]), S
    D = synthetic.synthetic_daa(g, inp) * 30
    o = oracle_built.RamOracle(g, inp, DTs=5.0)
    o.set_array("ATAC", D)
    o.op("wpadif", 4)
    W = ind.wpadif(g, inp, inp.F2[3], np.zeros_like(D), D, 5.0)
    assert np.array_equal(W, o.F2[3]) and not np.array_equal(W, inp.F2[3])
    for kp in (2.0, 5.0):                                    # WAVELO: both wave amplitudes
        inpk = synthetic.make_inputs(g, f2_kind="noisy", Kp=kp)
        o = oracle_built.RamOracle(g, inpk, DTs=5.0)
        o.op("wavelo", 4)
        W = ind.wavelo(g, inpk, inpk.F2[3], 5.0)
        assert np.array_equal(W, o.F2[3]) and not np.array_equal(W, inpk.F2[3])
    for S in (1, 2, 4):                                      # COULPARA tables, COULEN, COULMU
        o = oracle_built.RamOracle(g, inp, DTs=5.0)
        o.set_scalar("T", 50.0)
        o.op("coulpara", S)
        CE, CI, AT, GT = ind.coulpara(g, S, 5.0, grids.gcoul)
        assert np.array_equal(CE, o.COULE[S - 1]) and np.array_equal(CI, o.COULI[S - 1]), S
        assert np.array_equal(AT, o.ATA[S - 1]) and np.array_equal(GT, o.GTA[S - 1]), S
        o.op("coulen", S)
        F1 = ind.coulen(g, inp, inp.F2[S - 1], S, CE, CI, 1.5)
        assert np.array_equal(F1, o.F2[S - 1]) and not np.array_equal(F1, inp.F2[S - 1]), S
        o.op("coulmu", S)
        F2 = ind.coulmu(g, inp, F1, S, AT, GT, 50.0)
        assert np.array_equal(F2, o.F2[S - 1]) and not np.array_equal(F2, F1), S


@pytest.mark.parametrize("flags", [0, 2, 1 | 4])
def test_independent_numpy_restatement_of_ram_run(oracle_built, flags):
    """The whole species loop + epilogue of ram_run (src/ModRamRun.f90:64-222) composed from the
    independent numpy routines in the reference's call order -- default operators, Coulomb, WPI+EMIC --
    against the C++ oracle's ram_run: F2 of all species, DtsNext, SETRC, the six loss accumulators and
    PPERT/PPART bit for bit."""
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import independent_ram as ind
    g = grids.build_grids(NR=7, NT=7, NE=35)
    inp = synthetic.make_inputs(g, f2_kind="noisy", inductive=True, mgnp=True)
    D = synthetic.synthetic_daa(g, inp)
    Z = np.zeros_like(D)
    kw = {0: {}, 2: dict(coulomb=True), 5: dict(wpi=True, emic=True, DAA=(Z, D, D, Z))}[flags]
    o = oracle_built.RamOracle(g, inp, DTs=5.0)
    o.set_scalar("T", 25.0)
    o.set_array("ATAC", D)
    o.set_array("ATAW_emic_h", D)
    dtn = o.ram_run(flags=flags)
    F2, dn, SETRC, loss, pe, pa = ind.ram_run(g, inp, inp.F2, 5.0, 1.5, grids.gcoul, T=25.0, **kw)
    assert np.array_equal(F2, o.F2) and dn == dtn and np.array_equal(SETRC, o.SETRC)
    assert np.array_equal(pe[:, 1:], o.PPERT[:, 1:]) and np.array_equal(pa[:, 1:], o.PPART[:, 1:])
    for k, n in (("DR", "LSDR"), ("CHA", "LSCHA"), ("ATM", "LSATM"), ("WAE", "LSWAE"), ("COE", "LSCOE"), ("CSC", "LSCSC")):
        assert np.array_equal(loss[k], o.arr[n]), n
    assert (np.abs(o.arr["LSCOE"]).max() > 0) == bool(flags & 2)


def test_flcscatter_matches_wpadif_with_one_coefficient(oracle_built, small):
    """FLCscatter (src/ModRamLoss.f90:513-575) is WPADIF's tridiagonal with FLC_coef as the only
    coefficient array: with the same array in ATAW_emic_h (and ATAW_emic_he = 0) the two restatements
    must agree bit for bit; before the first boundary cycle (T < Dt_bc) it is a no-op."""
    g, inp = small
    D = synthetic.synthetic_daa(g, inp) * 20.0
    a = oracle_built.RamOracle(g, inp, DTs=5.0)
    b = oracle_built.RamOracle(g, inp, DTs=5.0)
    a.set_array("FLC_coef", D)
    b.set_array("ATAW_emic_h", D)
    a.set_scalar("T", 10.0)
    assert a.op("flcscatter", 1) == 0 and np.array_equal(a.F2, inp.F2)
    a.set_scalar("T", 600.0)
    nv_a = a.op("flcscatter", 1)
    nv_b = b.op("wpadif", 1)
    assert nv_a == nv_b
    assert np.array_equal(a.F2, b.F2) and not np.array_equal(a.F2[0], inp.F2[0])


@pytest.mark.parametrize("S", [1, 4])
def test_para_flc_oracle_vs_independent_numpy(oracle_built, small, S):
    """PARA_FLC (src/ModRamLoss.f90:342-455): the loop-nest oracle against the whole-array numpy restatement
    (tests/independent_ram.py).  Same formulas, different evaluation of the powers (libm pow vs numpy): <= 1e-13;
    properties: zero at L = NPA and where epsilon < 0.1, non-negative, scales as 1/tau_bounce ~ V (same epsilon)."""
    import independent_ram as ind
    import test_zz_late_additions_gpu as T
    g, inp = small
    o = oracle_built.RamOracle(g, inp, DTs=5.0)
    rc, z1, z2 = T.flc_radius_inputs(g, S)
    for n, a in (("r_curvEq", rc), ("zeta1Eq", z1), ("zeta2Eq", z2)):
        o.set_array(n, a)
    o.op("para_flc", S)
    ref = o.arr["FLC_coef"].copy()
    mine = ind.para_flc(g, inp, S, rc, z1, z2)
    assert np.array_equal(ref == 0, mine == 0)
    nz = ref != 0
    assert 0.05 < nz.mean() < 0.95
    assert np.max(np.abs(mine[nz] - ref[nz]) / np.abs(ref[nz])) <= 1e-13
    assert (ref >= 0).all() and (ref[..., -1] == 0).all()
    # a larger curvature radius everywhere -> smaller epsilon -> fewer (never more) scattering lines
    o.set_array("r_curvEq", 3.0 * rc)
    o.op("para_flc", S)
    assert ((o.arr["FLC_coef"] != 0) <= nz).all() and (o.arr["FLC_coef"] != 0).sum() < nz.sum()


def test_coulomb_operators_properties(oracle_built, small):
    """COULPARA/COULEN/COULMU (src/ModRamCoul.f90): the drag coefficients are negative (energy
    loss), vanish at L = NPA (never assigned by the reference), scale linearly with DTs; without
    plasmasphere electrons (NECR = 0) both operators leave F2 unchanged up to the pitch-angle
    solve's own round trip; with them F2 stays positive."""
    g, inp = small
    o = oracle_built.RamOracle(g, inp, DTs=5.0)
    o.set_scalar("T", 100.0)
    o.op("coulpara", 1)
    ce = o.COULE[0].copy()
    assert np.all(ce[:, :-1] < 0) and np.all(ce[:, -1] == 0) and np.all(o.ATA[0][:, 1:-1] > 0)
    o.set_scalar("DTs", 10.0)
    o.op("coulpara", 1)
    assert np.allclose(o.COULE[0], 2 * ce, rtol=1e-13, atol=0)
    o.set_scalar("DTs", 5.0)
    o.op("coulpara", 1)
    o.op("coulen", 1)
    o.op("coulmu", 1)
    assert np.all(o.F2[0][1:, :, 1:, 1:] > 0) and np.array_equal(o.F2[1:], inp.F2[1:])
    # no cold electrons -> no drag, and the implicit solve with zero coefficients is the identity
    import copy
    inp0 = copy.copy(inp)
    inp0.NECR = np.zeros_like(inp.NECR)
    z = oracle_built.RamOracle(g, inp0, DTs=5.0)
    z.set_scalar("T", 100.0)
    z.op("coulpara", 1)
    z.op("coulen", 1)
    assert np.array_equal(z.F2, inp.F2)
    z.op("coulmu", 1)
    f, f0 = z.F2[0][1:, :, 1:, 1:-1], inp.F2[0][1:, :, 1:, 1:-1]
    assert np.allclose(f, f0, rtol=1e-13, atol=0)


def test_independent_numpy_restatement_of_scb(oracle_built):
    """tests/independent_scb.py: `computeBandJacob`, `metrica` / `metric` (nine stencil coefficients from
    the 27-point neighbourhood of x, y, z), `newk` (anisotropic) and three lexicographic SOR sweeps of
    `iterateAlpha`, restated a second time in numpy from the Fortran text -- the C++ oracle agrees
    bit for bit."""
    import math
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import independent_scb as ind
    from ramscb_b200 import scb_synthetic as S
    inp = S.build_scb(nthe=31, npsi=11, nzeta=17, warp=0.3)
    o = oracle_built.ScbOracle(inp)
    o.bandjacob(); o.metrica(); o.newk()
    v = ind.metrica(inp.x, inp.y, inp.z, inp.nthe, inp.npsi, inp.nzeta)
    for n, a in v.items():
        assert np.array_equal(getattr(o, n), a), n
        assert np.abs(a).max() > 0
    m = ind.bandjacob(inp, S.derivs3d)                      # computeBandJacob on the numpy Steffen derivatives
    for n, a in m.items():
        assert np.array_equal(a, getattr(o, n)[:, :, :a.shape[2]]), n
    assert np.array_equal(ind.newk_aniso(inp, m, inp.nthe, inp.npsi, inp.nzeta), o.vecx[:, :, :inp.nzeta])
    nthe, npsi, nzeta, nT = inp.nthe, inp.npsi, inp.nzeta, 4
    rjac = 1.0 - 2.0 * math.pi * math.pi / (nzeta * nzeta + nthe * nthe)
    omopt = 2.0 / (1.0 + math.sqrt(1.0 - rjac * rjac))
    alfa0 = o.alfa.copy()
    o.set_int("nimax", 3)
    o.set_scalar("InConAlpha", 1e-300)
    fail, ni = o.iterate_alpha()
    assert fail == 0 and ni[1:-1].max() == 4              # three sweeps done, counter past nimax
    a, rm = ind.sor_alpha_sweeps(alfa0, {n: getattr(o, n) for n in v}, o.vecx, nthe, npsi, nzeta, nT, 3, [1.0, omopt, omopt])
    core = (slice(nT, nthe - nT), slice(1, npsi - 1), slice(1, nzeta))       # untouched by the post-processing
    assert np.array_equal(a[core], o.alfa[core]) and not np.array_equal(a[core], alfa0[core])
    assert rm.max() == o.get("diffmx")
    o.metric()                                               # the psi equation's coefficients
    for n, a in ind.metric(inp.x, inp.y, inp.z, nthe, npsi, nzeta).items():
        assert np.array_equal(getattr(o, n), a), "metric " + n


def test_scb_steffen_properties(oracle_built):
    """Steffen spline derivative: exact on linear data, zero at local extrema (monotonicity
    preserving), one-sided at the ends; numpy and C++ restatements agree bit for bit."""
    from ramscb_b200 import scb_synthetic as S
    lib = oracle_built.scb_lib()
    x = np.linspace(0.0, 2.0, 17)
    for y in (3.0 * x - 1.0, np.sin(3 * x), np.abs(x - 1.0), x ** 3):
        d = np.zeros_like(x)
        lib.scbo_steffen(len(x), x.ctypes.data, np.ascontiguousarray(y).ctypes.data, d.ctypes.data)
        assert np.array_equal(d, S.steffen_derivs(x, y, 0))
    d = S.steffen_derivs(x, 3.0 * x - 1.0, 0)
    assert np.allclose(d, 3.0, rtol=0, atol=1e-13)
    y = np.abs(x - 1.0)
    d = S.steffen_derivs(x, y, 0)
    assert d[8] == 1e-31                                     # extremum: slope 0, nudged (src/RamGSL.c:276)


def test_scb_oracle_dipole_and_sor(oracle_built):
    from ramscb_b200 import scb_synthetic as S
    inp = S.build_scb(nthe=41, npsi=17, nzeta=33, warp=0.0)
    o = oracle_built.ScbOracle(inp)
    assert o.bandjacob() == 0
    sl = (slice(6, -6), slice(2, -2), slice(1, 33))
    rel = np.abs(o.bsq[sl] - inp.bsq0[sl]) / inp.bsq0[sl]
    assert np.median(rel) < 5e-2          # coarse 41x17x33 grid: truncation error of the spline derivatives
    assert np.all(o.jacobian > 0)
    o.metrica(); o.newk()
    fail, ni = o.iterate_alpha()
    assert fail == 0 and 1 < ni.max() < 5001 and o.get("diffmx") < 1e-6
    # the converged solution satisfies the 9-point equations (Write_convergence_anisotropic's
    # definition of the linear residual, src/ModScbIO.f90:970-1119)
    u = o.alfa
    nthe, nzeta = inp.nthe, inp.nzeta
    c = slice(4, nthe - 4); kk = slice(1, nzeta)
    sh = lambda di, dk: u[4 + di:nthe - 4 + di, 1:-1, 1 + dk:nzeta + dk]
    v = lambda a: a[c, 1:-1, kk]
    res = (-v(o.vecd) * sh(0, 0) + v(o.vec1) * sh(-1, -1) + v(o.vec2) * sh(0, -1) + v(o.vec3) * sh(1, -1) + v(o.vec4) * sh(-1, 0)
           + v(o.vec6) * sh(1, 0) + v(o.vec7) * sh(-1, 1) + v(o.vec8) * sh(0, 1) + v(o.vec9) * sh(1, 1) - v(o.vecx))
    # (edge rows/planes excluded: their neighbours were rewritten by the post-processing :262-292)
    assert np.max(np.abs(res[1:-1, :, 1:-1])) < 5e-5


def test_scb_oracle_theChange_one_branch(oracle_built):
    """theChange <= 1 (src/ModScbEuler.f90:272-279 / :585-591): the theta end points are extrapolated with extap
    (src/ModScbFunctions.f90:57-76) instead of the linear fill.  The oracle used to abort() here.  Checked:
    the post-processing against a literal numpy transcription applied to the same pre-image (the interior rows the SOR
    sweeps leave are untouched by the end-point rule, so the pre-image is recoverable), for iteratePsi (k = 2..nzeta) and
    iterateAlpha (the reference's k = 2..nthe-1 over the zeta index, clipped to the planes that exist: oracle header)."""
    from ramscb_b200 import scb_synthetic as S

    def extap(x1, x2, x3):
        x4 = 3. * x3 - 3. * x2 + x1
        ddx1, ddx2 = x3 - x2, x2 - x1
        if (x4 - x3) * ddx1 > 0.:
            return x4
        if abs(ddx2) <= 1e-9:
            return 2. * x3 - x2
        return x3 + (ddx1 * ddx1) / ddx2

    inp = S.build_scb(nthe=25, npsi=13, nzeta=17, warp=0.1)
    nthe, npsi, nzeta = inp.nthe, inp.npsi, inp.nzeta
    for which in ("psi", "alpha"):
        o = oracle_built.ScbOracle(inp)
        assert o.bandjacob() == 0
        o.set_int("theChange", 1)
        o.set_int("nimax", 30)
        if which == "psi":
            o.metric(); o.newj()
            o.iterate_psi()
            u = o.psi.copy()
            kend, wrap = nzeta, 0.0
        else:
            o.metrica(); o.newk()
            o.iterate_alpha()
            u = o.alfa.copy()
            kend, wrap = min(nthe - 1, nzeta + 1), 2.0 * np.pi
        assert np.all(np.isfinite(u))
        v = u.copy()
        for k in range(2, kend + 1):                       # Fortran k; arrays are (nthe, npsi, nzeta+1)
            for j in range(1, npsi + 1):
                v[nthe - 1, j - 1, k - 1] = extap(v[nthe - 4, j - 1, k - 1], v[nthe - 3, j - 1, k - 1], v[nthe - 2, j - 1, k - 1])
                v[0, j - 1, k - 1] = extap(v[3, j - 1, k - 1], v[2, j - 1, k - 1], v[1, j - 1, k - 1])
        v[:, :, 0] = v[:, :, nzeta - 1] - wrap
        v[:, :, nzeta] = v[:, :, 1] + wrap
        # the end points are a pure function of rows 2..4 / nthe-3..nthe-1, which the rule does not touch: idempotent image
        assert np.array_equal(u, v), which
        # and it is NOT the linear fill of the default branch (which with nT = 1 would leave the end points unchanged)
        o4 = oracle_built.ScbOracle(inp)
        o4.bandjacob()
        o4.set_int("nimax", 30)
        if which == "psi":
            o4.metric(); o4.newj(); o4.iterate_psi()
            assert not np.array_equal(o4.psi[0], u[0])
        else:
            o4.metrica(); o4.newk(); o4.iterate_alpha()
            assert not np.array_equal(o4.alfa[0], u[0])


# ---- the C-ABI library -------------------------------------------------------------------------
def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (the CPU arm the driver runs next to ours) needs no GPU and prints
    one JSON line with the contract's keys; under torchrun only rank 0 works."""
    import json
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--workload", "default", "--steps", "1",
                        "--warmup", "1"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["higher_is_better"] is True and line["unit"] == "cell-updates/s"
    assert line["value"] > 0 and line["steps"] == 1 and 1 <= line["warmup"] <= 3
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["value"] == line["value"] and line["e2e"]["h2d_bytes_per_step"] == 0
    assert "configs[1]" in line["config"]["workload"]
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1"],
                       capture_output=True, text=True, timeout=120, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_c_abi_exports_every_declared_symbol():
    from ramscb_b200 import build, host
    build.build()
    L = ctypes.CDLL(host.LIB_PATH)
    hdr = open(os.path.join(ROOT, "include", "ramscb_gpu.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    names = set(re.findall(r"\b(rsg_[a-z0-9_]+)\s*\(", hdr))
    assert len(names) > 60
    missing = [n for n in sorted(names) if not hasattr(L, n)]
    assert not missing, missing


def test_header_is_plain_c(tmp_path):
    """The drop-in boundary is a C ABI: include/ramscb_gpu.h must compile as C99 (no C++ or torch types in the
    signatures) and as C++."""
    import subprocess
    src = tmp_path / "hdr.c"
    src.write_text('#include "ramscb_gpu.h"\nint main(void) { rsg_scb_run_params p; (void)p; return 0; }\n')
    inc = os.path.join(ROOT, "include")
    for cmd in (["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-fsyntax-only"],
                ["g++", "-std=c++17", "-fsyntax-only", "-x", "c++"]):
        r = subprocess.run(cmd + ["-I", inc, str(src)], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr


def test_no_gpu_means_loud_failure():
    """There is no CPU fallback: without a device, creating a handle fails with RSG_ERR_CUDA."""
    from ramscb_b200 import host
    L = host.lib()
    if L.rsg_device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(host.RsgError):
        host.RamGpu(grids.build_grids())


def test_product_code_does_not_import_the_oracle():
    """oracle/ is test infrastructure: nothing under ramscb_b200/ may reference it."""
    pkg = os.path.join(ROOT, "ramscb_b200")
    for dp, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                txt = open(os.path.join(dp, fn)).read()
                assert "import oracle" not in txt and "from oracle" not in txt and "oracle/" not in txt.replace("the CPU oracle", ""), fn


# ---------------------------------------------------------------------------------------------
# GSL_Interpolation_1D (Steffen) and mapAlpha / mapPsi / mapTheta of the SCB oracle
def _interp(oracle_mod, x1, f1, x2):
    import ctypes as C
    lib = oracle_mod.scb_lib()
    x1, f1, x2 = (np.ascontiguousarray(a, dtype=np.float64) for a in (x1, f1, x2))
    out = np.zeros_like(x2)
    rc = lib.scbo_interp1d(len(x1), x1.ctypes.data, f1.ctypes.data, len(x2), x2.ctypes.data, out.ctypes.data)
    return rc, out


def test_steffen_interpolation_properties(oracle_built):
    """interp1d of the oracle (src/ModRamGSL.f90:240-311 + src/RamGSL.c:111-174): reproduces the nodes,
    is exact on straight lines, never overshoots monotone data (Steffen 1990), extrapolates linearly,
    drops non-increasing abscissae, and agrees bit for bit with the independent numpy restatement."""
    import independent_scb as ind
    rng = np.random.default_rng(11)
    xa = np.cumsum(rng.uniform(0.1, 1.0, 40))
    fa = np.cumsum(rng.uniform(0.0, 1.0, 40)) ** 1.5                    # monotone data
    rc, at_nodes = _interp(oracle_built, xa, fa, xa)
    assert rc == 0 and np.array_equal(at_nodes, fa)
    xb = np.sort(rng.uniform(xa[0] - 1.0, xa[-1] + 1.0, 500))
    rc, fb = _interp(oracle_built, xa, fa, xb)
    assert rc == 0
    inside = (xb > xa[0]) & (xb < xa[-1])
    assert np.all(np.diff(fb[inside]) >= 0), "Steffen's interpolant must be monotone on monotone data"
    assert np.array_equal(fb, ind.interp1d_steffen(xa, fa, xb))
    lo = xb <= xa[0]
    assert np.allclose(fb[lo], fa[0] + (xb[lo] - xa[0]) * (fa[1] - fa[0]) / (xa[1] - xa[0]), rtol=1e-14)
    rc, lin = _interp(oracle_built, xa, 3.0 * xa - 2.0, xb)
    assert np.allclose(lin, 3.0 * xb - 2.0, rtol=1e-13, atol=1e-13)
    # the wrapper's filter: repeated / decreasing abscissae are skipped
    xd = np.insert(xa, [5, 5, 17], [xa[4], xa[3], xa[16] - 1e-3])
    fd = np.insert(fa, [5, 5, 17], [1e9, -1e9, 1e9])
    rc, fb2 = _interp(oracle_built, xd, fd, xb)
    assert rc == 0 and np.array_equal(fb2, fb)
    rc, _ = _interp(oracle_built, [0.0, 0.0, 0.0, -1.0], [1.0, 2.0, 3.0, 4.0], [0.5])
    assert rc != 0                                                    # fewer than 3 usable points: GSLerr


def test_map_routines_match_independent_restatement(oracle_built):
    """mapAlpha / mapPsi / mapTheta of the C++ oracle vs tests/independent_scb.py (bit for bit), on a
    warped grid with perturbed potentials; plus their defining property: re-gridding with the
    unperturbed potentials (alfa == alphaVal on every line) leaves x, y, z where they are."""
    import independent_scb as ind
    from ramscb_b200 import scb_synthetic as S
    inp = S.build_scb(nthe=31, npsi=13, nzeta=25, warp=0.3)
    n = (inp.nthe, inp.npsi, inp.nzeta)
    o = oracle_built.ScbOracle(inp)
    x0, y0, z0 = o.x.copy(), o.y.copy(), o.z.copy()
    assert o.map_alpha() == 0 and o.map_psi() == 0
    for a, b in ((o.x, x0), (o.y, y0), (o.z, z0)):
        assert np.allclose(a, b, rtol=0, atol=1e-12), "identity re-gridding moved the points"
    rng = np.random.default_rng(3)
    o = oracle_built.ScbOracle(inp)
    o.alfa[...] = np.asfortranarray(inp.alfa + 0.08 * rng.standard_normal(inp.alfa.shape))
    o.psi[...] = np.asfortranarray(inp.psi * (1.0 + 0.03 * rng.standard_normal(inp.psi.shape)))
    rx, ry, rz, ra = ind.map_alpha(o.x, o.y, o.z, o.alfa, inp.alphaVal, *n)
    assert o.map_alpha() == 0
    for a, b in ((o.x, rx), (o.y, ry), (o.z, rz), (o.alfa, ra)):
        assert np.array_equal(a, b)
    assert np.abs(o.x - x0).max() > 1e-3
    rx, ry, rz, rp = ind.map_psi(o.x, o.y, o.z, o.psi, inp.psiVal, *n)
    assert o.map_psi() == 0
    for a, b in ((o.x, rx), (o.y, ry), (o.z, rz), (o.psi, rp)):
        assert np.array_equal(a, b)
    rx, ry, rz = ind.map_theta(o.x, o.y, o.z, inp.chiVal, *n)
    assert o.map_theta() == 0
    for a, b in ((o.x, rx), (o.y, ry), (o.z, rz)):
        assert np.array_equal(a, b)
    # after mapTheta the points of a line sit at the prescribed arc-length fractions (to the spline's accuracy)
    j, k = 5, 7
    seg = np.sqrt(np.diff(o.x[:, j, k]) ** 2 + np.diff(o.y[:, j, k]) ** 2 + np.diff(o.z[:, j, k]) ** 2)
    frac = np.concatenate([[0.0], np.cumsum(seg)]) / seg.sum()
    assert np.max(np.abs(frac - inp.chiVal / np.pi)) < 2e-2


@pytest.mark.parametrize("iLossCone,iReduce", [(1, 0), (2, 0), (1, 1), (2, 1)])
def test_pressure_aniso_oracle_vs_independent_numpy(oracle_built, iLossCone, iReduce):
    """The oracle's anisotropic pressure mapping (src/ModScbRun.f90:1087-1175) against the whole-array numpy
    restatement (bit for bit), against the generator of the synthetic inputs (which applies the iLossCone = 1
    formulas on its own), and the chain rule of the Euler-potential derivatives."""
    import independent_scb as ind
    import test_scb_parity_gpu as TS
    from ramscb_b200 import scb_synthetic as S
    inp = S.build_scb(nthe=31, npsi=13, nzeta=25, warp=0.3)
    n = (inp.nthe, inp.npsi, inp.nzeta)
    o = oracle_built.ScbOracle(inp)
    o.bandjacob()
    pe, pa = TS._equatorial_pressures(inp, hot=bool(iReduce))
    o.pressure_aniso(pe, pa, iLossCone, iReduce)
    ref = ind.pressure_aniso(pe, pa, o.bf, o.bsq, *n, iLossCone=iLossCone, iReduce=iReduce)
    nz = inp.nzeta
    for name, r in zip(("pper", "ppar", "sigma", "tau"), ref):
        assert np.array_equal(getattr(o, name)[:, :, :nz], r), name
    assert np.all(o.pper[:, :, :nz] > 0) and np.all(o.ppar[:, :, :nz] > 0)
    assert np.array_equal(o.dPPerdPsi, (1.0 / inp.f)[None, :, None] * o.dPPerdRho)
    assert np.array_equal(o.dBsqdAlpha, (1.0 / inp.fzet[:nz])[None, None, :] * o.dBsqdZeta)
    dT, dR, dZ = S.derivs3d(inp.thetaVal, inp.rhoVal, inp.zetaVal, o.pper[:, :, :nz])
    assert np.array_equal(o.dPPerdTheta, dT) and np.array_equal(o.dPPerdRho, dR) and np.array_equal(o.dPPerdZeta, dZ)
    if iLossCone == 1 and iReduce == 0:
        # the synthetic generator's own 3-D pressure: feed its equatorial values back and recover it
        ieq = (inp.nthe + 1) // 2 - 1
        o2 = oracle_built.ScbOracle(inp)
        o2.bf[...] = np.sqrt(inp.bsq0)
        o2.bsq[...] = inp.bsq0
        o2.pressure_aniso(inp.pper[ieq], inp.ppar[ieq], 1, 0)
        assert np.allclose(o2.pper[:, :, :nz], inp.pper[:, :, :nz], rtol=1e-13, atol=0)
        assert np.allclose(o2.ppar[:, :, :nz], inp.ppar[:, :, :nz], rtol=1e-13, atol=0)


def _c_prototypes():
    """name -> list of (ctype, is_pointer) from include/ramscb_gpu.h"""
    import re
    src = open(os.path.join(ROOT, "include", "ramscb_gpu.h")).read()
    src = re.sub(r"/\*.*?\*/", " ", src, flags=re.S)
    out = {}
    for m in re.finditer(r"\b(?:int|double|long long|const char\s*\*)\s+(rsg_\w+)\s*\(([^)]*)\)\s*;", src):
        params = []
        for p in m.group(2).split(","):
            p = " ".join(p.split())
            if p in ("", "void"):
                continue
            ptr = "*" in p or "[" in p
            base = re.sub(r"\bconst\b|\*|\[.*?\]", " ", p).split()
            ctype = " ".join(base[:-1])          # drop the parameter name
            params.append((ctype, ptr))
        out[m.group(1)] = params
    return out


def test_fortran_shims_match_the_c_header():
    """There is no Fortran compiler in this image, so the ISO_C_BINDING interface blocks of
    ramscb_b200/fortran/ are checked statically against include/ramscb_gpu.h: every bound name exists
    in the header and is exported by the library, the argument counts agree, C scalars are passed by
    `value`, C pointers are arrays / intent(out) scalars / `type(c_ptr), value`, and the kinds agree
    (int <-> c_int, double <-> c_double, long long <-> c_long_long, char* <-> c_char)."""
    import glob
    import re
    protos = _c_prototypes()
    hdr_nc = re.sub(r"/\*.*?\*/", " ", open(os.path.join(ROOT, "include", "ramscb_gpu.h")).read(), flags=re.S)
    assert len(protos) > 90
    from ramscb_b200 import build, host
    build.build()
    L = ctypes.CDLL(host.LIB_PATH)
    nbound = 0
    for path in sorted(glob.glob(os.path.join(ROOT, "ramscb_b200", "fortran", "*.f90"))):
        txt = re.sub(r"&\s*\n\s*", " ", open(path).read())           # join continuation lines
        txt = "\n".join(l.split("!")[0] if "'" not in l.split("!")[0] or l.split("!")[0].count("'") % 2 == 0 else l
                        for l in txt.splitlines())
        for m in re.finditer(r"function\s+(\w+)\s*\(([^)]*)\)\s*bind\(C,\s*name='(\w+)'\)\s*result\(\w+\)(.*?)end function",
                             txt, flags=re.S | re.I):
            fname, args, cname, body = m.group(1), [a.strip() for a in m.group(2).split(",") if a.strip()], m.group(3), m.group(4)
            where = f"{os.path.basename(path)}:{cname}"
            assert fname == cname, where
            assert cname in protos, where + " not declared in include/ramscb_gpu.h"
            assert hasattr(L, cname), where + " not exported by the library"
            cpar = protos[cname]
            assert len(cpar) == len(args), f"{where}: {len(args)} dummy arguments, C prototype has {len(cpar)}"
            decl = {}
            for line in body.splitlines():
                if "::" not in line:
                    continue
                spec, names = line.split("::", 1)
                for n in re.split(r",(?![^(]*\))", names):
                    n = n.strip()
                    if n:
                        decl[re.sub(r"\(.*\)", "", n).strip().lower()] = (spec.lower(), "(" in n)
            for a, (ctype, ptr) in zip(args, cpar):
                assert a.lower() in decl, f"{where}: dummy {a} is not declared"
                spec, is_array = decl[a.lower()]
                by_value = "value" in spec
                if not ptr and ctype.endswith("_fn"):       # function-pointer typedef
                    assert by_value and "type(c_funptr)" in spec, f"{where}: callback {a} must be type(c_funptr), value"
                    continue
                if not ptr:
                    assert by_value and not is_array, f"{where}: C scalar {ctype} {a} must be passed by value"
                    kind = {"int": "c_int", "double": "c_double", "long long": "c_long_long"}[ctype]
                    assert kind in spec, f"{where}: {a} should be {kind}"
                else:
                    if by_value:
                        assert "type(c_ptr)" in spec, f"{where}: pointer {a} passed by value must be type(c_ptr)"
                    else:
                        kind = {"int": "c_int", "double": "c_double", "long long": "c_long_long", "char": "c_char",
                                "void": "c_ptr"}.get(ctype)
                        if kind:                      # opaque handles (rsg_ram**, rsg_scb**) are type(c_ptr), intent(out)
                            assert kind in spec, f"{where}: {a} should be {kind} ({spec.strip()})"
                        elif ctype in ("rsg_ram", "rsg_scb", "rsg_hi"):
                            assert "type(c_ptr)" in spec, f"{where}: handle {a}"
                        else:                         # struct passed by reference: a bind(C) derived type of the same name
                            assert f"type({ctype})" in spec, f"{where}: {a} should be type({ctype})"
                            m2 = re.search(r"typedef struct " + ctype + r"\s*\{(.*?)\}", hdr_nc, flags=re.S)
                            cmem = [(t.strip(), [x.strip() for x in names.split(",")]) for t, names in
                                    re.findall(r"\b(int|double)\s+([^;]+);", m2.group(1))]
                            f2 = re.search(r"type, bind\(C\) :: " + ctype + r"(.*?)end type", txt, flags=re.S | re.I)
                            fmem = [("int" if "c_int" in l.split("::")[0] else "double", [x.strip() for x in l.split("::")[1].split(",")])
                                    for l in f2.group(1).splitlines() if "::" in l]
                            flat = lambda mem: [(t, n) for t, ns in mem for n in ns]
                            assert flat(cmem) == flat(fmem), f"{where}: members of {ctype} differ"
            nbound += 1
    assert nbound >= 55, nbound


def _hi_line_by_quadrature(mirror, cVal, bf, var, quad):
    """integrator_c / bounceaverage_c (src/RamGSL.c:479-602) restated in Python with the LITERAL integrands
    (f_I = sqrt(Bm - B), f_h = 1/sqrt(Bm - B), f_D = n/sqrt(Bm - B), zero where B >= Bm; B, n by linear table
    look-up, :295-448) handed to an adaptive quadrature, as the reference hands them to cquad."""
    nT, nPa = len(cVal), len(mirror)
    mirror = mirror.copy()
    B = lambda t: np.interp(t, cVal, bf)
    V = lambda t: np.interp(t, cVal, var)

    def integrals(a, b, bm):
        pts = [c for c in cVal if a < c < b]
        for k in range(nT - 1):                      # mirror points inside [a, b]: integrable singularities of f_h, f_D
            if (bf[k] - bm) * (bf[k + 1] - bm) < 0.0:
                t = cVal[k] + (bm - bf[k]) / (bf[k + 1] - bf[k]) * (cVal[k + 1] - cVal[k])
                if a < t < b:
                    pts.append(t)
        fI = lambda t: 0.0 if B(t) >= bm else np.sqrt(bm - B(t))
        fH = lambda t: 0.0 if B(t) >= bm else np.sqrt(1.0 / (bm - B(t)))
        fD = lambda t: 0.0 if B(t) >= bm else V(t) * np.sqrt(1.0 / (bm - B(t)))
        return [quad(f, a, b, points=sorted(pts), limit=400, epsabs=1e-9, epsrel=1e-9)[0] for f in (fI, fH, fD)]

    a, b, LH, RH = np.zeros(nPa), np.zeros(nPa), np.zeros(nPa, int), np.zeros(nPa, int)
    for L in range(1, nPa - 1):
        for i in range(1, nT - 1):
            if bf[i] <= mirror[L] <= bf[i - 1]:
                a[L], LH[L] = cVal[i - 1], i - 1
                break
        for i in range(nT - 2, 0, -1):
            if bf[i - 1] <= mirror[L] <= bf[i]:
                b[L], RH[L] = cVal[i], i
                break
    yI, yH, yV = np.zeros(nPa), np.zeros(nPa), np.zeros(nPa)
    yI[-1], yH[-1], yV[-1] = integrals(cVal[0], cVal[-1], mirror[-1])
    computed = []
    for L in range(nPa - 2, 0, -1):
        if mirror[L] >= bf[1] or a[L] == 0:
            a[L] = cVal[0]
        if mirror[L] >= bf[nT - 1] or b[L] == 0:
            b[L] = cVal[-1]
        if a[L] <= cVal[0] or b[L] >= cVal[-1]:
            mirror[L] = mirror[-1]
            yI[L], yH[L], yV[L] = yI[-1], yH[-1], yV[-1]
        elif RH[L] - LH[L] <= 4:
            yI[L], yH[L], yV[L] = yI[L + 1], yH[L + 1], yV[L + 1]
        else:
            yI[L], yH[L], yV[L] = integrals(a[L], b[L], mirror[L])
            computed.append(L)
            if yI[L] <= 0: yI[L] = yI[L + 1]
            if yH[L] <= 0: yH[L] = yH[L + 1]
            if yV[L] <= 0: yV[L] = yV[L + 1]
    yI[0], yH[0], yV[0] = 0.0, yH[1], yV[1]
    return mirror, yI, yH, yV, computed


@pytest.mark.parametrize("wiggle", [0.0, 0.1])
def test_hi_oracle_vs_adaptive_quadrature(oracle_built, default_grids, wiggle):
    """The closed-form segment sums of oracle/hi_oracle.cpp against an independent adaptive quadrature (QUADPACK
    via scipy) of the reference's literal integrands, with the mirror search and fall-back chain restated a second
    time in Python: the bar is cquad's own tolerance in the reference (epsrel 1e-3, src/RamGSL.c:351, :402, :455);
    the closed forms agree with the converged quadrature far inside it."""
    from scipy.integrate import quad
    import warnings
    from ramscb_b200 import scb_synthetic
    g = default_grids
    mu = g.MU[::3].copy()
    mu[-1] = g.MU[-1]
    d = scb_synthetic.ram_field_lines(g.LZ[1:g.NR + 1:9], g.MLT[:2], nthe=41, wiggle=wiggle, seed=9)
    _, _, _, _, M = oracle_built.hi_integrals(mu=mu, **d)
    ncomputed = 0
    for i in range(d["bRAM"].shape[1]):
        bf = d["bRAM"][:, i, 1].copy()
        ke = d["nThetaEquator"] - 1
        if abs(bf[ke] - bf.min()) > 1e-9:                                 # src/ModRamScb.f90:388-394
            bf[ke] = 2.0 * bf.min() - bf[ke] if 2.0 * bf.min() - bf[ke] > 0.0 else bf.min() - 0.01
        mir = np.append(bf[ke] / (1.0 - mu[:-1] ** 2), bf[-1])
        m1, yI, yH, yV = oracle_built.hi_line(mir, d["chiVal"], bf, d["density"][:, i, 1])
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            m2, qI, qH, qV, computed = _hi_line_by_quadrature(mir, d["chiVal"], bf, d["density"][:, i, 1], quad)
        assert np.array_equal(m1, m2) and np.array_equal(m1, M[i, 1, :])
        for name, y, q in (("yI", yI, qI), ("yH", yH, qH), ("yD", yV, qV)):
            rel = np.abs(y - q) / np.maximum(np.abs(q), 1e-300)
            assert np.all(rel[1:] < 1e-3), (name, i, rel.max())          # the reference's own quadrature tolerance
            assert np.all(rel[1:] < 1e-6), (name, i, rel.max())          # what the closed forms actually reach
        ncomputed += len(computed)
    assert ncomputed > 20


def _hi_tail_numpy(I, H, D, bz, ScaleAt, outside, Lz, PA, PAbn, smooth, DthI, ram):
    """computehI after the integral block (src/ModRamScb.f90:413-637) a second time, in whole-array numpy straight from
    the Fortran array statements; smoothing through scipy.ndimage (mode 'reflect' = the 3 x 3 reflected tiling of
    srcExternal/gaussian_filter.f90, which says it mirrors a Python implementation)."""
    import independent_scb as ind
    from scipy import ndimage
    I, H, D, bz = (np.array(a, order="F") for a in (I, H, D, bz))
    nR, nT, nPa = I.shape
    for j in range(1, nT):
        ii = ScaleAt[j]
        if ii == 0:
            continue
        fr = (Lz[ii - 1] - Lz[ii - 2]) / (Lz[ii - 3] - Lz[ii - 2])
        for a in (I, H, D):
            t = a[ii - 2, j, 1:] + fr * (a[ii - 3, j, 1:] - a[ii - 2, j, 1:])
            s = np.where(t <= 0, a[ii - 2, j, 1:] / a[ii - 1, j, 1:], t / a[ii - 1, j, 1:])
            for i in range(ii - 1, nR):
                a[i, j, 1:] = a[i, j, 1:] * s if outside[i, j] == 0 else a[i - 1, j, 1:]
        bz[ii - 1:, j] = bz[ii - 2, j]
    for a in (I, H, D):
        a[:, 0, :] = a[:, nT - 1, :]
    bz[:, 0] = bz[:, nT - 1]
    I[:, :, 2] = 0.50 * I[:, :, 3]; I[:, :, 1] = 0.20 * I[:, :, 2]; I[:, :, 0] = 0.0
    H[:, :, 2] = 0.99 * H[:, :, 3]; H[:, :, 1] = 0.99 * H[:, :, 2]; H[:, :, 0] = 0.99 * H[:, :, 1]
    D[:, :, 2] = 0.999 * D[:, :, 3]; D[:, :, 1] = 0.999 * D[:, :, 2]; D[:, :, 0] = 0.999 * D[:, :, 1]
    if min(np.nanmin(H), np.nanmin(I), np.nanmin(D)) < 0:
        for i in range(1, nR):
            for a in (H, I, D):
                a[i] = np.where(a[i] < 0, a[i - 1], a[i])
    for L in range(nPa - 2, -1, -1):
        for a, f in ((I, 0.99), (H, 0.99), (D, 0.999)):
            a[:, :, L] = np.where(a[:, :, L] > a[:, :, L + 1], f * a[:, :, L + 1], a[:, :, L])
    hI, iI = np.zeros_like(H), np.zeros_like(I)
    for i in range(nR):
        for j in range(nT):
            hI[i, j, nPa - 2:0:-1] = ind.interp1d_steffen(PA[::-1], H[i, j, ::-1], PAbn[nPa - 2:0:-1])
            iI[i, j, nPa - 2:0:-1] = ind.interp1d_steffen(PA[::-1], I[i, j, ::-1], PAbn[nPa - 2:0:-1])
    for a in (hI, iI):
        a[:, :, nPa - 1] = a[:, :, nPa - 2]
        a[:, :, 0] = a[:, :, 1]
    if smooth:
        xx, yy = np.meshgrid(np.arange(-4, 5.0), np.arange(-4, 5.0), indexing="ij")
        k = 2.0 * np.exp(-0.5 * (xx ** 2 + yy ** 2) / 1.0)
        k = k / k.sum()
        for L in range(1, nPa):
            for a in (H, I, hI, iI, D):
                a[:, :, L] = ndimage.correlate(a[:, :, L], k, mode="reflect")
    out = {n: np.array(ram[n], order="F") for n in ("FNHS", "FNIS", "BOUNHS", "BOUNIS", "HDNS", "BNES")}
    prev = {n: out[n].copy() for n in out}
    out["FNHS"][1:], out["FNIS"][1:], out["HDNS"][1:], out["BOUNHS"][1:], out["BOUNIS"][1:] = H, I, D, hI, iI
    out["BNES"][1:] = bz / 1e9
    still = abs(DthI) <= 1e-9
    for dn, n in (("dIdt", "FNIS"), ("dHdt", "FNHS"), ("dIbndt", "BOUNIS"), ("dBdt", "BNES")):
        out[dn] = np.zeros_like(out[n]) if still else (out[n] - prev[n]) / DthI
        out[dn][0] = 0.0
    for n in ("FNHS", "FNIS", "BOUNHS", "BOUNIS", "HDNS"):
        out[n][0] = out[n][1]
    out["BNES"][0] = 0.32 / Lz[0] ** 3 / 1.e4
    for i in range(1, nR + 1):
        for n in ("FNIS", "FNHS", "BOUNIS", "BOUNHS", "HDNS"):
            out[n][i] = np.where(np.isnan(out[n][i]), out[n][i - 1], out[n][i])
    for n in ("dIdt", "dIbndt"):
        out[n][1:] = np.where(np.isnan(out[n][1:]), 0.0, out[n][1:])
    out.update(I_cart=I, H_cart=H, HDens_cart=D, bZEq_cart=bz, h_interp=hI, I_interp=iI)
    return out


@pytest.mark.parametrize("smooth,DthI,variant", [(0, 300.0, "scaled"), (1, 300.0, "repairs"), (1, 0.0, "plain")])
def test_hi_tail_oracle_vs_independent_numpy(oracle_built, default_grids, smooth, DthI, variant):
    """scbo_hi_tail (scalar loops in the reference's statement order) against a whole-array numpy restatement written
    from the Fortran array statements, with scipy.ndimage as the smoothing: bit-identical without smoothing, within
    3e-11 of each array's maximum with it (ndimage sums the 81 products in another order; the time derivatives
    divide differences of smoothed values)."""
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import test_zz_late_additions_gpu as TZ
    g = default_grids
    d, ScaleAt, rng = TZ.hi_tail_inputs(g, seed=33, scale_cols=variant != "plain")
    I, H, D, bz, _ = oracle_built.hi_integrals(mu=g.MU, **d)
    if variant == "repairs":
        for a in (I, H, D):
            m = rng.random(a.shape) < 0.01
            a[m] = -a[m]
            m = rng.random(a.shape) < 0.01
            a[m] = 3.0 * a[m]
        if not smooth:
            H[rng.random(H.shape) < 0.002] = np.nan
    shape3 = (g.NR + 1, g.NT, g.NPA)
    ram = {n: np.asfortranarray(rng.random(shape3)) for n in ("FNHS", "FNIS", "BOUNHS", "BOUNIS", "HDNS")}
    ram["BNES"] = np.asfortranarray(1e-7 * rng.random((g.NR + 1, g.NT)))
    Lz = g.LZ[:g.NR + 1] if len(g.LZ) > g.NR else np.append(2 * g.LZ[0] - g.LZ[1], g.LZ)
    args = (I, H, D, bz, ScaleAt, d["outsideMGNP"], Lz, g.PA, g.PAbn, smooth, DthI, ram)
    ref = oracle_built.hi_tail(*args)
    ind = _hi_tail_numpy(*args)
    for n in ("I_cart", "H_cart", "HDens_cart", "bZEq_cart", "h_interp", "I_interp", "FNHS", "FNIS", "BOUNHS", "BOUNIS", "HDNS",
              "BNES", "dIdt", "dHdt", "dIbndt", "dBdt"):
        if smooth:
            scale = np.nanmax(np.abs(ref[n])) + 1e-300
            assert np.allclose(ref[n], ind[n], rtol=0, atol=1e-13 * scale * 300, equal_nan=True), (n, float(np.nanmax(np.abs(ref[n] - ind[n])) / scale))
        else:
            assert np.array_equal(ref[n], ind[n], equal_nan=True), (n, float(np.nanmax(np.abs(ref[n] - ind[n]))))


def test_hi_convert_oracle_vs_independent_numpy(oracle_built):
    """hio_convert_lines (nine MINLOC passes, literal) against a numpy restatement that takes the nine nearest scattered
    points from ONE stable argsort of the squared distances -- the same sequence MINLOC + overwrite produces -- and
    NN_Interpolation_2D's weights (src/ModRamGSL.f90:368-422, :872-917); winding number from src/ModRamScb.f90:262-280
    in array form.  Bit-identical."""
    from ramscb_b200 import scb_synthetic
    nthe, npsi, nzeta, nR, nT = 9, 13, 15, 7, 9
    inp = scb_synthetic.build_scb(nthe=nthe, npsi=npsi, nzeta=nzeta, warp=0.2)
    ke = nthe // 2 + 1
    r = np.sqrt(inp.x ** 2 + inp.y ** 2 + inp.z ** 2)
    bf = np.asfortranarray(30574.0 / r ** 3)
    Lz = np.linspace(1.75, 8.5, nR + 1)
    MLT = np.linspace(0.0, 24.0, nT)
    xR, yR, zR, bR, outside, psiR = oracle_built.hi_convert_lines(inp.x, inp.y, inp.z, bf, inp.psi, inp.alfa, Lz, MLT, ke)

    def nn9(X, Y, Fs, x2, y2):               # X, Y, F: (npsi, nzeta-1); scatter order = C order of that shape
        xs, ys = X.ravel(order="C"), Y.ravel(order="C")
        d2 = (xs - x2) ** 2 + (ys - y2) ** 2
        near = np.argsort(d2, kind="stable")[:9]
        w, wsum = np.zeros(9), 0.0
        for i, n in enumerate(near):
            d = np.sqrt((xs[n] - x2) ** 2 + (ys[n] - y2) ** 2)
            if abs(d) <= 1e-9:
                w[:] = 0.0
                w[i] = 1.0
                wsum = 1.0
                break
            w[i] = 1 / (d * d)               # gfortran expands d**2 to d*d; Python's pow(d, 2) is not always the rounded product
            wsum = wsum + w[i]
        res = []
        for F in Fs:
            fs, v = F.ravel(order="C"), 0.0
            for i, n in enumerate(near):
                v = v + fs[n] * w[i] / wsum
            res.append(v)
        return res

    pi = np.pi
    for i in range(nR):
        for j in range(nT):
            xo = Lz[i + 1] * np.cos(MLT[j] * 2.0 * pi / 24.0 - pi)
            yo = Lz[i + 1] * np.sin(MLT[j] * 2.0 * pi / 24.0 - pi)
            xn, yn = inp.x[ke - 1, npsi - 2, :nzeta], inp.y[ke - 1, npsi - 2, :nzeta]
            xp, yp = inp.x[ke - 1, npsi - 2, 1:nzeta + 1], inp.y[ke - 1, npsi - 2, 1:nzeta + 1]
            cross = (xp - xn) * (yo - yn) - (yp - yn) * (xo - xn)
            wn = np.sum((yn <= yo) & (yp > yo) & (cross > 0)) - np.sum((yn > yo) & (yp <= yo) & (cross < 0))
            assert outside[i, j] == (0 if wn != 0 else 1)
            if wn == 0:
                assert not xR[:, i, j].any() and not bR[:, i, j].any()
                continue
            al = MLT[j] * pi / 12.0 + pi
            if al > 2.0 * pi:
                al = al - 2.0 * pi
            sl = (slice(None), slice(1, nzeta))
            psiRAM, = nn9(inp.x[ke - 1][sl], inp.y[ke - 1][sl], [inp.psi[ke - 1][sl]], xo, yo)
            assert psiRAM == psiR[i, j]
            for k in range(nthe):
                v = nn9(inp.psi[k][sl], inp.alfa[k][sl], [inp.x[k][sl], inp.y[k][sl], inp.z[k][sl], bf[k][sl]], psiRAM, al)
                assert v == [xR[k, i, j], yR[k, i, j], zR[k, i, j], bR[k, i, j]], (i, j, k)


def test_hi_integrals_reproduce_reference_dipole_functions(oracle_built, default_grids):
    """Known-answer pin of the computehI integral block against the reference's OWN dipole closed forms: funt(mu) = h and
    funi(mu) = I of Ejiri (1978) (src/ModRamFunctions.f90:90-143, what the reference initialises FNHS / FNIS with).  On
    dipole lines with nodes at equal arc-length fractions of chiVal (as mapTheta leaves them) H_cart and I_cart of
    src/ModRamScb.f90:372-410 converge to them as nthe grows; the residual 1e-3 is the accuracy of Ejiri's fit."""
    from ramscb_b200 import scb_synthetic
    g = default_grids

    def funt(x):
        y = np.sqrt(1 - x * x)
        al = 1. + np.log(2. + np.sqrt(3.)) / 2. / np.sqrt(3.)
        be = al / 2. - np.pi * np.sqrt(2.) / 12.
        return al - be * (y + np.sqrt(y)) + 0.055 * y ** (1. / 3.) - 0.037 * y ** (2. / 3.) - 0.074 * y + 0.056 * y ** (4. / 3.)

    def funi(x):
        y = np.sqrt(1 - x * x)
        yl = np.log(y)
        al = 1. + np.log(2. + np.sqrt(3.)) / 2. / np.sqrt(3.)
        be = al / 2. - np.pi * np.sqrt(2.) / 12.
        return (2. * al * (1. - y) + 2. * be * y * yl + 4. * be * (y - np.sqrt(y)) + 3. * 0.055 * (y ** (1. / 3.) - y)
                + 6. * (-0.037) * (y ** (2. / 3.) - y) + 6. * 0.056 * (y - y ** (4. / 3.)) - 2. * (-0.074) * y * yl)

    Ls = np.array([3.0, 5.0, 6.5])
    for nthe, tol_max, tol_med in ((101, 3e-2, 4e-3), (801, 3e-3, 1.5e-3)):
        d = scb_synthetic.ram_field_lines(Ls, g.MLT[:2], nthe=nthe)
        I, H, _, _, _ = oracle_built.hi_integrals(mu=g.MU, **d)
        for i in range(len(Ls)):
            mu_lc = np.sqrt(1.0 - d["bRAM"][nthe // 2, i, 0] / d["bRAM"][-1, i, 0])       # mirror point at the foot of the line
            sel = (g.MU < 0.95 * mu_lc) & (np.arange(g.NPA) >= 4)
            assert sel.sum() > 40
            eH = np.abs(H[i, 0, sel] / funt(g.MU[sel]) - 1.0)
            eI = np.abs(I[i, 0, sel] / funi(g.MU[sel]) - 1.0)
            assert eH.max() < tol_max and eI.max() < tol_max and np.median(eH) < tol_med and np.median(eI) < tol_med, (nthe, Ls[i], eH.max(), eI.max())
