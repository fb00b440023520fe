"""The RAM kernels, executed thread for thread by the host-CPU CUDA emulator (tests/emu/), against
the oracle -- the `-m gpu` parity tests of tests/test_ram_parity_gpu.py re-run on a reduced grid in
a container without a GPU.  TEST INFRASTRUCTURE: the emulator library is built from a scratch copy
of ramscb_b200/csrc by tests/emu/build_emu.py and is loaded only here (the product library has no
CPU path and still fails loudly without a device).  What this buys: indexing, barriers, shuffles,
shared-memory staging and the arithmetic of every RAM kernel are checked before GPU time is spent.
"""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def emu():
    """Point ramscb_b200.host at the emulator build for this module only."""
    import conftest
    from ramscb_b200 import host
    saved = (host.LIB_PATH, host._lib)
    conftest.use_emulator()
    yield host
    host.LIB_PATH, host._lib = saved


@pytest.fixture(scope="module")
def small_grids():
    from ramscb_b200 import grids
    return grids.build_grids(NR=9, NT=11, NE=35)      # ragged planes; NE stays 35: the ANISCH energy bands (khi, ModRamRun.f90:352) are hard-wired up to K=30


@pytest.fixture(scope="module")
def T():
    import test_ram_parity_gpu as t
    return t


@pytest.mark.parametrize("op", ["DRIFTR", "DRIFTP", "DRIFTE", "DRIFTMU"])
def test_emulated_exact_sweeps_bit_identical(emu, T, small_grids, oracle_built, op):
    T.test_drift_sweeps_bit_exact(small_grids, oracle_built, "adversarial_mgnp", op)


@pytest.mark.parametrize("flags", [0, 1 | 2 | 4])
def test_emulated_full_ram_run_exact(emu, T, small_grids, oracle_built, flags):
    T.test_full_ram_run(small_grids, oracle_built, flags)


def test_emulated_fast_full_ram_run(emu, T, small_grids, oracle_built):
    T.test_fast_mode_full_ram_run(small_grids, oracle_built)


@pytest.mark.parametrize("mode", ["exact", "fast", "fast_unfused"])
def test_emulated_driftp_wrap_nonperiodic_fields(emu, T, small_grids, oracle_built, mode):
    T.test_driftp_wrap_with_fields_not_periodic_in_mlt(small_grids, oracle_built, mode)


def test_emulated_ram_run_host_pipeline(emu, T, small_grids):
    T.test_ram_run_host_pipelined_equals_three_calls(small_grids, "fast", 5)


def test_emulated_losses_wpadif_coulomb(emu, T, small_grids, oracle_built):
    T.test_losses(small_grids, oracle_built, "noisy")
    T.test_sumrc_and_anisch(small_grids, oracle_built)
    T.test_wpadif_bit_exact(small_grids, oracle_built, "chorus")
    T.test_coulomb_operators_bit_exact(small_grids, oracle_built, 1)
    import test_zz_late_additions_gpu as TZ
    TZ.test_para_flc_on_device(small_grids, oracle_built, 1)


def test_emulated_fused_equals_unfused_and_graph_replay(emu, T, small_grids):
    from ramscb_b200 import host, synthetic
    g = small_grids
    inp = synthetic.make_inputs(g, f2_kind="noisy", inductive=True, mgnp=True)
    runs = []
    for fused, graph in ((False, False), (True, False), (True, True)):
        gpu = host.RamGpu(g)
        gpu.set_mode(host.MODE_FAST)
        gpu.set_inputs(inp)
        gpu.use_fused(fused)
        gpu.use_graph(graph)
        outs = [gpu.ram_run(dts, DtsMin=1.0, flags=0) for dts in (5.0, 5.0, 2.5)]
        runs.append((gpu.f2_d2h(), outs))
        gpu.close()
    for f_b, o_b in runs[1:]:
        assert np.array_equal(runs[0][0], f_b)
        for a, b in zip(runs[0][1], o_b):
            assert np.array_equal(a["DtDrift"], b["DtDrift"]) and a["DtsNext"] == b["DtsNext"]
            for k in ("PPERT", "PPART", "SETRC"):
                assert np.allclose(a[k], b[k], rtol=1e-12, atol=0), k


@pytest.mark.parametrize("flags", [1 | 4])
def test_emulated_fused_wpadif_step(emu, T, small_grids, oracle_built, flags):
    T.test_fused_wpadif_fast_step(small_grids, oracle_built, "default", flags)


@pytest.mark.parametrize("flags", [1 | 2 | 4])
def test_emulated_fused_coulomb_step(emu, T, small_grids, oracle_built, flags):
    T.test_fused_coulomb_fast_step(small_grids, oracle_built, flags)


def test_emulated_scb_maps_and_geometry(emu, oracle_built):
    """SCB kernels in the emulator: computeBandJacob, metrica/newk, the lexicographic SOR wavefront and
    mapAlpha / mapPsi / mapTheta, all bit-identical to the oracle (tests/test_scb_parity_gpu.py)."""
    import test_scb_parity_gpu as TS
    TS.test_map_alpha_psi_theta_bit_exact(oracle_built)
    TS.test_outer_iteration_alpha_half_stays_on_device(oracle_built)
    TS.test_pressure_anisotropic_mapping_bit_exact(oracle_built, 1, 1)
    TS.test_pressure_anisotropic_mapping_bit_exact(oracle_built, 2, 0)


@pytest.mark.parametrize("nthe,wiggle,outside", [(51, 0.0, 0.0), (51, 0.15, 0.1)])
def test_emulated_hI_integrals(emu, default_grids, oracle_built, nthe, wiggle, outside):
    import test_zz_late_additions_gpu as TZ
    TZ.test_hI_integrals_on_device(default_grids, oracle_built, nthe, wiggle, outside)


@pytest.mark.parametrize("smooth,DthI,variant", [(1, 300.0, "scaled"), (0, 0.0, "plain"), (1, 300.0, "repairs")])
def test_emulated_hI_tail(emu, default_grids, oracle_built, smooth, DthI, variant):
    import test_zz_late_additions_gpu as TZ
    TZ.test_hI_tail_on_device(default_grids, oracle_built, smooth, DthI, variant)


def test_emulated_hI_convert_lines(emu, oracle_built):
    import test_zz_late_additions_gpu as TZ
    TZ.test_hI_convert_lines_on_device(oracle_built, (11, 15, 17, 6, 7))


def test_emulated_computehI_composed(emu, default_grids, oracle_built):
    import test_zz_late_additions_gpu as TZ
    TZ.test_computehI_composed_on_device(default_grids, oracle_built, (21, 19, 25))


def test_emulated_computehI_resident(emu, small_grids, oracle_built):
    import test_zz_late_additions_gpu as TZ
    TZ.test_computehI_resident_handle(small_grids, oracle_built, (21, 15, 25))      # (the GPU run uses the default RAM grid)


def test_emulated_coupled_cycle(emu, oracle_built):
    import test_zz_late_additions_gpu as TZ
    from ramscb_b200 import grids
    g = grids.build_grids(NR=11, NT=11, NE=35)      # the pressure front end needs NR >= 10 (GPU: default RAM grid)
    TZ.test_coupled_ram_scb_cycle_stays_on_the_device(g, oracle_built, (21, 15, 25))


def test_emulated_scb_run_outer_iterations(emu, oracle_built):
    """rsg_scb_run -- the whole outer iteration of scb_run in one C call, 3-D arrays resident, pressure front
    end as a host callback -- against the oracle's composition, incl. the SORFail restore path."""
    import test_zz_late_additions_gpu as TZ
    TZ.test_scb_run_outer_iterations_resident(oracle_built, numit=2, color4=False)     # (the GPU run does 3 iterations and the 4-colour ordering)


def test_emulated_results_do_not_depend_on_thread_order():
    """Race check: EMU_ORDER=random runs the runnable threads of every block in a fresh random order
    between synchronisation points.  A kernel whose result depended on the order (a missing barrier
    between a producer and a consumer stage) would break the bit-exact comparisons of the fused-step
    tests; they must pass unchanged."""
    import subprocess
    env = dict(os.environ, EMU_ORDER="random")
    here = os.path.dirname(os.path.abspath(__file__))
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(here, "test_emu_cpu.py"), "-x", "-q", "-k",
                        "(fused_wpadif and 5) or (exact_sweeps and DRIFTP) or hI_tail or hI_convert"], env=env, capture_output=True, text=True, cwd=os.path.dirname(here))
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]


@pytest.mark.parametrize("use_bas", [True, False])
def test_anisch_diffusion_coefficient_rebuild(emu, use_bas):
    """ANISCH, second half (src/ModRamRun.f90:422-605): chorus (Steffen 1-D), hiss and EMIC (bilinear 2-D)
    coefficients rebuilt by the device kernels (emulated) against the oracle -- the kernels keep the reference's operation
    order, so ATAW, ATAC, ATAW_emic_h, ATAW_emic_he are BIT-IDENTICAL -- and the oracle against a second, whole-array numpy
    restatement (<= 1e-13: np.log10 / ** vs libm differ in the last bit)."""
    import independent_ram as ir
    from oracle import oracle
    from ramscb_b200 import grids, synthetic
    g = grids.build_grids(NR=9, NT=7, NE=8)
    inp = synthetic.make_inputs(g, f2_kind="smooth")
    t = synthetic.synthetic_wave_tables(g, inp)
    assert 0 < (t["XNE"] > 50).sum() < t["XNE"].size          # both sides of the plasmapause are exercised
    o = oracle.RamOracle(g, inp)
    gpu = emu.RamGpu(g)
    gpu.set_inputs(inp)
    with pytest.raises(emu.RsgError, match="set_wave_tables"):
        gpu.ANISCH_diffcoef(4, 1, t["XNE"])
    gpu.set_wave_tables(t, use_bas=use_bas)
    for S, fl in ((4, emu.F_WPI), (1, emu.F_EMIC)):
        assert o.anisch_diffcoef(S, fl, t, AE=150, use_bas=use_bas) == 0
        assert gpu.ANISCH_diffcoef(S, fl, t["XNE"], AE=150) == 0
    ref2 = {}
    ref2.update(ir.anisch_diffcoef(g, inp, t, 4, True, False, inp.Kp, AE=150, use_bas=use_bas))
    ref2.update(ir.anisch_diffcoef(g, inp, t, 1, False, True, inp.Kp, AE=150, use_bas=use_bas))
    for which, name in enumerate(("ATAW", "ATAC", "ATAW_emic_h", "ATAW_emic_he")):
        a, b = gpu.get_diffcoef(which), getattr(o, name)
        assert (b != 0).sum() > 1000
        assert np.array_equal(a, b), name
        c = ref2[name]
        assert np.array_equal(c == 0, b == 0), name
        assert np.max(np.abs(c - b) / np.maximum(np.abs(b), 1e-300)) <= 1e-13, name
    assert (o.ATAW_emic_h == 1e-31).sum() > 0                 # the 1e-20 floor of :585 is hit
    # species the flags do not select are left alone
    before = gpu.get_diffcoef(1)
    assert gpu.ANISCH_diffcoef(2, emu.F_WPI | emu.F_EMIC, t["XNE"]) == 0
    assert np.array_equal(gpu.get_diffcoef(1), before)
    gpu.close()


def _ram_pressures(g):
    from ramscb_b200 import scb_synthetic
    return scb_synthetic.synthetic_ram_pressures(g)


def test_pressure_front_end_on_device(emu):
    """SURVEY 8(f)-2, the 2-D front end of `pressure` (src/ModScbRun.f90:838-1086): RAM pressures summed over the SCB species,
    radial extension (all four PressModes), SavGol7 / Gaussian smoothing, bilinear interpolation to the equatorial foot points,
    extap / floor / periodic columns.  Device kernels (emulated) BIT-IDENTICAL to the oracle; the oracle's pieces against
    scipy (Savitzky-Golay interior points, RegularGridInterpolator); then scb_run with NO host callback against the oracle's
    composition of the same loop."""
    from oracle import oracle
    from scipy.interpolate import RegularGridInterpolator
    from scipy.signal import savgol_coeffs
    from ramscb_b200 import grids, scb_synthetic as S
    g = grids.build_grids()
    PPerT, PParT, scb, LZ, PHI = _ram_pressures(g)
    sinp = S.build_scb(nthe=51, npsi=23, nzeta=49, warp=0.3)
    o, gpu = oracle.ScbOracle(sinp), emu.ScbGpu(sinp)
    with pytest.raises(emu.RsgError, match="set_ram_pressure"):
        gpu.pressure_front()
    for mode, ism in (("ROE", 1), ("EXT", 3), ("FLT", 0), ("SKD", 4)):
        r = o.pressure_raw(PPerT, PParT, scb, LZ, PHI, PressMode=mode, iSm2=ism)
        gpu.set_ram_pressure(PPerT, PParT, scb, LZ, PHI, PressMode=mode, iSm2=ism)
        for a, b in zip(r, gpu.get_ram_pressure()):
            assert np.array_equal(a, b), mode
        o.bandjacob(); gpu.computeBandJacob()
        pe, pa = o.pressure_front()
        ge, ga = gpu.pressure_front()
        assert np.array_equal(pe, ge) and np.array_equal(pa, ga), mode
        assert pe.min() > 0 and np.array_equal(pe[:, 0], pe[:, -2]) and np.array_equal(pe[:, -1], pe[:, 1])
        o.pressure_aniso(pe, pa)                 # the 3-D tail (gpu.pressure_front ran it already)
        _same = lambda n: np.array_equal(gpu.get_field(n), getattr(o, n))
        assert all(_same(n) for n in ("pper", "ppar", "sigma", "dPPerdTheta", "dBsqdRho"))
    # independent checks of the oracle's pieces: one Savitzky-Golay pass (interior points) and the bilinear rule
    r2, az, per0, _ = o.pressure_raw(PPerT, PParT, scb, LZ, PHI, PressMode="FLT", iSm2=0)
    _, _, per1, _ = o.pressure_raw(PPerT, PParT, scb, LZ, PHI, PressMode="FLT", iSm2=1, SavGolIters=1)
    c = savgol_coeffs(7, 2)
    rad_pass = per0.copy()
    for j in range(3, per0.shape[0] - 3):
        rad_pass[j] = sum(c[m] * per0[j - 3 + m] for m in range(7))
    want = sum(c[m] * rad_pass[5:-5, 10 - 3 + m] for m in range(7))
    assert np.max(np.abs(per1[5:-5, 10] - want)) <= 1e-13 * np.abs(want).max()
    o.pressure_raw(PPerT, PParT, scb, LZ, PHI, PressMode="FLT", iSm2=0)
    pe, _ = o.pressure_front()
    ieq = (sinp.nthe + 1) // 2 - 1
    xe, ye = o.x[ieq][:, 1:-1], o.y[ieq][:, 1:-1]
    ang = np.where(xe > 0, np.arcsin(ye / np.hypot(xe, ye)) + np.pi, np.where(ye >= 0, 2 * np.pi - np.arcsin(ye / np.hypot(xe, ye)),
                                                                              -np.arcsin(ye / np.hypot(xe, ye))))
    rgi = RegularGridInterpolator((r2, az), per0, bounds_error=False, fill_value=None)
    want = rgi(np.stack([(xe ** 2 + ye ** 2).ravel(), ang.ravel()], axis=1)).reshape(xe.shape) / o.get("pnormal")
    far = np.hypot(xe, ye) >= 2.0                                      # inside 2 RE the extap repair takes over
    sel = far & (want > 0)
    assert np.max(np.abs(pe[:, 1:-1][sel] - want[sel]) / want[sel]) <= 1e-12
    # the whole outer iteration without a host callback
    o.pressure_raw(PPerT, PParT, scb, LZ, PHI)
    gpu.set_ram_pressure(PPerT, PParT, scb, LZ, PHI)
    gpu.set_map_targets(sinp.alphaVal, sinp.psiVal, sinp.chiVal)
    kw = dict(numit=2, MinSCBIterations=2, decreaseConvAlpha=1e-30, decreaseConvPsi=1e-30)
    ro = o.scb_run(None, **kw)
    rg = gpu.scb_run(None, ordering=emu.SOR_LEX, **kw)
    assert ro["SORFail"] == 0 and rg["SORFail"] == 0 and rg["iterations"] == ro["iterations"] == 2
    for k in ("blendAlpha", "blendPsi", "errorAlpha", "errorPsi", "nisaveAlpha", "nisavePsi"):
        assert rg[k] == ro[k], k
    for n in ("x", "y", "z", "alfa", "psi", "pper", "sigma"):
        assert np.array_equal(gpu.get_field(n), getattr(o, n)), n
    gpu.close()


def test_flc_radius_on_device(emu):
    """R12, FLC_Radius (src/ModRamLoss.f90:176-336): field-line curvature radius and zeta parameters from the resident SCB
    geometry / field, interpolated to the RAM points by the 9-nearest-neighbour rule.  Device (emulated) BIT-IDENTICAL to
    the oracle; known answer: on dipole field lines the equatorial curvature radius is r / 3."""
    from oracle import oracle
    from ramscb_b200 import grids, scb_synthetic as S
    g = grids.build_grids()
    radRaw = 1.75 + (6.75 - 1.75) * np.arange(1, g.NR + 1) / g.NR          # radRaw(1:nR), src/ModRamEField.f90:100-103
    azimRaw = 24.0 * np.arange(g.NT) / (g.NT - 1)
    for warp in (0.3, 0.0):
        sinp = S.build_scb(nthe=51, npsi=23, nzeta=49, warp=warp)
        o, gpu = oracle.ScbOracle(sinp), emu.ScbGpu(sinp)
        with pytest.raises(emu.RsgError, match="computeBandJacob"):
            gpu.FLC_Radius(radRaw, azimRaw)
        o.bandjacob(); gpu.computeBandJacob()
        ref = oracle.flc_radius(o.x, o.y, o.z, o.Bx, o.By, o.Bz, radRaw, azimRaw, (sinp.nthe + 1) // 2, o.get("bnormal"))
        got = gpu.FLC_Radius(radRaw, azimRaw)
        for a, b in zip(got, ref):
            assert np.array_equal(a, b)
            assert np.array_equal(b[:, -1], b[:, 0]) and np.array_equal(b[0], b[1])
        if warp == 0.0:
            rc = ref[0][1:, :-1] / 6.4e6                     # in RE
            assert np.max(np.abs(rc / (radRaw[1:, None] / 3.0) - 1.0)) < 0.05, np.max(np.abs(rc / (radRaw[1:, None] / 3.0) - 1.0))
        gpu.close()


def test_geosb_and_electric_field_on_device(emu, small_grids):
    """GEOSB (src/ModRamBoundary.f90:241-319) and get_electric_field (src/ModRamEField.f90:14-63) on the
    device (emulated), BIT-IDENTICAL to the oracle; the Volland-Stern branch also against ram_run's own formula
    (synthetic.volland_stern, src/ModRamRun.f90:45-51)."""
    from oracle import oracle
    from ramscb_b200 import synthetic
    g = small_grids
    inp = synthetic.make_inputs(g, f2_kind="smooth")
    o = oracle.RamOracle(g, inp)
    gpu = emu.RamGpu(g)
    gpu.set_inputs(inp)
    rng = np.random.default_rng(5)
    flux = np.asfortranarray(10.0 ** (2 + 3 * rng.random((g.NT, g.NE))))
    for S, comp in ((1, 0.7), (4, 1.0)):
        o.geosb(S, flux, comp)
        gpu.GEOSB(S, flux, comp)
        got = gpu.get_boundary(S)
        assert np.array_equal(got, o.FGEOS[S - 1]) and (got != 0).sum() > 100
        assert np.array_equal(got[0], got[-1])                         # J = 1 carries the J = NT flux
    # the sweeps see the new boundary: one DRIFTR against the oracle
    for S in (1, 4):
        o.op("driftpara", S); o.op("driftr", S)
        gpu.DRIFTPARA(S, 5.0); gpu.DRIFTR(S)
    assert np.array_equal(gpu.f2_d2h()[[0, 3]], o.F2[[0, 3]])
    VTOL, VTN = inp.VT.copy(order="F"), np.asfortranarray(1.3 * inp.VT + 5.0)
    o.get_electric_field(False, VTOL=VTOL, VTN=VTN, t=450.0, TOLV=300.0, DtEfi=300.0)
    assert np.array_equal(gpu.get_electric_field(False, VTOL=VTOL, VTN=VTN, t=450.0, TOLV=300.0, DtEfi=300.0), o.VT)
    o.set_scalar("Kp", 4.3)
    o.get_electric_field(True, PHI=g.PHI[:g.NT], PHIOFS=0.1)
    vt = gpu.get_electric_field(True, Kp=4.3, PHI=g.PHI[:g.NT], PHIOFS=0.1)
    assert np.array_equal(vt, o.VT)
    assert np.allclose(vt, synthetic.volland_stern(g, 4.3, 0.1), rtol=1e-14, atol=0)
    gpu.close()
