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


@pytest.mark.parametrize("flags", [1 | 4, 1])
def test_emulated_fused_wpadif_step(emu, T, small_grids, oracle_built, flags):
    T.test_fused_wpadif_fast_step(small_grids, oracle_built, "default", flags)


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


def test_emulated_scb_run_outer_iterations(emu, oracle_built):
    """rsg_scb_run -- the whole outer iteration of scb_run in one C call, 3-D arrays resident, pressure front
    end as a host callback -- against the oracle's composition, incl. the SORFail restore path."""
    import test_zz_late_additions_gpu as TZ
    TZ.test_scb_run_outer_iterations_resident(oracle_built)


def test_emulated_results_do_not_depend_on_thread_order():
    """Race check: EMU_ORDER=random runs the runnable threads of every block in a fresh random order
    between synchronisation points.  A kernel whose result depended on the order (a missing barrier
    between a producer and a consumer stage) would break the bit-exact comparisons of the fused-step
    tests; they must pass unchanged."""
    import subprocess
    env = dict(os.environ, EMU_ORDER="random")
    here = os.path.dirname(os.path.abspath(__file__))
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(here, "test_emu_cpu.py"), "-x", "-q", "-k",
                        "fused_equals or (fused_wpadif and 5) or exact_sweeps or scb_maps or hI_integrals or hI_tail or hI_convert"], env=env, capture_output=True, text=True, cwd=os.path.dirname(here))
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
