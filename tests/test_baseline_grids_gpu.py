"""Parity on BASELINE.json's own grids (the other parity files use reduced grids where the oracle is needed):
  * SCB, configs[3] grid 101 x 45 x 97: metrica / newk / metric / newj bit-exact, the lexicographic SOR bit-exact
    (potentials, sweep counts, residual maxima), the production 4-colour cluster kernel within 1e-8 of the converged
    reference-order solution with a RELATIVE denominator, Compute_convergence, and three outer iterations of rsg_scb_run;
  * RAM, configs[2] grid 80 x 49 x 70 x 72 with WPI + EMIC pitch-angle diffusion: one ram_run in EXACT mode (<= 1e-12
    strict, CFL limits bit-equal) and one in FAST mode (the production path; strict per-cell bar of test_ram_parity_gpu).
The oracle needs ~10-20 s of CPU per RAM step at this size and ~5 s for the SCB outer iterations."""
import numpy as np
import pytest

from ramscb_b200 import grids, scb_synthetic as S, synthetic

pytestmark = pytest.mark.gpu

DEFAULT_SCB = dict(nthe=101, npsi=45, nzeta=97, warp=0.2)
VECS = ("vecd", "vec1", "vec2", "vec3", "vec4", "vec6", "vec7", "vec8", "vec9")


def _same(gpu, o, names):
    for n in names:
        a, b = gpu.get_field(n), getattr(o, n)
        bad = int(np.sum(a != b))
        assert bad == 0, f"{n}: {bad} of {a.size} entries differ (max abs {np.max(np.abs(a - b)):.3e})"


@pytest.fixture(scope="module")
def scb_pair(oracle_built):
    from ramscb_b200.host import ScbGpu
    inp = S.build_scb(**DEFAULT_SCB)
    return inp, oracle_built.ScbOracle(inp), ScbGpu(inp)


def test_scb_default_grid_coefficients_and_lex_sor(scb_pair):
    inp, o, gpu = scb_pair
    o.bandjacob(); gpu.computeBandJacob()
    o.metrica(); o.newk(); gpu.metrica(); gpu.newk()
    _same(gpu, o, VECS + ("vecx",))
    fail, ni = o.iterate_alpha()
    r = gpu.iterateAlpha(1e-6, ordering=0)
    assert fail == 0 and r["SORFail"] == 0 and np.array_equal(r["ni"], ni) and ni.max() > 50
    assert r["nisave"] == int(o.get("nisave")) and r["diffmx"] == o.get("diffmx")
    assert np.array_equal(gpu.get_field("alfa"), o.alfa)
    o.metric(); o.newj(); gpu.metric(); gpu.newj()
    _same(gpu, o, VECS + ("vecr",))
    fail, ni = o.iterate_psi()
    r = gpu.iteratePsi(1e-6, ordering=0)
    assert fail == 0 and r["SORFail"] == 0 and np.array_equal(r["ni"], ni)
    assert r["diffmx"] == o.get("diffmx")
    assert np.array_equal(gpu.get_field("psi"), o.psi)
    # Compute_convergence on the solved state
    o.bandjacob(); gpu.computeBandJacob()
    assert o.convergence() == 0
    rc = gpu.Compute_convergence()
    assert rc["SORFail"] == 0
    _same(gpu, o, ("jGradRho", "jGradZeta", "jGradTheta", "Jx", "Jy", "Jz", "GradPx", "GradPy", "GradPz", "jCrossB", "GradP"))
    for n in ("normDiff", "normJxB", "normGradP"):
        assert abs(rc[n] - o.get(n)) <= 1e-12 * max(abs(o.get(n)), 1e-300), n


def test_scb_default_grid_color4_cluster_kernel_vs_oracle(oracle_built):
    """the production kernel of configs[3] (k_scb_sor_cluster_reg: 4- and 2-CTA clusters) against the oracle's
    reference-order solve, both converged tightly.  Denominator: the point's own |value|, floored at 1 % of the
    field's largest magnitude (alfa passes through zero)."""
    from ramscb_b200.host import ScbGpu
    inp = S.build_scb(**DEFAULT_SCB)
    o, gpu = oracle_built.ScbOracle(inp), ScbGpu(inp)
    o.bandjacob(); gpu.computeBandJacob()
    o.metrica(); o.newk(); gpu.metrica(); gpu.newk()
    # 1e-11 is below the rounding floor of the residual on this grid (|alfa| ~ 6, |vecd| ~ 1e3: the 4-colour solve runs
    # into nimax); 1e-10 is reached by both orderings
    o.set_scalar("InConAlpha", 1e-10)
    fail, ni = o.iterate_alpha()
    r = gpu.iterateAlpha(1e-10, ordering=1)
    print(f"\nalpha: oracle sweeps {int(ni.max())}, 4-colour cluster sweeps {r['nisave']}")
    assert fail == 0 and r["SORFail"] == 0 and r["nisave"] < 5001 and ni.max() < 5001
    assert gpu.last_cluster() == 3
    a, b = gpu.get_field("alfa"), o.alfa
    rel = np.abs(a - b) / np.maximum(np.abs(b), 1e-2 * np.abs(b).max())
    assert rel.max() <= 1e-8, rel.max()
    o.metric(); o.newj(); gpu.metric(); gpu.newj()
    o.set_scalar("InConPsi", 1e-9)            # |psi| ~ 1e2, |vecd| ~ 1e3: 1e-11 is below the rounding floor of the residual
    fail, ni = o.iterate_psi()
    r = gpu.iteratePsi(1e-9, ordering=1)
    assert fail == 0 and r["SORFail"] == 0 and r["nisave"] < 5001
    assert gpu.last_cluster() == 2
    a, b = gpu.get_field("psi"), o.psi
    rel = np.abs(a - b) / np.maximum(np.abs(b), 1e-2 * np.abs(b).max())
    assert rel.max() <= 1e-8, rel.max()
    gpu.close()


def test_scb_default_grid_three_outer_iterations(oracle_built):
    """rsg_scb_run on the configs[3] grid, reference sweep order: same decisions and bit-identical state as the
    oracle's composition of the same loop after three outer iterations"""
    from ramscb_b200 import host
    inp = S.build_scb(**DEFAULT_SCB)
    o, gpu = oracle_built.ScbOracle(inp), host.ScbGpu(inp)
    fn = S.equatorial_pressure_fn()
    kw = dict(numit=3, MinSCBIterations=2, decreaseConvAlpha=1e-30, decreaseConvPsi=1e-30)
    gpu.set_map_targets(inp.alphaVal, inp.psiVal, inp.chiVal)
    ro = o.scb_run(fn, **kw)
    rg = gpu.scb_run(fn, ordering=host.SOR_LEX, **kw)
    assert ro["SORFail"] == 0 and rg["SORFail"] == 0 and rg["iterations"] == ro["iterations"] == 3
    for k in ("blendAlpha", "blendPsi", "errorAlpha", "errorPsi", "nisaveAlpha", "nisavePsi", "blendRetries"):
        assert rg[k] == ro[k], (k, rg[k], ro[k])
    _same(gpu, o, ("x", "y", "z", "alfa", "psi", "jacobian", "bsq", "pper", "sigma"))
    for a, b in zip((rg["normDiff"], rg["normJxB"], rg["normGradP"]), ro["norm"]):
        assert abs(a - b) <= 1e-12 * abs(b)
    gpu.close()


# ---- RAM on the configs[2] grid -------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def ram_x4(oracle_built):
    g = grids.build_grids(NR=80, NT=49, NE=70, energy_refine=2)
    inp = synthetic.make_inputs(g, f2_kind="noisy", inductive=True)
    D = synthetic.synthetic_daa(g, inp)
    o = oracle_built.RamOracle(g, inp, DTs=5.0)
    o.set_array("ATAC", D)
    o.set_array("ATAW_emic_h", D)
    dtn = o.ram_run(flags=5)
    return g, inp, D, o, dtn


@pytest.mark.parametrize("mode", ["exact", "fast"])
def test_ram_configs2_grid_full_step_vs_oracle(ram_x4, mode):
    from ramscb_b200 import host
    g, inp, D, o, dtn = ram_x4
    gpu = host.RamGpu(g, mode=host.MODE_EXACT if mode == "exact" else host.MODE_FAST)
    gpu.set_inputs(inp)
    gpu.set_diffcoef(1, D)
    gpu.set_diffcoef(2, D)
    n0 = gpu.launch_count()
    out = gpu.ram_run(5.0, DtsMin=1.0, flags=5)
    launches = gpu.launch_count() - n0
    got = gpu.f2_d2h()
    gpu.close()
    strict = np.abs(got - o.F2) / np.maximum(np.abs(o.F2), 1e-300)
    n = int((strict > 1e-12).sum())
    print(f"\nconfigs[2] grid, {mode}: strict per-cell max {strict.max():.2e}, cells > 1e-12: {n} of {strict.size}; launches {launches}")
    if mode == "exact":
        assert strict.max() <= 1e-12
        assert out["DtsNext"] == dtn
        DtR, DtP, DtE, DtM = o.DtDriftR, o.DtDriftP, o.DtDriftE, o.DtDriftMu
        assert np.array_equal(out["DtDrift"], np.stack([DtR, DtP, DtE, DtM]))
    else:
        assert launches <= 16, "the fused path was not taken"        # 6 per step + the one-off table / CFL / inflow kernels
        assert n <= max(20, int(1e-5 * strict.size)) and strict.max() <= 1e-11
        assert abs(out["DtsNext"] - dtn) <= 1e-13 * dtn
    # SUMRC adds 19.8 M terms per species: the reference's serial sum and the tree sum each carry ~ sqrt(N) ulp
    print(f"   SETRC rel diff {np.max(np.abs(out['SETRC'] - o.SETRC) / np.abs(o.SETRC)):.2e}")
    assert np.allclose(out["SETRC"], o.SETRC, rtol=2e-11, atol=0)
    assert np.max(np.abs(out["PPERT"][:, 1:] - o.PPERT[:, 1:]) / np.maximum(np.abs(o.PPERT[:, 1:]), 1e-300)) <= 1e-12
    assert np.max(np.abs(out["PPART"][:, 1:] - o.PPART[:, 1:]) / np.maximum(np.abs(o.PPART[:, 1:]), 1e-300)) <= 1e-12


# ---- RAM at the size of configs[4]'s grid (8 x the default: 4 x NR, 4 x NT, 2 x NE = 156 M cells, 1.25 GB) -----------
def test_ram_configs4_size_properties(dims=(80, 97, 70, 2)):
    """No oracle run at this size (the CPU restatement needs minutes per step): the size-independent properties the path
    offers instead.  (i) default operators: the fused FAST step and the one-kernel-per-operator FAST step give bit-identical
    F2 and CFL limits (same per-cell arithmetic, different kernels, launch shapes and staging); (ii) flags 5: rsg_ram_run_host
    (12 pipelined chunks) equals the three calls bit for bit; (iii) F2 stays positive and finite, J = NT repeats J = 1; (iv) a checksum of
    checksums: the SETRC the device returns equals the moment recomputed on the host from the returned F2
    (src/ModRamRun.f90:246-253) to 1e-11."""
    from ramscb_b200 import host
    g = grids.build_grids(NR=dims[0], NT=dims[1], NE=dims[2], energy_refine=dims[3])
    inp = synthetic.make_inputs(g, f2_kind="noisy", inductive=True)
    D = synthetic.synthetic_daa(g, inp)
    res = {}
    for name, flags in (("unfused", 0), ("fused0", 0), ("fused", 5), ("host", 5)):
        gpu = host.RamGpu(g, mode=host.MODE_FAST)
        gpu.set_inputs(inp)
        gpu.set_diffcoef(1, D)
        gpu.set_diffcoef(2, D)
        gpu.use_fused(name != "unfused")
        if name == "host":
            F = inp.F2.copy(order="F")
            out = gpu.ram_run_host(F, 5.0, DtsMin=1.0, flags=flags)
        else:
            out = gpu.ram_run(5.0, DtsMin=1.0, flags=flags)
            F = gpu.f2_d2h()
        res[name] = (F, out)
        gpu.close()
    Fu, ou = res.pop("unfused")
    F0, o0 = res.pop("fused0")
    # (with WPADIF the fused stage applies tabulated Thomas factors: bit-identity holds for the default operator set)
    assert np.array_equal(Fu, F0), f"fused vs unfused: {int((Fu != F0).sum())} of {Fu.size} cells differ"
    assert np.array_equal(ou["DtDrift"], o0["DtDrift"]) and ou["DtsNext"] == o0["DtsNext"]
    del Fu, F0
    Ff, of = res["fused"]
    Fh, oh = res["host"]
    assert np.array_equal(Ff, Fh) and np.array_equal(of["SETRC"], oh["SETRC"]) and np.array_equal(of["PPERT"], oh["PPERT"])
    assert np.all(np.isfinite(Ff)) and Ff.min() > 0.0
    assert np.array_equal(Ff[:, :, -1], Ff[:, :, 0])
    assert not np.array_equal(Ff, inp.F2)
    # SUMRC on the host: sum over I = 2..NR, J = 1..NT-1, K = 2..NE, L = 2..NPA of F2 * WE(K) * WMU(L) * EKEV(K)
    w = (g.WE[1:g.NE] * g.EKEV[1:g.NE])[:, None] * g.WMU[1:g.NPA][None, :]
    for s in range(g.nS):
        want = float(np.einsum("ijkl,kl->", Ff[s, 1:, :-1, 1:, 1:], w))
        assert abs(of["SETRC"][s] - want) <= 1e-11 * abs(want), (s, of["SETRC"][s], want)
