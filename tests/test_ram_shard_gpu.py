"""The multi-GPU RAM step of the library (rsg_ram_run_sharded: peer-memory write-backs, device-side barriers, one CUDA
graph per rank) on ONE device: `world` handles of one process are wired together with rsg_ram_peer_attach_local and run
in lock-step (enqueue every rank, then collect every rank).  Same kernels, barriers and graphs as the multi-process path
of tests/multi_gpu_check.py, which differs only in how the peer pointers are obtained (CUDA IPC).

Bars: the re-assembled F2 of the ranks is BIT-IDENTICAL to rsg_ram_run on one handle (the kernels do the same per-cell
arithmetic; only the owner of a cell changes), CFL limits identical, SUMRC moments / pressures <= 1e-12 (partial sums
are added in rank order instead of block order), and every rank returns the same bits."""
import os

import numpy as np
import pytest

from ramscb_b200 import grids, synthetic

pytestmark = pytest.mark.gpu

STEPS = (5.0, 7.5, 7.5)


def _single(g, inp, mode, flags, D=None):
    from ramscb_b200 import host
    ref = host.RamGpu(g, mode=mode)
    ref.set_inputs(inp)
    if D is not None:
        ref.set_diffcoef(1, D)
        ref.set_diffcoef(2, D)
    outs = [ref.ram_run(dts, flags=flags) for dts in STEPS]
    F = ref.f2_d2h()
    ref.close()
    return F, outs


def _sharded(g, inp, mode, flags, world, policy, D=None):
    from ramscb_b200 import host
    ranks = []
    for r in range(world):
        gpu = host.RamGpu(g, mode=mode)
        gpu.set_fields(inp)
        gpu.set_efield(inp.VT, inp.EIR, inp.EIP)
        gpu.set_boundary(inp.FGEOS)
        gpu.set_wavelo(inp.WALOS1, inp.WALOS2, inp.WALOS3, inp.Kp, inp.Kpmax12)
        gpu.set_plasmasphere(inp.NECR)
        if D is not None:
            gpu.set_diffcoef(1, D)
            gpu.set_diffcoef(2, D)
        ranks.append(gpu)
    for r, gpu in enumerate(ranks):
        gpu.peer_attach_local(r, ranks, policy)
        gpu.f2_h2d_shard(inp.F2)               # only the rank's share goes up
    outs = []
    for dts in STEPS:
        for gpu in ranks:
            gpu.run_sharded_enqueue(dts, flags=flags)
        outs.append([gpu.run_sharded_collect() for gpu in ranks])
    F = np.full(inp.F2.shape, np.nan, order="F")
    for gpu in ranks:
        gpu.f2_d2h_shard(F)
    plans = [gpu.shard_info().as_dict() for gpu in ranks]
    for gpu in ranks:
        gpu.close()
    return F, outs, plans


def _check(F1, o1, FN, oN, plans):
    assert not np.isnan(FN).any(), "the ranks' shares do not cover F2"
    assert np.array_equal(FN, F1), f"F2 differs in {int((FN != F1).sum())} cells"
    for step, per_rank in enumerate(oN):
        a = per_rank[0]
        for b in per_rank[1:]:
            for k in ("DtDrift", "losses", "SETRC", "PPERT", "PPART"):
                assert np.array_equal(a[k], b[k]), f"ranks disagree on {k}"
            assert a["DtsNext"] == b["DtsNext"]
        r = o1[step]
        assert np.array_equal(a["DtDrift"], r["DtDrift"]) and a["DtsNext"] == r["DtsNext"]
        np.testing.assert_allclose(a["PPERT"], r["PPERT"], rtol=1e-12, atol=0)
        np.testing.assert_allclose(a["PPART"], r["PPART"], rtol=1e-12, atol=0)
        np.testing.assert_allclose(a["SETRC"], r["SETRC"], rtol=1e-12, atol=0)
        if step > 0:     # the first step's increments start from SETRC = 0 on both sides
            np.testing.assert_allclose(a["losses"], r["losses"], rtol=1e-9, atol=1e-12 * np.abs(r["SETRC"]).max())


@pytest.mark.parametrize("world,policy", [(2, 0), (4, 0), (8, 0), (2, 1), (4, 1), (8, 1), (3, 1)])
@pytest.mark.parametrize("flags", [0, 5])
def test_sharded_step_matches_one_gpu_fast(default_grids, world, policy, flags):
    from ramscb_b200 import host
    g = default_grids
    inp = synthetic.make_inputs(g, f2_kind="noisy", inductive=True, mgnp=True)
    D = synthetic.synthetic_daa(g, inp) if flags else None
    F1, o1 = _single(g, inp, host.MODE_FAST, flags, D)
    FN, oN, plans = _sharded(g, inp, host.MODE_FAST, flags, world, policy, D)
    _check(F1, o1, FN, oN, plans)


@pytest.mark.parametrize("world,policy", [(8, 0), (4, 1)])
def test_sharded_step_with_coulomb_stages(default_grids, world, policy):
    """flags 7: WPI + Coulomb + EMIC, ranks sharing a species: the Coulomb stages run inside the sharded column kernel,
    their four moments are reduced with the others"""
    from ramscb_b200 import host
    g = default_grids
    inp = synthetic.make_inputs(g, f2_kind="noisy", inductive=True, mgnp=True)
    D = synthetic.synthetic_daa(g, inp)
    F1, o1 = _single(g, inp, host.MODE_FAST, 7, D)
    FN, oN, plans = _sharded(g, inp, host.MODE_FAST, 7, world, policy, D)
    _check(F1, o1, FN, oN, plans)


@pytest.mark.parametrize("push", ["1", "2"])
@pytest.mark.parametrize("world,policy", [(8, 1), (8, 0), (3, 1)])
def test_sharded_step_bulk_push_of_the_resharding(default_grids, monkeypatch, world, policy, push):
    """RSG_PEER_PUSH=1: the column kernel writes its block range locally and k_peer_push sends every (l, k) row to the
    pitch angle's owner as one contiguous run; =2: both re-shardings that way, in chunks, the push of a chunk on a second
    stream beside the kernel of the next (fork / join inside the step's graph) -- same bits as the kernels' own peer
    write-backs"""
    from ramscb_b200 import host
    monkeypatch.setenv("RSG_PEER_PUSH", push)
    g = default_grids
    inp = synthetic.make_inputs(g, f2_kind="noisy", inductive=True, mgnp=True)
    D = synthetic.synthetic_daa(g, inp)
    F1, o1 = _single(g, inp, host.MODE_FAST, 5, D)
    FN, oN, plans = _sharded(g, inp, host.MODE_FAST, 5, world, policy, D)
    _check(F1, o1, FN, oN, plans)


def test_sharded_step_ragged_grid():
    """odd sizes: NR odd (8-byte staging path of the plane kernel), slabs and column ranges that do not divide evenly"""
    from ramscb_b200 import host
    g = grids.build_grids(NR=11, NT=9, NE=13, NPA=72)
    inp = synthetic.make_inputs(g, f2_kind="noisy", inductive=True)
    F1, o1 = _single(g, inp, host.MODE_FAST, 0)
    for world, policy in ((8, 0), (5, 1)):
        FN, oN, plans = _sharded(g, inp, host.MODE_FAST, 0, world, policy)
        _check(F1, o1, FN, oN, plans)
    for push in ("1", "2"):                      # odd column cuts: the scalar head / tail of the bulk push
        os.environ["RSG_PEER_PUSH"] = push
        try:
            FN, oN, plans = _sharded(g, inp, host.MODE_FAST, 0, 5, 1)
            _check(F1, o1, FN, oN, plans)
        finally:
            del os.environ["RSG_PEER_PUSH"]


@pytest.mark.parametrize("world", [2, 4])
def test_species_sharded_exact_mode(default_grids, world):
    """whole species per rank: any mode and flag set (here EXACT + Coulomb: one kernel per operator), results gathered
    on the device"""
    from ramscb_b200 import host
    g = default_grids
    inp = synthetic.make_inputs(g, f2_kind="noisy", inductive=True)
    F1, o1 = _single(g, inp, host.MODE_EXACT, host.F_COULOMB)
    FN, oN, plans = _sharded(g, inp, host.MODE_EXACT, host.F_COULOMB, world, host.SHARD_SPECIES)
    _check(F1, o1, FN, oN, plans)


def test_sharded_errors(default_grids):
    from ramscb_b200 import host
    g = default_grids
    inp = synthetic.make_inputs(g, f2_kind="smooth")
    a, b = host.RamGpu(g), host.RamGpu(g)
    with pytest.raises(host.RsgError, match="before rsg_ram_peer_attach"):
        a.run_sharded(5.0)
    for r, gpu in enumerate((a, b)):
        gpu.set_inputs(inp)
        gpu.peer_attach_local(r, [a, b], host.SHARD_SLABS)
    with pytest.raises(host.RsgError, match="fused FAST"):          # EXACT mode cannot share a species
        a.run_sharded(5.0)
    with pytest.raises(host.RsgError, match="already attached"):
        a.peer_attach_local(0, [a, b], host.SHARD_SLABS)
    a.close(); b.close()


@pytest.mark.parametrize("pipes", ["1", "2"])
def test_multi_gpu_peer_memory(pipes):
    """>= 2 GPUs: one process per GPU, peer pointers through CUDA IPC (tests/multi_gpu_peer_check.py).  Skipped on a
    1-GPU box, where the in-process tests above run the same kernels, barriers and graphs.  pipes = 2: the rank's species
    as two pipelines on two streams (RSG_PEER_PIPES; own barrier ids, fork / join in the step's graph) -- only testable
    with one process per GPU: several in-process "ranks" with forked graphs on ONE device stall each other's barriers."""
    import os
    import subprocess
    import sys
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs 2 GPUs")
    n = 8 if n >= 8 else (4 if n >= 4 else 2)
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr",
                        "127.0.0.1", "--master-port", "2951" + pipes, os.path.join(root, "tests", "multi_gpu_peer_check.py")],
                       capture_output=True, text=True, timeout=900, env=dict(os.environ, RSG_PEER_PIPES=pipes))
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "MULTI_GPU_PEER_CHECK_OK" in r.stdout
