"""Host-side logic of the multi-GPU path on CPU: the sharding plan, and the (L,K)-slab
re-sharding exchanged over a world_size-2 gloo process group (the same code path the
nccl backend drives on GPUs: ramscb_b200.parallel.exchange)."""
import os
import socket
import sys

import numpy as np
import pytest

from ramscb_b200 import parallel


def test_plan_species_sharding():
    for world in (1, 2, 4):
        owned = []
        for r in range(world):
            p = parallel.make_plan(world, r, 4, 72, 35)
            assert p.G == 1 and (p.l0, p.nl, p.k0, p.nk) == (0, 72, 0, 35)
            owned += list(range(p.s0, p.s0 + p.ns))
        assert owned == [0, 1, 2, 3]
    with pytest.raises(ValueError):
        parallel.make_plan(3, 0, 4, 72, 35)


def test_plan_slab_sharding_covers_everything():
    for world, NE in ((8, 35), (16, 70)):
        G = world // 4
        for s in range(4):
            Ls, Ks = [], []
            for r in range(s * G, (s + 1) * G):
                p = parallel.make_plan(world, r, 4, 72, NE)
                assert (p.s0, p.ns) == (s, 1) and p.group == tuple(range(s * G, (s + 1) * G))
                Ls += list(range(p.l0, p.l0 + p.nl))
                Ks += list(range(p.k0, p.k0 + p.nk))
            assert Ls == list(range(72)) and Ks == list(range(NE))
        # what a sends to b is what b receives from a
        pa, pb = parallel.make_plan(world, 0, 4, 72, NE), parallel.make_plan(world, 1, 4, 72, NE)
        for to_k in (True, False):
            sa = [b for b in parallel.exchange_blocks(pa, to_k) if b[0] == 1][0]
            sb = [b for b in parallel.exchange_blocks(pb, to_k) if b[0] == 0][0]
            assert sa[1] == sb[2] and sa[2] == sb[1]


def test_small_grids_leave_extra_ranks_idle():
    """A species is split among ranks only above parallel.SPLIT_MIN_CELLS cells: below it the
    first nS ranks take one species each and the others hold nothing (bench.py at N = 8 on the
    default and 4x grids); at or above it the species x slab groups are used."""
    small, big = parallel.SPLIT_MIN_CELLS - 1, parallel.SPLIT_MIN_CELLS
    plans = [parallel.make_plan(8, r, 4, 72, 35, cells_per_species=small) for r in range(8)]
    assert [p.ns for p in plans] == [1, 1, 1, 1, 0, 0, 0, 0]
    assert [p.s0 for p in plans[:4]] == [0, 1, 2, 3]
    assert all(p.G == 1 and p.active == (0, 1, 2, 3) and (p.l0, p.nl, p.k0, p.nk) == (0, 72, 0, 35) for p in plans)
    plans = [parallel.make_plan(8, r, 4, 72, 35, cells_per_species=big) for r in range(8)]
    assert all(p.G == 2 and p.ns == 1 and len(p.active) == 8 for p in plans)
    # never idle when there are enough species
    assert all(parallel.make_plan(4, r, 4, 72, 35, cells_per_species=small).ns == 1 for r in range(4))
    assert all(parallel.make_plan(2, r, 4, 72, 35, cells_per_species=small).ns == 2 for r in range(2))


def _worker(rank, world, port, NPA, NE, Pp, q):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    # one species shared by the two ranks (nS = 1, world = 2)
    p = parallel.make_plan(world, rank, 1, NPA, NE)
    full = torch.arange(NPA * NE * Pp, dtype=torch.float64)          # the global species block [L][K][Pp]
    buf = torch.full_like(full, -1.0)
    v = buf.view(NPA, NE, Pp)
    v[p.l0:p.l0 + p.nl] = full.view(NPA, NE, Pp)[p.l0:p.l0 + p.nl]   # L-slab layout: I hold my pitch angles, all energies
    parallel.exchange(p, [buf], Pp, True, dist)
    ok1 = bool(torch.equal(v[:, p.k0:p.k0 + p.nk], full.view(NPA, NE, Pp)[:, p.k0:p.k0 + p.nk]))  # all L of my energies
    # "pitch-angle block": touch my K-slab, then go back
    v[:, p.k0:p.k0 + p.nk] += 0.5
    parallel.exchange(p, [buf], Pp, False, dist)
    ok2 = bool(torch.equal(v[p.l0:p.l0 + p.nl], full.view(NPA, NE, Pp)[p.l0:p.l0 + p.nl] + 0.5))  # my L-slab, all energies updated
    # the fused step's re-sharding: pitch-angle slabs <-> ranges of plane positions (blocks of 4)
    nblocks, per = (Pp - 3 + 3) // 4, 4          # P = Pp - 3 positions in use, as a ragged case
    buf2 = torch.full_like(full, -1.0)
    w = buf2.view(NPA * NE, Pp)
    fw = full.view(NPA * NE, Pp)
    w[p.l0 * NE:(p.l0 + p.nl) * NE] = fw[p.l0 * NE:(p.l0 + p.nl) * NE]      # my rows, all positions
    parallel.exchange_lp(p, [buf2], Pp, nblocks, per, True, dist)
    b0, nb = parallel._split(nblocks, p.G, p.gidx)
    c0, c1 = b0 * per, min((b0 + nb) * per, Pp)
    ok3 = bool(torch.equal(w[:, c0:c1], fw[:, c0:c1]))                        # all rows of my positions
    w[:, c0:c1] += 0.25
    parallel.exchange_lp(p, [buf2], Pp, nblocks, per, False, dist)
    last = nblocks * per                                                      # columns beyond the last block are padding
    ok4 = bool(torch.equal(w[p.l0 * NE:(p.l0 + p.nl) * NE, :last], fw[p.l0 * NE:(p.l0 + p.nl) * NE, :last] + 0.25))
    q.put((rank, ok1 and ok3, ok2 and ok4))
    dist.destroy_process_group()


def test_slab_exchange_gloo_world2():
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, 10, 7, 16, q)) for r in range(2)]
    for pr in procs:
        pr.start()
    res = [q.get(timeout=120) for _ in procs]
    for pr in procs:
        pr.join(timeout=60)
    assert sorted(res) == [(0, True, True), (1, True, True)]


def _zeta_worker(rank, world, port, q):
    """One rank of the zeta-sharded iterateAlpha: the real kernels (host-CPU emulator build, test
    infrastructure) + the real ScbZetaSharded protocol over gloo."""
    import numpy as np
    import torch.distributed as dist
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import conftest
    conftest.use_emulator()
    from ramscb_b200 import host, scb_synthetic as S
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    inp = S.build_scb(nthe=41, npsi=13, nzeta=22, warp=0.2)      # 21 relaxed planes: 7+7+7 / 11+10, edges of both parities
    ref = host.ScbGpu(inp)
    ref.computeBandJacob(); ref.metrica(); ref.newk()
    r1 = ref.iterateAlpha(1e-6, ordering=host.SOR_COLOR4)
    g = host.ScbGpu(inp)
    g.computeBandJacob(); g.metrica(); g.newk()
    z = parallel.ScbZetaSharded(g, dist, rank, world, on_cuda=False, poll=7)
    r = z.iterate(1e-6)
    same = bool(np.array_equal(g.get_field("alfa"), ref.get_field("alfa")))
    meta = bool(np.array_equal(r["ni"], r1["ni"]) and r["diffmx"] == r1["diffmx"] and r["sumb"] == r1["sumb"]
                and r["sumdb"] == r1["sumdb"] and r["SORFail"] == r1["SORFail"] == 0 and r["nisave"] == r1["nisave"])
    # a sweep limit below convergence: the loop control (ni = nimax + 1 on exit) is the reference's
    ref.set_field("alfa", inp.alfa); g.set_field("alfa", inp.alfa)
    r2 = ref.iterateAlpha(1e-6, nimax=9, ordering=host.SOR_COLOR4)
    r3 = z.iterate(1e-6, nimax=9)
    capped = bool(np.array_equal(g.get_field("alfa"), ref.get_field("alfa")) and np.array_equal(r3["ni"], r2["ni"])
                  and r3["diffmx"] == r2["diffmx"] and int(r2["ni"].max()) == 10)
    q.put((rank, same, meta, capped, z.messages > 0 or world == 1))
    dist.destroy_process_group()


def _ram_worker(rank, world, port, q):
    """One rank of the sharded RAM step (parallel.RamSharded) over gloo, kernels in the emulator, on a reduced
    ragged grid: nS = 1 -> the two ranks share the species (pitch-angle slabs <-> energy slabs / position ranges,
    two re-shardings per step); nS = 2 -> one species per rank, no data-path exchange.  Same checks as
    tests/multi_gpu_check.py (the NCCL run on hardware)."""
    import numpy as np
    import torch.distributed as dist
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import conftest
    conftest.use_emulator()
    from ramscb_b200 import grids, host, synthetic
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    res = []
    for mode, flags in ((host.MODE_EXACT, 0), (host.MODE_FAST, 0), (host.MODE_FAST, 4)):     # 4: EMIC pitch-angle diffusion (H+)
        for nS in (1, 2):
            g = grids.build_grids(nS=nS, NR=9, NT=11, NE=35)
            inp = synthetic.make_inputs(g, f2_kind="noisy", inductive=True)
            gpu = host.RamGpu(g, mode=mode)
            gpu.set_inputs(inp)
            plan = parallel.make_plan(world, rank, nS, g.NPA, g.NE)
            sh = parallel.RamSharded(gpu, plan, dist, on_cuda=False)
            ref = host.RamGpu(g, mode=mode)
            ref.set_inputs(inp)
            if flags:
                # ranks that share a species run WPADIF as its own kernel on their energy slab; one rank by default
                # runs it inside the fused column kernel with tabulated elimination factors (<= 1e-12 apart):
                # the bit-exact comparison is against the one-rank step with that stage unfused
                D = synthetic.synthetic_daa(g, inp)
                for h in (gpu, ref):
                    h.set_diffcoef(2, D)
                if plan.G > 1:
                    ref.use_fused(True, wpadif=False)
            for dts in (5.0, 7.5):
                out = sh.ram_run(dts, flags=flags)
                r = ref.ram_run(dts, flags=flags)
            mine, full = gpu.f2_d2h(), ref.f2_d2h()
            sl, lsl = slice(plan.s0, plan.s0 + plan.ns), slice(plan.l0, plan.l0 + plan.nl)
            same = bool(np.array_equal(mine[sl][..., lsl], full[sl][..., lsl]))
            dt_ok = bool(np.array_equal(out["DtDrift"], r["DtDrift"]) and out["DtsNext"] == r["DtsNext"])
            pp_ok = bool(np.allclose(out["PPERT"][:, 1:], r["PPERT"][:, 1:], rtol=1e-12, atol=0)
                         and np.allclose(out["PPART"][:, 1:], r["PPART"][:, 1:], rtol=1e-12, atol=0))
            res.append((mode, nS, plan.G, same, dt_ok, pp_ok))
            gpu.close(); ref.close()
    q.put((rank, res))
    dist.destroy_process_group()


def test_ram_sharded_step_gloo_world2():
    """The whole N > 1 RAM path on CPU: RamSharded.ram_run at world_size 2 over gloo with the real kernels
    (emulator): every rank's slab of F2 bit-identical to the one-rank step, CFL steps equal, pressures <= 1e-12."""
    import torch.multiprocessing as mp
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "emu"))
    import build_emu
    build_emu.build()
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_ram_worker, args=(r, 2, port, q)) for r in range(2)]
    for pr in procs:
        pr.start()
    res = dict(q.get(timeout=900) for _ in procs)
    for pr in procs:
        pr.join(timeout=60)
    for rank in (0, 1):
        assert len(res[rank]) == 6
        for mode, nS, G, same, dt_ok, pp_ok in res[rank]:
            assert G == (2 if nS == 1 else 1)
            assert same and dt_ok and pp_ok, (rank, mode, nS, G, same, dt_ok, pp_ok)


def _subproblem_worker(rank, world, port, q):
    """One rank of the sub-problem-sharded iterateAlpha / iteratePsi (parallel.ScbSharded) over gloo, kernels in
    the emulator: bit-identical to the one-rank solves."""
    import numpy as np
    import torch.distributed as dist
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import conftest
    conftest.use_emulator()
    from ramscb_b200 import host, scb_synthetic as S
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    inp = S.build_scb(nthe=41, npsi=14, nzeta=22, warp=0.2)       # 12 psi surfaces, 21 zeta planes: uneven splits
    ref = host.ScbGpu(inp)
    ref.computeBandJacob(); ref.metrica(); ref.newk()
    ra = ref.iterateAlpha(1e-6, ordering=host.SOR_COLOR4)
    ref.metric(); ref.newj()
    rp = ref.iteratePsi(1e-6, ordering=host.SOR_COLOR4)
    g = host.ScbGpu(inp)
    sh = parallel.ScbSharded(g, dist, rank, world, on_cuda=False)
    g.computeBandJacob(); g.metrica(); g.newk()
    ok = True
    for alpha, r1, fld in ((True, ra, "alfa"), (False, rp, "psi")):
        if not alpha:
            g.metric(); g.newj()
        r = sh.iterate(alpha, 1e-6)
        ok = ok and bool(np.array_equal(g.get_field(fld), ref.get_field(fld)) and np.array_equal(r["ni"], r1["ni"])
                         and r["diffmx"] == r1["diffmx"] and r["sumb"] == r1["sumb"] and r["sumdb"] == r1["sumdb"]
                         and r["SORFail"] == 0 and r["nisave"] == r1["nisave"])
    q.put((rank, ok))
    dist.destroy_process_group()


def test_scb_sub_problem_sharding_gloo_world2():
    """SURVEY 8(e), the zero-communication alternative: psi surfaces (alpha) / zeta planes (psi) split between two
    ranks, solved planes all-gathered over gloo -- the code path tests/multi_gpu_scb_check.py drives over NCCL."""
    import torch.multiprocessing as mp
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "emu"))
    import build_emu
    build_emu.build()
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_subproblem_worker, args=(r, 2, port, q)) for r in range(2)]
    for pr in procs:
        pr.start()
    res = [q.get(timeout=600) for _ in procs]
    for pr in procs:
        pr.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]


@pytest.mark.parametrize("world", [2, 3])
def test_scb_zeta_sharded_alpha_gloo(world):
    """iterateAlpha sharded along zeta -- halo planes per half-sweep + residual all-reduce over
    gloo, kernels run by the emulator: alfa, ni, diffmx, sumb, sumdb bit-identical to the one-rank 4-colour solve."""
    import torch.multiprocessing as mp
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "emu"))
    import build_emu
    build_emu.build()                                    # once, before the ranks start
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_zeta_worker, args=(r, world, port, q)) for r in range(world)]
    for pr in procs:
        pr.start()
    res = [q.get(timeout=600) for _ in procs]
    for pr in procs:
        pr.join(timeout=60)
    assert sorted(res) == [(r, True, True, True, True) for r in range(world)]


# ---- the library's own multi-GPU step (ramscb_b200/csrc/ram_shard.inl): host-side logic ---------------------------
def test_shard_plan_covers_every_cell_once():
    """rsg_shard_plan (no device needed): for every world size and policy each (species, pitch angle) and each
    (species, column block) has exactly one owner, groups are contiguous rank ranges, slabs hold >= 2 pitch angles."""
    from ramscb_b200 import host
    nS, NPA, P = 4, 72, 20 * 25
    nblocks = (P + 3) // 4
    for world in range(1, 9):
        for policy in (host.SHARD_SPECIES, host.SHARD_SLABS):
            if policy == host.SHARD_SPECIES and not (world % nS == 0 or nS % world == 0):
                with pytest.raises(host.RsgError):
                    host.shard_plan(world, 0, policy, nS, NPA, P)
                continue
            own_l = np.zeros((nS, NPA), dtype=int)
            own_b = np.zeros((nS, nblocks), dtype=int)
            for r in range(world):
                p = host.shard_plan(world, r, policy, nS, NPA, P)
                assert p.world == world and p.rank == r and p.per == 4
                assert p.g0 <= r < p.g0 + p.G and p.gidx == r - p.g0
                assert p.nl >= 2 or p.G == 1
                own_l[p.s0:p.s0 + p.ns, p.l0:p.l0 + p.nl] += 1
                own_b[p.s0:p.s0 + p.ns, p.b0:p.b0 + p.nb] += 1
                q = host.shard_plan(world, p.g0, policy, nS, NPA, P)          # the group shares one species range
                assert (q.s0, q.ns, q.G) == (p.s0, p.ns, p.G)
            assert (own_l == 1).all() and (own_b == 1).all()
    with pytest.raises(host.RsgError):
        host.shard_plan(9, 0, host.SHARD_SLABS, nS, NPA, P)
    with pytest.raises(host.RsgError):
        host.shard_plan(8, 0, host.SHARD_SLABS, nS, 8, P)                     # slabs of one pitch angle


def _peer_blob_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from ramscb_b200 import host, parallel
    blob = np.full(host.PEER_BLOB_BYTES, rank + 1, dtype=np.uint8)            # stands in for rsg_ram_peer_export
    allb = parallel.gather_blobs(dist, blob, world)
    ok = allb.shape == (world, host.PEER_BLOB_BYTES) and all((allb[r] == r + 1).all() for r in range(world))
    # without a CUDA device the library refuses to build a handle: the product path fails loudly, no fallback
    try:
        from ramscb_b200 import grids
        host.RamGpu(grids.build_grids(NR=8, NT=9, NE=8))
        ok = False
    except host.RsgError:
        pass
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


def test_peer_blob_all_gather_gloo():
    """world_size 2 over gloo: the one host-side message of rsg_ram_run_sharded's set-up (the ranks' opaque blobs,
    all-gathered in rank order) -- what a Fortran host does with one MPI_Allgather."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    world, port = 2, 29541
    ps = [ctx.Process(target=_peer_blob_worker, args=(r, world, port, q)) for r in range(world)]
    for p in ps:
        p.start()
    res = sorted(q.get(timeout=180) for _ in range(world))
    for p in ps:
        p.join(timeout=60)
    assert res == [(0, True), (1, True)]
