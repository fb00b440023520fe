"""Multi-GPU execution of the RAM step: one process per GPU, torch.distributed for the plumbing.

The reference has no domain decomposition (MPI is init/finalize only,
src/Main.f90:54-57); this sharding is a new design:

* species are independent inside ``ram_run`` (src/ModRamRun.f90:64-185), so up to
  nS ranks need **no data-path communication at all**;
* beyond nS ranks a species is shared by a group of G ranks.  DRIFTR/P/E and the
  pointwise losses are independent across pitch angle L, DRIFTMU / WPADIF couple all
  L but are independent across energy K: a rank owns an L-slab for the R,P,E sweeps
  and a K-slab for the pitch-angle block.  The step is palindromic
  (R,P,E | MU, losses, MU | E,P,R), so there are exactly **two** re-shardings per
  step, each an all-to-all *inside the group* done with NCCL send/recv of the
  complementary (L,K) blocks straight out of / into the resident F2 buffer (the
  device layout [L][K][plane] makes every block a run of contiguous planes).

Control-plane reductions (CFL minima, SUMRC partial sums, partial pressures: a few
KB) are all-reduces after the step.
"""
from __future__ import annotations

import os
from dataclasses import dataclass

import numpy as np


def _split(n, parts, idx):
    """[start, count) of slab ``idx`` when n items are split into ``parts`` near-equal slabs."""
    base, rem = divmod(n, parts)
    start = idx * base + min(idx, rem)
    return start, base + (1 if idx < rem else 0)


@dataclass
class ShardPlan:
    world: int
    rank: int
    nS: int
    NPA: int
    NE: int
    s0: int = 0          # first owned species (0-based)
    ns: int = 0          # number of owned species
    group: tuple = ()    # ranks sharing this rank's species (sorted); len 1 => no exchange
    gidx: int = 0        # index of this rank in its group
    l0: int = 0
    nl: int = 0
    k0: int = 0
    nk: int = 0
    active: tuple = ()   # ranks that hold species (all of them unless small grids leave ranks idle)

    @property
    def G(self):
        return len(self.group)


# A species is split among several ranks only when it has at least this many cells: the two
# re-shardings per step move half of a rank's data each and cost ~0.1 ms of launch/NCCL latency,
# which only pays off for grids well beyond the 4x one (19.8 M cells per species); measured in
# profiles/r1/scaling_r1.txt.  Below it, ranks beyond nS stay idle.
SPLIT_MIN_CELLS = int(float(os.environ.get("RSG_SPLIT_MIN_CELLS", "5e7")))


def make_plan(world: int, rank: int, nS: int, NPA: int, NE: int, cells_per_species=None) -> ShardPlan:
    """cells_per_species=None: always split a species when world > nS (tests, large grids)."""
    p = ShardPlan(world=world, rank=rank, nS=nS, NPA=NPA, NE=NE)
    p.active = tuple(range(world))
    if world > nS and cells_per_species is not None and cells_per_species < SPLIT_MIN_CELLS:
        # one species per rank on the first nS ranks, the others idle
        p.active = tuple(range(nS))
        p.group, p.gidx = (rank,), 0
        p.s0, p.ns = (rank, 1) if rank < nS else (0, 0)
        p.l0, p.nl, p.k0, p.nk = 0, NPA, 0, NE
        return p
    if world <= nS:
        if nS % world != 0:
            raise ValueError(f"{nS} species cannot be split evenly over {world} ranks")
        per = nS // world
        p.s0, p.ns = rank * per, per
        p.group, p.gidx = (rank,), 0
    else:
        if world % nS != 0:
            raise ValueError(f"{world} ranks must be a multiple of the {nS} species")
        G = world // nS
        if G > min(NPA // 2, NE):
            raise ValueError("too many ranks per species for the (L,K) slabs")
        p.s0, p.ns = rank // G, 1
        p.group = tuple(range((rank // G) * G, (rank // G + 1) * G))
        p.gidx = rank % G
    p.l0, p.nl = _split(NPA, p.G, p.gidx)
    p.k0, p.nk = _split(NE, p.G, p.gidx)
    return p


def exchange_blocks(p: ShardPlan, to_kslab: bool):
    """Blocks to swap with every other rank of the group.

    Returns a list of (peer_rank, send (l0,nl,k0,nk), recv (l0,nl,k0,nk)).
    to_kslab=True : L-slab -> K-slab layout (before DRIFTMU): send my pitch angles of the
                    peer's energies, receive the peer's pitch angles of my energies.
    to_kslab=False: the way back (after the second DRIFTMU).
    """
    out = []
    for gi, peer in enumerate(p.group):
        if peer == p.rank:
            continue
        pl0, pnl = _split(p.NPA, p.G, gi)
        pk0, pnk = _split(p.NE, p.G, gi)
        mine_L_peer_K = (p.l0, p.nl, pk0, pnk)
        peer_L_my_K = (pl0, pnl, p.k0, p.nk)
        if to_kslab:
            out.append((peer, mine_L_peer_K, peer_L_my_K))
        else:
            out.append((peer, peer_L_my_K, mine_L_peer_K))
    return out


def block_chunks(block, NE, Pp):
    """Contiguous runs (offset, length in doubles) of an (l0,nl,k0,nk) block of a species
    buffer laid out [L][K][Pp]."""
    l0, nl, k0, nk = block
    return [((l * NE + k0) * Pp, nk * Pp) for l in range(l0, l0 + nl)]


def exchange(p: ShardPlan, bufs, Pp: int, to_kslab: bool, dist):
    """Re-shard the species buffers of this rank inside its group.

    ``bufs``: list of 1-D torch tensors (one per owned species) aliasing the resident F2
    buffers, on whatever device the process group's backend moves (CUDA for nccl, CPU for
    gloo).  An (l0,nl,k0,nk) block is nl runs of nk*Pp contiguous doubles: it is packed into one
    message per peer (a strided 2-D copy), all messages are posted as one batch
    (ncclGroupStart/End underneath) and the received blocks are scattered back.
    """
    if p.G == 1:
        return
    import torch
    import torch.distributed as td
    ops, pending = [], []
    for peer, sblk, rblk in exchange_blocks(p, to_kslab):
        for buf in bufs:
            v = buf[:p.NPA * p.NE * Pp].view(p.NPA, p.NE * Pp)
            sl0, snl, sk0, snk = sblk
            rl0, rnl, rk0, rnk = rblk
            send = v[sl0:sl0 + snl, sk0 * Pp:(sk0 + snk) * Pp].contiguous()
            recv = torch.empty((rnl, rnk * Pp), dtype=buf.dtype, device=buf.device)
            ops.append(td.P2POp(td.isend, send, peer))
            ops.append(td.P2POp(td.irecv, recv, peer))
            pending.append((v[rl0:rl0 + rnl, rk0 * Pp:(rk0 + rnk) * Pp], recv))
    for w in td.batch_isend_irecv(ops):
        w.wait()
    for dst, recv in pending:
        dst.copy_(recv)


def exchange_lp(p: ShardPlan, bufs, Pp: int, nblocks: int, per: int, to_columns: bool, dist):
    """Re-sharding of the FUSED step inside a group: pitch-angle slabs (plane kernels) <-> ranges of
    plane positions (column kernel).  The species buffer is [NPA*NE][Pp]; rank g owns rows
    l in slab_g (all positions) in the plane layout and columns p in range_g (all rows) in the
    column layout.  to_columns: send my rows of the peer's columns, receive the peer's rows of my
    columns; the way back swaps the roles.  One packed message per peer (strided 2-D copies)."""
    if p.G == 1:
        return
    import torch
    import torch.distributed as td

    def cols(gi):
        b0, nb = _split(nblocks, p.G, gi)
        return b0 * per, min((b0 + nb) * per, Pp)

    def rows(gi):
        l0, nl = _split(p.NPA, p.G, gi)
        return l0 * p.NE, (l0 + nl) * p.NE

    ops, pending = [], []
    my_r, my_c = rows(p.gidx), cols(p.gidx)
    for gi, peer in enumerate(p.group):
        if peer == p.rank:
            continue
        pr, pc = rows(gi), cols(gi)
        (sr, sc), (rr, rc) = ((my_r, pc), (pr, my_c)) if to_columns else ((pr, my_c), (my_r, pc))
        for buf in bufs:
            v = buf[:p.NPA * p.NE * Pp].view(p.NPA * p.NE, Pp)
            send = v[sr[0]:sr[1], sc[0]:sc[1]].contiguous()
            recv = torch.empty((rr[1] - rr[0], rc[1] - rc[0]), dtype=buf.dtype, device=buf.device)
            ops.append(td.P2POp(td.isend, send, peer))
            ops.append(td.P2POp(td.irecv, recv, peer))
            pending.append((v[rr[0]:rr[1], rc[0]:rc[1]], recv))
    for w in td.batch_isend_irecv(ops):
        w.wait()
    for dst, recv in pending:
        dst.copy_(recv)


def _dev_tensor(ptr, n, on_cuda=True):
    """torch view of library-owned memory: device memory on the GPU; with the host-CPU emulator of the
    test-suite (tests/emu) the "device" pointer is host memory and the view is a CPU tensor."""
    import torch
    if on_cuda:
        return torch.as_tensor(_DevBuf(ptr, n), device="cuda")
    import ctypes
    return torch.from_numpy(np.ctypeslib.as_array(ctypes.cast(ptr, ctypes.POINTER(ctypes.c_double)), shape=(n,)))


class _DevBuf:
    """Zero-copy view of library-owned device memory for torch (CUDA array interface)."""

    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f8", "data": (ptr, False), "version": 2}


def _bind_stream(gpu, on_cuda):
    """The library's kernels and torch.distributed's NCCL traffic must share ONE stream: the exchanges read / write the
    library's buffers through zero-copy views, and only stream order keeps a send behind the kernels that produce its data
    and the next kernels behind the scatter of what was received.  The classes below bind the handle to torch's current
    stream when they are built and check that it still is the current one whenever they run (the host-CPU emulator of the
    test-suite is synchronous: nothing to bind)."""
    if not on_cuda:
        return None
    import torch
    ptr = torch.cuda.current_stream().cuda_stream
    gpu.set_stream(ptr)
    return ptr


def _check_stream(ptr):
    if ptr is None:
        return
    import torch
    if torch.cuda.current_stream().cuda_stream != ptr:
        raise RuntimeError("the handle was bound to another CUDA stream: build and run the sharded step under the same "
                           "torch.cuda.stream(...) (see ramscb_b200.parallel._bind_stream)")


class RamSharded:
    """One rank's share of the RAM step (drop-in for ``RamGpu.ram_run`` at N > 1).  Python-driven exchange over
    torch.distributed (round 1); the library's own multi-GPU step is ``RamPeerSharded`` below."""

    def __init__(self, gpu, plan: ShardPlan, dist=None, on_cuda=True):
        self.gpu, self.p, self.dist, self.on_cuda = gpu, plan, dist, on_cuda
        self.setrc = np.zeros(gpu.g.nS)
        self._views = None
        self._agroup = None
        self._stream = _bind_stream(gpu, on_cuda and dist is not None and plan.world > 1)
        if dist is not None and plan.world > 1 and len(plan.active) < plan.world:
            self._agroup = dist.new_group(ranks=list(plan.active))   # collective: every rank calls it

    def _bufs(self):
        import torch
        out = []
        for s in range(self.p.s0, self.p.s0 + self.p.ns):
            ptr, n, pp = self.gpu.f2_device(s + 1)
            out.append(_dev_tensor(ptr, n, self.on_cuda))
        return out, pp

    def _result_views(self):
        if self._views is None:
            import torch
            r, rn, q, qn = self.gpu.results_device()
            nS = self.gpu.g.nS
            self._views = (torch.as_tensor(_DevBuf(r, nS * rn), device="cuda"), torch.as_tensor(_DevBuf(q, nS * qn), device="cuda"), rn, qn)
        return self._views

    def ram_run(self, DTs, DtsMin=1.0, flags=0):
        import torch
        g, p, gpu = self.gpu.g, self.p, self.gpu
        if p.ns == 0:
            return None                      # idle rank (small grid, more ranks than species): takes part in NO collective
        _check_stream(self._stream)
        if p.G == 1:
            # species-sharded: the whole step is local (fused kernels, graph replay)
            gpu.part_all(DTs, flags, p.s0, p.ns)
        elif gpu.fused_available(flags):
            # ranks sharing a species, fused kernels: pitch-angle slabs for the plane kernels,
            # ranges of plane positions for the column kernel, two re-shardings per step
            nblocks, per = gpu.col_blocks()
            b0, nb = _split(nblocks, p.G, p.gidx)
            gpu.fpart_planes_fwd(DTs, flags, p.s0, p.ns, p.l0, p.nl)
            bufs, pp = self._bufs()
            exchange_lp(p, bufs, pp, nblocks, per, True, self.dist)
            gpu.fpart_columns(DTs, flags, p.s0, p.ns, b0, nb)
            exchange_lp(p, bufs, pp, nblocks, per, False, self.dist)
            gpu.fpart_planes_rev(p.s0, p.ns, p.l0, p.nl)
        else:
            gpu.part_fwd(DTs, flags, p.s0, p.ns, p.l0, p.nl)
            bufs, pp = self._bufs()
            exchange(p, bufs, pp, True, self.dist)
            gpu.part_mid(DTs, flags, p.s0, p.ns, p.k0, p.nk)
            bufs, pp = self._bufs()
            exchange(p, bufs, pp, False, self.dist)
            gpu.part_rev(p.s0, p.ns, p.l0, p.nl)
        if p.G == 1 and len(p.active) > 1 and self.dist is not None and self.dist.get_backend() == "nccl":
            # every rank owns whole species: the result blocks of all ranks are concatenated by two
            # in-place all-gathers on the device (same stream as the step), then decoded once
            res, pp_t, rn, qn = self._result_views()
            self.dist.all_gather_into_tensor(res, res[p.s0 * rn:(p.s0 + p.ns) * rn], group=self._agroup)
            self.dist.all_gather_into_tensor(pp_t, pp_t[p.s0 * qn:(p.s0 + p.ns) * qn], group=self._agroup)
            DT, MOM, PE, PA = gpu.part_results(0, g.nS)
            return {"DtDrift": DT, "DtsNext": max(float(DT.min()), DtsMin), "moments": MOM,
                    "PPERT": np.asfortranarray(np.moveaxis(PE, 2, 0)), "PPART": np.asfortranarray(np.moveaxis(PA, 2, 0))}
        dt, mom, pper, ppar = gpu.part_results(p.s0, p.ns)
        # control-plane reductions over all ranks (a few KB)
        DT = np.full((4, g.nS), np.inf)
        MOM = np.zeros((14, g.nS))
        PE = np.zeros((g.NR, g.NT, g.nS))
        PA = np.zeros((g.NR, g.NT, g.nS))
        sl = slice(p.s0, p.s0 + p.ns)
        DT[:, sl], MOM[:, sl], PE[:, :, sl], PA[:, :, sl] = dt, mom, pper, ppar
        if self.dist is not None and p.world > 1 and len(p.active) > 1:
            # only the ranks that hold species take part (idle ranks returned above): the group of the active ranks
            dev = "cuda" if self.dist.get_backend() == "nccl" else "cpu"
            t_min = torch.as_tensor(DT, device=dev)
            t_sum = torch.as_tensor(np.concatenate([MOM.ravel(), PE.ravel(), PA.ravel()]), device=dev)
            self.dist.all_reduce(t_min, op=self.dist.ReduceOp.MIN, group=self._agroup)
            self.dist.all_reduce(t_sum, op=self.dist.ReduceOp.SUM, group=self._agroup)
            DT = t_min.cpu().numpy()
            v = t_sum.cpu().numpy()
            MOM = v[:MOM.size].reshape(MOM.shape)
            PE = v[MOM.size:MOM.size + PE.size].reshape(PE.shape)
            PA = v[MOM.size + PE.size:].reshape(PA.shape)
        return {"DtDrift": DT, "DtsNext": max(float(DT.min()), DtsMin), "moments": MOM,
                "PPERT": np.asfortranarray(np.moveaxis(PE, 2, 0)), "PPART": np.asfortranarray(np.moveaxis(PA, 2, 0))}


def bind_to_gpu_numa(device_index: int) -> dict:
    """Pin the calling process to the host cores next to its GPU (NVML's ideal CPU affinity for the device, found through
    the PCI bus id so that CUDA_VISIBLE_DEVICES remapping does not matter).  Call it BEFORE allocating the host arrays a
    rank hands to rsg_ram_f2_h2d(_shard): first touch then places them on the GPU's NUMA node, and the pinned copies do
    not cross the socket interconnect.  The reference's MPI ranks get the same from `mpirun --bind-to`; torchrun binds
    nothing.  Returns what was done (never raises: an unbound process is slower, not wrong)."""
    import os
    try:
        import pynvml
        import torch
        pr = torch.cuda.get_device_properties(device_index)
        bus = f"{pr.pci_domain_id:08X}:{pr.pci_bus_id:02X}:{pr.pci_device_id:02X}.0"
        pynvml.nvmlInit()
        hdl = pynvml.nvmlDeviceGetHandleByPciBusId(bus.encode())
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(hdl, (ncpu + 63) // 64)
        cpus = {64 * w + b for w, m in enumerate(words) for b in range(64) if (int(m) >> b) & 1}
        allowed = os.sched_getaffinity(0)
        cpus = sorted(cpus & allowed)
        if not cpus:
            return {"bound": False, "why": "NVML affinity mask empty within the allowed cpus"}
        if len(cpus) < len(allowed):
            os.sched_setaffinity(0, cpus)
        return {"bound": len(cpus) < len(allowed), "cpus": len(cpus), "first_cpu": cpus[0], "last_cpu": cpus[-1], "pci": bus}
    except Exception as e:                                  # no NVML, container without the right, ...
        return {"bound": False, "why": f"{type(e).__name__}: {e}"[:160]}


def gather_blobs(dist, blob, world, device=None):
    """All-gather of the ranks' peer blobs (rsg_ram_peer_export) in rank order: the ONLY host-side communication of the
    library's own multi-GPU step (a Fortran host does the same with one MPI_Allgather).  Works on any backend: the
    bytes travel as a uint8 tensor on `device` ("cuda" for nccl, "cpu" for gloo)."""
    import torch
    if device is None:
        device = "cuda" if dist.get_backend() == "nccl" else "cpu"
    mine = torch.from_numpy(np.ascontiguousarray(blob, dtype=np.uint8)).to(device)
    out = torch.empty(world * mine.numel(), dtype=torch.uint8, device=device)
    dist.all_gather_into_tensor(out, mine)
    return out.cpu().numpy().reshape(world, -1)


class RamPeerSharded:
    """One rank of the sharded RAM step that lives INSIDE the library (include/ramscb_gpu.h: rsg_ram_run_sharded).

    The ranks map each other's F2 buffer over NVLink (CUDA IPC); the kernels store their results straight into the
    buffer of the rank that reads them next, barriers and the result reduction run on the device, one CUDA graph per
    step.  Python's part is the one-off exchange of the 192-byte blobs.  ``ram_run`` returns what ``RamGpu.ram_run``
    returns (results of ALL species, identical on every rank)."""

    def __init__(self, gpu, dist, rank, world, policy=None):
        from . import host
        self.gpu, self.rank, self.world = gpu, rank, world
        self.policy = host.SHARD_SPECIES if policy is None else policy
        blobs = gather_blobs(dist, gpu.peer_export(), world)
        gpu.peer_attach(rank, world, self.policy, blobs)
        self.plan = gpu.shard_info()

    def load(self, F2):
        """this rank's share of the host array F2(nS,NR,NT,NE,NPA)"""
        self.gpu.f2_h2d_shard(F2)

    def store(self, F2):
        return self.gpu.f2_d2h_shard(F2)

    def ram_run(self, DTs, DtsMin=1.0, T=0.0, flags=0):
        return self.gpu.run_sharded(DTs, DtsMin=DtsMin, T=T, flags=flags)


class ScbSharded:
    """iterateAlpha / iteratePsi with the independent sub-problems split among ranks.

    The sub-problems of one solve (psi surfaces jz for alpha, zeta planes k for psi,
    src/ModScbEuler.f90:204, :514) do not talk to each other, so each rank solves a contiguous
    range with no in-solve communication; afterwards the solved planes are all-gathered
    (one packed message per rank, NCCL) and every rank runs the cheap post-processing on the
    complete field.  At the default grid the 43 x 4 / 96 x 2 cluster CTAs of one solve need two
    waves on one GPU and one wave on two: that is where the speed-up comes from."""

    def __init__(self, gpu, dist, rank, world, on_cuda=True):
        self.gpu, self.dist, self.rank, self.world, self.on_cuda = gpu, dist, rank, world, on_cuda
        self._fields = {}
        self._stream = _bind_stream(gpu, on_cuda and dist is not None and world > 1)

    def _field(self, name):
        if name not in self._fields:
            ptr, n = self.gpu.field_device(name)
            g = self.gpu
            self._fields[name] = _dev_tensor(ptr, n, self.on_cuda).view(g.nzeta + 1, g.npsi, g.nthe)
        return self._fields[name]

    def iterate(self, alpha, tol, nimax=5001, theChange=4, psiChange=0, ordering=1):
        import torch
        _check_stream(self._stream)
        g = self.gpu
        nP = max(psiChange, 1)
        nsub = (g.npsi - nP - 1) if alpha else (g.nzeta - 1)
        s0, ns = _split(nsub, self.world, self.rank)
        g.iterate_part(alpha, tol, s0, ns, nimax=nimax, theChange=theChange, psiChange=psiChange, ordering=ordering)
        f = self._field("alfa" if alpha else "psi")

        def planes(a0, n):                      # sub-problem q lives at index q+1 of its axis
            return f[:, 1 + a0:1 + a0 + n, :] if alpha else f[1 + a0:1 + a0 + n]

        if self.world > 1:
            nmax = (nsub + self.world - 1) // self.world
            per = planes(0, 1).numel()
            send = torch.zeros(nmax * per, dtype=f.dtype, device=f.device)
            if alpha:
                send[:ns * per].view(g.nzeta + 1, ns, g.nthe).copy_(planes(s0, ns))
            else:
                send[:ns * per].copy_(planes(s0, ns).reshape(-1))
            out = torch.empty(self.world * nmax * per, dtype=f.dtype, device=f.device)
            self.dist.all_gather_into_tensor(out, send)
            for r in range(self.world):
                if r == self.rank:
                    continue
                r0, rn = _split(nsub, self.world, r)
                blk = out[r * nmax * per:r * nmax * per + rn * per]
                if alpha:
                    planes(r0, rn).copy_(blk.view(g.nzeta + 1, rn, g.nthe))
                else:
                    planes(r0, rn).copy_(blk.view(rn, g.npsi, g.nthe))
        res = g.iterate_finish(alpha, theChange=theChange, psiChange=psiChange)
        if self.world > 1:
            dev = "cuda" if self.on_cuda else "cpu"
            ni = torch.as_tensor(res["ni"].astype(np.int64), device=dev)
            sc = torch.tensor([res["diffmx"], float(res["SORFail"])], dtype=torch.float64, device=dev)
            self.dist.all_reduce(ni, op=self.dist.ReduceOp.MAX)
            self.dist.all_reduce(sc, op=self.dist.ReduceOp.MAX)
            res["ni"] = ni.cpu().numpy().astype(np.int32)
            res["nisave"] = int(res["ni"].max())
            res["diffmx"], res["SORFail"] = float(sc[0].item()), int(sc[1].item())
        return res


class ScbZetaSharded:
    """iterateAlpha sharded along the periodic azimuthal axis zeta (north-star scheme).

    Rank p relaxes a contiguous range of the zeta planes 1..nzeta-1 (0-based) of EVERY psi surface
    (src/ModScbEuler.f90:204-262 couples a surface's planes k-1, k, k+1).  Per sweep: half-sweep of the
    even planes, the even edge planes go to the zeta neighbours (one plane = nthe*npsi contiguous
    doubles), half-sweep of the odd planes, the odd edge planes go over, then ONE all-reduce(MAX) of
    the per-surface residual maxima + failure flags, and the device applies the loop control.  The host
    only learns every `poll` sweeps whether anything still iterates; surfaces that have finished are
    skipped on the device, so ni / alfa are bit-identical to the one-GPU RSG_SOR_COLOR4 solve whatever
    `poll` is.  The periodic images (k = 0, nzeta) are fixed during a solve, as in the reference (they
    are refreshed by the wrap in rsg_scb_iterate_finish), so there is no wrap-around message.
    iteratePsi's sub-problems are the zeta planes themselves: ScbSharded covers it with the same
    partition and no in-solve communication."""

    def __init__(self, gpu, dist, rank, world, on_cuda=True, poll=8):
        self.gpu, self.dist, self.rank, self.world, self.on_cuda, self.poll = gpu, dist, rank, world, on_cuda, poll
        self._stream = _bind_stream(gpu, on_cuda and dist is not None and world > 1)
        self.k0, self.nk = _split(gpu.nzeta - 1, world, rank)
        self.k0 += 1                                        # planes 1..nzeta-1 are relaxed
        ptr, n = gpu.field_device("alfa")
        self.alfa = _dev_tensor(ptr, n, on_cuda).view(gpu.nzeta + 1, gpu.npsi * gpu.nthe)
        self.messages = 0

    def _edges(self, parity):
        """P2P ops for the edge planes of this parity: send mine, receive the neighbours' into my halo"""
        import torch.distributed as td
        ops = []
        lo, hi = self.k0, self.k0 + self.nk - 1
        if self.nk == 0:
            return ops
        dn = self._neighbour(-1)
        up = self._neighbour(+1)
        if dn is not None:
            if lo % 2 == parity:
                ops.append(td.P2POp(td.isend, self.alfa[lo], dn))
            if (lo - 1) % 2 == parity:
                ops.append(td.P2POp(td.irecv, self.alfa[lo - 1], dn))
        if up is not None:
            if hi % 2 == parity:
                ops.append(td.P2POp(td.isend, self.alfa[hi], up))
            if (hi + 1) % 2 == parity:
                ops.append(td.P2POp(td.irecv, self.alfa[hi + 1], up))
        return ops

    def _neighbour(self, step):
        """next rank in that direction that owns planes (ranks beyond nzeta-1 planes own none)"""
        r = self.rank + step
        while 0 <= r < self.world:
            if _split(self.gpu.nzeta - 1, self.world, r)[1] > 0:
                return r
            r += step
        return None

    def iterate(self, tol, nimax=5001, theChange=4, psiChange=0):
        import torch.distributed as td
        _check_stream(self._stream)
        g = self.gpu
        g.zsolve_begin(tol, self.k0, self.nk, nimax=nimax, theChange=theChange, psiChange=psiChange)
        ptr, n = g.zsolve_state_device()
        state = _dev_tensor(ptr, n, self.on_cuda)
        sweeps = 0
        while sweeps < nimax:
            for _ in range(min(self.poll, nimax - sweeps)):
                for parity in (0, 1):
                    g.zsolve_half(parity)
                    if self.world > 1:
                        ops = self._edges(parity)
                        if ops:
                            self.messages += len(ops)
                            for w in td.batch_isend_irecv(ops):
                                w.wait()
                if self.world > 1:
                    self.dist.all_reduce(state, op=self.dist.ReduceOp.MAX)
                g.zsolve_commit()
                sweeps += 1
            if g.zsolve_pending() == 0:
                break
        # every rank gets the relaxed planes of the others, then the shared post-processing
        if self.world > 1:
            for r in range(self.world):
                r0, rn = _split(g.nzeta - 1, self.world, r)
                if rn:
                    self.dist.broadcast(self.alfa[1 + r0:1 + r0 + rn], src=r)
        res = g.iterate_finish(True, theChange=theChange, psiChange=psiChange)
        res["sweeps_launched"] = sweeps
        return res
