"""torchrun entry: the library's own multi-GPU RAM step (rsg_ram_run_sharded: CUDA IPC peer memory, device-side
barriers, one CUDA graph per rank) on N GPUs vs rsg_ram_run on one GPU.  Run by bench.py's N > 1 arm before timing
(`--check`), by tests/test_ram_parity_gpu.py::test_multi_gpu_peer_memory when the box has >= 2 GPUs, and by hand:
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tests/multi_gpu_peer_check.py [x4]
Bars: every rank's share of F2 bit-identical to the one-GPU step, CFL limits identical, pressures / SUMRC <= 1e-12."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ramscb_b200 import grids, host, parallel, synthetic  # noqa: E402

STEPS = (5.0, 7.5, 7.5)


def check(rank, world, lr, g, inp, policy, flags, mode=host.MODE_FAST):
    D = synthetic.synthetic_daa(g, inp) if flags & 5 else None
    gpu = host.RamGpu(g, device=lr, mode=mode)
    gpu.set_fields(inp)
    gpu.set_efield(inp.VT, inp.EIR, inp.EIP)
    gpu.set_boundary(inp.FGEOS)
    gpu.set_wavelo(inp.WALOS1, inp.WALOS2, inp.WALOS3, inp.Kp, inp.Kpmax12)
    gpu.set_plasmasphere(inp.NECR)
    if D is not None:
        gpu.set_diffcoef(1, D)
        gpu.set_diffcoef(2, D)
    sh = parallel.RamPeerSharded(gpu, dist, rank, world, policy)
    sh.load(inp.F2)
    outs = [sh.ram_run(dts, flags=flags) for dts in STEPS]
    mine = sh.store(np.full(inp.F2.shape, np.nan, order="F"))
    p = sh.plan
    dist.barrier()
    gpu.close()
    ref = host.RamGpu(g, device=lr, mode=mode)
    ref.set_inputs(inp)
    if D is not None:
        ref.set_diffcoef(1, D)
        ref.set_diffcoef(2, D)
    refs = [ref.ram_run(dts, flags=flags) for dts in STEPS]
    full = ref.f2_d2h()
    ref.close()
    sl, lsl = slice(p.s0, p.s0 + p.ns), slice(p.l0, p.l0 + p.nl)
    same = bool(np.array_equal(mine[sl][..., lsl], full[sl][..., lsl]))
    ok = same
    for a, r in zip(outs, refs):
        ok = ok and np.array_equal(a["DtDrift"], r["DtDrift"]) and a["DtsNext"] == r["DtsNext"]
        ok = ok and np.allclose(a["PPERT"], r["PPERT"], rtol=1e-12, atol=0) and np.allclose(a["PPART"], r["PPART"], rtol=1e-12, atol=0)
        ok = ok and np.allclose(a["SETRC"], r["SETRC"], rtol=1e-12, atol=0)
    print(f"rank {rank} policy={policy} flags={flags} mode={mode} species [{p.s0},{p.s0 + p.ns}) G={p.G} l=[{p.l0},{p.l0 + p.nl}) "
          f"blocks=[{p.b0},{p.b0 + p.nb}): F2 share identical={same} results ok={bool(ok)}", flush=True)
    return bool(ok)


def main():
    rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    big = len(sys.argv) > 1 and sys.argv[1] == "x4"
    g = grids.build_grids(NR=80, NT=49, NE=70, energy_refine=2) if big else grids.build_grids()
    inp = synthetic.make_inputs(g, f2_kind="noisy", inductive=True)
    ok = True
    for policy in (host.SHARD_SPECIES, host.SHARD_SLABS):
        if policy == host.SHARD_SPECIES and not (world % g.nS == 0 or g.nS % world == 0):
            continue
        for flags in ((5,) if big else (0, 5)):
            ok = check(rank, world, lr, g, inp, policy, flags) and ok
    if not big and g.nS % world == 0:       # whole species per rank: any mode / flag set
        ok = check(rank, world, lr, g, inp, host.SHARD_SPECIES, host.F_COULOMB, host.MODE_EXACT) and ok
    t = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    if rank == 0 and int(t.item()) == 1:
        print("MULTI_GPU_PEER_CHECK_OK", flush=True)
    dist.destroy_process_group()
    sys.exit(0 if int(t.item()) == 1 else 1)


if __name__ == "__main__":
    main()
