"""torchrun entry: SCB solves with the sub-problems sharded over N GPUs, and iterateAlpha sharded along
zeta (halo exchange per half-sweep), vs the single-GPU solve (bit-identical potentials, sweep counts,
residual maxima, sums).
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tests/multi_gpu_scb_check.py
"""
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ramscb_b200 import host, parallel, scb_synthetic  # noqa: E402


def main():
    rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    st = torch.cuda.Stream()
    torch.cuda.set_stream(st)
    inp = scb_synthetic.build_scb(nthe=101, npsi=45, nzeta=97, warp=0.2)
    ok = True
    ref = host.ScbGpu(inp, device=lr)
    ref.computeBandJacob(); ref.metrica(); ref.newk()
    ra = ref.iterateAlpha(1e-6, ordering=host.SOR_COLOR4)
    ref.metric(); ref.newj()
    rp = ref.iteratePsi(1e-6, ordering=host.SOR_COLOR4)
    gpu = host.ScbGpu(inp, device=lr)
    gpu.set_stream(st.cuda_stream)
    sh = parallel.ScbSharded(gpu, dist, rank, world)
    gpu.computeBandJacob(); gpu.metrica(); gpu.newk()
    times = {}
    for name, alpha, r1 in (("alpha", True, ra), ("psi", False, rp)):
        if not alpha:
            gpu.metric(); gpu.newj()
        torch.cuda.synchronize(); dist.barrier()
        t0 = time.perf_counter()
        r = sh.iterate(alpha, 1e-6)
        torch.cuda.synchronize()
        times[name] = (time.perf_counter() - t0) * 1e3
        fld = "alfa" if alpha else "psi"
        same = np.array_equal(gpu.get_field(fld), ref.get_field(fld))
        meta = (np.array_equal(r["ni"], r1["ni"]) and r["diffmx"] == r1["diffmx"] and r["sumb"] == r1["sumb"]
                and r["sumdb"] == r1["sumdb"] and r["SORFail"] == r1["SORFail"] == 0)
        print(f"rank {rank} {name}: field identical={same} counts/sums identical={meta} "
              f"wall {times[name]:.2f} ms (1 GPU kernel {r1['ms']:.2f} ms)", flush=True)
        ok = ok and same and meta
    # iterateAlpha sharded along zeta (halo planes per half-sweep + MAX all-reduce per sweep), same bar
    gpu.set_field("alfa", inp.alfa)
    gpu.metrica(); gpu.newk()
    zs = parallel.ScbZetaSharded(gpu, dist, rank, world, on_cuda=True, poll=16)
    torch.cuda.synchronize(); dist.barrier()
    t0 = time.perf_counter()
    r = zs.iterate(1e-6)
    torch.cuda.synchronize()
    tz = (time.perf_counter() - t0) * 1e3
    same = np.array_equal(gpu.get_field("alfa"), ref.get_field("alfa"))
    meta = (np.array_equal(r["ni"], ra["ni"]) and r["diffmx"] == ra["diffmx"] and r["sumb"] == ra["sumb"]
            and r["sumdb"] == ra["sumdb"] and r["SORFail"] == 0)
    print(f"rank {rank} alpha, zeta-sharded: field identical={same} counts/sums identical={meta} wall {tz:.2f} ms "
          f"({r['sweeps_launched']} sweeps launched, {zs.messages} halo messages)", flush=True)
    ok = ok and same and meta
    t = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    if rank == 0 and int(t.item()) == 1:
        print("MULTI_GPU_SCB_CHECK_OK", flush=True)
    dist.destroy_process_group()
    sys.exit(0 if int(t.item()) == 1 else 1)


if __name__ == "__main__":
    main()
