#!/usr/bin/env python
"""bench.py -- RAM phase-space cell-updates/s per full RAM step (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--workload x4|default] [--flags F] [--policy species|slabs]

One "step" = one pass of the RAM hot path (`ram_run`, src/ModRamRun.f90:64-222) over one synthetic input set: for each of
the 4 species CEPARA, DRIFTPARA, DRIFTR/P/E/MU, SUMRC, [WPADIF], [WAVELO | CHAREXCHANGE], ATMOL x2, the same in reverse
back to DRIFTR, then the epilogue, the ANISCH pressures and the CFL time step.  Work unit: one cell-update = one F2 cell
advanced by one operator call; a step applies 12 operators (8 drift sweeps + 2 ATMOL + 2 CHAREXCHANGE-or-WAVELO) plus 2
WPADIF for the species that diffuse (electrons: WPI, H+: EMIC) to nS*NR*NT*NE*NPA cells.

Workload (the same at every N, so the driver's scaling efficiency is a strong-scaling figure): BASELINE configs[2] -- the
full step with WPI / EMIC pitch-angle diffusion on the 4x grid (NR=80 NT=49 NE=70 NPA=72, 79 M cells, 632 MB), which fits
one GPU.  configs[1] (default grid, drift + loss step) is measured beside it at N = 1 (`configs1`, also copied into
`roofline.configs1_default_grid`).

Printed JSON keys follow the driver contract: `value` is measured with F2 resident in HBM (CUDA events over the library's
run stream, L2 flushed between steps, max over ranks); `e2e` is the same metric through the C ABI with HOST buffers
(pinned), the rank's share of F2 going host->device and device->host inside the timed region (wall clock).
`--impl reference` times the reference's CPU algorithm (the C++ oracle with the reference's own OpenMP-over-species
parallelism) on a bounded sample of the same workload.

N > 1 (one process per GPU under torchrun): the library's own sharded step, rsg_ram_run_sharded
(ramscb_b200/csrc/ram_shard.inl) -- species over ranks, pitch-angle slabs / plane-position blocks inside a species group,
the two re-shardings done by the kernels' write-backs into the peer's buffer over NVLink (CUDA IPC), device-side barriers
and result reduction, one CUDA graph per rank.  Before timing, every rank's share is compared bit for bit with the one-GPU
step (`config.sharded_check`).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

OPS_PER_STEP = 12  # operator applications per cell per ram_run (see module docstring)
DTS = 5.0


def workload(name):
    from ramscb_b200 import grids, synthetic
    if name == "default":
        g = grids.build_grids()                                     # BASELINE configs[1]
        desc = "configs[1]: RAM drift+loss step, reference default grid nS=4 NR=20 NT=25 NE=35 NPA=72 (5.04 M cells, 40 MB), synthetic F2"
    elif name == "x4":
        g = grids.build_grids(NR=80, NT=49, NE=70, energy_refine=2)  # BASELINE configs[2] grid
        desc = "configs[2] grid: nS=4 NR=80 NT=49 NE=70 NPA=72 (79.0 M cells, 632 MB), synthetic F2"
    else:
        raise SystemExit("unknown workload " + name)
    inp = synthetic.make_inputs(g, f2_kind="noisy", inductive=True)
    return g, inp, desc


class ClockSampler:
    """SM clock and throttle reasons during the timed region.  NVML is polled in-process once per
    timed step (right after the step, before the untimed L2 flush: tens of microseconds, outside the
    device-timed interval); nvidia-smi in a side process is the fallback."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.rows = []
        self.proc = None
        self.index = index
        self.nv = None
        self.sm, self.reasons, self.mx = [], set(), None
        try:
            import pynvml
            pynvml.nvmlInit()
            # torch's device index follows CUDA_VISIBLE_DEVICES, NVML's does not: match by PCI address
            self.h = None
            try:
                import torch
                pr = torch.cuda.get_device_properties(index)
                bus = f"{pr.pci_domain_id:08X}:{pr.pci_bus_id:02X}:{pr.pci_device_id:02X}.0"
                self.h = pynvml.nvmlDeviceGetHandleByPciBusId(bus.encode())
            except Exception:
                self.h = None
            if self.h is None:
                vis = os.environ.get("CUDA_VISIBLE_DEVICES")
                phys = index
                if vis:
                    ids = [v.strip() for v in vis.split(",") if v.strip()]
                    if index < len(ids) and ids[index].isdigit():
                        phys = int(ids[index])
                self.h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.mx = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.nv = pynvml
        except Exception:
            self.nv = None

    def sample(self):
        if self.nv is None:
            return
        nv = self.nv
        try:
            self.sm.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
            r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
            for name, bit in (("hw_slowdown", nv.nvmlClocksEventReasonHwSlowdown),
                              ("hw_thermal_slowdown", nv.nvmlClocksEventReasonHwThermalSlowdown),
                              ("sw_thermal_slowdown", nv.nvmlClocksEventReasonSwThermalSlowdown),
                              ("sw_power_cap", nv.nvmlClocksEventReasonSwPowerCap)):
                if r & bit:
                    self.reasons.add(name)
        except Exception:
            pass

    def start(self):
        if self.nv is not None:
            return
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.nv is not None:
            return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": self.mx,
                    "reasons": sorted(self.reasons), "samples": len(self.sm), "source": "nvml, one sample per timed step"}
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = float(r[1])
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm), "source": "nvidia-smi -lms 100"}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def cpu_reference(g, inp, steps, warmup, flags=0, nthreads=None):
    """The reference's CPU algorithm (C++ oracle, oracle/ram_oracle.cpp) on the same workload: one `ram_run` per step,
    OpenMP over species like src/ModRamRun.f90:64 (the reference's own decomposition: at most nS threads).
    Returns (cell-updates/s, threads, s/step)."""
    from oracle import oracle
    from ramscb_b200 import synthetic
    oracle.build()
    o = oracle.RamOracle(g, inp, DTs=DTS)
    if flags & 5:
        D = synthetic.synthetic_daa(g, inp)
        o.set_array("ATAC", D)
        o.set_array("ATAW_emic_h", D)
    ncpu = os.cpu_count() or 1
    if nthreads is None:
        nthreads = min(g.nS, ncpu)
    for _ in range(warmup):
        o.ram_run(flags=flags, nthreads=nthreads)
    t0 = time.perf_counter()
    for _ in range(steps):
        o.ram_run(flags=flags, nthreads=nthreads)
    dt = (time.perf_counter() - t0) / steps
    cells = g.nS * g.NR * g.NT * g.NE * g.NPA
    return ops_per_cell(g, flags) * cells / dt, nthreads, dt


def cpu_all_cores(steps=2):
    """BASELINE.md 3.2's second figure: every host core busy with the reference's algorithm.  The reference parallelises
    ram_run over species only (<= nS = 4 threads per run), so all cores = as many concurrent default-grid runs as fit
    (cores // 4 instances of the oracle, 4 threads each); value = their aggregate cell-updates/s."""
    from concurrent.futures import ThreadPoolExecutor
    from oracle import oracle
    g, inp, desc = workload("default")
    ncpu = os.cpu_count() or 1
    groups = max(1, ncpu // g.nS)
    os_ = [oracle.RamOracle(g, inp, DTs=DTS) for _ in range(groups)]
    nthreads = min(g.nS, ncpu)

    def one(o):
        o.ram_run(flags=0, nthreads=nthreads)          # ctypes releases the GIL inside the call

    def step():
        with ThreadPoolExecutor(groups) as ex:
            list(ex.map(one, os_))

    step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / steps
    cells = groups * g.nS * g.NR * g.NT * g.NE * g.NPA
    return {"value": OPS_PER_STEP * cells / dt, "unit": "cell-updates/s", "cores": nthreads * groups, "host_cores": ncpu,
            "concurrent_runs": groups, "s_per_step": dt, "workload": desc,
            "note": "the reference's own decomposition (OpenMP over the 4 species) replicated over all host cores"}


def ops_per_cell(g, flags):
    """operator applications per cell per ram_run, averaged over the species: 8 drift sweeps + 2 ATMOL + 2 CHAREXCHANGE /
    WAVELO, + 2 WPADIF for each species that diffuses (electrons with WPI, H+ with EMIC), + 4 with the Coulomb operators"""
    nw = sum(1 for sp in g.species if (flags & 1 and sp.WPI) or (flags & 4 and sp.EMIC))
    return OPS_PER_STEP + 2.0 * nw / g.nS + (4.0 if flags & 2 else 0.0)


def scb_metrics(device):
    """The second half of BASELINE.json's metric: SCB SOR sweeps/s.  One Euler-potential solve on
    the default SCB grid (configs[3]: 101 x 45 x 97, synthetic anisotropic pressure) through the
    C ABI: computeBandJacob, metrica+newk, iterateAlpha, metric+newj, iteratePsi to the reference's
    tolerance (InConAlpha = InConPsi = 1e-6), 4-colour ordering.  Device times of the calls
    (rsg_scb_last_ms, CUDA events), median of 3 solves from the same start."""
    from ramscb_b200 import host, scb_synthetic
    inp = scb_synthetic.build_scb(nthe=101, npsi=45, nzeta=97, warp=0.2)
    gpu = host.ScbGpu(inp, device=device)
    alfa0, psi0 = gpu.get_field("alfa").copy(), gpu.get_field("psi").copy()
    npts_a = (inp.npsi - 2) * (inp.nzeta - 1) * (inp.nthe - 8)   # jz=2..npsi-1, k=2..nzeta, iz=1+nT..nthe-nT (nT=4)
    npts_p = (inp.nzeta - 1) * (inp.npsi - 2) * (inp.nthe - 8)
    runs = []
    for _ in range(3):
        gpu.set_field("alfa", alfa0)
        gpu.set_field("psi", psi0)
        r = {}
        gpu.computeBandJacob(); r["bandjacob_ms"] = gpu.last_ms()
        gpu.metrica(); r["metrica_ms"] = gpu.last_ms()
        gpu.newk()
        ra = gpu.iterateAlpha(1e-6, ordering=host.SOR_COLOR4)
        gpu.metric(); r["metric_ms"] = gpu.last_ms()
        gpu.newj()
        rp = gpu.iteratePsi(1e-6, ordering=host.SOR_COLOR4)
        r["alpha"], r["psi"] = ra, rp
        runs.append(r)
    runs.sort(key=lambda r: r["alpha"]["ms"] + r["psi"]["ms"])
    r = runs[1]
    out = {"grid": "nthe=101 npsi=45 nzeta=97 (configs[3]), synthetic anisotropic pressure, tol 1e-6, 4-colour SOR",
           "bandjacob_ms": r["bandjacob_ms"], "metrica_ms": r["metrica_ms"], "metric_ms": r["metric_ms"]}
    for name, rr, npts in (("iterate_alpha", r["alpha"], npts_a), ("iterate_psi", r["psi"], npts_p)):
        sweeps = int(np.sum(rr["ni"]))          # sub-problem sweeps, summed over the independent sub-problems
        nsub = int(np.count_nonzero(rr["ni"]))
        full = sweeps / max(nsub, 1)             # equivalent sweeps over all sub-problems
        out[name] = {"ms": rr["ms"], "max_sweeps": int(rr["nisave"]), "mean_sweeps": full, "SORFail": int(rr["SORFail"]),
                     "sweeps_per_s": full / (rr["ms"] * 1e-3), "point_updates_per_s": full * npts / (rr["ms"] * 1e-3)}
    # the re-gridding steps that follow the solves in the outer iteration (mapAlpha / mapPsi / mapTheta,
    # src/ModScbEuler.f90): timed once, last, because they move x, y, z
    try:
        m = {}
        for name, fn in (("map_alpha_ms", gpu.mapAlpha), ("map_psi_ms", gpu.mapPsi), ("map_theta_ms", gpu.mapTheta)):
            fail = fn()
            m[name] = gpu.last_ms() if not fail else None
        out.update(m)
    except Exception as e:      # informational entry: never take the bench line down
        out["map_error"] = str(e)[:200]
    gpu.close()
    return out


def measure_ram(workload_name, flags, steps, warmup, device, mode="fast", dist=None, rank=0, world=1, policy=0,
                do_e2e=True, do_profile=True, check=False, e2e_steps=None):
    """One RAM workload, measured the way the headline is: `value` with F2 resident (CUDA events over the library's run
    stream, L2 flushed between steps, max over ranks), `e2e` through the C ABI with the pinned HOST array going up and
    coming back every step (wall clock), per-kernel device times and the dominant kernel's roofline (1 GPU).  world > 1:
    the library's own sharded step (rsg_ram_run_sharded, ramscb_b200/csrc/ram_shard.inl), strong scaling."""
    import torch
    from ramscb_b200 import host, parallel, synthetic
    g, inp, desc = workload(workload_name)
    cells = g.nS * g.NR * g.NT * g.NE * g.NPA
    ops = ops_per_cell(g, flags)
    gpu = host.RamGpu(g, device=device, mode=host.MODE_FAST if mode == "fast" else host.MODE_EXACT)
    gpu.set_fields(inp)
    gpu.set_efield(inp.VT, inp.EIR, inp.EIP)
    gpu.set_boundary(inp.FGEOS)
    gpu.set_wavelo(inp.WALOS1, inp.WALOS2, inp.WALOS3, inp.Kp, inp.Kpmax12)
    gpu.set_plasmasphere(inp.NECR)
    D = None
    if flags & 5:
        D = synthetic.synthetic_daa(g, inp)          # synthetic Daa in ATAC / ATAW_emic_h
        gpu.set_diffcoef(1, D)
        gpu.set_diffcoef(2, D)
    sh, plan = None, None
    if world > 1:
        sh = parallel.RamPeerSharded(gpu, dist, rank, world, policy)
        plan = sh.plan
        sh.load(inp.F2)
    else:
        gpu.f2_h2d(inp.F2)
    F2_host = inp.F2.copy(order="F")
    host.host_register(F2_host)
    flush = torch.empty(512 * 1024 * 1024 // 4, dtype=torch.float32, device="cuda")

    def step():
        return sh.ram_run(DTS, flags=flags) if sh else gpu.ram_run(DTS, DtsMin=1.0, flags=flags)

    def barrier():
        if dist is not None and world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(warmup):
        step()
    out = {"workload": desc, "flags": flags, "ops_per_cell_per_step": ops, "cells": cells}
    if check and sh:
        # inside the same lease as the timing: this rank's share after `warmup` sharded steps against the one-GPU step
        ref = host.RamGpu(g, device=device, mode=host.MODE_FAST if mode == "fast" else host.MODE_EXACT)
        ref.set_inputs(inp)
        if D is not None:
            ref.set_diffcoef(1, D)
            ref.set_diffcoef(2, D)
        for _ in range(warmup):
            r1 = ref.ram_run(DTS, DtsMin=1.0, flags=flags)
        full = ref.f2_d2h()
        ref.close()
        mine = sh.store(np.full(inp.F2.shape, np.nan, order="F"))
        sl, lsl = slice(plan.s0, plan.s0 + plan.ns), slice(plan.l0, plan.l0 + plan.nl)
        same = bool(np.array_equal(mine[sl][..., lsl], full[sl][..., lsl]))
        t = torch.tensor([1 if same else 0], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        out["sharded_check"] = {"every_rank_share_of_F2_bit_identical_to_one_gpu_step": bool(int(t.item())),
                                "after_steps": warmup, "script": "same comparison as tests/multi_gpu_peer_check.py"}
        del full, mine
    launches0 = gpu.launch_count()
    clocks = ClockSampler(device)
    clocks.start()
    barrier()
    dev_ms = 0.0
    t_wall0 = time.perf_counter()
    for _ in range(steps):
        flush.zero_()                      # evict F2 from the 126 MB L2 (untimed)
        barrier()                          # ranks start the step together: the step itself contains device-side barriers
        gpu.timer_begin()
        step()
        dev_ms += gpu.timer_end()
        clocks.sample()
    barrier()
    out["wall_s_timed_region"] = time.perf_counter() - t_wall0
    out["clocks"] = clocks.stop()
    launches = gpu.launch_count() - launches0
    if dist is not None and world > 1:
        t = torch.tensor([dev_ms, float(launches)], dtype=torch.float64, device="cuda")
        dist.all_reduce(t[:1], op=dist.ReduceOp.MAX)
        dist.all_reduce(t[1:], op=dist.ReduceOp.SUM)
        dev_ms, launches = float(t[0].item()), int(t[1].item())
    ms = dev_ms / steps
    out.update(ms_per_step=ms, value=ops * cells / (ms * 1e-3), gpu_launches=int(launches), launches_per_step=launches / steps)

    if do_e2e:
        # ---- end to end through the C ABI with host buffers: the (share of the) host array goes up, the step
        # runs, the share comes back -- every step, as the routine-level drop-in does
        n = e2e_steps or max(3, min(steps, 10))
        per_l = g.nS * g.NR * g.NT * g.NE * 8
        if sh:
            up = plan.nl * per_l
            down = plan.nl * per_l
            if plan.ns < g.nS:
                up += plan.nl * per_l          # d2h of a species subset starts from the host image (other species kept)
        else:
            up = down = F2_host.nbytes
        up += 3 * inp.VT.nbytes
        down += (4 + 6 + 1) * g.nS * 8 + 2 * g.nS * g.NR * g.NT * 8
        def e2e_step(pipelined):
            gpu.set_efield(inp.VT, inp.EIR, inp.EIP)
            if sh:
                gpu.f2_h2d_shard(F2_host)
                step()
                gpu.f2_d2h_shard(F2_host)
            elif pipelined:       # one C call: upload | DRIFTR, DRIFTP per chunk of pitch angles ... | download
                gpu.ram_run_host(F2_host, DTS, DtsMin=1.0, flags=flags)
            else:
                gpu.f2_h2d(F2_host)
                step()
                gpu.f2_d2h(F2_host)

        seq_s = None
        if not sh:                # the three calls one after the other (round 1's e2e), for comparison
            e2e_step(False)
            barrier()
            t0 = time.perf_counter()
            for _ in range(max(3, n // 2)):
                e2e_step(False)
            barrier()
            seq_s = (time.perf_counter() - t0) / max(3, n // 2)
            F2_host[...] = inp.F2
            e2e_step(True)        # warm the pipelined path (stream, events)
        barrier()
        t0 = time.perf_counter()
        for _ in range(n):
            e2e_step(True)
        barrier()
        e2e_s = (time.perf_counter() - t0) / n
        tot = [float(up), float(down)]
        if dist is not None and world > 1:
            t = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            e2e_s = float(t.item())
            t = torch.tensor(tot, dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
            tot = [float(t[0].item()), float(t[1].item())]
        out["e2e"] = {"value": ops * cells / e2e_s, "unit": "cell-updates/s", "h2d_bytes_per_step": int(tot[0]),
                      "d2h_bytes_per_step": int(tot[1]), "ms_per_step": e2e_s * 1e3, "steps": n,
                      "timer": "wall clock around the C-ABI calls, max over ranks; pinned host F2 goes host->device and back "
                               "EVERY step (routine-level drop-in, INTEGRATION.md 3a): PCIe bound",
                      "call": "rsg_ram_f2_h2d_shard + rsg_ram_run_sharded + rsg_ram_f2_d2h_shard" if sh else
                              "rsg_ram_run_host (upload, step and download pipelined over chunks of pitch angles)"}
        if seq_s is not None:
            out["e2e"]["three_calls_ms_per_step"] = seq_s * 1e3
        if not sh:
            t0 = time.perf_counter()
            for _ in range(n):
                gpu.set_efield(inp.VT, inp.EIR, inp.EIP)
                step()
            torch.cuda.synchronize()
            res_s = (time.perf_counter() - t0) / n
            out["e2e"]["resident_state"] = {"ms_per_step": res_s * 1e3, "value": ops * cells / res_s,
                                            "h2d_bytes_per_step": int(3 * inp.VT.nbytes),
                                            "d2h_bytes_per_step": int((4 + 6 + 1) * g.nS * 8 + 2 * g.nS * g.NR * g.NT * 8),
                                            "note": "F2 stays on the device (the fused integration, INTEGRATION.md 3b); not the headline e2e"}

    peak, peak_src = measured_peak()
    if do_profile and not sh:
        # per-kernel device times: CUDA events recorded on the run stream between the stages of rsg_ram_run (graph
        # replay off for this pass), L2 flushed per step
        gpu.profile(True)
        npf = max(3, min(steps, 10))
        for _ in range(npf):
            flush.zero_()
            torch.cuda.synchronize()
            step()
        stages = gpu.profile_get()
        gpu.profile(False)
        per_kernel_ms = {k: v[0] / v[1] for k, v in stages.items() if v[1] and k != "end"}
        nw = ops - OPS_PER_STEP - (4.0 if flags & 2 else 0.0)          # WPADIF applications per cell, species average
        # cell-updates per cell one launch performs (16 B = one FP64 read + one write per cell-update)
        ops_per_launch = {"k_driftr": 1, "k_driftp": 1, "k_drifte": 1, "k_driftmu": 1,
                          "k_plane_rp": 2,          # DRIFTR + DRIFTP of every plane
                          "k_col_fused": 8 + nw}    # DRIFTE, DRIFTMU, [WPADIF], CHAREX, ATMOL, ATMOL, CHAREX, [WPADIF], DRIFTMU, DRIFTE
        sweeps = {k: per_kernel_ms[k] for k in ops_per_launch if k in per_kernel_ms}
        dom = max(sweeps, key=lambda n: sweeps[n])
        alg_bytes = 16.0 * cells * ops_per_launch[dom]
        achieved = alg_bytes / (sweeps[dom] * 1e-3) / 1e9
        step_sum = sum(v[0] for k, v in stages.items() if k != "end") / npf
        out["roofline"] = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                           "frac": achieved / peak, "traffic": profiled_traffic(workload_name, flags, dom), "peak_source": peak_src,
                           "algorithmic_bytes_per_launch": int(alg_bytes),
                           "cell_updates_per_cell_per_launch": ops_per_launch[dom],
                           "note": "achieved = 16 B per cell-update (SURVEY 8(d)) x the cell-updates of one launch (all cells of the 4 "
                                   "species x the operators the kernel fuses) / CUDA-event duration on the launching stream, L2 flushed "
                                   "per step.  A fused kernel MOVES 16 B per cell once for all its operators, so this is an "
                                   "HBM-equivalent throughput, not a bandwidth utilisation: `traffic` (ncu dram bytes, when the "
                                   "committed capture matches the current sources) and `whole_step` say what the DRAM sees.",
                           "per_kernel_ms": per_kernel_ms,
                           "per_kernel_frac_of_peak": {k: 16.0 * cells * ops_per_launch[k] / (v * 1e-3) / 1e9 / peak for k, v in sweeps.items()},
                           "kernel_share_of_step": {k: (v[0] / npf) / step_sum for k, v in stages.items() if k != "end"},
                           "whole_step": {"algorithmic_GBps": 16.0 * cells * ops / (ms * 1e-3) / 1e9,
                                          "frac": 16.0 * cells * ops / (ms * 1e-3) / 1e9 / peak,
                                          "three_pass_floor_ms": 3 * 16.0 * cells / (peak * 1e9) * 1e3}}
    elif sh:
        per_stage = None
        if do_profile:
            # per-stage device times of the sharded step (CUDA events on the run stream between its launches, graph replay
            # off for this pass); a stage that ends in a device-side barrier includes the wait for the slowest peer
            gpu.profile(True)
            npf = 5
            for _ in range(npf):
                flush.zero_()
                barrier()
                step()
            stages = gpu.profile_get()
            gpu.profile(False)
            names = [k for k in stages if k != "end"]
            t = torch.tensor([stages[k][0] / max(stages[k][1], 1) for k in names], dtype=torch.float64, device="cuda")
            tmax = t.clone()
            dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
            per_stage = {"rank0_ms": {k: float(v) for k, v in zip(names, t.tolist())},
                         "max_over_ranks_ms": {k: float(v) for k, v in zip(names, tmax.tolist())},
                         "column_resharding": "bulk push after the column kernel (RSG_PEER_PUSH=1)" if os.environ.get("RSG_PEER_PUSH", "0") not in ("", "0")
                         else "the column kernel's own write-back into the peers"}
        per_gpu = 16.0 * cells * ops / world / (ms * 1e-3) / 1e9
        sent = 0.0                                                # bytes this rank stores into peers per step (2 re-shardings)
        if plan.G > 1:
            mine = plan.ns * plan.nl * g.NE * g.NR * g.NT * 8.0
            sent = 2.0 * mine * (plan.G - 1) / plan.G
        out["roofline"] = {"bound": "hbm", "kernel": "whole sharded step (per GPU)", "achieved": per_gpu, "peak": peak, "unit": "GB/s",
                           "frac": per_gpu / peak, "traffic": None, "peak_source": peak_src,
                           "note": "16 B per cell-update x this GPU's share of the cell-updates / max-over-ranks step time",
                           "per_stage": per_stage,
                           "nvlink": {"peer_store_bytes_per_rank_per_step": int(sent),
                                      "link_floor_ms": sent / 770e9 * 1e3,
                                      "note": "F2 values stored straight into the consuming rank's buffer by the producing kernel's "
                                              "write-back (no separate exchange pass); floor = bytes / 770 GB/s measured peer bandwidth "
                                              "(B200_PROFILING.md)"}}
    out["plan"] = plan.as_dict() if plan is not None else None
    host.host_unregister(F2_host)
    gpu.close()
    del flush
    torch.cuda.empty_cache()
    return out, g, inp


def profiled_traffic(workload_name, flags, kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of `kernel` from the committed `ncu --set full` capture
    (profiles/traffic.json) -- only when that capture was taken from the kernel sources as they are now (sha of
    ramscb_b200/csrc/ram_*.cuh recorded beside it); else None (a stale number is no evidence about this run)."""
    import hashlib
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            t = json.load(f)
        h = hashlib.sha1()
        for fn in ("ram_kernels.cuh", "ram_fused.cuh"):
            with open(os.path.join(ROOT, "ramscb_b200", "csrc", fn), "rb") as f:
                h.update(f.read())
        if t.get("sources_sha1") != h.hexdigest():
            return None
        return t.get(f"{workload_name}_flags{flags}", {}).get(kernel)
    except (OSError, ValueError):
        return None


def time_ram_step(workload_name, flags, steps, warmup, device):
    """informational extras: resident step + per-kernel times of another workload / flag set"""
    out, _, _ = measure_ram(workload_name, flags, steps, warmup, device, do_e2e=False)
    keep = ("workload", "flags", "ms_per_step", "value", "ops_per_cell_per_step", "launches_per_step")
    r = {k: out[k] for k in keep}
    r["unit"] = "cell-updates/s"
    r["per_kernel_ms"] = out.get("roofline", {}).get("per_kernel_ms")
    return r


def scb_zeta_metrics(device):
    """iterateAlpha through the zeta-sharded protocol (rsg_scb_zsolve_*, one launch per half-sweep) on ONE
    rank, next to the on-chip cluster solve: what the host-driven scheme costs per sweep before any halo
    traffic.  Wall clock around the whole solve (the protocol is host-driven); results checked identical."""
    import torch
    from ramscb_b200 import host, parallel, scb_synthetic
    inp = scb_synthetic.build_scb(nthe=101, npsi=45, nzeta=97, warp=0.2)
    ref = host.ScbGpu(inp, device=device)
    ref.computeBandJacob(); ref.metrica(); ref.newk()
    r1 = ref.iterateAlpha(1e-6, ordering=host.SOR_COLOR4)
    gpu = host.ScbGpu(inp, device=device)
    st = torch.cuda.Stream()
    out = {}
    with torch.cuda.stream(st):
        gpu.set_stream(st.cuda_stream)
        gpu.computeBandJacob(); gpu.metrica(); gpu.newk()
        alfa0 = gpu.get_field("alfa").copy()
        for poll in (16, 64):
            gpu.set_field("alfa", alfa0)
            z = parallel.ScbZetaSharded(gpu, None, 0, 1, on_cuda=True, poll=poll)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            r = z.iterate(1e-6)
            torch.cuda.synchronize()
            ms = (time.perf_counter() - t0) * 1e3
            same = bool(np.array_equal(gpu.get_field("alfa"), ref.get_field("alfa")) and np.array_equal(r["ni"], r1["ni"]))
            out[f"poll{poll}"] = {"wall_ms": ms, "sweeps_launched": r["sweeps_launched"], "us_per_sweep": ms * 1e3 / r["sweeps_launched"],
                                  "identical_to_cluster_solve": same}
    out["cluster_solve_ms"] = r1["ms"]
    out["max_sweeps"] = int(r1["nisave"])
    gpu.close(); ref.close()
    return out


def scb_run_metrics(device):
    """configs[3] end to end: scb_run (src/ModScbRun.f90:149-440) with the reference's parameters (InCon 1e-6,
    MinSCBIterations 11, blend 0.5) on the default SCB grid through ONE rsg_scb_run call, 4-colour ordering.  Two variants:
    `device_front_end` -- `pressure` entirely on the device from synthetic RAM pressures (rsg_scb_set_ram_pressure; no host
    hop inside the outer iteration) -- and `host_callback` (round 1: the 2-D front end as a host callback, a synthetic
    analytic pressure).  Wall clock of the call: one cold run, then the fastest of three warm ones."""
    from ramscb_b200 import grids, host, scb_synthetic
    inp = scb_synthetic.build_scb(nthe=101, npsi=45, nzeta=97, warp=0.2)
    fn = scb_synthetic.equatorial_pressure_fn()
    g = grids.build_grids()
    PPerT, PParT, scb, LZ, PHI = scb_synthetic.synthetic_ram_pressures(g)
    res = {}
    for name, cb in (("device_front_end", None), ("host_callback", fn)):
        out = {}
        for rep in range(4):                       # one cold run, then the fastest of three warm ones (wall clock on a shared host)
            gpu = host.ScbGpu(inp, device=device)
            gpu.set_map_targets(inp.alphaVal, inp.psiVal, inp.chiVal)
            if cb is None:
                gpu.set_ram_pressure(PPerT, PParT, scb, LZ, PHI)
            n0 = gpu.launch_count()
            t0 = time.perf_counter()
            r = gpu.scb_run(cb, ordering=host.SOR_COLOR4)
            ms = (time.perf_counter() - t0) * 1e3
            if rep >= 2 and ms >= out["wall_ms"]:
                gpu.close()
                continue
            out = {"wall_ms": ms, "outer_iterations": r["iterations"], "SORFail": r["SORFail"], "iConvGlobal": r["iConvGlobal"],
                   "nisaveAlpha_last": r["nisaveAlpha"], "nisavePsi_last": r["nisavePsi"], "blendRetries": r["blendRetries"],
                   "normDiff_start_end": [r["normDiffStart"], r["normDiff"]], "normJxB_start_end": [r["normJxBStart"], r["normJxB"]],
                   "normGradP_start_end": [r["normGradPStart"], r["normGradP"]], "launches": int(gpu.launch_count() - n0),
                   "ms_per_outer_iteration": ms / max(r["iterations"], 1)}
            gpu.close()
        res[name] = out
    res.update(res["device_front_end"])            # the headline entries = the callback-free run
    return res


def hi_metrics(device):
    """computehI's integral block (src/ModRamScb.f90:372-410) through rsg_hI_integrals on the field lines of the
    default RAM grid (20 x 25 lines) and of the configs[2] grid (80 x 49), nthe = 101, NPA = 72: device time of the
    kernel (CUDA events inside the call) and wall clock of the whole call (host arrays in and out)."""
    from ramscb_b200 import grids, host, scb_synthetic
    out = {}
    for name, kw in (("default_20x25", {}), ("x4_80x49", dict(NR=80, NT=49, NE=70, energy_refine=2))):
        g = grids.build_grids(**kw)
        d = scb_synthetic.ram_field_lines(g.LZ[1:g.NR + 1] if len(g.LZ) > g.NR else g.LZ, g.MLT[:g.NT], nthe=101, wiggle=0.05)
        ms, wall = [], []
        for _ in range(5):
            t0 = time.perf_counter()
            r = host.hI_integrals(mu=g.MU, device=device, **d)
            wall.append((time.perf_counter() - t0) * 1e3)
            ms.append(r[4])
        lines = g.NR * g.NT
        bytes_alg = lines * (5 * 101 + 3 * g.NPA + 1) * 8
        k = sorted(ms)[len(ms) // 2]
        try:        # the rest of computehI (rsg_hI_tail: 4 kernels) on the arrays just computed
            rng = np.random.default_rng(1)
            shape3 = (g.NR + 1, g.NT, g.NPA)
            ram = {n: np.asfortranarray(rng.random(shape3)) for n in ("FNHS", "FNIS", "BOUNHS", "BOUNIS", "HDNS")}
            ram["BNES"] = np.asfortranarray(1e-7 * rng.random((g.NR + 1, g.NT)))
            Lz = g.LZ[:g.NR + 1] if len(g.LZ) > g.NR else np.append(2 * g.LZ[0] - g.LZ[1], g.LZ)
            tms, twall = [], []
            for _ in range(5):
                t0 = time.perf_counter()
                tr = host.hI_tail(r[0], r[1], np.where(np.isfinite(r[2]), r[2], 1.0), r[3], np.zeros(g.NT, dtype=np.int32),
                                  d["outsideMGNP"], Lz, g.PA, g.PAbn, 1, 300.0, ram, device=device)
                twall.append((time.perf_counter() - t0) * 1e3)
                tms.append(tr["ms"])
            tail = {"kernels_ms": sorted(tms)[2], "call_wall_ms": sorted(twall)[2], "launches": 4, "integral_smooth": 1}
        except Exception as e:
            tail = {"error": str(e)[:200]}
        out[name] = {"tail": tail, "lines": lines, "nthe": 101, "NPA": g.NPA, "kernel_ms": k, "call_wall_ms": sorted(wall)[len(wall) // 2],
                     "integrals_per_s": 3 * lines * g.NPA / (k * 1e-3) if k > 0 else None, "algorithmic_bytes": bytes_alg,
                     "GBps": bytes_alg / (k * 1e-3) / 1e9 if k > 0 else None,
                     "bound": "latency (one CTA per line, 500 / 3920 CTAs, serial chain over NPA): far below the HBM roofline by construction"}
    try:        # the first block of computehI: SCB field lines -> RAM field lines (rsg_hI_convert_lines), configs[3] SCB grid
        inp = scb_synthetic.build_scb(nthe=101, npsi=45, nzeta=97, warp=0.2)
        rr = np.sqrt(inp.x ** 2 + inp.y ** 2 + inp.z ** 2)
        bf = np.asfortranarray(30574.0 / rr ** 3 * np.sqrt(1.0 + 3.0 * (inp.z / rr) ** 2))
        g = grids.build_grids()
        Lz = g.LZ[:g.NR + 1] if len(g.LZ) > g.NR else np.append(2 * g.LZ[0] - g.LZ[1], g.LZ)
        ms, wall = [], []
        for _ in range(3):
            t0 = time.perf_counter()
            r = host.hI_convert_lines(inp.x, inp.y, inp.z, bf, inp.psi, inp.alfa, Lz, g.MLT[:g.NT], 51, device=device)
            wall.append((time.perf_counter() - t0) * 1e3)
            ms.append(r[5])
        nq = int((r[4] == 0).sum()) * 101
        k = sorted(ms)[1]
        out["convert_lines_default"] = {"kernels_ms": k, "call_wall_ms": sorted(wall)[1], "queries": nq, "candidates_per_query": 45 * 96,
                                        "passes": 1, "distance_evaluations_per_s": nq * 45 * 96 / (k * 1e-3) if k > 0 else None,
                                        "bound": "shared-memory bandwidth / FP64 issue (candidates resident in shared memory, no HBM traffic to speak of)"}
    except Exception as e:
        out["convert_lines_default"] = {"error": str(e)[:200]}
    try:        # the whole routine resident on the device (rsg_hi): nothing but the SCB arrays go in, or nothing at all
        from ramscb_b200 import synthetic
        inp = scb_synthetic.build_scb(nthe=101, npsi=45, nzeta=97, warp=0.2)
        g = grids.build_grids()
        Lz = g.LZ[:g.NR + 1] if len(g.LZ) > g.NR else np.append(2 * g.LZ[0] - g.LZ[1], g.LZ)
        rng = np.random.default_rng(1)
        shape3 = (g.NR + 1, g.NT, g.NPA)
        ram = {n: np.asfortranarray(rng.random(shape3)) for n in ("FNHS", "FNIS", "BOUNHS", "BOUNIS", "HDNS")}
        ram["BNES"] = np.asfortranarray(1e-7 * rng.random((g.NR + 1, g.NT)))
        sg = host.ScbGpu(inp, device=device)
        sg.computeBandJacob()                                     # leaves bf on the device
        scb = dict(x=inp.x, y=inp.y, z=inp.z, psi=inp.psi, alfa=inp.alfa, bf=sg.get_field("bf"))
        hi = host.HiGpu(101, 45, 97, Lz, g.MLT[:g.NT], g.MU, g.PA, g.PAbn, inp.chiVal, 51, 1.0, device=device)
        hi.set_ram_fields(ram)
        res = {}
        for label, src in (("scb_arrays_from_host", scb), ("scb_arrays_from_rsg_scb_handle", sg)):
            wall, ms = [], []
            for _ in range(7):
                t0 = time.perf_counter()
                hi.computehI(src, 300.0, True)
                wall.append((time.perf_counter() - t0) * 1e3)
                ms.append(hi.last_ms())
            res[label] = {"call_wall_ms": sorted(wall)[3], "kernels_ms": sorted(ms)[3]}
        rg = host.RamGpu(g, device=device)
        rg.set_inputs(synthetic.make_inputs(g, f2_kind="smooth"))
        t0 = time.perf_counter()
        hi.push_to_ram(rg)
        res["push_to_rsg_ram_wall_ms"] = (time.perf_counter() - t0) * 1e3
        res["launches_per_call"] = 9
        res["note"] = ("rsg_computehI: convert -> ScaleAt -> RAIRDEN -> integrals -> tail on one stream, intermediates resident; wall "
                       "includes the final synchronisation (and the 21 MB upload of the six SCB arrays in the host variant)")
        out["computehI_resident_default"] = res
        rg.close(); hi.close(); sg.close()
    except Exception as e:
        out["computehI_resident_default"] = {"error": str(e)[:300]}
    return out


def coupled_cycle_metrics():
    """BASELINE configs[4]'s physics in miniature: one coupled RAM <-> SCB cycle composed from the device entry points on ONE
    GPU at the default grids (scripts/time_coupled_cycle.py: 300 s of RAM steps -> pressure -> scb_run -> computehI -> new
    fields into the RAM state, three consecutive cycles; wall clock per phase).  The last cycle is reported."""
    r = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "time_coupled_cycle.py")], capture_output=True, text=True, timeout=120)
    d = json.loads(r.stdout.strip().splitlines()[-1])
    out = dict(d["cycles"][-1])
    out["cycles_run"] = len(d["cycles"])
    out["all_cycles_SORFail_0"] = all(c["SORFail"] == 0 for c in d["cycles"])
    out["F2_finite"] = d["F2_finite"]
    return out


def extras_main(device):
    """`bench.py --extras-only`: informational measurements beside the headline line (run by the N = 1 bench in a
    child process, so that nothing here can take the headline down).  One JSON object on stdout."""
    out = {}
    jobs = (("ram_default_wpi_emic", lambda: time_ram_step("default", 5, 10, 3, device)),
            ("ram_x4_no_wpi", lambda: time_ram_step("x4", 0, 5, 3, device)),
            ("ram_default_coulomb", lambda: time_ram_step("default", 2, 10, 3, device)),
            ("scb_alpha_zeta_protocol_one_rank", lambda: scb_zeta_metrics(device)),
            ("scb_run_configs3", lambda: scb_run_metrics(device)),
            ("computehI_integrals", lambda: hi_metrics(device)),
            ("coupled_cycle_default_grids", lambda: coupled_cycle_metrics()))
    for name, fn in jobs:
        t0 = time.perf_counter()
        try:
            out[name] = fn()
        except Exception as e:           # informational
            out[name] = {"error": str(e)[:300]}
        out[name]["wall_s"] = time.perf_counter() - t0
    print("EXTRAS_JSON " + json.dumps(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="x4", choices=["default", "x4"],
                    help="x4 = BASELINE configs[2] grid (the headline at every N), default = configs[1] grid")
    ap.add_argument("--flags", type=int, default=None,
                    help="rsg_ram_run operator flags: 1 = WPI pitch-angle diffusion (electrons), 4 = EMIC (H+), 2 = Coulomb; "
                         "default 5 on the x4 grid (configs[2]: full step with WPADIF), 0 on the default grid (configs[1])")
    ap.add_argument("--policy", default=os.environ.get("RSG_SHARD_POLICY", "slabs"), choices=["species", "slabs"],
                    help="N > 1: slabs (default) = every rank holds a pitch-angle slab of all species: an even split whatever "
                         "operators each species runs, and the rank's share of the host array F2(nS,NR,NT,NE,NPA) is one contiguous "
                         "run (e2e 12 ms at N = 8 against 55 ms); species = whole species per rank up to 4 ranks, 2 ranks per "
                         "species at 8: 4 %% faster resident, but the host copies are strided (profiles/r2/scaling_r2.txt)")
    ap.add_argument("--no-scb", action="store_true", help="skip the SCB solve metrics")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the informational extras")
    ap.add_argument("--no-configs1", action="store_true", help="skip the default-grid (configs[1]) secondary measurement")
    ap.add_argument("--no-check", action="store_true", help="N > 1: skip the in-run comparison of the sharded step with the one-GPU step")
    ap.add_argument("--extras-only", action="store_true", help=argparse.SUPPRESS)
    ap.add_argument("--mode", default="fast", choices=["fast", "exact"],
                    help="arithmetic mode of the sweeps (include/ramscb_gpu.h rsg_mode)")
    a = ap.parse_args()
    a.warmup = max(a.warmup, 3)
    if a.flags is None:
        a.flags = 5 if a.workload == "x4" else 0

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if a.extras_only:
        import torch
        torch.cuda.set_device(local_rank)
        extras_main(local_rank)
        return

    unit = "cell-updates/s"
    metric = "RAM phase-space cell-updates/s per full RAM step (8 drift sweeps + losses + WPI/EMIC pitch-angle diffusion, 4 species)"

    if a.impl == "reference":
        if rank != 0:
            return
        g, inp, desc = workload(a.workload)
        # the 4x grid takes ~8 s per step on the reference's 4 species threads: a bounded sample of whole steps
        cap = 40 if a.workload == "default" else 2
        steps, wu = max(1, min(a.steps, cap)), (max(1, min(a.warmup, 3)) if a.workload == "default" else 1)
        v, nthreads, dt = cpu_reference(g, inp, steps, wu, flags=a.flags)
        line = {"impl": "reference", "metric": metric, "value": v, "unit": unit, "n_gpus": a.gpus, "steps": steps,
                "warmup": wu, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": desc + f"; flags={a.flags}"},
                "cpu_baseline": {"value": v, "unit": unit, "cores": nthreads, "kind": "port",
                                 "sample": f"{steps} full ram_run steps of the same workload (flags {a.flags}), OpenMP over species "
                                           f"({nthreads} threads: the reference's own decomposition, src/ModRamRun.f90:64); the "
                                           "reference Fortran cannot be built in this image, so this is the C++ restatement "
                                           "oracle/ram_oracle.cpp (-O3 -march=native)"},
                "e2e": {"value": v, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return

    import torch
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local_rank)
    from ramscb_b200 import parallel as _par
    numa = {"bound": False, "why": "RSG_NO_NUMA_BIND"} if os.environ.get("RSG_NO_NUMA_BIND") else _par.bind_to_gpu_numa(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    from ramscb_b200 import host
    policy = host.SHARD_SPECIES if a.policy == "species" else host.SHARD_SLABS

    r, g, inp = measure_ram(a.workload, a.flags, a.steps, a.warmup, local_rank, mode=a.mode, dist=dist, rank=rank, world=world,
                            policy=policy, check=(world > 1 and not a.no_check))
    plan = r["plan"]
    if world == 1:
        par = "1 GPU, all species per launch"
    elif plan["G"] == 1:
        par = (f"{world} ranks, {plan['ns']} whole species each (rsg_ram_run_sharded, policy species): no F2 exchange; result "
               "blocks gathered over NVLink peer memory and reduced on the device")
    else:
        par = (f"{world} ranks, {plan['G']} per species group ({'all species' if plan['ns'] > 1 else 'one species'} per group): "
               "pitch-angle slabs for DRIFTR/DRIFTP, plane-position blocks for the column kernel; the 2 re-shardings per step are "
               "the kernels' own write-backs into the consuming rank's buffer over NVLink peer memory (CUDA IPC), device-side "
               "barriers, one CUDA graph per rank, no NCCL on the data path")
    line = {"metric": metric, "value": r["value"], "unit": unit, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": r["workload"] + f"; flags={a.flags}", "ops_per_cell_per_step": r["ops_per_cell_per_step"],
                       "mode": ("fast: separable coefficients + FMA + division-free limiter, fused shared-memory kernels (parity: "
                                "PARITY.md, tests/test_ram_parity_gpu.py)") if a.mode == "fast"
                       else "exact: reference operation order, bit-identical to the oracle",
                       "l2": "flushed between timed steps (512 MB memset, untimed)", "DTs": DTS, "parallelism": par,
                       "host_numa_binding_rank0": numa,
                       "same_workload_at_every_N": "strong scaling: bench.py --gpus 1 runs this same workload on one GPU"},
            "clocks": r["clocks"], "e2e": r.get("e2e"), "gpu_launches": r["gpu_launches"], "roofline": r.get("roofline"),
            "wall_s_timed_region": r["wall_s_timed_region"]}
    if "sharded_check" in r:
        line["config"]["sharded_check"] = r["sharded_check"]
    if plan is not None:
        line["config"]["shard_plan_rank0"] = plan

    if rank == 0 and world == 1 and a.mode == "fast":
        # the bit-faithful arithmetic mode (reference operation order, no FMA contraction) on the same workload
        try:
            ex, _, _ = measure_ram(a.workload, a.flags, 5, 3, local_rank, mode="exact", do_e2e=False, do_profile=False)
            line["config"]["modes"] = {"fast_ms_per_step": r["ms_per_step"], "exact_ms_per_step": ex["ms_per_step"],
                                       "exact": "reference operation order, sweeps bit-identical to the oracle (PARITY.md)"}
            if line.get("roofline") is not None:
                line["roofline"]["exact_mode_ms_per_step"] = ex["ms_per_step"]
        except Exception as e:
            line["config"]["modes"] = {"error": str(e)[:200]}
    if rank == 0 and world == 1 and not a.no_configs1 and a.workload == "x4":
        # BASELINE configs[1] (default grid, drift + loss step) measured the same way; secondary to the headline.  Kept
        # inside `roofline` / `e2e` too, because the driver's parsed line keeps those objects.
        try:
            c1, _, _ = measure_ram("default", 0, a.steps, a.warmup, local_rank, mode=a.mode)
            c1s = {k: c1[k] for k in ("workload", "flags", "ms_per_step", "value", "ops_per_cell_per_step", "launches_per_step")}
            c1s["unit"] = unit
            c1s["e2e"] = c1.get("e2e")
            c1s["roofline"] = c1.get("roofline")
            line["configs1"] = c1s
            if line.get("roofline") is not None:
                rf = c1.get("roofline") or {}
                line["roofline"]["configs1_default_grid"] = {"ms_per_step": c1["ms_per_step"], "value": c1["value"], "unit": unit,
                                                             "kernel": rf.get("kernel"), "frac": rf.get("frac"),
                                                             "per_kernel_ms": rf.get("per_kernel_ms"),
                                                             "per_kernel_frac_of_peak": rf.get("per_kernel_frac_of_peak"),
                                                             "e2e_value": (c1.get("e2e") or {}).get("value"),
                                                             "e2e_ms_per_step": (c1.get("e2e") or {}).get("ms_per_step")}
        except Exception as e:
            line["configs1"] = {"error": str(e)[:300]}
    if rank == 0 and world == 1 and not a.no_scb:
        line["scb"] = scb_metrics(local_rank)
        if line.get("roofline") is not None:
            # the second half of BASELINE's metric where the driver's parsed line keeps it
            sc = line["scb"]
            line["roofline"]["scb_sor"] = {k: sc[k] for k in ("grid", "iterate_alpha", "iterate_psi", "bandjacob_ms", "metrica_ms", "metric_ms")
                                           if k in sc}
    if rank == 0 and world == 1 and not a.no_cpu_baseline and not a.no_scb:
        # CPU side of configs[3]: the oracle's scb_run (same loop, same parameters, OpenMP over the sub-problems like
        # the reference) on the host cores, once -- next to extras.scb_run_configs3
        try:
            from oracle import oracle
            from ramscb_b200 import scb_synthetic
            oracle.build()
            sinp = scb_synthetic.build_scb(nthe=101, npsi=45, nzeta=97, warp=0.2)
            so = oracle.ScbOracle(sinp)
            t0 = time.perf_counter()
            sr = so.scb_run(scb_synthetic.equatorial_pressure_fn())
            line["scb"]["cpu_scb_run"] = {"wall_s": time.perf_counter() - t0, "outer_iterations": sr["iterations"],
                                          "SORFail": sr["SORFail"], "cores": os.cpu_count(), "kind": "port"}
        except Exception as e:
            line["scb"]["cpu_scb_run"] = {"error": str(e)[:200]}
    if rank == 0 and world == 1 and not a.no_cpu_baseline:      # reported at N = 1 only
        big = a.workload == "x4"
        v, nthreads, dt = cpu_reference(g, inp, 1 if big else 3, 1, flags=a.flags)
        line["cpu_baseline"] = {"value": v, "unit": unit, "cores": nthreads, "kind": "port",
                                "sample": f"{1 if big else 3} full ram_run step(s) of the same workload on the host cores after 1 warm-up "
                                          f"({dt:.3f} s/step, OpenMP over species = the reference's decomposition, {nthreads} threads)"}
        try:
            line["cpu_baseline"]["all_cores"] = cpu_all_cores()
        except Exception as e:
            line["cpu_baseline"]["all_cores"] = {"error": str(e)[:200]}
    if rank == 0 and world == 1 and not a.no_extras and a.workload == "x4" and a.flags == 5:
        # informational, in a child process: other flag sets, the zeta-sharded SOR protocol, scb_run, computehI
        try:
            r2 = subprocess.run([sys.executable, os.path.abspath(__file__), "--extras-only"], capture_output=True, text=True,
                                timeout=420, env=dict(os.environ, LOCAL_RANK=str(local_rank)))
            tag = [l for l in r2.stdout.splitlines() if l.startswith("EXTRAS_JSON ")]
            line["extras"] = json.loads(tag[-1][len("EXTRAS_JSON "):]) if tag else {"error": (r2.stderr or r2.stdout)[-300:]}
            if line.get("roofline") is not None and "scb_run_configs3" in line["extras"]:
                line["roofline"].setdefault("scb_sor", {})["scb_run_configs3"] = line["extras"]["scb_run_configs3"]
            if line.get("roofline") is not None and "coupled_cycle_default_grids" in line["extras"]:
                line["roofline"].setdefault("scb_sor", {})["coupled_cycle_default_grids"] = line["extras"]["coupled_cycle_default_grids"]
        except Exception as e:
            line["extras"] = {"error": str(e)[:300]}
    if rank == 0:
        print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
