#!/usr/bin/env python
"""bench.py -- RAM phase-space cell-updates/s per full RAM step (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--workload default|x4] [--scaling weak|strong]

One "step" = one pass of the RAM hot path (`ram_run`, src/ModRamRun.f90:64-222)
over one synthetic input set: for each of the 4 species CEPARA, DRIFTPARA,
DRIFTR/P/E/MU, SUMRC, [WAVELO | CHAREXCHANGE], ATMOL x2, reversed order back to
DRIFTR, then the epilogue, ANISCH pressures and the CFL time step.  Work unit:
one cell-update = one F2 cell advanced by one operator call; a step applies 12
operators (8 drift sweeps + 2 ATMOL + 2 CHAREXCHANGE-or-WAVELO) to nS*NR*NT*NE*NPA
cells.

Printed JSON keys follow the driver contract: `value` is measured with F2 resident
in HBM (CUDA events across the library's streams, L2 flushed between steps);
`e2e` is the same metric through the C-ABI with HOST buffers (pinned), F2
host->device and device->host copies inside the timed region (wall clock).
`--impl reference` times the reference's CPU algorithm (the C++ oracle, the
reference's own OpenMP-over-species parallelism) on the same workload.

N > 1 (one process per GPU under torchrun).  Species are the independent unit of `ram_run`
(the OpenMP loop of src/ModRamRun.f90:64), so the default is WEAK scaling with no data-path
collective: the job advances 4*N species of the named grid, 4 per rank, every rank running exactly
the N = 1 path on its own species; `value` = cell-updates of all ranks / max-over-ranks device time.
`--scaling strong` keeps the 4 species fixed and shards them (and, beyond 4 ranks, pitch-angle /
energy slabs of a species with two NCCL re-shardings per step; ramscb_b200/parallel.py) -- the
numbers of that mode are in profiles/r1/scaling_r1.txt.
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


def cpu_reference(g, inp, steps, warmup, groups=1):
    """The reference's CPU algorithm (C++ oracle; OpenMP over species like
    src/ModRamRun.f90:64) on the same workload.  `groups` > 1 is the weak-scaling job of N GPUs
    (4*N species): one oracle instance per group of 4 species, run concurrently, as the reference's
    species loop would spread 4*N species over the host threads.
    Returns (cell-updates/s, threads, s/step)."""
    from concurrent.futures import ThreadPoolExecutor
    from oracle import oracle
    oracle.build()
    ncpu = os.cpu_count() or 1
    groups = max(1, min(groups, max(1, ncpu // g.nS)))   # no more instances than the cores can hold
    os_ = [oracle.RamOracle(g, inp, DTs=DTS) for _ in range(groups)]
    nthreads = min(g.nS, ncpu)

    def one(o):
        o.ram_run(flags=0, nthreads=nthreads)

    def step():
        if groups == 1:
            one(os_[0])
        else:
            with ThreadPoolExecutor(groups) as ex:      # ctypes releases the GIL inside the call
                list(ex.map(one, os_))

    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / steps
    cells = groups * g.nS * g.NR * g.NT * g.NE * g.NPA
    return OPS_PER_STEP * cells / dt, nthreads * groups, dt


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


def time_ram_step(workload_name, flags, steps, warmup, device):
    """Device time of one resident `ram_run` of the named workload with the given operator flags, measured
    exactly like the headline value (L2 flushed between steps, CUDA events over the library's streams)."""
    import torch
    from ramscb_b200 import host, synthetic
    g, inp, desc = workload(workload_name)
    gpu = host.RamGpu(g, device=device, mode=host.MODE_FAST)
    gpu.set_inputs(inp)
    nw = 0
    if flags:
        D = synthetic.synthetic_daa(g, inp)
        gpu.set_diffcoef(1, D)
        gpu.set_diffcoef(2, D)
        nw = sum(1 for sp in g.species if (flags & 1 and sp.WPI) or (flags & 4 and sp.EMIC))
    flush = torch.empty(512 * 1024 * 1024 // 4, dtype=torch.float32, device="cuda")
    for _ in range(warmup):
        gpu.ram_run(DTS, DtsMin=1.0, flags=flags)
    n0 = gpu.launch_count()
    ms = 0.0
    for _ in range(steps):
        flush.zero_()
        torch.cuda.synchronize()
        gpu.timer_begin()
        gpu.ram_run(DTS, DtsMin=1.0, flags=flags)
        ms += gpu.timer_end()
    ms /= steps
    launches = (gpu.launch_count() - n0) / steps
    gpu.profile(True)
    for _ in range(3):
        flush.zero_()
        torch.cuda.synchronize()
        gpu.ram_run(DTS, DtsMin=1.0, flags=flags)
    stages = gpu.profile_get()
    gpu.profile(False)
    cells = g.nS * g.NR * g.NT * g.NE * g.NPA
    ops = OPS_PER_STEP + 2.0 * nw / g.nS + (4.0 if flags & 2 else 0.0)     # Coulomb: COULEN + COULMU in both half steps, all species
    out = {"workload": desc, "flags": flags, "ms_per_step": ms, "value": ops * cells / (max(ms, 1e-9) * 1e-3), "unit": "cell-updates/s",
           "ops_per_cell_per_step": ops, "launches_per_step": launches, "steps": steps, "warmup": warmup,
           "per_kernel_ms": {k: v[0] / v[1] for k, v in stages.items() if v[1] and k != "end"}}
    gpu.close()
    del flush
    torch.cuda.empty_cache()
    return out


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
    MinSCBIterations 11, blend 0.5) on the default SCB grid through ONE rsg_scb_run call; the synthetic
    pressure front end is the host callback.  Wall clock of the call (it includes the callback and the
    2-D transfers), 4-colour ordering."""
    from ramscb_b200 import host, scb_synthetic
    inp = scb_synthetic.build_scb(nthe=101, npsi=45, nzeta=97, warp=0.2)
    fn = scb_synthetic.equatorial_pressure_fn()
    out = {}
    for rep in range(2):                       # the second run is the warm one
        gpu = host.ScbGpu(inp, device=device)
        gpu.set_map_targets(inp.alphaVal, inp.psiVal, inp.chiVal)
        n0 = gpu.launch_count()
        t0 = time.perf_counter()
        r = gpu.scb_run(fn, ordering=host.SOR_COLOR4)
        ms = (time.perf_counter() - t0) * 1e3
        out = {"wall_ms": ms, "outer_iterations": r["iterations"], "SORFail": r["SORFail"], "iConvGlobal": r["iConvGlobal"],
               "nisaveAlpha_last": r["nisaveAlpha"], "nisavePsi_last": r["nisavePsi"], "blendRetries": r["blendRetries"],
               "normDiff_start_end": [r["normDiffStart"], r["normDiff"]], "normJxB_start_end": [r["normJxBStart"], r["normJxB"]],
               "normGradP_start_end": [r["normGradPStart"], r["normGradP"]], "launches": int(gpu.launch_count() - n0),
               "ms_per_outer_iteration": ms / max(r["iterations"], 1)}
        gpu.close()
    return out


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
    return out


def extras_main(device):
    """`bench.py --extras-only`: informational measurements beside the headline line (run by the N = 1 bench in a
    child process, so that nothing here can take the headline down).  One JSON object on stdout."""
    out = {}
    jobs = (("ram_default_wpi_emic", lambda: time_ram_step("default", 5, 10, 3, device)),
            ("ram_x4_configs2_wpi_emic", lambda: time_ram_step("x4", 5, 5, 3, device)),
            ("ram_x4_no_wpi", lambda: time_ram_step("x4", 0, 5, 3, device)),
            ("ram_default_coulomb", lambda: time_ram_step("default", 2, 10, 3, device)),   # one kernel per operator (no fused Coulomb stage)
            ("scb_alpha_zeta_protocol_one_rank", lambda: scb_zeta_metrics(device)),
            ("scb_run_configs3", lambda: scb_run_metrics(device)),
            ("computehI_integrals", lambda: hi_metrics(device)))
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
    ap.add_argument("--workload", default="default", choices=["default", "x4"])
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="N > 1: weak = 4 species per rank (4*N in the job), strong = the 4 species sharded over the ranks")
    ap.add_argument("--flags", type=int, default=0,
                    help="rsg_ram_run operator flags: 1 = WPI pitch-angle diffusion (electrons), 4 = EMIC (H+); 5 with "
                         "--workload x4 is BASELINE configs[2] (full step with WPADIF); default 0 = configs[1]")
    ap.add_argument("--no-scb", action="store_true", help="skip the SCB solve metrics")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the informational extras (WPI/EMIC steps, 4x grid, zeta protocol)")
    ap.add_argument("--extras-only", action="store_true", help=argparse.SUPPRESS)
    ap.add_argument("--mode", default="fast", choices=["fast", "exact"],
                    help="arithmetic mode of the sweeps (include/ramscb_gpu.h rsg_mode)")
    a = ap.parse_args()
    a.warmup = max(a.warmup, 3)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if a.extras_only:
        import torch
        torch.cuda.set_device(local_rank)
        extras_main(local_rank)
        return

    g, inp, desc = workload(a.workload)
    weak = a.scaling == "weak"
    jobs = world if weak else 1                      # independent 4-species sets in the job
    cells = g.nS * g.NR * g.NT * g.NE * g.NPA        # cells one rank's species set holds
    ops_per_step = float(OPS_PER_STEP)               # operator applications per cell, averaged over the species
    if a.flags:
        if a.flags & ~5 or a.impl == "reference" or a.scaling == "strong" and world > 1:
            raise SystemExit("--flags supports 1 (WPI) and 4 (EMIC) on our arm, 1 GPU or weak scaling")
        nw = sum(1 for sp in g.species if (a.flags & 1 and sp.WPI) or (a.flags & 4 and sp.EMIC))
        ops_per_step += 2.0 * nw / g.nS               # two WPADIF applications for each species that diffuses
        desc += f"; flags={a.flags}: WPADIF twice per step for {nw} of the {g.nS} species"
    if weak and world > 1:
        desc += f"; weak scaling: {world} x 4 species, 4 per rank"
    unit = "cell-updates/s"
    metric = "RAM phase-space cell-updates/s per full RAM step (8 drift sweeps + losses, 4 species)"

    if a.impl == "reference":
        if rank != 0:
            return
        # ~0.5 s per step on 4 cores at the default grid: bound the run to about a minute
        cap = 40 if a.workload == "default" else 4
        steps, wu = max(1, min(a.steps, cap)), max(1, min(a.warmup, 3 if a.workload == "default" else 1))
        v, nthreads, dt = cpu_reference(g, inp, steps, wu, groups=jobs)
        line = {"impl": "reference", "metric": metric, "value": v, "unit": unit, "n_gpus": a.gpus, "steps": steps,
                "warmup": wu, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": a.scaling,
                "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": desc},
                "cpu_baseline": {"value": v, "unit": unit, "cores": nthreads, "kind": "port",
                                 "sample": f"{steps} full ram_run steps of the same workload, OpenMP over species "
                                           f"({nthreads} threads, the reference's own decomposition"
                                           + (f"; {nthreads // max(1, min(g.nS, os.cpu_count() or 1))} concurrent 4-species sets" if jobs > 1 else "") + ")"},
                "e2e": {"value": v, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return

    import torch
    from ramscb_b200 import host, synthetic
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    # ---- sharding (ramscb_b200/parallel.py): species over ranks (no data-path collective);
    # beyond nS ranks, (L,K) slabs inside a species with two NCCL re-shardings per step
    from ramscb_b200 import parallel
    split = world > 1 and not weak                   # strong scaling: the 4 species are sharded over the ranks
    plan = parallel.make_plan(world if split else 1, rank if split else 0, g.nS, g.NPA, g.NE,
                              cells_per_species=g.NR * g.NT * g.NE * g.NPA)
    idle = plan.ns == 0     # more ranks than species on a grid too small to split a species
    gpu = host.RamGpu(g, device=local_rank, mode=host.MODE_FAST if a.mode == "fast" else host.MODE_EXACT)
    gpu.set_inputs(inp)
    if a.flags:
        D = synthetic.synthetic_daa(g, inp)          # synthetic Daa in ATAC / ATAW_emic_h
        gpu.set_diffcoef(1, D)
        gpu.set_diffcoef(2, D)
    F2_host = inp.F2.copy(order="F")
    host.host_register(F2_host)
    flush = torch.empty(512 * 1024 * 1024 // 4, dtype=torch.float32, device="cuda")
    sharded = None
    if split:
        # a non-default torch stream carries both the library's kernels and the NCCL traffic
        run_stream = torch.cuda.Stream()
        torch.cuda.set_stream(run_stream)
        gpu.set_stream(run_stream.cuda_stream)
        sharded = parallel.RamSharded(gpu, plan, dist)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def step_resident():
        return gpu.ram_run(DTS, DtsMin=1.0, flags=a.flags) if not split else sharded.ram_run(DTS)

    for _ in range(a.warmup):
        step_resident()
    launches0 = gpu.launch_count()
    clocks = ClockSampler(local_rank)
    clocks.start()
    barrier()
    dev_ms = 0.0
    t_wall0 = time.perf_counter()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(a.steps):
        flush.zero_()                      # evict F2 from the 126 MB L2 (untimed)
        torch.cuda.synchronize()
        if not split:
            gpu.timer_begin()
            step_resident()
            dev_ms += gpu.timer_end()
        else:                              # library work and NCCL are on / ordered with torch's current stream
            ev0.record()
            step_resident()
            ev1.record()
            torch.cuda.synchronize()
            dev_ms += ev0.elapsed_time(ev1)
        clocks.sample()
    barrier()
    wall_s = time.perf_counter() - t_wall0
    clk = clocks.stop()
    launches = gpu.launch_count() - launches0
    if dist is not None:
        t = torch.tensor([dev_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_ms = float(t.item())
    ms_per_step = dev_ms / a.steps
    value = ops_per_step * cells * jobs / (ms_per_step * 1e-3)

    # ---- end to end through the C ABI with host buffers ---------------------------
    e2e_steps = max(3, min(a.steps, 10))
    VT = inp.VT
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        if idle:
            continue
        gpu.f2_h2d(F2_host)
        gpu.set_efield(VT, inp.EIR, inp.EIP)
        out = step_resident()
        gpu.f2_d2h(F2_host)
    barrier()
    e2e_s = (time.perf_counter() - t0) / e2e_steps
    if dist is not None:
        t = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    h2d = F2_host.nbytes + 3 * VT.nbytes
    d2h = F2_host.nbytes + (4 + 6 + 1) * g.nS * 8 + 2 * g.nS * g.NR * g.NT * 8
    e2e = {"value": ops_per_step * cells * jobs / e2e_s, "unit": unit, "h2d_bytes_per_step": int(h2d) * jobs,
           "d2h_bytes_per_step": int(d2h) * jobs, "ms_per_step": e2e_s * 1e3,
           "timer": "wall clock, pinned host F2; F2 goes host->device and back EVERY step (routine-level drop-in, "
                    "INTEGRATION.md 3a): PCIe bound"}
    # for information: the fused integration keeps F2 resident; per step only the
    # E-field arrays go up and the step's results (DtsNext, DtDrift, losses, SETRC, PPERT, PPART) come back
    if not split:
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            gpu.set_efield(VT, inp.EIR, inp.EIP)
            out = step_resident()
        torch.cuda.synchronize()
        res_s = (time.perf_counter() - t0) / e2e_steps
        e2e["resident_state"] = {"ms_per_step": res_s * 1e3, "value": ops_per_step * cells / res_s, "per": "rank",
                                 "h2d_bytes_per_step": int(3 * VT.nbytes),
                                 "d2h_bytes_per_step": int((4 + 6 + 1) * g.nS * 8 + 2 * g.nS * g.NR * g.NT * 8),
                                 "note": "F2 stays on the device; not the headline e2e"}

    # ---- per-kernel device times over timed steps (CUDA events recorded on the run
    # stream between the stages of rsg_ram_run), dominant kernel roofline ------------
    peak, peak_src = measured_peak()
    roofline = None
    if not split:                                    # per rank; rank 0's is printed
        gpu.profile(True)
        for _ in range(a.steps):
            flush.zero_()
            torch.cuda.synchronize()
            step_resident()
        stages = gpu.profile_get()
        gpu.profile(False)
        per_kernel_ms = {k: v[0] / v[1] for k, v in stages.items() if v[1] and k != "end"}
        # cell-updates one launch performs (16 B = one FP64 read + one write per cell-update)
        ops_per_launch = {"k_driftr": 1, "k_driftp": 1, "k_drifte": 1, "k_driftmu": 1,
                          "k_plane_rp": 2,     # DRIFTR + DRIFTP of every plane
                          "k_col_fused": 8}    # DRIFTE, DRIFTMU, CHAREX, ATMOL, ATMOL, CHAREX, DRIFTMU, DRIFTE
        sweeps = {k: per_kernel_ms[k] for k in ops_per_launch if k in per_kernel_ms}
        dom = max(sweeps, key=lambda n: sweeps[n])
        alg_bytes = 16.0 * cells * ops_per_launch[dom]
        achieved = alg_bytes / (sweeps[dom] * 1e-3) / 1e9
        step_sum = sum(v[0] for k, v in stages.items() if k != "end") / a.steps
        traffic = None   # dram read+write bytes per launch from the committed `ncu --set full` capture
        try:
            with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "traffic.json")) as f:
                traffic = json.load(f).get(a.workload, {}).get(dom)
        except OSError:
            pass
        roofline = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                    "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                    "algorithmic_bytes_per_launch": int(alg_bytes),
                    "cell_updates_per_cell_per_launch": ops_per_launch[dom],
                    "limiter": ("instruction issue, not HBM: the fused kernels move 16 B per cell for "
                                f"{ops_per_launch[dom]} cell-updates (see traffic) and spend ~42-47 warp instructions per "
                                "cell-update (profiles/README.md)") if dom in ("k_col_fused", "k_plane_rp") else "see profiles/README.md",
                    "note": "16 B per cell-update (SURVEY 8(d)) x cell-updates of one launch (all cells of the 4 species x the "
                            "operators the kernel fuses); duration = CUDA events on the launching stream inside rsg_ram_run, "
                            "L2 flushed per step.  A fused kernel moves fewer bytes than its algorithmic figure.",
                    "per_kernel_ms": per_kernel_ms,
                    "per_kernel_frac_of_peak": {k: 16.0 * cells * ops_per_launch[k] / (v * 1e-3) / 1e9 / peak for k, v in sweeps.items()},
                    "kernel_share_of_step": {k: (v[0] / a.steps) / step_sum for k, v in stages.items() if k != "end"}}

    line = {"metric": metric, "value": value, "unit": unit, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": a.scaling, "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": desc, "ops_per_cell_per_step": ops_per_step, "mode": ("fast: separable coefficients + FMA + division-free limiter, fused shared-memory kernels, <=1e-12 of the oracle relative to the "
                                "stencil neighbourhood (tests/test_ram_parity_gpu.py)") if a.mode == "fast"
                       else "exact: reference operation order, bit-identical to the oracle",
                       "l2": "flushed between timed steps (512 MB memset, untimed)", "DTs": DTS,
                       "parallelism": (f"{world} ranks: species x slab groups of {plan.G}, NCCL re-sharding twice per step"
                                       if plan.G > 1 else
                                       (f"{world} ranks, {len(plan.active)} active (one species each, no data-path collective), "
                                        f"{world - len(plan.active)} idle: a species is split only above "
                                        f"{parallel.SPLIT_MIN_CELLS:.0e} cells (ramscb_b200/parallel.py)"
                                        if len(plan.active) < world else
                                        f"{world} ranks, species-sharded, no data-path collective"))
                       if split else ("1 GPU, all species per launch" if world == 1 else
                                      f"{world} ranks x 4 species (species are independent in ram_run, src/ModRamRun.f90:64): every rank runs "
                                      "the 1-GPU path on its own species set, no data-path collective; strong-scaling mode "
                                      "(species / slab sharding with NCCL re-sharding): --scaling strong, profiles/r1/scaling_r1.txt")},
            "clocks": clk, "e2e": e2e, "gpu_launches": int(launches) * jobs, "roofline": roofline,
            "wall_s_timed_region": wall_s}
    if rank == 0 and world == 1 and not a.no_scb:
        line["scb"] = scb_metrics(local_rank)
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
        try:        # CPU side of extras.computehI_integrals: the oracle's closed-form loop nest, one thread (the reference
            # runs adaptive cquad there, ~50-100x more integrand evaluations per integral), default-grid lines
            from ramscb_b200 import scb_synthetic
            hd = scb_synthetic.ram_field_lines(g.LZ[1:g.NR + 1] if len(g.LZ) > g.NR else g.LZ, g.MLT[:g.NT], nthe=101, wiggle=0.05)
            t0 = time.perf_counter()
            oracle.hi_integrals(mu=g.MU, **hd)
            line["scb"]["cpu_hI_integrals"] = {"wall_ms": (time.perf_counter() - t0) * 1e3, "lines": g.NR * g.NT, "cores": 1, "kind": "port"}
        except Exception as e:
            line["scb"]["cpu_hI_integrals"] = {"error": str(e)[:200]}
        try:        # CPU side of extras.computehI_integrals.convert_lines_default (OpenMP over the RAM points like the reference)
            rr = np.sqrt(sinp.x ** 2 + sinp.y ** 2 + sinp.z ** 2)
            t0 = time.perf_counter()
            oracle.hi_convert_lines(sinp.x, sinp.y, sinp.z, np.asfortranarray(30574.0 / rr ** 3), sinp.psi, sinp.alfa,
                                    g.LZ[:g.NR + 1] if len(g.LZ) > g.NR else np.append(2 * g.LZ[0] - g.LZ[1], g.LZ), g.MLT[:g.NT], 51)
            line["scb"]["cpu_hI_convert_lines"] = {"wall_ms": (time.perf_counter() - t0) * 1e3, "cores": os.cpu_count(), "kind": "port"}
        except Exception as e:
            line["scb"]["cpu_hI_convert_lines"] = {"error": str(e)[:200]}
    if rank == 0 and world == 1 and not a.no_cpu_baseline:      # reported at N = 1 only
        v, nthreads, dt = cpu_reference(g, inp, 3 if a.workload == "default" else 1, 1)
        line["cpu_baseline"] = {"value": v, "unit": unit, "cores": nthreads, "kind": "port",
                                "sample": "full ram_run steps of the same workload on the host cores "
                                          f"({dt:.3f} s/step, OpenMP over species = the reference's decomposition)"}
    if rank == 0 and world == 1 and not a.no_extras and a.workload == "default" and a.flags == 0:
        # informational, in a child process (this process's device state is released first): the WPI/EMIC step
        # (configs[2] on the 4x grid) and the zeta-sharded SOR protocol; never part of `value`
        try:
            gpu.close()
            del flush
            torch.cuda.empty_cache()
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--extras-only"], capture_output=True, text=True,
                               timeout=420, env=dict(os.environ, LOCAL_RANK=str(local_rank)))
            tag = [l for l in r.stdout.splitlines() if l.startswith("EXTRAS_JSON ")]
            line["extras"] = json.loads(tag[-1][len("EXTRAS_JSON "):]) if tag else {"error": (r.stderr or r.stdout)[-300:]}
        except Exception as e:
            line["extras"] = {"error": str(e)[:300]}
    if rank == 0:
        print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
