"""Digest of `ncu -i X.ncu-rep --page raw --csv` files: the per-launch metrics the profiles/ summaries quote, and
profiles/traffic.json (dram bytes per launch of every RAM step kernel, keyed by workload, stamped with the sha1 of the
kernel sources bench.py checks before quoting it as `roofline.traffic`).

    python scripts/ncu_digest.py gpurun_out/r2d profiles/r2
"""
import csv
import hashlib
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "lts__t_sector_hit_rate.pct"]
STALL = re.compile(r"smsp__average_warps_issue_stalled_(\w+)_per_issue_active\.ratio")
SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}


def read(path):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    ix = {n: i for i, n in enumerate(hdr)}
    out = []
    for r in rows[2:]:
        if len(r) < len(hdr):
            continue
        name = r[ix["Kernel Name"]]
        m = {}
        for n, i in ix.items():
            if n in KEEP or STALL.match(n):
                try:
                    m[n] = (float(r[i].replace(",", "")), units[i])
                except ValueError:
                    pass
        out.append((name, r[ix["Grid Size"]] if "Grid Size" in ix else "", r[ix["Block Size"]] if "Block Size" in ix else "", m))
    return out


def short(name):
    return re.sub(r"^void ", "", name).split("(")[0]


def base(name):
    return re.sub(r"<.*", "", short(name))


def digest(rows):
    lines = []
    for name, grid, block, m in rows:
        lines.append(f"--- {short(name)}  grid {grid} block {block}")
        for k in KEEP:
            if k in m:
                lines.append(f"   {k:88s} {m[k][0]:16.6f} {m[k][1]}")
        st = sorted(((v[0], STALL.match(k).group(1)) for k, v in m.items() if STALL.match(k)), reverse=True)[:6]
        lines.append("   top stalls (warps per issue): " + ", ".join(f"{n} {v:.2f}" for v, n in st))
    return "\n".join(lines) + "\n"


def main(src, dst):
    os.makedirs(dst, exist_ok=True)
    traffic = {}
    for wl, key in (("x4", "x4_flags5"), ("default", "default_flags0")):
        p = os.path.join(src, f"full_{wl}_raw.csv")
        if not os.path.exists(p):
            continue
        rows = read(p)
        open(os.path.join(dst, f"ncu_full_{wl}_digest.txt"), "w").write(digest(rows))
        t = {}
        for name, _, _, m in rows:
            b = sum(m[k][0] * SCALE[m[k][1]] for k in ("dram__bytes_read.sum", "dram__bytes_write.sum") if k in m)
            t.setdefault(base(name), []).append(b)
        traffic[key] = {k: sum(v) / len(v) for k, v in t.items()}
    p = os.path.join(src, "full_computehI_raw.csv")
    if os.path.exists(p):
        open(os.path.join(dst, "ncu_full_computehI_digest.txt"), "w").write(digest(read(p)))
    if traffic:
        h = hashlib.sha1()
        for fn in ("ram_kernels.cuh", "ram_fused.cuh"):
            h.update(open(os.path.join(ROOT, "ramscb_b200", "csrc", fn), "rb").read())
        traffic["sources_sha1"] = h.hexdigest()
        traffic["how"] = ("ncu --set full --clock-control none, dram__bytes_read.sum + dram__bytes_write.sum per launch, mean over the "
                          "captured launches of each kernel (scripts/gpu_r2_call6.sh, scripts/ncu_digest.py); valid for the kernel "
                          "sources with this sha1 (bench.py checks it)")
        json.dump(traffic, open(os.path.join(ROOT, "profiles", "traffic.json"), "w"), indent=1)
    print(json.dumps(traffic, indent=1))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
