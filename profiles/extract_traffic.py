"""dram read+write bytes per launch of each kernel from an `ncu --set full` raw CSV -> JSON
(bench.py reads profiles/traffic.json for `roofline.traffic`)."""
import csv
import json
import sys

UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def traffic(path):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    ir, iw, ik = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum"), hdr.index("Kernel Name")
    acc = {}
    for r in rows[2:]:
        k = r[ik].split("(")[0].replace("void ", "").split("<")[0]
        b = float(r[ir].replace(",", "")) * UNIT[units[ir]] + float(r[iw].replace(",", "")) * UNIT[units[iw]]
        acc.setdefault(k, []).append(b)
    return {k: sum(v) / len(v) for k, v in acc.items()}


if __name__ == "__main__":
    out = {}
    for arg in sys.argv[1:]:
        name, path = arg.split("=")
        out[name] = traffic(path)
    json.dump(out, sys.stdout, indent=1)
