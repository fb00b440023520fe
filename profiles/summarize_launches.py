"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel."""
import collections
import csv
import sys


def main(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    tot = collections.defaultdict(lambda: [0, 0.0])
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        k = row["Kernel Name"].split("(")[0]
        v = float(row["Metric Value"].replace(",", ""))
        u = row["Metric Unit"]
        v = v / 1000 if u == "ns" else v * 1000 if u == "ms" else v
        tot[k][0] += 1
        tot[k][1] += v
    s = sum(v[1] for v in tot.values())
    print(f"{'kernel':28s} {'n':>4s} {'total_us':>10s} {'avg_us':>9s} {'share':>6s}")
    for k, v in sorted(tot.items(), key=lambda kv: -kv[1][1]):
        print(f"{k:28s} {v[0]:4d} {v[1]:10.1f} {v[1] / v[0]:9.2f} {v[1] / s * 100:5.1f}%")
    print(f"{'TOTAL':28s} {sum(v[0] for v in tot.values()):4d} {s:10.1f}")


if __name__ == "__main__":
    main(sys.argv[1])
