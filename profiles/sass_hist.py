"""Opcode histogram (weighted by executed warp instructions) from `ncu --page source --csv`."""
import collections
import csv
import sys


def main(path, top=28):
    rows = list(csv.reader(open(path)))
    hdr = rows[1]
    ie, src, ws = hdr.index('Instructions Executed'), hdr.index('Source'), hdr.index('Warp Stall Sampling (All Samples)')
    data = [r for r in rows[2:] if len(r) > 10 and r[ie].isdigit()]
    tot = sum(int(r[ie]) for r in data)
    tots = sum(int(r[ws]) for r in data)
    print('total warp instr', tot, 'sass lines', len(data), 'stall samples', tots)
    h, hs = collections.Counter(), collections.Counter()
    for r in data:
        t = r[src].split()
        op = (t[1] if t[0].startswith('@') else t[0]).split('.')[0]
        h[op] += int(r[ie])
        hs[op] += int(r[ws])
    for op, c in h.most_common(top):
        print(f"{op:10s} {c:12d} {c / tot * 100:5.1f}%  stall-samples {hs[op] / max(tots, 1) * 100:5.1f}%")


if __name__ == "__main__":
    main(sys.argv[1])
