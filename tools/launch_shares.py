#!/usr/bin/env python
"""aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel name.
usage: python tools/launch_shares.py profiles/r01_launches_final.csv > profiles/r01_launch_shares.md"""
import collections
import csv
import math
import sys


def main(path):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
    h, data = rows[hi], rows[hi + 1:]
    kn, mv, mu = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in data:
        if len(r) <= mv:
            continue
        try:
            v = float(r[mv].replace(",", ""))
        except ValueError:
            continue
        if math.isnan(v):
            continue
        v *= {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3, "nsecond": 1e-3}.get(r[mu], 1.0)
        a = agg.setdefault(r[kn].split("(")[0][:48], [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    print(f"# ncu launch list `{path}` aggregated per kernel ({len(data)} launches; cold-cache, serialised replays)\n")
    print("| kernel | launches | total us | mean us | share |\n|---|---|---|---|---|")
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        if t / tot >= 0.002:
            print(f"| {k} | {n} | {t:.1f} | {t / n:.1f} | {t / tot:.3f} |")


if __name__ == "__main__":
    main(sys.argv[1])
