#!/usr/bin/env python3
"""Aggregate an ncu launch list (--metrics gpu__time_duration.sum --csv) into per-kernel shares.

    python tools/launch_shares.py profiles/xxx_launches.csv
"""
import collections
import csv
import sys


def main(path):
    lines = [l for l in open(path) if l.startswith('"')]
    rows = list(csv.DictReader(lines))
    agg = collections.OrderedDict()
    for r in rows:
        k = r["Kernel Name"].split("(")[0][:70]
        v = float(r["Metric Value"].replace(",", "")) * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r["Metric Unit"], 1.0)
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    print("%d launches, %.3f ms of kernel time (cold-cache, serialised: compare shares)" % (len(rows), tot))
    for k, a in sorted(agg.items(), key=lambda x: -x[1][1]):
        print("%-72s n=%3d %10.3f ms %5.1f%%" % (k, a[0], a[1], 100 * a[1] / tot))


if __name__ == "__main__":
    main(sys.argv[1])
