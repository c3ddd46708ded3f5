#!/usr/bin/env python3
"""Warp instructions / active lanes / stall samples of an ncu source page summed over source-line ranges.
   python tools/ncu_regions.py x.csv file:first-last[:name] ..."""
import csv
import sys


def load(path):
    rows = list(csv.reader(open(path)))
    cur, ix, out = None, None, []
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            cur = r[1].split("/")[-1]
        elif r[0] == "Line No":
            ix = {}
            for i, h in enumerate(r):
                ix.setdefault(h, i)
        elif r[0].isdigit() and ix and len(r) > ix.get("# Samples", 1 << 30) and r[2] == "-":
            try:
                out.append((cur, int(r[0]), int(r[ix["Instructions Executed"]]), int(r[ix["Thread Instructions Executed"]]), int(r[ix["# Samples"]]),
                            int(r[ix["stall_long_sb"]])))
            except ValueError:
                pass
    return out


def main():
    out = load(sys.argv[1])
    tot = sum(o[4] for o in out) or 1
    tins = sum(o[2] for o in out)
    print("total: %.2f G warp instructions, %.1f lanes" % (tins / 1e9, sum(o[3] for o in out) / max(tins, 1)))
    for spec in sys.argv[2:]:
        parts = spec.split(":")
        f, (a, b) = parts[0], [int(x) for x in parts[1].split("-")]
        name = parts[2] if len(parts) > 2 else spec
        sel = [o for o in out if o[0] == f and a <= o[1] <= b]
        ins = sum(o[2] for o in sel)
        print("%-28s ins %6.2f G  lanes %5.1f  samples %5.1f %%  long_sb %5.1f %%" % (name, ins / 1e9, sum(o[3] for o in sel) / max(ins, 1),
                                                                               100.0 * sum(o[4] for o in sel) / tot, 100.0 * sum(o[5] for o in sel) / tot))


if __name__ == "__main__":
    main()
