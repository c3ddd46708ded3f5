#!/usr/bin/env python3
"""Per-source-line summary of an ncu report's source page (needs -lineinfo at compile time):
   ncu -i X.ncu-rep --page source --csv --print-source cuda,sass > x.csv ; python tools/ncu_source_lines.py x.csv [N]"""
import csv
import sys


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    topn = int(sys.argv[2]) if len(sys.argv) > 2 else 40
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
            def g(k):
                try:
                    return int(r[ix[k]])
                except (ValueError, KeyError, IndexError):
                    return 0
            out.append(dict(file=cur, line=int(r[0]), src=r[1].strip(), samples=g("# Samples"), ins=g("Instructions Executed"),
                            thr=r[ix["Avg. Threads Executed"]], lsb=g("stall_long_sb"), wait=g("stall_wait"),
                            br=g("stall_branch_resolving"), ssb=g("stall_short_sb"), sel=g("stall_selected") + g("stall_not_selected"),
                            math=g("stall_math")))
    tot = sum(o["samples"] for o in out) or 1
    tins = sum(o["ins"] for o in out)
    print("samples %d  warp instructions %.2f G" % (tot, tins / 1e9))
    for o in sorted(out, key=lambda o: -o["samples"])[:topn]:
        print("%-14s:%4d %5.1f%% ins %5.2fG thr %5s | long_sb %4.1f wait %4.1f branch %4.1f short_sb %4.1f sel %4.1f math %4.1f | %s" % (
            o["file"], o["line"], 100.0 * o["samples"] / tot, o["ins"] / 1e9, o["thr"], 100.0 * o["lsb"] / tot, 100.0 * o["wait"] / tot,
            100.0 * o["br"] / tot, 100.0 * o["ssb"] / tot, 100.0 * o["sel"] / tot, 100.0 * o["math"] / tot, o["src"][:80]))


if __name__ == "__main__":
    main()
