#!/usr/bin/env python3
"""GPU probe for the SignedDistance query kernel on the C2 workload: kernel ms and work counters,
optionally swept over AXB_SD_* tuning variables.   python tools/sd_probe.py [--grid 256] [--freq 316] [--sweep VAR=a,b,c]"""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--grid", type=int, default=256)
    ap.add_argument("--freq", type=int, default=316)
    ap.add_argument("--sweep", default="")
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--check", action="store_true", help="compare phi with mode 0 (reference-order kernel) on a 64^3 sub-lattice")
    args = ap.parse_args()
    import torch
    from axom_b200 import SignedDistance, synth
    x, y, z, conn = synth.icosphere(args.freq)
    dev = torch.device("cuda", 0)
    ax = torch.linspace(-1.0, 1.0, args.grid, dtype=torch.float64, device=dev)
    zz, yy, xx = torch.meshgrid(ax, ax, ax, indexing="ij")
    q = torch.stack([xx.reshape(-1), yy.reshape(-1), zz.reshape(-1)], dim=1).contiguous()
    phi = torch.empty(q.shape[0], dtype=torch.float64, device=dev)
    var, vals = None, [None]
    if args.sweep:
        var, v = args.sweep.split("=")
        vals = v.split(",")
    ref = None
    for val in vals:
        if var:
            os.environ[var] = val
        sd = SignedDistance(x, y, z, conn, 3, True, True, device=0)
        sd.computeDistances(q, out=phi)
        sd.setProfiling(1)
        for _ in range(args.reps):
            sd.computeDistances(q, out=phi)
        ms = sd.phase_ms("query.kernel")
        tot = sd.phase_ms("query.total")
        sd.setProfiling(2)
        sd.computeDistances(q, out=phi)
        lt, iv = sd.work_counters()
        sd.setProfiling(0)
        same = None
        if ref is None:
            ref = phi.clone()
        else:
            same = bool(torch.equal(ref, phi))
        out = {"var": var, "val": val, "kernel_ms": ms, "total_ms": tot, "leaf_tests_per_query": lt / q.shape[0],
               "inner_visits_per_query": iv / q.shape[0], "same_as_first": same}
        if args.check:
            sub = q.reshape(args.grid, args.grid, args.grid, 3)[::4, ::4, ::4].reshape(-1, 3).contiguous()
            a, _, _ = sd.computeDistances(sub)
            sd.setMode(0)
            b, _, _ = sd.computeDistances(sub)
            sd.setMode(1)
            out["matches_reference_order_kernel"] = bool(torch.equal(a, b))
        print(json.dumps(out), flush=True)
        del sd


if __name__ == "__main__":
    main()
