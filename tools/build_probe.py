"""Build-time probe: spin::BVH::initialize over device-resident triangle AABBs at several sizes, with the
fused bottom-up build (default) and the legacy tree_kernel + refit_kernel pair (AXB_BUILD_LEGACY=1), phase by
phase, plus the 156 B/box HBM roofline fraction.  python tools/build_probe.py [--sizes 1000000,20000000]"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sizes", default="1000000,2000000,10000000,20000000")
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--check", action="store_true", help="compare legacy vs fused outputs bit for bit")
    args = ap.parse_args()
    import numpy as np
    import torch
    from axom_b200 import BVH, synth
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm = float(peaks.get("hbm_gbs", 7700.0))
    for n in [int(s) for s in args.sizes.split(",")]:
        boxes = torch.from_numpy(synth.triangle_aabbs(n, seed=12345)).cuda()
        outs = {}
        for legacy in (1, 0):
            os.environ["AXB_BUILD_LEGACY"] = str(legacy)
            b = BVH(3)
            b.initialize(boxes)
            b.setProfiling(True)
            best = None
            for _ in range(args.reps):
                b.initialize(boxes)
                ph = b.phases_ms("build.", ("total", "bounds", "morton", "sort", "tree", "refit", "agglo"))
                if best is None or ph["total"] < best["total"]:
                    best = ph
            line = {"boxes": n, "legacy": bool(legacy), "build_ms": best, "achieved_gbs": round(156.0 * n / best["total"] / 1e6, 1),
                    "frac_of_hbm": round(156.0 * n / best["total"] / 1e6 / hbm, 4)}
            print(json.dumps(line), flush=True)
            if args.check:
                outs[legacy] = b.arrays()
            del b
        if args.check:
            for k in outs[0]:
                assert np.array_equal(outs[0][k], outs[1][k]), k
            print(json.dumps({"boxes": n, "legacy_vs_fused": "bit-identical"}), flush=True)
    os.environ.pop("AXB_BUILD_LEGACY", None)


if __name__ == "__main__":
    main()
