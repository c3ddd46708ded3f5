"""Bring-up / debugging driver (not a pytest file): runs small parity cases with verbose
diagnostics.  Usage on the GPU box:  python tests/gpu_bringup.py [n ...]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from axom_b200 import BVH, SignedDistance, synth  # noqa: E402
from oracle import oracle as O  # noqa: E402


def diff(name, a, b):
    a, b = np.asarray(a), np.asarray(b)
    if a.shape != b.shape:
        print("   MISMATCH", name, "shape", a.shape, b.shape)
        return False
    if not np.array_equal(a, b):
        bad = np.nonzero((a != b).reshape(len(a), -1).any(axis=1))[0]
        print("   MISMATCH", name, "count", len(bad), "first", bad[:5], "ref", a[bad[:3]], "gpu", b[bad[:3]])
        return False
    return True


def run(n, ndims=3):
    boxes = synth.triangle_aabbs(max(n, 1), seed=100 + n, ndims=ndims)[:n]
    ref = O.Bvh(boxes, ndims=ndims)
    g = BVH(ndims)
    g.setProfiling(True)
    t = time.time()
    g.initialize(boxes)
    dt = time.time() - t
    A, G = ref.arrays(), g.arrays()
    ok = all([diff(k, A[k], G[k]) for k in ("bounds", "mcodes", "leafs", "inner_children", "inner_nodes")])
    pts = synth.random_points(2000, seed=n, ndims=ndims)
    r, q = ref.find_points(pts), g.findPoints(pts)
    ok &= all([diff("pts." + k, u, v) for k, u, v in zip(("off", "cnt", "cand"), r, q)])
    qb = synth.triangle_aabbs(1500, seed=n + 99, ndims=ndims)
    r, q = ref.find_boxes(qb), g.findBoundingBoxes(qb)
    ok &= all([diff("box." + k, u, v) for k, u, v in zip(("off", "cnt", "cand"), r, q)])
    o, d = synth.random_rays(700, seed=n + 3, lo=-0.5, hi=1.5, ndims=ndims)
    r, q = ref.find_rays(o, d * 1.7, True), g.findRays(o, d * 1.7)
    ok &= all([diff("ray." + k, u, v) for k, u, v in zip(("off", "cnt", "cand"), r, q)])
    ph = {k: round(g.phase_ms("build." + k), 4) for k in ("total", "bounds", "morton", "sort", "tree", "refit")}
    print("bvh n=%d D=%d %s host %.1f ms phases(ms) %s" % (n, ndims, "OK" if ok else "FAIL", dt * 1e3, ph), flush=True)
    return ok


def run_sd(freq, grid):
    x, y, z, conn = synth.icosphere(freq)
    q = synth.uniform_grid_points(-1, 1, grid)
    ref = O.SignedDistance(x, y, z, conn)
    t = time.time()
    rphi, rcp, rn = ref.compute(q, True, True, nthreads=0)
    tc = time.time() - t
    g = SignedDistance(x, y, z, conn)
    ok = True
    for mode in (0, 1):
        g.setMode(mode)
        g.setProfiling(2)
        t = time.time()
        gphi, gcp, gn = g.computeDistances(q, True, True)
        tg = time.time() - t
        okm = diff("phi", rphi, gphi)
        cperr = np.abs(rcp - gcp).max()
        okm &= cperr <= 1e-12
        nerr = np.abs(rn - gn).max()
        leaf, inner = g.work_counters()
        print("sd mode %d tris=%d q=%d %s cp maxdiff %.1e normals maxdiff %.2e cpu(omp) %.2fs gpu host %.3fs kernel %.3f ms leaf/q %.1f inner/q %.1f" % (
            mode, len(conn), len(q), "OK" if okm else "FAIL", cperr, nerr, tc, tg, g.phase_ms("query.kernel"), leaf / len(q), inner / len(q)), flush=True)
        ok &= okm
    return ok


if __name__ == "__main__":
    sizes = [int(a) for a in sys.argv[1:] if not a.startswith('-')] or [0, 1, 2, 3, 27, 1000, 4097, 50000]
    ok = True
    for n in sizes:
        ok &= run(n, 3)
        if n <= 50000:
            ok &= run(n, 2)
    ok &= run_sd(8, 12)
    ok &= run_sd(40, 24)
    if "--big" in sys.argv:
        ok &= run_sd(316, 64)
    print("ALL OK" if ok else "SOME FAILED")
    sys.exit(0 if ok else 1)
