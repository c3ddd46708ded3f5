#!/usr/bin/env python3
"""Measure the BASELINE.json configs that are NOT bench.py's headline line (C1, C3, C4, C5).

    python tools/bench_configs.py c1|c3|c3n|c4|c5|c5p|c2mc [--scale S] [--steps K]
    python -m torch.distributed.run --nproc-per-node N ... tools/bench_configs.py c5      (N-way partitioned surface)

Prints one JSON line per config with the same vocabulary as bench.py (value / roofline / cpu_baseline).
`--scale` shrinks the sizes (scale 0.1 -> 10x fewer boxes and queries) for quick runs.
The CPU baseline legs use oracle/ (test infrastructure) on a bounded sample, like bench.py.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return json.load(open(p))["hbm_gbs"], "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def device_time_ms(fn, steps, warmup=2):
    import torch
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps, out


def find_config(name, args):
    import torch
    from axom_b200 import BVH, synth
    from oracle import oracle as O
    dev = torch.device("cuda", 0)
    hbm, src = peaks()
    if name == "c1":
        n = q = int(1_000_000 * args.scale)
        boxes = synth.triangle_aabbs(n, seed=12345)
        prim = synth.random_points(q, seed=12346)
        kind, P, label = "points", 24, "C1: spin::BVH<3> build over %d triangle AABBs + findPoints for %d random points" % (n, q)
    elif name == "c3":
        n = q = int(10_000_000 * args.scale)
        h = float(n) ** (-1.0 / 3.0)
        boxes = synth.triangle_aabbs(n, seed=12345)
        prim = synth.triangle_aabbs(q, seed=54321, shift=(h / 2, h / 2, h / 2))
        kind, P, label = "boxes", 48, "C3: findBoundingBoxes, two %d-triangle meshes (B shifted by h/2)" % n
    else:  # c4
        freq = max(2, int(round(1000 * args.scale ** 0.5)))
        x, y, z, conn = synth.icosphere(freq)
        boxes = synth.mesh_cell_boxes(x, y, z, conn)
        n = len(boxes)
        q = int(100_000_000 * args.scale)
        o, d = synth.random_rays(q, seed=777)
        prim = np.ascontiguousarray(np.concatenate([o, d], axis=1))
        kind, P, label = "rays", 48, "C4: findRays, %d rays vs %d-triangle icosphere (freq %d), chunks of <= 16M rays" % (q, n, freq)
    boxes_d = torch.from_numpy(boxes).to(dev)
    b = BVH(3)
    b.initialize(boxes_d)
    b.setProfiling(True)
    for _ in range(3):
        b.initialize(boxes_d)
    build_ms = b.phase_ms("build.total")
    build_ph = b.phases_ms("build.", ("bounds", "morton", "sort", "tree", "refit", "agglo"))
    fn = {"points": b.findPoints, "boxes": b.findBoundingBoxes, "rays": lambda r: b.findRays(r, normalized=False)}[kind]
    chunk = 16_000_000
    chunks = [torch.from_numpy(prim[i:i + chunk]).to(dev) for i in range(0, q, chunk)]

    def step():
        tot = 0
        for c in chunks:
            off, cnt, cand = fn(c)
            tot += cand.numel()
        return tot

    step()
    b.setProfiling(True)
    ms, total = device_time_ms(step, args.steps)
    ph = {k: round(b.phase_ms("find." + k), 4) for k in ("total", "sortq", "count", "scan", "fill")}
    alg_bytes = q * (P + 8) + 4 * total + 108 * n * len(chunks)
    # CPU baseline: reference SEQ_EXEC build (full size when small) + OpenMP count over a sample of the queries
    cpu = None
    if not args.no_cpu:
        kind_ref = "reference" if O.have_reference() else "port"
        t0 = time.perf_counter()
        rb = O.Bvh(boxes, ndims=3, kind=kind_ref)
        cpu_build_s = time.perf_counter() - t0
        ns = min(q, 1_000_000)
        t0 = time.perf_counter()
        if kind == "points":
            ctot, _ = rb.count_points_omp(prim[:ns], nthreads=0)
            cores = O.max_threads(kind_ref)
        elif kind == "boxes":
            ctot, _ = rb.count_boxes_omp(prim[:ns], nthreads=0)
            cores = O.max_threads(kind_ref)
        else:
            ctot, _ = rb.count_rays_omp(prim[:ns, :3], prim[:ns, 3:], nthreads=0)
            cores = O.max_threads(kind_ref)
        dt = time.perf_counter() - t0
        cpu = {"value": ns / dt, "unit": "queries/s", "cores": cores, "kind": kind_ref,
               "sample": "%d of the %d queries, the reference's traverse_tree under an external OpenMP loop (RAJA absent); SEQ_EXEC build of all %d boxes took %.2f s" % (ns, q, n, cpu_build_s),
               "build_ms": cpu_build_s * 1e3}
    line = {
        "metric": "find%s queries/s; BVH build ms beside it" % kind.capitalize(), "value": q / (ms * 1e-3), "unit": "queries/s",
        "n_gpus": 1, "steps": args.steps, "ms_per_step": ms, "higher_is_better": True, "dtype": "f64", "data": "synthetic",
        "config": {"workload": label, "boxes": n, "queries": q, "candidates": int(total), "candidates_per_query": total / q},
        "find_phases_ms_per_call": ph, "build_ms": build_ms, "build_phases_ms": build_ph,
        "roofline": {"bound": "hbm", "achieved": alg_bytes / (ms * 1e-3) / 1e9, "peak": hbm, "unit": "GB/s",
                     "frac": alg_bytes / (ms * 1e-3) / 1e9 / hbm, "algorithmic_bytes": alg_bytes, "peak_source": src, "traffic": None},
        "build_roofline": {"bound": "hbm", "achieved": 156.0 * n / (build_ms * 1e-3) / 1e9, "peak": hbm, "unit": "GB/s",
                           "frac": 156.0 * n / (build_ms * 1e-3) / 1e9 / hbm},
        "cpu_baseline": cpu,
    }
    print(json.dumps(line), flush=True)


def c3n(args):
    """SURVEY 8(f) rank 1, the step downstream of C3: quest::findTriMeshIntersectionsBVH on two interpenetrating
    icospheres (10 M triangles at scale 1): BVH build + ONE fused walk (findBoundingBoxes with the mesh's own AABBs,
    i < j, exact primal::intersect) + pair scatter.  CPU baseline: the reference's SEQ_EXEC implementation on a smaller
    mesh of the same shape (triangles/s)."""
    import torch
    from axom_b200 import MeshTester, synth
    from oracle import oracle as O
    hbm, src = peaks()
    dev = torch.device("cuda", 0)

    def two_spheres(freq):
        # coordinates scaled by the geodesic frequency so that triangle edges are O(1): the reference's fuzzy
        # comparators are absolute (1e-8 on unnormalised cross products), a unit sphere at this resolution
        # would report no intersections at all
        x, y, z, c = synth.icosphere(freq)
        x, y, z = x * freq, y * freq, z * freq
        X = np.concatenate([x, x * 0.9 + 0.3 * freq])
        Y = np.concatenate([y, y * 0.9])
        Z = np.concatenate([z, z * 0.9])
        return X, Y, Z, np.concatenate([c, c + len(x)]).astype(np.int32)

    freq = max(2, int(round(500 * args.scale ** 0.5)))
    X, Y, Z, C = two_spheres(freq)
    n = len(C)
    Xd, Yd, Zd, Cd = (torch.from_numpy(a).to(dev) for a in (X, Y, Z, C))
    t0 = time.perf_counter()
    mt = MeshTester(Xd, Yd, Zd, Cd)
    torch.cuda.synchronize()
    setup_wall_ms = (time.perf_counter() - t0) * 1e3
    b = mt.getBVH()

    def step():
        return mt.findTriMeshIntersections(1e-8)

    step()
    b.setProfiling(True)
    ms, pairs = device_time_ms(step, args.steps)
    ph = b.phases_ms("find.", ("total", "sortq", "count", "scan", "fill"))
    # end to end from host arrays: mesh upload + setup + BVH build + fused walk + pairs back on the host
    t0 = time.perf_counter()
    mt2 = MeshTester(X, Y, Z, C)
    hp = mt2.findTriMeshIntersections(1e-8)
    e2e_ms = (time.perf_counter() - t0) * 1e3
    assert np.array_equal(hp, pairs.cpu().numpy())
    # algorithmic bytes: per triangle its 72 B vertices + 48 B AABB, the node array once, 8 B per reported pair
    alg_bytes = n * (72 + 48) + 108 * n + 8 * int(pairs.shape[0])
    cpu = None
    if not args.no_cpu:
        kind_ref = "reference" if O.have_reference() else "port"
        cf = max(2, int(round(freq * 0.2)))
        x, y, z, c = two_spheres(cf)
        t0 = time.perf_counter()
        rp, _ = O.find_tri_mesh_intersections(x, y, z, c, 1e-8, kind_ref)
        dt = time.perf_counter() - t0
        gp = MeshTester(x, y, z, c).findTriMeshIntersections(1e-8)
        cpu = {"value": len(c) / dt, "unit": "triangles/s", "cores": 1, "kind": kind_ref,
               "sample": "same two-sphere mesh at icosphere frequency %d (%d triangles, %d intersecting pairs), SEQ_EXEC, %.2f s"
                         % (cf, len(c), len(rp), dt), "matches_gpu_bit_exact": bool(np.array_equal(rp, gp))}
    print(json.dumps({
        "metric": "findTriMeshIntersectionsBVH triangles/s (fused broad + narrow phase); setup + BVH build beside it",
        "value": n / (ms * 1e-3), "unit": "triangles/s", "n_gpus": 1, "steps": args.steps, "ms_per_step": ms, "higher_is_better": True,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": "two interpenetrating icospheres (freq %d), %d triangles" % (freq, n), "triangles": n,
                   "intersecting_pairs": int(pairs.shape[0])},
        "find_phases_ms_per_call": ph, "setup_wall_ms": setup_wall_ms,
        "e2e": {"value": n / (e2e_ms * 1e-3), "unit": "triangles/s", "ms": e2e_ms,
                "h2d_bytes_per_step": int(X.nbytes * 3 + C.nbytes), "d2h_bytes_per_step": int(hp.nbytes),
                "note": "host mesh -> upload, triangle/AABB setup, BVH build, fused walk, pairs back on the host (wall clock)"},
        "roofline": {"bound": "hbm", "achieved": alg_bytes / (ms * 1e-3) / 1e9, "peak": hbm, "unit": "GB/s",
                     "frac": alg_bytes / (ms * 1e-3) / 1e9 / hbm, "algorithmic_bytes": alg_bytes, "peak_source": src, "traffic": None},
        "cpu_baseline": cpu}), flush=True)


def c2mc(args):
    """SURVEY 8(f) rank 4, the consumer of C2: quest::MarchingCubes iso-contour of the 256^3 nodal distance field
    (255^3 = 16.6 M cells, a sphere of radius 0.5) with the field resident in HBM.  One pass = computeIsocontour on a cleared
    contour: mark (one read of the field) + count/scan + emit.  Metric: cells/s.  The dominant kernel is mc.mark, HBM-bound:
    algorithmic bytes = 8 B per node read once + 1 B per cell of case ids written.  CPU baseline: the REAL reference
    (quest::MarchingCubes, seq policy, through the Conduit mock) on the same field -- its OpenMP policy needs RAJA, so 1 core."""
    import torch
    from axom_b200 import MarchingCubes, synth
    from oracle import oracle as O
    hbm, src = peaks()
    n = max(8, int(round(255 * args.scale ** (1.0 / 3.0))))
    mesh = synth.blueprint_structured_mesh(cells=(n, n, n), center=(0.01, -0.02, 0.03))
    dmesh = synth.blueprint_to_device(mesh)
    cells, nodes = n ** 3, (n + 1) ** 3
    mc = MarchingCubes(device=0)
    mc.setMesh(dmesh, "mesh")
    mc.setFunctionField("dist")

    def step():
        mc.clearOutput()
        mc.computeIsocontour(0.5)
        return mc.getContourCellCount()

    step()
    # the field (134 MB) is larger than L2, so every pass reads it from HBM
    ms, facets = device_time_ms(step, args.steps, warmup=3)
    # per-kernel times in a second loop: the event pairs around every launch cost host time, so they stay out of `ms`
    mc.set_profiling(True)
    for _ in range(args.steps):
        step()
    ph = {k: mc.phase_ms("mc." + k) for k in ("mark", "count", "emit")}
    mc.set_profiling(False)
    # end to end through the public API with HOST arrays: upload of coordinates + field, contour, contour back on the host
    t0 = time.perf_counter()
    mh = MarchingCubes(device=0)
    mh.setMesh(mesh, "mesh")
    mh.setFunctionField("dist")
    mh.computeIsocontour(0.5)
    host = mh.relinquishContourData()
    e2e_ms = (time.perf_counter() - t0) * 1e3
    dev_out = [a.cpu().numpy() for a in mc.relinquishContourData(device_out=True)]
    same = all(np.array_equal(a, b) for a, b in zip(host, dev_out))
    alg_mark = 8 * nodes + cells
    traffic, traffic_source = None, None
    try:  # dram bytes of the mark kernel from the committed ncu capture (full-size workload only)
        prof = json.load(open(os.path.join(ROOT, "profiles", "mc_mark_kernel_ncu.json")))
        if n == 255:
            traffic, traffic_source = prof["dram_bytes_per_launch"], prof["source"]
    except Exception:
        pass
    cpu = None
    if not args.no_cpu:
        kind_ref = "reference" if O.have_reference() else "port"
        t0 = time.perf_counter()
        if kind_ref == "reference":
            ref = O.ref_mc_isocontour(mesh, "mesh", "dist", "", 1, (0.5,), 1)
        else:
            from axom_b200.marching_cubes import domain_views
            ref = O.mc_isocontour(domain_views(mesh, "mesh", "dist"), 0.5)
        dt = time.perf_counter() - t0
        cpu = {"value": cells / dt, "unit": "cells/s", "cores": 1, "kind": kind_ref,
               "sample": "the full %d^3-cell field, seq policy (hybridParallel), %.2f s" % (n, dt),
               "matches_gpu_bit_exact": bool(all(a.size == b.size and np.array_equal(a.reshape(-1), b.reshape(-1)) for a, b in zip(ref, dev_out)))}
    print(json.dumps({
        "metric": "MarchingCubes cells/s (iso-contour of the nodal distance field)", "value": cells / (ms * 1e-3), "unit": "cells/s",
        "n_gpus": 1, "steps": args.steps, "ms_per_step": ms, "higher_is_better": True, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "C2 consumer: %d^3 nodes, distance to a point, contour 0.5" % (n + 1), "cells": cells, "facets": int(facets),
                   "l2_policy": "field larger than L2 (%d MB)" % (8 * nodes // 2 ** 20)},
        "phases_ms_per_call": ph, "gpu_launches_per_step": 3, "mark_kernel": "plain" if os.environ.get("AXB_MC_MARK_PLAIN") == "1" else "rows",
        "e2e": {"value": cells / (e2e_ms * 1e-3), "unit": "cells/s", "ms": e2e_ms, "h2d_bytes_per_step": int(4 * 8 * nodes),
                "d2h_bytes_per_step": int(sum(a.nbytes for a in host)), "matches_device_path": bool(same),
                "note": "host mesh -> upload of x, y, z and the field, contour, contour arrays back on the host (wall clock, first call)"},
        "roofline": {"bound": "hbm", "kernel": "mc.mark", "achieved": alg_mark / (ph["mark"] * 1e-3) / 1e9, "peak": hbm, "unit": "GB/s",
                     "frac": alg_mark / (ph["mark"] * 1e-3) / 1e9 / hbm, "algorithmic_bytes_per_launch": alg_mark, "peak_source": src,
                     "kernel_ms": ph["mark"], "traffic": traffic, "traffic_source": traffic_source},
        "cpu_baseline": cpu}), flush=True)


def c5(args):
    """distributed closest point: surface split into G Morton ranges (one per rank; 8 sequential partitions
    when run on one GPU), unsigned distance per partition, elementwise MIN (NCCL all-reduce when G > 1)."""
    import torch
    import torch.distributed as dist
    from axom_b200 import SignedDistance, synth
    from axom_b200 import dist as D
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    freq = max(2, int(round(1000 * args.scale ** 0.5)))
    x, y, z, conn = synth.icosphere(freq)
    q = int(50_000_000 * args.scale)
    pts = synth.random_points(q, seed=999, lo=-1.0, hi=1.0)
    parts_total = world if world > 1 else 8
    P = np.stack([x, y, z], 1)
    cen = P[conn].mean(axis=1)
    parts = D.morton_partition(cen, parts_total)
    mine = [rank] if world > 1 else list(range(parts_total))
    qd = torch.from_numpy(pts).to(dev)
    sds = [SignedDistance(x, y, z, conn[parts[p]], 3, False, False, device=local) for p in mine]
    out = torch.empty(q, dtype=torch.float64, device=dev)
    tmp = torch.empty(q, dtype=torch.float64, device=dev)

    def step():
        for k, sd in enumerate(sds):
            sd.computeDistances(qd, out=(out if k == 0 else tmp))
            if k:
                torch.minimum(out, tmp, out=out)
        if world > 1:
            dist.all_reduce(out, op=dist.ReduceOp.MIN)
        return out

    step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    ok = None
    if rank == 0 and not args.no_check:
        full = SignedDistance(x, y, z, conn, 3, False, False, device=local)
        ref, _, _ = full.computeDistances(qd)
        ok = bool(torch.equal(ref, out))
    if rank == 0:
        print(json.dumps({
            "metric": "distributed closest point queries/s (partitioned surface, MIN-reduce)", "value": q / (ms * 1e-3),
            "unit": "queries/s", "n_gpus": world, "steps": args.steps, "ms_per_step": ms, "higher_is_better": True, "dtype": "f64",
            "data": "synthetic", "scaling": "strong",
            "config": {"workload": "C5: %d-triangle icosphere split into %d Morton ranges, %d queries on every rank" % (len(conn), parts_total, q),
                       "collective": "ncclAllReduce(MIN, f64) over %d x 8 B" % q if world > 1 else "none (partitions evaluated sequentially on one GPU)"},
            "matches_single_bvh_bit_exact": ok}), flush=True)
    if world > 1:
        dist.destroy_process_group()


def c5p(args):
    """quest::DistributedClosestPoint proper: the object is a POINT CLOUD (the vertices of the C4
    icosphere) split into one Morton range per rank; every rank owns 1/N of the 50 M query points.  One call =
    all-gather of the query blocks, one search kernel per block on every rank, MIN / ring-position / payload
    all-reduces per block (NCCL).  Checked against a single-handle search of the whole cloud (closest coordinates and
    distances must be bit-identical; rank / index are partition-relative)."""
    import torch
    import torch.distributed as dist
    from axom_b200 import DistributedClosestPoint, synth
    from axom_b200 import dist as D
    from oracle import oracle as O
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    freq = max(2, int(round(1000 * args.scale ** 0.5)))
    x, y, z, _ = synth.icosphere(freq)
    P = np.stack([x, y, z], 1)
    parts = D.morton_partition(P, world)
    q_total = int(50_000_000 * args.scale)
    lo, hi = D.slab_range(q_total, rank, world)
    pts = synth.random_points(q_total, seed=999, lo=-1.0, hi=1.0)
    myq = torch.from_numpy(pts[lo:hi]).to(dev)
    d = DistributedClosestPoint(3, device=local)
    d.setObjectMesh([P[parts[rank]]])
    d.generateBVHTree()

    def step():
        return d.computeClosestPoints(myq)

    step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        got = step()
    e1.record()
    torch.cuda.synchronize()
    ms = D.allreduce_max_scalar(e0.elapsed_time(e1) / args.steps, device=dev)
    ok = None
    if not args.no_check:
        full = DistributedClosestPoint(3, device=local)
        full.setObjectMesh([P])
        full.generateBVHTree()
        ns = min(hi - lo, 2_000_000)
        dist_was = dist.is_initialized()
        # single-handle answer for this rank's first ns queries (no collectives: call the backend directly)
        ref = full._b.compute_local(0, myq[:ns].contiguous())
        ok = bool(torch.equal(ref["cp_coords"], got["cp_coords"][:ns]) and torch.equal(ref["cp_distance"], got["cp_distance"][:ns]))
        if world > 1:
            t = torch.tensor([1.0 if ok else 0.0], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MIN)
            ok = bool(t.item() == 1.0)
    cpu = None
    if rank == 0 and not args.no_cpu:
        kind_ref = "reference" if O.have_reference() else "port"
        ns = min(q_total, 200_000)
        r = O.DistributedClosestPointRank(P, None, 3, kind_ref)
        t0 = time.perf_counter()
        r.compute_local(0, pts[:ns])
        dt = time.perf_counter() - t0
        cpu = {"value": ns / dt, "unit": "queries/s", "cores": 1, "kind": kind_ref,
               "sample": "%d of the %d queries against the whole %d-point cloud, one rank, %s BVH traversal, %.2f s" % (ns, q_total, len(P), kind_ref, dt)}
    if rank == 0:
        hbm, src = peaks()
        alg = q_total * (24 + 4 + 4 + 4 + 24 + 8) + 108 * len(P)
        print(json.dumps({
            "metric": "DistributedClosestPoint queries/s (point cloud partitioned over ranks, NCCL MIN-reduce with ring-order ties)",
            "value": q_total / (ms * 1e-3), "unit": "queries/s", "n_gpus": world, "steps": args.steps, "ms_per_step": ms,
            "higher_is_better": True, "dtype": "f64", "data": "synthetic", "scaling": "strong",
            "config": {"workload": "%d object points (icosphere freq %d vertices) in %d Morton ranges, %d queries in [-1,1]^3 split over ranks"
                                   % (len(P), freq, world, q_total),
                       "collective": ("all_gather(queries) + per block all_reduce MIN f64, MIN i64, SUM i64 x 7 (56 B/query)" if world > 1 else "none")},
            "roofline": {"bound": "hbm", "achieved": alg / (ms * 1e-3) / 1e9, "peak": hbm, "unit": "GB/s", "frac": alg / (ms * 1e-3) / 1e9 / hbm,
                         "algorithmic_bytes": alg, "peak_source": src, "traffic": None},
            "matches_single_handle_bit_exact": ok, "cpu_baseline": cpu}), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("config", choices=["c1", "c3", "c3n", "c4", "c5", "c5p", "c2mc"])
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-check", action="store_true")
    args = ap.parse_args()
    if args.config == "c5":
        c5(args)
    elif args.config == "c3n":
        c3n(args)
    elif args.config == "c5p":
        c5p(args)
    elif args.config == "c2mc":
        c2mc(args)
    else:
        find_config(args.config, args)


if __name__ == "__main__":
    main()
