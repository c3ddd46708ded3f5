#!/usr/bin/env python3
"""bench.py -- headline benchmark of the B200-native spin::BVH / quest::SignedDistance path.

Workload (BASELINE.json configs[1], "C2"): quest::SignedDistance on a synthetic 2M-triangle
icosphere (geodesic frequency 316 -> 1 997 120 triangles, radius 0.5, watertight, signs on)
evaluated on the 256^3 uniform grid spanning [-1,1]^3.  With N GPUs the grid is sharded into
contiguous z-slabs (rank r takes planes [256 r / N, 256 (r+1) / N)); --sharding planes
deals the planes round-robin instead.  Measured on one GPU (tools/sd_shard_probe.py) a 1/8 shard costs
15.9-17.1 ms as a slab and 16.9-19.4 ms as every 8th plane, so slabs are the default.  One process per
GPU, surface BVH replicated and built per GPU, no data-path collective; total work is fixed, so scaling
is "strong".

A "step" is one computeDistances() pass over the rank's shard.
  value : whole-job points/s with queries and results resident in HBM (CUDA events, max over ranks)
  e2e   : the same through the C ABI with HOST buffers (pinned), H2D and D2H inside the timed region
  roofline     : the dominant kernel (the distance kernel) against the measured HBM peak
  cpu_baseline : the reference's CPU path (oracle/_ref if present, else the oracle port) on the
                 box's host cores, on a bounded sample of the same workload

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "SignedDistance query points/s (2M-triangle icosphere, 256^3 grid); BVH build ms reported beside it"
UNIT = "points/s"
FREQ = 316
GRID = 256
WORKLOAD = "C2: quest::SignedDistance, icosphere freq=%d (%d triangles), %d^3 grid on [-1,1]^3" % (FREQ, 20 * FREQ * FREQ, GRID)


def grid_axis():
    lo, hi, n = -1.0, 1.0, GRID
    return lo + np.arange(n, dtype=np.float64) * ((hi - lo) / (n - 1))


def sublattice(step):
    """query points of the sub-lattice i,j,k = 0 (mod step): bit-identical coordinates to the full grid"""
    ax = grid_axis()[::step]
    zz, yy, xx = np.meshgrid(ax, ax, ax, indexing="ij")
    return np.ascontiguousarray(np.stack([xx.ravel(), yy.ravel(), zz.ravel()], axis=1))


def ncu_profile_summary():
    """dram traffic of the distance kernel from the committed ncu capture (profiles/), if any"""
    p = os.path.join(ROOT, "profiles", "sd_min_kernel_ncu.json")
    try:
        return json.load(open(p))
    except Exception:
        return None


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return {"hbm_gbs": 6650.0}, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """samples nvidia-smi clocks / throttle reasons while the timed region runs"""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                continue
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_reference_rate(x, y, z, conn, step, repeats=1, warm=0):
    """reference CPU path on all host cores over the sub-lattice `step`; returns (points/s, info)"""
    from oracle import oracle as O
    kind = "reference" if O.have_reference() else "port"
    sd = O.SignedDistance(x, y, z, conn, 3, True, True, kind=kind)
    q = sublattice(step)
    # all host cores this process may use; explicit, because torchrun exports OMP_NUM_THREADS=1
    try:
        cores = len(os.sched_getaffinity(0))
    except AttributeError:
        cores = os.cpu_count() or 1
    cores = max(cores, O.max_threads(kind))
    for _ in range(warm):
        sd.compute(q, nthreads=cores)
    ts = []
    phi = None
    for _ in range(repeats):
        t = time.perf_counter()
        phi, _, _ = sd.compute(q, nthreads=cores)
        ts.append(time.perf_counter() - t)
    return len(q) / (sum(ts) / len(ts)), dict(kind=kind, cores=cores, npts=len(q), seconds=ts, phi=phi, q=q)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from axom_b200 import synth
    x, y, z, conn = synth.icosphere(FREQ)
    step = 4  # 64^3 sub-lattice per step: ~1-3 s of CPU work on 16 cores
    rate, info = cpu_reference_rate(x, y, z, conn, step, repeats=args.steps, warm=args.warmup)
    ms = 1e3 * sum(info["seconds"]) / len(info["seconds"])
    sample = "%d^3 sub-lattice (i,j,k = 0 mod %d) of the %d^3 grid per step, OpenMP over queries on %d host threads" % (
        GRID // step, step, GRID, info["cores"])
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "sample": sample},
        "cpu_baseline": {"value": rate, "unit": UNIT, "cores": info["cores"], "kind": info["kind"], "sample": sample},
        "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def marching_cubes_leg(phi_d, ax, device, check=True, steps=10):
    """quest::MarchingCubes::computeIsocontour(0.0) on the signed-distance field the timed region just produced (256^3 nodes,
    x fastest = Blueprint's default layout), everything resident in HBM: ms per contour, per-kernel times, and -- through
    Blueprint strides / offsets into the SAME arrays -- a 64^3-cell sub-domain checked bit for bit against the CPU oracle."""
    import torch
    from axom_b200 import MarchingCubes
    n = GRID - 1
    zz, yy, xx = torch.meshgrid(ax, ax, ax, indexing="ij")  # [k][j][i] contiguous = i fastest
    coords = {"x": xx.reshape(-1).contiguous(), "y": yy.reshape(-1).contiguous(), "z": zz.reshape(-1).contiguous()}
    del zz, yy, xx

    def domain(cells, offset):
        dims = {"i": cells, "j": cells, "k": cells}
        fld = {"association": "vertex", "topology": "mesh", "values": phi_d}
        if offset:
            lay = {"offsets": np.array([offset] * 3, np.int32), "strides": np.array([1, GRID, GRID * GRID], np.int32)}
            dims.update(lay)
            fld.update(lay)
        return {"domain_000000": {"coordsets": {"coords": {"type": "explicit", "values": coords}},
                                  "topologies": {"mesh": {"type": "structured", "coordset": "coords", "elements": {"dims": dims}}},
                                  "fields": {"phi": fld}}}

    mc = MarchingCubes(device=device)
    mc.setMesh(domain(n, 0), "mesh")
    mc.setFunctionField("phi")
    mc.computeIsocontour(0.0)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        mc.clearOutput()
        mc.computeIsocontour(0.0)  # synchronous: the facet count comes back to the host inside the call
    ms = (time.perf_counter() - t0) * 1e3 / steps
    facets = mc.getContourCellCount()
    mc.set_profiling(True)  # per-kernel event pairs cost host time: measured in a second loop, outside `ms`
    for _ in range(steps):
        mc.clearOutput()
        mc.computeIsocontour(0.0)
    phases = {k: mc.phase_ms("mc." + k) for k in ("mark", "count", "emit")}
    mc.set_profiling(False)
    info = {"ms_per_contour": ms, "cells": n ** 3, "facets": int(facets), "cells_per_s": n ** 3 / (ms * 1e-3), "contour_value": 0.0,
            "phases_ms": phases, "launches_per_contour": 3,
            "mark_kernel_hbm_frac": (8.0 * GRID ** 3 + n ** 3) / (phases["mark"] * 1e-3) / 1e9 / measured_peaks()[0]["hbm_gbs"]}
    if check:
        from axom_b200.marching_cubes import domain_views
        from oracle import oracle as O
        sub = domain(64, 32)  # nodes 32..96 per direction: a block that straddles the sphere
        ms_ = MarchingCubes(device=device)
        ms_.setMesh(sub, "mesh")
        ms_.setFunctionField("phi")
        ms_.computeIsocontour(0.0)
        got = ms_.relinquishContourData()
        host = {"domain_000000": dict(sub["domain_000000"])}
        host["domain_000000"]["coordsets"] = {"coords": {"type": "explicit", "values": {k: v.cpu().numpy() for k, v in coords.items()}}}
        host["domain_000000"]["fields"] = {"phi": dict(sub["domain_000000"]["fields"]["phi"], values=phi_d.cpu().numpy())}
        want = O.mc_isocontour(domain_views(host, "mesh", "phi"), 0.0)
        info["sub_domain_check"] = {"cells": 64 ** 3, "facets": int(want[2].size), "kind": "port",
                                    "matches_gpu_bit_exact": bool(want[2].size > 0 and all(np.array_equal(a.reshape(-1), b.reshape(-1))
                                                                                          for a, b in zip(want, got)))}
    return info


def marching_cubes_sharded_leg(phi_d, ax, p0, p1, rank, world, device, dist, steps=10):
    """N > 1: every rank contours ITS z-slab of phi (the slab the timed region just filled) as one Blueprint domain with
    domain id = rank.  The cells between two slabs need the first node plane of the next rank: one NCCL all-gather of a
    512 KB plane per rank over NVLink -- the only exchange; facets stay sharded.  Checked when the slabs are equal-sized:
    the slabs' facets, taken in rank order, ARE the single-domain contour of the whole field bit for bit (every rank
    recomputes that contour from the all-gathered field and compares its own range).
    Every collective below is reached by every rank unconditionally; only the local work sits in try blocks."""
    import torch
    from axom_b200 import MarchingCubes
    from axom_b200.marching_cubes import slab_domain
    dev = phi_d.device
    plane = GRID * GRID
    halo = torch.empty(world * plane, dtype=torch.float64, device=dev)
    dist.all_gather_into_tensor(halo, phi_d[:plane].contiguous())  # (1) halo exchange
    ok, err, facets_local, ms, got = 1, None, 0, 0.0, None
    nplanes = (p1 - p0) + (1 if rank < world - 1 else 0)
    try:
        dom, _ = slab_domain(phi_d, halo[(rank + 1) * plane:(rank + 2) * plane] if rank < world - 1 else None, (ax, ax, ax), p0, rank)
        mc = MarchingCubes(device=device)
        mc.setMesh(dom, "mesh")
        mc.setFunctionField("phi")
        mc.computeIsocontour(0.0)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(steps):
            mc.clearOutput()
            mc.computeIsocontour(0.0)
        ms = (time.perf_counter() - t0) * 1e3 / steps
        facets_local = mc.getContourCellCount()
        got = mc.relinquishContourData(device_out=True)
    except Exception as e:
        ok, err = 0, "%s: %s" % (type(e).__name__, e)
    t = torch.tensor([float(ok), float(facets_local), ms], dtype=torch.float64, device=dev)
    tsum, tmax = t.clone(), t.clone()
    dist.all_reduce(tsum, op=dist.ReduceOp.SUM)  # (2) totals
    dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    all_ok = int(round(float(tsum[0].item()))) == world
    match = None
    if GRID % world == 0:  # equal slabs: the whole field can be all-gathered into one tensor
        full = torch.empty(world * phi_d.numel(), dtype=torch.float64, device=dev)
        dist.all_gather_into_tensor(full, phi_d)  # (3) the whole field on every rank, for the check only
        good = 0
        try:
            if all_ok:
                zz, yy, xx = torch.meshgrid(ax, ax, ax, indexing="ij")
                fc = {"x": xx.reshape(-1).contiguous(), "y": yy.reshape(-1).contiguous(), "z": zz.reshape(-1).contiguous()}
                del zz, yy, xx
                fd = {"domain_000000": {"coordsets": {"coords": {"type": "explicit", "values": fc}},
                                        "topologies": {"mesh": {"type": "structured", "coordset": "coords",
                                                                "elements": {"dims": {"i": GRID - 1, "j": GRID - 1, "k": GRID - 1}}}},
                                        "fields": {"phi": {"association": "vertex", "topology": "mesh", "values": full}}}}
                mf = MarchingCubes(device=device)
                mf.setMesh(fd, "mesh")
                mf.setFunctionField("phi")
                mf.computeIsocontour(0.0)
                _, fxyz, fpar, _ = mf.relinquishContourData(device_out=True)
                cpp = (GRID - 1) ** 2  # cells per cell plane; parents are sorted, x fastest, z slowest
                edges = torch.tensor([p0 * cpp, (p0 + nplanes - 1) * cpp], dtype=torch.int32, device=dev)
                lo, hi = (int(v) for v in torch.searchsorted(fpar, edges).tolist())
                good = int(hi - lo == facets_local and torch.equal(fxyz[3 * lo:3 * hi], got[1])
                           and torch.equal(fpar[lo:hi] - p0 * cpp, got[2]) and bool((got[3] == rank).all()))
        except Exception as e:
            err = err or "%s: %s" % (type(e).__name__, e)
        g = torch.tensor([good], dtype=torch.int32, device=dev)
        dist.all_reduce(g, op=dist.ReduceOp.MIN)
        match = bool(int(g.item()) == 1)
    ms_max = float(tmax[2].item())
    info = {"sharding": "one Blueprint domain per rank = its z-slab + the next rank's first node plane (NCCL all-gather, %d B per rank)" % (plane * 8),
            "ms_per_contour_max_over_ranks": ms_max, "cells": (GRID - 1) ** 3, "facets_total": int(round(float(tsum[1].item()))),
            "cells_per_s": ((GRID - 1) ** 3 / (ms_max * 1e-3)) if ms_max > 0 else None, "contour_value": 0.0,
            "matches_single_domain_bit_exact": match, "all_ranks_ok": all_ok}
    if err:
        info["error"] = err
    return info



# ======================================================================================================================
# The other BASELINE.json configs, each with its parity bit against the UNMODIFIED reference (oracle/_ref):
#   C1 findPoints        1 M triangle AABBs, 1 M random points            (spin/BVH.hpp:480-506)
#   C3 findBoundingBoxes two 10 M-triangle meshes, B shifted by h/2       (spin/BVH.hpp:539-571)
#   C4 findRays          20 M-triangle icosphere x 100 M random rays      (spin/BVH.hpp:508-537), calls of <= 16 M rays
#   C5 distributed closest point: surface in G Morton ranges (one per rank), 50 M queries on every rank, elementwise MIN
# Query sets are generated on the device (torch generator, fixed seeds), sharded over the ranks as contiguous slices with
# the BVH replicated per GPU and no data-path collective; the first queries go to the host for the reference.
# ======================================================================================================================
def _aabbs_device(n, seed, shift, dev):
    """C1/C3 generator on the device: centres U[0,1)^3, three vertices = centre + h U(-1/2,1/2)^3, h = n^(-1/3)"""
    import torch
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    h = float(n) ** (-1.0 / 3.0)
    c = torch.rand((n, 1, 3), generator=g, dtype=torch.float64, device=dev)
    v = c + h * (torch.rand((n, 3, 3), generator=g, dtype=torch.float64, device=dev) - 0.5)
    v = v + torch.tensor(shift, dtype=torch.float64, device=dev)
    return torch.cat([v.amin(dim=1), v.amax(dim=1)], dim=1).contiguous()


def _points_device(n, seed, lo, hi, dev):
    import torch
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    return (lo + (hi - lo) * torch.rand((n, 3), generator=g, dtype=torch.float64, device=dev)).contiguous()


def _rays_device(n, seed, dev):
    """C4: origins U[-1,1]^3, directions uniform on the sphere (a normalised Gaussian triple; the reference normalises again)"""
    import torch
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    out = torch.empty((n, 6), dtype=torch.float64, device=dev)
    step = 10_000_000
    for i in range(0, n, step):
        m = min(step, n - i)
        out[i:i + m, :3] = -1.0 + 2.0 * torch.rand((m, 3), generator=g, dtype=torch.float64, device=dev)
        d = torch.randn((m, 3), generator=g, dtype=torch.float64, device=dev)
        out[i:i + m, 3:] = d / d.norm(dim=1, keepdim=True)
    return out


def _surface_boxes_device(x, y, z, conn, dev):
    import torch
    P = torch.from_numpy(np.stack([x, y, z], axis=1)).to(dev)
    c = torch.from_numpy(conn.astype(np.int64)).to(dev)
    out = torch.empty((len(conn), 6), dtype=torch.float64, device=dev)
    step = 4_000_000
    for i in range(0, len(conn), step):
        t = P[c[i:i + step]]
        out[i:i + step, :3] = t.amin(dim=1)
        out[i:i + step, 3:] = t.amax(dim=1)
    return out


def _host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def find_leg(name, ctx, surface=None):
    """One find* config: BVH build + the candidate query over this rank's slice of the query set.  Returns the leg's
    dict on rank 0 (None elsewhere).  Collectives (max over ranks) are reached by every rank unconditionally."""
    import torch
    from axom_b200 import BVH
    rank, world, dev, local, args = ctx["rank"], ctx["world"], ctx["dev"], ctx["local"], ctx["args"]
    s = args.config_scale
    err = None
    res = {}
    ms = 0.0
    total_local = 0
    q_local = 0
    try:
        if name == "C1":
            n = q = max(1000, int(1_000_000 * s))
            boxes_d = _aabbs_device(n, 12345, (0.0, 0.0, 0.0), dev)
            prim_d = _points_device(q, 12346, 0.0, 1.0, dev)
            kind, P = "points", 24
            label = "C1: spin::BVH<3> build over %d triangle AABBs + findPoints for %d random points" % (n, q)
        elif name == "C3":
            n = q = max(1000, int(10_000_000 * s))
            h = float(n) ** (-1.0 / 3.0)
            boxes_d = _aabbs_device(n, 12345, (0.0, 0.0, 0.0), dev)
            prim_d = _aabbs_device(q, 54321, (h / 2, h / 2, h / 2), dev)
            kind, P = "boxes", 48
            label = "C3: findBoundingBoxes, two %d-triangle meshes (B shifted by h/2)" % n
        else:
            x, y, z, conn = surface
            boxes_d = _surface_boxes_device(x, y, z, conn, dev)
            n = len(conn)
            q = max(1000, int(100_000_000 * s))
            prim_d = _rays_device(q, 777, dev)
            kind, P = "rays", 48
            label = "C4: findRays, %d random rays vs the %d-triangle icosphere, calls of <= 16 M rays" % (q, n)
        lo, hi = (q * rank) // world, (q * (rank + 1)) // world
        mine = prim_d[lo:hi]
        q_local = hi - lo
        b = BVH(3, device=local)
        b.setStream(torch.cuda.current_stream().cuda_stream)  # the CUDA events below are recorded on the launching stream
        b.initialize(boxes_d)
        b.setProfiling(True)
        for _ in range(3):
            b.initialize(boxes_d)
        build_ms = b.phase_ms("build.total")
        build_ph = b.phases_ms("build.", ("bounds", "morton", "sort", "tree", "refit", "agglo"))
        fn = {"points": b.findPoints, "boxes": b.findBoundingBoxes, "rays": lambda r: b.findRays(r, normalized=False)}[kind]
        chunk = 16_000_000
        chunks = [mine[i:i + chunk] for i in range(0, q_local, chunk)]

        def step():
            tot = 0
            for c in chunks:
                _, _, cand = fn(c)
                tot += cand.numel()
            return tot

        for _ in range(2):
            step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.config_steps):
            total_local = step()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.config_steps
        walk_ms = b.phase_ms("find.count")
        find_ph = {k: round(b.phase_ms("find." + k), 4) for k in ("total", "sortq", "count", "scan", "fill")}
        b.setProfiling(False)
        res = {"label": label, "n": n, "q": q, "kind": kind, "P": P, "build_ms": build_ms, "build_ph": build_ph, "find_ph": find_ph,
               "walk_ms": walk_ms, "calls": len(chunks)}
        # ---- end to end through the C ABI with pinned HOST buffers: one call over min(4 M, slice) queries ----
        ne = min(q_local, 4_000_000)
        host_q = torch.empty((ne, mine.shape[1]), dtype=torch.float64, pin_memory=True)
        host_q.copy_(mine[:ne])
        hq = host_q.numpy()
        fh = {"points": b.findPoints, "boxes": b.findBoundingBoxes, "rays": lambda r: b.findRays(r, normalized=False)}[kind]
        fh(hq)
        t0 = time.perf_counter()
        oh, ch, candh = fh(hq)
        e2e_s = time.perf_counter() - t0
        res["e2e"] = {"value": ne / e2e_s, "unit": "queries/s", "queries": ne, "ms": e2e_s * 1e3, "h2d_bytes_per_step": int(hq.nbytes),
                      "d2h_bytes_per_step": int(oh.nbytes + ch.nbytes + candh.nbytes)}
        # ---- parity against the unmodified reference (rank 0): counts on the first <= 1 M queries, lists on 100 k ----
        if rank == 0 and not args.no_cpu_baseline:
            from oracle import oracle as O
            kind_ref = "reference" if O.have_reference() else "port"
            boxes_h = boxes_d.cpu().numpy()
            t0 = time.perf_counter()
            rb = O.Bvh(boxes_h, ndims=3, kind=kind_ref)
            cpu_build_s = time.perf_counter() - t0
            ns = min(q_local, 1_000_000)
            sample = mine[:ns].cpu().numpy()
            cores = max(_host_cores(), O.max_threads(kind_ref))
            t0 = time.perf_counter()
            if kind == "points":
                _, rc = rb.count_points_omp(sample, nthreads=cores)
            elif kind == "boxes":
                _, rc = rb.count_boxes_omp(sample, nthreads=cores)
            else:
                _, rc = rb.count_rays_omp(sample[:, :3], sample[:, 3:], nthreads=cores)
            dt = time.perf_counter() - t0
            _, gc, _ = fn(mine[:ns])
            counts_ok = bool(np.array_equal(rc, gc.cpu().numpy()))
            nl = min(ns, 100_000)
            if kind == "points":
                ro, rcnt, rcand = rb.find_points(sample[:nl])
            elif kind == "boxes":
                ro, rcnt, rcand = rb.find_boxes(sample[:nl])
            else:
                ro, rcnt, rcand = rb.find_rays(sample[:nl, :3], sample[:nl, 3:], normalize=True)
            go, gcnt, gcand = fn(mine[:nl])
            lists_ok = bool(np.array_equal(ro, go.cpu().numpy()) and np.array_equal(rcnt, gcnt.cpu().numpy())
                            and np.array_equal(rcand, gcand.cpu().numpy()))
            ga = b.arrays() if n <= 2_000_000 else None
            res["cpu"] = {"value": ns / dt, "unit": "queries/s", "cores": cores, "kind": kind_ref,
                          "sample": "the first %d of the %d queries: the reference's traverse_tree under an external OpenMP loop (RAJA absent), "
                                    "%.2f s; SEQ_EXEC build of all %d boxes %.2f s" % (ns, q, dt, n, cpu_build_s),
                          "build_ms": cpu_build_s * 1e3, "counts_checked": ns, "counts_match": counts_ok,
                          "candidate_lists_checked": nl, "candidate_lists_match": lists_ok}
            if ga is not None:
                ra = rb.arrays()
                res["cpu"]["build_arrays_match"] = bool(all(np.array_equal(ra[k], ga[k]) for k in ("mcodes", "leafs", "inner_children", "inner_nodes", "bounds")))
            res["cpu"]["matches_reference_bit_exact"] = bool(counts_ok and lists_ok and res["cpu"].get("build_arrays_match", True))
            del rb
        del b, boxes_d, prim_d, mine, chunks
        torch.cuda.empty_cache()
    except Exception as e:  # never lose the headline to a secondary leg; the collectives below still run
        err = "%s: %s" % (type(e).__name__, e)
    ms_max = ctx["max_over_ranks"](ms)
    tot_all = ctx["sum_over_ranks"](float(total_local))
    bad = ctx["sum_over_ranks"](1.0 if err else 0.0)
    if rank != 0:
        return None
    if err or bad:
        return {"error": err or "a rank other than 0 failed"}
    peaks, peak_src = measured_peaks()
    n, q, P = res["n"], res["q"], res["P"]
    # per query sizeof(prim) + 8 B (offset, count) + 4 B per candidate, plus the node array once per call
    alg_local = q_local * (P + 8) + 4 * total_local + 108 * n * res["calls"]
    ach = alg_local / (ms * 1e-3) / 1e9 if ms > 0 else 0.0
    leg = {
        "workload": res["label"], "metric": "find%s queries/s" % res["kind"].capitalize(), "value": q / (ms_max * 1e-3), "unit": "queries/s",
        "ms_per_step": ms_max, "steps": args.config_steps, "boxes": n, "queries": q, "queries_per_gpu": q_local,
        "candidates": int(tot_all), "candidates_per_query": tot_all / q, "calls_per_step": res["calls"],
        "find_phases_ms_per_call": res["find_ph"], "build_ms": res["build_ms"], "build_phases_ms": res["build_ph"],
        "roofline": {"bound": "hbm", "achieved": ach, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": ach / peaks["hbm_gbs"],
                     "algorithmic_bytes_per_step": alg_local, "peak_source": peak_src, "traffic": None,
                     "note": "rank 0's slice; a latency-bound tree walk, not a stream (see DESIGN 3.2)"},
        "build_roofline": {"bound": "hbm", "achieved": 156.0 * n / (res["build_ms"] * 1e-3) / 1e9, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                           "frac": 156.0 * n / (res["build_ms"] * 1e-3) / 1e9 / peaks["hbm_gbs"], "algorithmic_bytes_per_box": 156},
        "e2e": res.get("e2e"), "cpu_baseline": res.get("cpu"),
        "matches_reference_bit_exact": (res.get("cpu") or {}).get("matches_reference_bit_exact"),
    }
    return leg


def c5_leg(ctx, surface):
    """BASELINE config 5: the surface split into G spatially coherent parts (Morton ranges of the triangle centroids), one
    per rank with its own BVH; every rank evaluates ALL queries unsigned against its part; one elementwise MIN over the
    ranks (NCCL all-reduce over NVLink) gives the distance to the whole surface -- bit-identical, a min of exact
    per-triangle values.  On one GPU the 8 parts are evaluated in turn into a running minimum that also bounds the
    search of the next part (axb_sd_update_min_distances; no collective).  Checked against the unmodified reference's SignedDistance over the WHOLE surface on a query sample."""
    import torch
    from axom_b200 import SignedDistance
    from axom_b200 import dist as D
    rank, world, dev, local, args = ctx["rank"], ctx["world"], ctx["dev"], ctx["local"], ctx["args"]
    x, y, z, conn = surface
    q = max(1000, int(50_000_000 * args.config_scale))
    parts_total = world if world > 1 else 8
    err, ms, coll_ms, out = None, 0.0, 0.0, None
    try:
        P = np.stack([x, y, z], 1)
        cen = (P[conn[:, 0]] + P[conn[:, 1]] + P[conn[:, 2]]) / 3.0
        parts = D.morton_partition(cen, parts_total)
        mine = [rank] if world > 1 else list(range(parts_total))
        qd = _points_device(q, 999, -1.0, 1.0, dev)
        sds = [SignedDistance(x, y, z, conn[parts[p]], 3, False, False, device=local) for p in mine]
        out = torch.empty(q, dtype=torch.float64, device=dev)
        tmp = torch.empty(q, dtype=torch.float64, device=dev) if len(sds) > 1 else None
        stream = torch.cuda.current_stream()
        for sd in sds:
            sd.setStream(stream.cuda_stream)
            sd.setAsync(True)
    except Exception as e:
        err = "%s: %s" % (type(e).__name__, e)
    comm = None
    if world > 1:
        comm = ctx["comm"]()  # the library's own NCCL communicator (axb_comm): the MIN runs inside libaxb200.so
    ok_all = ctx["sum_over_ranks"](0.0 if err else 1.0) == world
    if not ok_all:
        return {"error": err or "setup failed on another rank"} if rank == 0 else None

    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]

    def step():
        if world > 1:
            sds[0].computeDistancesMinReduce(comm, qd, out=out)  # kernel -> ncclAllReduce(MIN) in place, one stream, one C call
            return
        for k, sd in enumerate(sds):
            if k == 0:
                sd.computeDistances(qd, out=out)
            else:
                sd.updateMinDistances(qd, out)  # out = min(out, distance to part k); out bounds the search of part k

    step()
    ctx["barrier"]()
    ev[0].record()
    for _ in range(args.config_steps):
        step()
    ev[1].record()
    ctx["barrier"]()
    ms = ctx["max_over_ranks"](ev[0].elapsed_time(ev[1]) / args.config_steps)
    kern_min = kern_max = None
    if world > 1:
        sds[0].setProfiling(1)
        step()
        sds[0].synchronize()
        mine_coll, mine_kern = sds[0].phase_ms("query.minreduce"), sds[0].phase_ms("query.kernel")
        sds[0].setProfiling(0)
        # the all-reduce starts when this rank's kernel ends and ends when the SLOWEST rank's data has arrived: the shortest
        # one over the ranks is the transfer itself, the longest one includes waiting for the slowest kernel
        coll_ms = -ctx["max_over_ranks"](-mine_coll)
        coll_max = ctx["max_over_ranks"](mine_coll)
        kern_min, kern_max = -ctx["max_over_ranks"](-mine_kern), ctx["max_over_ranks"](mine_kern)
    if rank != 0:
        return None
    leg = {"workload": "C5: %d-triangle icosphere in %d Morton ranges (one BVH per %s), %d queries in [-1,1]^3 on every rank, unsigned distance, elementwise MIN"
                       % (len(conn), parts_total, "rank" if world > 1 else "part, evaluated in turn on one GPU", q),
           "metric": "distributed closest point queries/s", "value": q / (ms * 1e-3), "unit": "queries/s", "ms_per_step": ms,
           "steps": args.config_steps, "queries": q, "partitions": parts_total,
           "collective": ("ncclAllReduce(MIN, f64), %d B per rank, issued by axb_sd_compute_distances_minreduce on the kernel's stream" % (8 * q))
                         if world > 1 else "none (one GPU)"}
    if world > 1:
        bus = 2.0 * (world - 1) / world * 8.0 * q
        leg["nccl"] = {"allreduce_ms": coll_ms, "allreduce_ms_incl_wait_for_slowest_rank": coll_max, "bus_bytes": bus,
                       "bus_gbs": bus / (coll_ms * 1e-3) / 1e9 if coll_ms > 0 else None, "nvlink5_peak_gbs_per_direction": 900.0,
                       "library": comm.library(), "kernel_ms_per_rank": {"min": kern_min, "max": kern_max}}
    if not args.no_cpu_baseline:
        try:
            from oracle import oracle as O
            kind_ref = "reference" if O.have_reference() else "port"
            cores = max(_host_cores(), O.max_threads(kind_ref))
            t0 = time.perf_counter()
            ref = O.SignedDistance(x, y, z, conn, 3, False, False, kind=kind_ref)
            setup_s = time.perf_counter() - t0
            # bounded sample: grow in blocks until ~20 s of CPU work or 1 M queries
            got = out.cpu().numpy()
            done, block, t_q, match = 0, 50_000, 0.0, True
            while done < min(q, 1_000_000) and t_q < 20.0:
                m = min(block, q - done)
                qs = qd[done:done + m].cpu().numpy()
                t0 = time.perf_counter()
                phi, _, _ = ref.compute(qs, nthreads=cores)
                t_q += time.perf_counter() - t0
                match = match and bool(np.array_equal(phi, got[done:done + m]))
                done += m
                block = min(2 * block, 250_000)
            leg["cpu_baseline"] = {"value": done / t_q, "unit": "queries/s", "cores": cores, "kind": kind_ref,
                                   "sample": "the first %d of the %d queries against the WHOLE %d-triangle surface (quest::SignedDistance, computeSign off), "
                                             "OpenMP over queries, %.1f s (+ %.1f s setMesh)" % (done, q, len(conn), t_q, setup_s),
                                   "matches_gpu_bit_exact": match}
            leg["matches_reference_bit_exact"] = match
        except Exception as e:
            leg["cpu_baseline"] = {"error": "%s: %s" % (type(e).__name__, e)}
    return leg


def dcp_leg(ctx, surface):
    """quest::DistributedClosestPoint proper: the object is a POINT CLOUD (the vertices of the C4
    icosphere) in one Morton range per rank; every rank owns 1/N of the 50 M query points; one call =
    axb_dcp_compute_closest_points (gather, two searches, MIN all-reduces of 8 + 8 + 1 B/query, one all-to-all of 48-byte
    winner records -- all inside the library over its own NCCL communicator).  Checked against the unmodified reference's
    traversal of the WHOLE cloud on a sample of rank 0's queries (coordinates and distances; rank / index are
    partition-relative and are checked against the partition)."""
    import torch
    from axom_b200 import DistributedClosestPoint
    from axom_b200 import dist as D
    rank, world, dev, local, args = ctx["rank"], ctx["world"], ctx["dev"], ctx["local"], ctx["args"]
    x, y, z, _ = surface
    q_total = max(1000, int(50_000_000 * args.config_scale))
    err, d, myq, got = None, None, None, None
    try:
        P = np.stack([x, y, z], 1)
        parts = D.morton_partition(P, world)
        lo, hi = D.slab_range(q_total, rank, world)
        myq = _points_device(q_total, 999, -1.0, 1.0, dev)[lo:hi].contiguous()
        d = DistributedClosestPoint(3, device=local)
        d.setObjectMesh([P[parts[rank]]])
        d.generateBVHTree()
    except Exception as e:
        err = "%s: %s" % (type(e).__name__, e)
    comm = ctx["comm"]() if world > 1 else None
    ok_all = ctx["sum_over_ranks"](0.0 if err else 1.0) == world
    if not ok_all:
        return {"error": err or "setup failed on another rank"} if rank == 0 else None
    d.setComm(comm)
    got = d.computeClosestPoints(myq)
    ctx["barrier"]()
    b0 = comm.traffic()[0] if comm else 0
    t0 = time.perf_counter()
    for _ in range(args.config_steps):
        got = d.computeClosestPoints(myq)  # synchronous: results complete on return
    torch.cuda.synchronize()
    ms = ctx["max_over_ranks"]((time.perf_counter() - t0) * 1e3 / args.config_steps)
    sent = (comm.traffic()[0] - b0) / args.config_steps if comm else 0
    phases = None
    if world > 1:
        d._b.set_profiling(True)
        d.computeClosestPoints(myq)
        names = ("counts", "gather", "search1", "bound_allreduce", "search2", "combine", "exchange", "total")
        mine = [d._b.phase_ms("dcpx." + n) for n in names]
        d._b.set_profiling(False)
        phases = {n: ctx["max_over_ranks"](v) for n, v in zip(names, mine)}
    sent_total = ctx["sum_over_ranks"](float(sent))
    if rank != 0:
        return None
    leg = {"workload": "DistributedClosestPoint: %d object points (icosphere vertices) in %d Morton ranges, %d queries in [-1,1]^3 split over the ranks"
                       % (len(P), world, q_total),
           "metric": "DistributedClosestPoint queries/s", "value": q_total / (ms * 1e-3), "unit": "queries/s", "ms_per_step": ms,
           "steps": args.config_steps, "queries": q_total,
           "timing": "host wall clock around the synchronous C call, max over ranks (the call contains four host round trips)",
           "collective": ("axb_dcp_compute_closest_points over axb_comm: broadcast-gather 24 B/query, MIN f64 x2 + MIN u8 (17 B/query), "
                          "grouped send/recv of 48-byte winner records") if world > 1 else "none (one GPU)"}
    if world > 1:
        leg["nccl"] = {"payload_bytes_per_step_all_ranks": sent_total, "phases_ms_max_over_ranks": phases, "library": comm.library(),
                       "payload_gbs": sent_total / (ms * 1e-3) / 1e9}
    if not args.no_cpu_baseline:
        try:
            from oracle import oracle as O
            kind_ref = "reference" if O.have_reference() else "port"
            ns = min(hi - lo, 200_000)
            r = O.DistributedClosestPointRank(P, None, 3, kind_ref)
            qs = myq[:ns].cpu().numpy()
            t0 = time.perf_counter()
            want = r.compute_local(0, qs)
            dt = time.perf_counter() - t0
            g = {k: v[:ns].cpu().numpy() for k, v in got.items()}
            match = bool(np.array_equal(g["cp_coords"], want["cp_coords"]) and np.array_equal(g["cp_distance"], want["cp_distance"]))
            # rank / index against the partition: the winner's point is where it says it is
            back = np.full((ns, 3), np.nan)
            for rk in range(world):
                m = g["cp_rank"] == rk
                back[m] = P[parts[rk][g["cp_index"][m]]]
            match = match and bool(np.array_equal(back, g["cp_coords"]))
            leg["cpu_baseline"] = {"value": ns / dt, "unit": "queries/s", "cores": 1, "kind": kind_ref,
                                   "sample": "the first %d of rank 0's queries against the whole %d-point cloud, %s BVH traversal, one thread, %.1f s"
                                             % (ns, len(P), kind_ref, dt), "matches_gpu_bit_exact": match}
            leg["matches_reference_bit_exact"] = match
        except Exception as e:
            leg["cpu_baseline"] = {"error": "%s: %s" % (type(e).__name__, e)}
    return leg


def config_legs(ctx):
    """C1, C3, C4, C5 -> {"C1": {...}, ...} on rank 0"""
    from axom_b200 import synth
    args = ctx["args"]
    want = [c.strip().upper() for c in args.configs.split(",") if c.strip()]
    out = {}
    for name in ("C1", "C3"):
        if name in want:
            out[name] = find_leg(name, ctx)
    if "C4" in want or "C5" in want or "DCP" in want:
        freq = max(2, int(round(1000 * args.config_scale ** 0.5)))
        surface = synth.icosphere(freq)
        if "C4" in want:
            out["C4"] = find_leg("C4", ctx, surface)
        if "C5" in want:
            out["C5"] = c5_leg(ctx, surface)
        if "DCP" in want:
            out["DCP"] = dcp_leg(ctx, surface)
    return out if ctx["rank"] == 0 else None


def run_ours(args):
    import torch
    import torch.distributed as dist
    from axom_b200 import SignedDistance, synth

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(v):
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(v):
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    def min_over_ranks(v):
        return -max_over_ranks(-v)

    # ---- surface + BVH (replicated per GPU) ----
    x, y, z, conn = synth.icosphere(FREQ)
    ntri = len(conn)
    t0 = time.perf_counter()
    sd = SignedDistance(x, y, z, conn, 3, True, True, device=local)
    first_setmesh_wall_ms = (time.perf_counter() - t0) * 1e3  # includes CUDA context creation and module load
    del sd
    t0 = time.perf_counter()
    sd = SignedDistance(x, y, z, conn, 3, True, True, device=local)  # host mesh -> upload, cell boxes, BVH, leaf + OBB records
    setmesh_wall_ms = (time.perf_counter() - t0) * 1e3
    setmesh_device_ms = {}
    for ph in ("setmesh.total", "setmesh.upload", "setmesh.cell_boxes", "setmesh_build.total", "setmesh.gather_soup", "setmesh.obb_build"):
        try:
            setmesh_device_ms[ph.replace("setmesh.", "").replace("setmesh_build.total", "bvh_build")] = sd.phase_ms(ph)
        except Exception:
            pass
    bvh = sd.getBVHTree()
    # BVH build time (device): rebuild from the device-resident cell boxes a few times
    boxes_d = torch.from_numpy(synth.mesh_cell_boxes(x, y, z, conn)).to(dev)
    from axom_b200 import BVH
    tb = BVH(3, device=local)
    tb.initialize(boxes_d)
    tb.setProfiling(True)
    for _ in range(5):
        tb.initialize(boxes_d)
    build_ms = tb.phase_ms("build.total")
    build_phases = tb.phases_ms("build.", ("bounds", "morton", "sort", "tree", "refit", "agglo"))
    del tb, boxes_d

    # ---- this rank's z-slab of the 256^3 grid, generated on the device ----
    ax = torch.from_numpy(grid_axis()).to(dev)
    if args.sharding == "planes":
        planes = torch.arange(rank, GRID, world, device=dev)  # z-planes dealt round-robin
    elif args.sharding == "blocks" and world > 1:
        # blocks of 4 z-planes dealt round-robin: every rank gets the same mix of far-field, near-surface and
        # near-centre queries (a contiguous slab of the outer region costs 1.5x a slab through the centre: 147 against
        # 100 node visits per query), and 4 planes keep a query's Morton neighbours on the same rank
        planes = torch.cat([torch.arange(b, min(b + 4, GRID), device=dev) for b in range(4 * rank, GRID, 4 * world)])
    else:
        planes = torch.arange((GRID * rank) // world, (GRID * (rank + 1)) // world, device=dev)  # contiguous z-slab
    zz, yy, xx = torch.meshgrid(ax[planes], ax, ax, indexing="ij")
    q_d = torch.stack([xx.reshape(-1), yy.reshape(-1), zz.reshape(-1)], dim=1).contiguous()
    del zz, yy, xx
    nq_local = q_d.shape[0]
    nq_total = GRID ** 3
    phi_d = torch.empty(nq_local, dtype=torch.float64, device=dev)

    stream = torch.cuda.current_stream()
    sd.setStream(stream.cuda_stream)

    # ---- value: device-resident, K steps back to back, CUDA events on the launching stream ----
    sd.setAsync(True)
    for _ in range(args.warmup):
        sd.computeDistances(q_d, out=phi_d)
    sd.setProfiling(1)
    launches0 = sd.launch_count()
    sampler = ClockSampler(local)
    barrier()
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        sd.computeDistances(q_d, out=phi_d)
    e1.record(stream)
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    ms_step = max_over_ranks(e0.elapsed_time(e1) / args.steps)
    launches = sd.launch_count() - launches0
    kernel_ms = sd.phase_ms("query.kernel")  # mean over the K timed steps (events around the four launches of one call)
    min_ms = sd.phase_ms("query.min")        # the dominant kernel: sd_min_kernel (sample pass + search proper)
    resolve_ms = sd.phase_ms("query.resolve")  # sd_resolve_kernel + sd_solo_kernel
    sd.setProfiling(0)
    sd.setAsync(False)
    value = nq_total / (ms_step * 1e-3)

    # ---- e2e: host buffers through the C ABI (H2D of the queries + D2H of phi inside the timed region) ----
    q_h = torch.empty((nq_local, 3), dtype=torch.float64, pin_memory=True)
    q_h.copy_(q_d)
    phi_h = torch.empty(nq_local, dtype=torch.float64, pin_memory=True)
    qn, pn = q_h.numpy(), phi_h.numpy()
    e2e_steps = max(1, min(args.steps, 5))
    sd.computeDistances(qn, out=pn)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        sd.computeDistances(qn, out=pn)
    torch.cuda.synchronize()
    e2e_ms = max_over_ranks((time.perf_counter() - t0) * 1e3 / e2e_steps)
    e2e_value = nq_total / (e2e_ms * 1e-3)
    same = bool(np.array_equal(pn, phi_d.cpu().numpy()))

    # ---- work counters (profiling level 2; outside every timed region) ----
    sd.setProfiling(2)
    sd.computeDistances(q_d, out=phi_d)
    leaf_tests, inner_visits = sd.work_counters()
    sd.setProfiling(0)

    # ---- N > 1: the consumer of the field, sharded like the queries (every rank takes part: collectives inside) ----
    mc_sharded = None
    if world > 1 and not args.no_marching_cubes:
        # the collectives inside are reached by every rank unconditionally; local failures are reported, not raised.
        # The contour works on a contiguous z-slab of the field (+ a halo plane): when the distance queries were dealt
        # in interleaved blocks, the slab's field is computed here, outside every timed region.
        p0, p1 = (GRID * rank) // world, (GRID * (rank + 1)) // world
        phi_slab = phi_d
        if args.sharding != "slabs":
            zz, yy, xx = torch.meshgrid(ax[p0:p1], ax, ax, indexing="ij")
            q_slab = torch.stack([xx.reshape(-1), yy.reshape(-1), zz.reshape(-1)], dim=1).contiguous()
            del zz, yy, xx
            phi_slab = torch.empty(q_slab.shape[0], dtype=torch.float64, device=dev)
            sd.computeDistances(q_slab, out=phi_slab)
            del q_slab
        mc_sharded = marching_cubes_sharded_leg(phi_slab, ax, p0, p1, rank, world, local, dist)
        del phi_slab

    # ---- per-rank imbalance of the distance kernel (N > 1: the tail of the slowest rank sets the step) ----
    kernel_ms_min, kernel_ms_max = min_over_ranks(kernel_ms), max_over_ranks(kernel_ms)

    # ---- the other BASELINE configs (C1, C3, C4, C5), each checked against the reference; every rank takes part ----
    cfg = None
    if args.configs:
        host_phi = phi_d.cpu().numpy() if (world == 1 and not args.no_cpu_baseline) else None
        _comm = []

        def lib_comm():
            # the library's own NCCL communicator, made once (collective: every rank calls it at the same point)
            if not _comm:
                from axom_b200.comm import Comm
                _comm.append(Comm.from_torch(local))
            return _comm[0]

        ctx = {"rank": rank, "world": world, "dev": dev, "local": local, "args": args, "dist": dist, "barrier": barrier,
               "max_over_ranks": max_over_ranks, "sum_over_ranks": sum_over_ranks, "comm": lib_comm}
        cfg = config_legs(ctx)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks, peak_src = measured_peaks()
    # algorithmic bytes of one distance-kernel launch: 24 B in + 8 B out per query,
    # plus the node array and the leaf geometry once
    alg_bytes = nq_local * 32 + 128 * (ntri - 1) + 72 * ntri
    achieved = alg_bytes / (min_ms * 1e-3) / 1e9
    prof = ncu_profile_summary()
    pk = (prof or {}).get("kernels", {}).get("sd_min_kernel", {})
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": achieved / peaks["hbm_gbs"],
                "traffic": None, "kernel": "sd_min_kernel (exact-minimum search: sample pass + search proper, 2 launches)", "kernel_ms": min_ms,
                "all_kernels_ms": kernel_ms, "resolve_and_heavy_ms": resolve_ms,
                "algorithmic_bytes_per_launch": alg_bytes, "peak_source": peak_src,
                "note": "not HBM-bound: a tree search, every lane reads its own 64-byte node record per step; the limiter is instruction "
                        "issue (68 % of cycles) and step latency -- see l1; the HBM fraction is reported because the contract asks for it"}
    if prof:
        roofline["traffic"] = prof.get("dram_bytes_per_launch")
        roofline["traffic_source"] = prof.get("source")
    flops = 80.0 * leaf_tests + 50.0 * inner_visits  # SURVEY.md 8(d) convention
    fp64 = {"leaf_tests_per_query": leaf_tests / nq_local, "inner_visits_per_query": inner_visits / nq_local,
            "gflops_survey_convention": flops / (kernel_ms * 1e-3) / 1e9,
            "frac_of_nominal_fp64_peak": flops / (kernel_ms * 1e-3) / 37.2e12,
            "nominal_fp64_peak": "37.2 TFLOP/s = 148 SMs x 64 DFMA/clk x 1.965 GHz (not in MEASURED_PEAKS.json)",
            "note": "phase 1 evaluates the oriented bounds in binary32 (conservatively) and only the normal axis and the leaves in binary64"}
    # the L1 side: every lane reads its own 64-byte node record (2 sectors = 2 wavefronts per visit) and 96-byte leaf
    # record (3 per test); nothing coalesces across lanes.  1 wavefront per clock per SM at best.
    sectors = 2.0 * inner_visits + 3.0 * leaf_tests
    l1 = {"wavefronts_per_launch_modelled": sectors, "achieved_gwavefronts_per_s": sectors / (min_ms * 1e-3) / 1e9,
          "peak_gwavefronts_per_s": 148 * 1.965, "frac_modelled": sectors / (min_ms * 1e-3) / 1e9 / (148 * 1.965),
          "frac_ncu": (pk.get("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed") or 0.0) / 100.0 or None,
          "lanes_active_ncu": pk.get("smsp__thread_inst_executed_per_inst_executed.ratio"),
          "issue_slots_ncu": pk.get("smsp__issue_active.avg.pct_of_peak_sustained_active"),
          "note": "ncu (profiles/r2za_sd_two_phase_ncu_full.txt): sd_min_kernel issues in 68 % of its cycles (33.8 G warp instructions, "
                  "22.6 of 32 lanes active, 5 blocks of 93 registers per SM), L1 data pipe 55 % busy (4.4 G sectors of global loads with "
                  "the 64-byte records; 89 % with the 128-byte ones, profiles/r2w), FP64 pipe 15 %, L2 hit 94 %, DRAM 3.0 GB per "
                  "launch: bound by instruction issue and the dependent latency of a traversal step, not by any memory level"}

    # ---- CPU baseline (rank 0, N=1 only): bounded sample, checked against the GPU result ----
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        step = 2 if args.cpu_sample == "large" else 4
        rate, info = cpu_reference_rate(x, y, z, conn, step)
        sub = phi_d.reshape(GRID, GRID, GRID)[::step, ::step, ::step].reshape(-1).cpu().numpy()  # world == 1: all planes
        cpu = {"value": rate, "unit": UNIT, "cores": info["cores"], "kind": info["kind"],
               "sample": "%d^3 sub-lattice (i,j,k = 0 mod %d) of the %d^3 grid, %d points, OpenMP over queries, %.1f s" % (
                   GRID // step, step, GRID, info["npts"], info["seconds"][0]),
               "matches_gpu_bit_exact": bool(np.array_equal(sub, info["phi"]))}


    # ---- the consumer of the field: quest::MarchingCubes on phi, zero level set, device-resident ----
    mc_info = mc_sharded
    if world == 1 and not args.no_marching_cubes:
        try:
            mc_info = marching_cubes_leg(phi_d, ax, local, check=not args.no_cpu_baseline)
        except Exception as e:  # never lose the headline line to the secondary leg
            mc_info = {"error": "%s: %s" % (type(e).__name__, e)}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": WORKLOAD, "queries_total": nq_total, "queries_per_gpu": nq_local, "sharding": {"planes": "z-planes round-robin over ranks", "blocks": "blocks of 4 z-planes dealt round-robin over the ranks",
                                "slabs": "contiguous z-slabs, one per rank"}[args.sharding] + ", BVH replicated per GPU",
                   "l2_policy": "inputs larger than L2 (403 MB of queries per pass)", "mode": "mode 1: oriented-bound overlay, Morton-ordered queries; sample pass (1 query in 32) for first bounds, order-free exact-minimum search, ordered replay of the in-window leaves, heavy queries one warp each"},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": nq_local * 24, "d2h_bytes_per_step": nq_local * 8,
                "ms_per_step": e2e_ms, "steps": e2e_steps, "matches_device_path": same},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": roofline,
        "fp64": fp64,
        "l1": l1,
        "cpu_baseline": cpu,
        "marching_cubes": mc_info,
        "kernel_ms_per_rank": {"min": kernel_ms_min, "max": kernel_ms_max},
        "configs": cfg,
        "build_ms": build_ms, "build_phases_ms": build_phases, "setmesh_wall_ms": setmesh_wall_ms, "setmesh_device_ms": setmesh_device_ms, "first_setmesh_wall_ms": first_setmesh_wall_ms,
        "build_roofline": {"bound": "hbm", "achieved": 156.0 * ntri / (build_ms * 1e-3) / 1e9, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                           "frac": 156.0 * ntri / (build_ms * 1e-3) / 1e9 / peaks["hbm_gbs"], "algorithmic_bytes_per_box": 156},
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--no-marching-cubes", action="store_true", help="skip the MarchingCubes leg on the computed field")
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--sharding", default="blocks", choices=["blocks", "slabs", "planes"], help="how the 256 z-planes are split over ranks")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-sample", default="small", choices=["small", "large"])
    ap.add_argument("--configs", default="C1,C3,C4,C5,DCP", help="the other BASELINE configs to run after the headline (C2), and DCP = the full DistributedClosestPoint; '' = none")
    ap.add_argument("--config-scale", type=float, default=1.0, help="shrink C1/C3/C4/C5 (0.1 -> 10x fewer boxes and queries); 1.0 = BASELINE sizes")
    ap.add_argument("--config-steps", type=int, default=3)
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
