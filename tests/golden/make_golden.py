#!/usr/bin/env python3
"""Generate the golden fixtures in this directory from the REAL reference (LLNL/axom v0.11.0,
SEQ_EXEC) compiled by oracle/build_ref.py.  Run where /root/reference exists:

    python oracle/build_ref.py && python tests/golden/make_golden.py

The fixtures hold seeded inputs and the reference's outputs; tests/test_oracle_golden.py pins the
oracle port (oracle/axb_oracle.cpp) to them, and the GPU tests compare against the same files."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from axom_b200 import synth  # noqa: E402
from oracle import oracle as O  # noqa: E402


def bvh_case(name, n, ndims, seed):
    boxes = synth.triangle_aabbs(n, seed=seed, ndims=ndims)
    boxes[3] = boxes[4]  # tied Morton codes
    b = O.Bvh(boxes, ndims=ndims, kind="reference")
    A = b.arrays()
    pts = synth.random_points(400, seed=seed + 1, ndims=ndims)
    qb = synth.triangle_aabbs(300, seed=seed + 2, ndims=ndims)
    ro, rd = synth.random_rays(200, seed=seed + 3, lo=-0.5, hi=1.5, ndims=ndims)
    p = b.find_points(pts)
    x = b.find_boxes(qb)
    r = b.find_rays(ro, rd * 1.3, True)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), boxes=boxes, ndims=ndims, pts=pts, qboxes=qb, ray_o=ro, ray_d=rd * 1.3,
                        p_off=p[0], p_cnt=p[1], p_cand=p[2], b_off=x[0], b_cnt=x[1], b_cand=x[2], r_off=r[0], r_cnt=r[1],
                        r_cand=r[2], **{"a_" + k: v for k, v in A.items()})


def sd_case(name, freq, grid):
    x, y, z, conn = synth.icosphere(freq)
    rng = np.random.default_rng(11)
    P = np.stack([x, y, z], 1)
    qv = P[rng.integers(0, len(x), 60)] * rng.choice([1.0, 0.9, 1.1], 60)[:, None]
    e = 0.5 * (P[conn[:, 0]] + P[conn[:, 1]])
    qe = e[rng.integers(0, len(e), 60)] * rng.choice([1.0, 0.7, 1.3], 60)[:, None]
    q = np.concatenate([synth.uniform_grid_points(-1, 1, grid), qv, qe])
    phi, cp, nrm = O.SignedDistance(x, y, z, conn, kind="reference").compute(q, True, True)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), x=x, y=y, z=z, conn=conn, q=q, phi=phi, cp=cp, nrm=nrm)


def meshtester_case(name):
    """quest::findTriMeshIntersectionsBVH<SEQ_EXEC,double> on two interpenetrating spheres + 2 degenerate cells,
    and primal::intersect on seeded random / lattice triangle pairs"""
    x, y, z, c = synth.icosphere(8)
    x2, y2, z2, c2 = synth.icosphere(6)
    X, Y, Z = np.concatenate([x, x2 * 0.9 + 0.3]), np.concatenate([y, y2 * 0.9]), np.concatenate([z, z2 * 0.9])
    C = np.concatenate([c, c2 + len(x), [[0, 0, 1], [2, 3, 2]]]).astype(np.int32)
    pairs, deg = O.find_tri_mesh_intersections(X, Y, Z, C, 1e-8, "reference")
    rng = np.random.default_rng(77)
    a = rng.random((3000, 3, 3))
    b = rng.random((3000, 3, 3)) * 0.4 + 0.3
    g = rng.integers(0, 3, (3000, 3, 3)).astype(np.float64)
    h = rng.integers(0, 3, (3000, 3, 3)).astype(np.float64)
    g[:1000, :, 2] = 0
    h[:1000, :, 2] = 0
    t1, t2 = np.concatenate([a, g]), np.concatenate([b, h])
    np.savez_compressed(os.path.join(HERE, name + ".npz"), pairs=pairs, degenerate=deg, t1=t1.astype(np.float32), t2=t2.astype(np.float32),
                        hit_open=O.tri_tri_intersect(t1.astype(np.float32), t2.astype(np.float32), False, 1e-8, "reference"),
                        hit_closed=O.tri_tri_intersect(t1.astype(np.float32), t2.astype(np.float32), True, 1e-8, "reference"))


def stl_weld_case(name):
    """quest::STLReader + quest::weldTriMeshVertices on a jittered icosphere soup, and the legacy process-global
    signed-distance API (signed_distance_init(file) ... evaluate) on an ASCII STL: the inputs are re-created by
    tests/test_quest_interface.py with the same seeds"""
    import tempfile
    out = {}
    with tempfile.TemporaryDirectory() as d:
        x, y, z, conn = synth.icosphere(6)
        jit = np.random.default_rng(3).uniform(-2e-8, 2e-8, (len(conn), 3, 3))
        p = os.path.join(d, "s.stl")
        synth.write_stl(p, x, y, z, conn, binary=False, jitter=jit)
        for eps in (1e-7, 1e-3, 0.1):
            wx, wy, wz, wc = O.ref_stl_read_weld(p, eps)
            out["eps_%g_xyz" % eps] = np.stack([wx, wy, wz], 1)
            out["eps_%g_conn" % eps] = wc
        x, y, z, conn = synth.icosphere(8)
        p = os.path.join(d, "t.stl")
        synth.write_stl(p, x, y, z, conn, binary=False)
        q = synth.uniform_grid_points(-1, 1, 12)
        phi, lo, hi = O.ref_legacy_signed_distance(p, q[:, 0].copy(), q[:, 1].copy(), q[:, 2].copy())
        out["legacy_phi_ascii_f8"] = phi
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)


def mc_cases(name):
    """quest::MarchingCubes (the real reference through oracle/ref_mc_driver.cpp) on tests/mc_cases.py"""
    import hashlib
    sys.path.insert(0, os.path.dirname(HERE))
    import mc_cases
    out = {}
    for cname in mc_cases.CASES:
        mesh, mask_field, mask_val, contours = mc_cases.build(cname)
        ids, xyz, par, dom = O.ref_mc_isocontour(mesh, "mesh", "dist", mask_field, mask_val, contours, data_parallelism=1)
        full = O.ref_mc_isocontour(mesh, "mesh", "dist", mask_field, mask_val, contours, data_parallelism=2)
        assert all(np.array_equal(a, b) for a, b in zip((ids, xyz, par, dom), full)), cname  # hybridParallel == fullParallel
        sha = hashlib.sha256()
        for d in mesh.values():
            sha.update(np.ascontiguousarray(d["fields"]["dist"]["values"]).tobytes())
        out[cname + "/ids"], out[cname + "/xyz"], out[cname + "/par"], out[cname + "/dom"] = ids, xyz, par, dom
        out[cname + "/fcn_sha"] = np.frombuffer(sha.digest(), np.uint8)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)


def vtk_case():
    """BVH::writeVtkFile of the reference for the small trees of tests/kats.vtk_boxes -> bvh_vtk_{3d,2d}.vtk"""
    sys.path.insert(0, os.path.join(HERE, ".."))
    import kats
    for nd in (3, 2):
        O.Bvh(kats.vtk_boxes(nd), ndims=nd, kind="reference").write_vtk(os.path.join(HERE, "bvh_vtk_%dd.vtk" % nd))


if __name__ == "__main__":
    assert O.have_reference(), "build the reference first: python oracle/build_ref.py"
    vtk_case()
    bvh_case("bvh3d_n600", 600, 3, 41)
    bvh_case("bvh2d_n400", 400, 2, 43)
    sd_case("sd_icosphere5", 5, 9)
    meshtester_case("meshtester_spheres")
    stl_weld_case("stl_weld")
    mc_cases("mc_contours")
    print("golden fixtures written to", HERE)
