"""GPU parity tests for quest::SignedDistance through the C ABI, against the CPU oracle.
Distances and closest points must be bit-identical (tolerance stated by north_star: 1e-12
relative); signs equal; unit normals to 1e-12 (acos of libm vs CUDA differs in the last ulp)."""
import numpy as np
import pytest

from axom_b200 import synth

pytestmark = pytest.mark.gpu


def _cmp(oracle, x, y, z, conn, q, npc=3, wt=True, cs=True):
    """both traversal modes against the oracle.
    mode 0 (reference visiting order): phi and closest points bit-identical.
    mode 1 (oriented bounds, Morton-ordered queries; the default): the evaluated leaves are a
    subsequence of the reference's visiting order, so phi AND closest points are bit-identical
    (ties between equidistant triangles resolve as in the reference); normals to 1e-12 (libm acos)."""
    from axom_b200 import SignedDistance
    ref = oracle.SignedDistance(x, y, z, conn, npc, wt, cs)
    gpu = SignedDistance(x, y, z, conn, npc, wt, cs)
    rphi, rcp, rn = ref.compute(q, True, True)
    for mode in (0, 1):
        gpu.setMode(mode)
        gphi, gcp, gn = gpu.computeDistances(q, True, True)
        assert np.array_equal(rphi, gphi), (mode, np.abs(rphi - gphi).max())
        assert np.array_equal(rcp, gcp), mode
        assert np.allclose(rn, gn, rtol=0, atol=1e-12), mode
    lo, hi = gpu.getMeshBounds()
    assert lo[0] == x.min() and hi[2] == z.max()
    return gphi


def test_reference_sphere_golden_norms(oracle):
    # quest/tests/quest_signed_distance.cpp:88-163
    x, y, z, conn = synth.latlong_sphere(0.5, 25, 25)
    lo = np.array([x.min(), y.min(), z.min()]) - 2.0
    hi = np.array([x.max(), y.max(), z.max()]) + 2.0
    q = synth.uniform_grid_points(lo, hi, 16)
    phi = _cmp(oracle, x, y, z, conn, q)
    d = phi - (np.linalg.norm(q, axis=1) - 0.5)
    assert np.abs(d).max() < 1e-2
    assert abs(np.abs(d).sum() - 6.7051997372579715) < 1e-3
    assert abs(np.sqrt(d.sum()) - 2.5894400431865519) < 1e-3
    assert abs(np.abs(d).max() - 0.00532092) < 1e-3


def test_icosphere_grid_and_features(oracle):
    x, y, z, conn = synth.icosphere(12)
    _cmp(oracle, x, y, z, conn, synth.uniform_grid_points(-1, 1, 24))
    rng = np.random.default_rng(5)
    P = np.stack([x, y, z], 1)
    qv = P[rng.integers(0, len(x), 500)] * rng.choice([1.0, 0.9, 1.1, 1.0000001], 500)[:, None]
    e = 0.5 * (P[conn[:, 0]] + P[conn[:, 1]])
    qe = e[rng.integers(0, len(e), 500)] * rng.choice([1.0, 0.7, 1.3], 500)[:, None]
    q = np.concatenate([qv, qe])
    _cmp(oracle, x, y, z, conn, q)
    _cmp(oracle, x, y, z, conn, q, wt=False)
    _cmp(oracle, x, y, z, conn, q, cs=False)


def test_quads_and_plane(oracle):
    g = np.linspace(-1, 1, 9)
    X, Y = np.meshgrid(g, g, indexing="ij")
    xs, ys = X.ravel(), Y.ravel()
    zs = 0.1 * np.sin(3 * xs) * np.cos(2 * ys)
    idx = lambda i, j: i * 9 + j
    quads = np.array([[idx(i, j), idx(i + 1, j), idx(i + 1, j + 1), idx(i, j + 1)] for i in range(8) for j in range(8)], np.int32)
    _cmp(oracle, xs, ys, zs, quads, synth.uniform_grid_points(-1.5, 1.5, 12), npc=4, wt=False)
    # 4-triangle z=0 plane, non-watertight: phi == z exactly (quest_signed_distance_interface.cpp:186-237)
    px = np.array([-5.0, 5.0, 5.0, -5.0, 0.0])
    py = np.array([-5.0, -5.0, 5.0, 5.0, 0.0])
    pz = np.zeros(5)
    tris = np.array([[0, 1, 4], [1, 2, 4], [2, 3, 4], [3, 0, 4]], np.int32)
    q = synth.uniform_grid_points(-4, 4, 16)
    phi = _cmp(oracle, px, py, pz, tris, q, wt=False)
    assert np.array_equal(phi, q[:, 2])


def test_medium_icosphere_device_queries(oracle):
    import torch
    from axom_b200 import SignedDistance
    x, y, z, conn = synth.icosphere(40)  # 32 000 triangles
    q = synth.uniform_grid_points(-1, 1, 20)
    ref = oracle.SignedDistance(x, y, z, conn).compute(q, True, False, nthreads=0)
    gpu = SignedDistance(torch.from_numpy(x).cuda(), torch.from_numpy(y).cuda(), torch.from_numpy(z).cuda(),
                         torch.from_numpy(conn).cuda())
    qd = torch.from_numpy(q).cuda()
    phi, cp, _ = gpu.computeDistances(qd, True, False)
    assert np.array_equal(ref[0], phi.cpu().numpy())
    assert np.array_equal(ref[1], cp.cpu().numpy())
    # SoA (ZipIndexable) queries give the same answer
    phi2, _, _ = gpu.computeDistances(tuple(qd[:, c].contiguous() for c in range(3)))
    assert torch.equal(phi, phi2)


def test_golden_fixture_on_gpu():
    import os
    from axom_b200 import SignedDistance
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "sd_icosphere5.npz"))
    sd = SignedDistance(g["x"], g["y"], g["z"], g["conn"])
    for mode in (0, 1):
        sd.setMode(mode)
        phi, cp, nrm = sd.computeDistances(g["q"], True, True)
        assert np.array_equal(phi, g["phi"]) and np.array_equal(cp, g["cp"])
        assert np.allclose(nrm, g["nrm"], rtol=0, atol=1e-12)


def test_scalar_overload_and_bvh_access(oracle):
    from axom_b200 import SignedDistance
    x, y, z, conn = synth.icosphere(4)
    sd = SignedDistance(x, y, z, conn)
    ref = oracle.SignedDistance(x, y, z, conn)
    for p in ([0.0, 0.0, 0.0], [0.3, -0.2, 0.9], [0.5, 0.0, 0.0]):
        assert sd.computeDistance(*p) == ref.compute(np.array([p]))[0][0]
    b = sd.getBVHTree()
    assert b.isInitialized() and b.numLeaves() == len(conn)
    assert b.getScaleFactor() == 1.000123  # SignedDistance keeps the default scale (:499-500)


def test_fast_mode_large_sorted_queries(oracle):
    """enough queries to take the Morton-sort path (>= 4096), random + grid, both modes agree with the oracle"""
    from axom_b200 import SignedDistance
    x, y, z, conn = synth.icosphere(30)
    rng = np.random.default_rng(8)
    q = np.concatenate([synth.uniform_grid_points(-1, 1, 20), rng.uniform(-2, 2, (6000, 3)), rng.normal(0, 0.02, (2000, 3))])
    ref = oracle.SignedDistance(x, y, z, conn).compute(q, True, False, nthreads=0)
    sd = SignedDistance(x, y, z, conn)
    phi, cp, _ = sd.computeDistances(q, True, False)
    assert np.array_equal(ref[0], phi)
    assert np.array_equal(ref[1], cp)


def _degenerate_mesh():
    """icosphere(10) + what a dirty STL holds: triangles collapsed to a point or to a segment, slivers of area
    ~1e-13 and ~1e-17, unwelded (coincident) vertex copies, a duplicated triangle and a 1e-16-thin triangle of
    the kind primal_closest_point.cpp:239-372 uses"""
    x, y, z, conn = synth.icosphere(10)
    P = np.stack([x, y, z], 1)
    nv = len(P)
    extra_pts, extra_tris = [], []

    def add(p):
        extra_pts.append(np.asarray(p, np.float64))
        return nv + len(extra_pts) - 1
    # collapsed to a point (three equal ids) and to a segment (two equal ids)
    extra_tris += [[5, 5, 5], [7, 9, 9], [11, 11, 40]]
    # unwelded copies of existing vertices: a triangle on copies of a real triangle's vertices
    a, b, c = conn[17]
    extra_tris.append([add(P[a]), add(P[b]), add(P[c])])
    # a duplicated triangle (same ids) and one with reversed winding
    extra_tris.append(list(conn[33]))
    extra_tris.append(list(conn[34][::-1]))
    # slivers on an edge: third vertex 1e-12 / 1e-16 off the edge's midpoint -> area ~1e-13 / ~1e-17
    for t, off in ((50, 1e-12), (51, 1e-16), (52, 0.0)):
        a, b, c = conn[t]
        m = 0.5 * (P[a] + P[b])
        n = np.cross(P[b] - P[a], P[c] - P[a])
        n /= np.linalg.norm(n)
        extra_tris.append([a, b, add(m + off * n)])
    # a needle: two vertices 1e-13 apart
    a, b, c = conn[70]
    extra_tris.append([a, add(P[a] + np.array([1e-13, 0, 0])), c])
    # the reference KAT's thin triangle, scaled into the mesh's neighbourhood
    extra_tris.append([add([0.6, 0.0, 0.0]), add([0.6, 1e-16, 0.0]), add([0.7, 0.0, 0.1])])
    P2 = np.concatenate([P, np.array(extra_pts)])
    conn2 = np.concatenate([conn, np.array(extra_tris, np.int32)]).astype(np.int32)
    return P2[:, 0].copy(), P2[:, 1].copy(), P2[:, 2].copy(), conn2, len(conn)


def test_degenerate_and_sliver_triangles(oracle, have_ref):
    """A16 / A18: degenerate, sliver and coincident-vertex triangles reach closest_point(Point,Triangle), the
    !degenerate() guard of the vertex pseudo-normal (quest/SignedDistance.hpp:722-728) and Triangle::angle on the
    GPU, and must come out as in the reference (both kernels; watertight on/off)"""
    from axom_b200 import SignedDistance
    x, y, z, conn, nclean = _degenerate_mesh()
    P = np.stack([x, y, z], 1)
    rng = np.random.default_rng(21)
    dirty = conn[nclean:]
    qs = [synth.uniform_grid_points(-0.8, 0.8, 14)]
    for t in dirty:                                   # on / next to every dirty triangle's vertices, edges, centroid
        V = P[t]
        cen = V.mean(axis=0)
        qs += [V, V * 1.05, V * 0.95, V + rng.normal(0, 1e-7, (3, 3)), V + rng.normal(0, 1e-13, (3, 3)),
               0.5 * (V + np.roll(V, 1, axis=0)), cen[None], cen[None] * 1.2, cen[None] + rng.normal(0, 1e-3, (8, 3))]
    q = np.concatenate(qs)
    kinds = (["reference"] if have_ref else []) + ["port"]
    for wt in (True, False):
        gpu = SignedDistance(x, y, z, conn, isWatertight=wt)
        for kind in kinds:
            rphi, rcp, rn = oracle.SignedDistance(x, y, z, conn, watertight=wt, kind=kind).compute(q, True, True)
            for mode in (0, 1):
                gpu.setMode(mode)
                gphi, gcp, gn = gpu.computeDistances(q, True, True)
                assert np.array_equal(rphi, gphi), (kind, wt, mode, int((rphi != gphi).sum()))
                assert np.array_equal(rcp, gcp), (kind, wt, mode)
                ok = np.isfinite(rn).all(axis=1)
                assert np.array_equal(ok, np.isfinite(gn).all(axis=1))
                assert np.allclose(rn[ok], gn[ok], rtol=0, atol=1e-12), (kind, wt, mode)


def _mixed_sheet():
    """wavy 8x8 sheet: quads on the even squares, two triangles on the odd ones (MIXED_SHAPE mesh)"""
    g = np.linspace(-1, 1, 9)
    X, Y = np.meshgrid(g, g, indexing="ij")
    xs, ys = X.ravel(), Y.ravel()
    zs = 0.1 * np.sin(3 * xs) * np.cos(2 * ys)
    idx = lambda i, j: i * 9 + j
    conn, off = [], [0]
    for i in range(8):
        for j in range(8):
            q = [idx(i, j), idx(i + 1, j), idx(i + 1, j + 1), idx(i, j + 1)]
            if (i + j) % 2 == 0:
                conn += q
                off.append(off[-1] + 4)
            else:
                conn += [q[0], q[1], q[2]]
                off.append(off[-1] + 3)
                conn += [q[0], q[2], q[3]]
                off.append(off[-1] + 3)
    return xs, ys, zs, np.array(conn, np.int32), np.array(off, np.int32)


def test_mixed_shape_mesh(oracle, have_ref):
    """mint::UnstructuredMesh<MIXED_SHAPE> (UcdMeshData with cell_node_offsets, quest/SignedDistance.hpp:44-96)"""
    from axom_b200 import SignedDistance
    xs, ys, zs, conn, off = _mixed_sheet()
    rng = np.random.default_rng(4)
    q = np.concatenate([synth.uniform_grid_points(-1.5, 1.5, 18), rng.uniform(-1, 1, (3000, 3)) * [1, 1, 0.2]])
    gpu = SignedDistance(xs, ys, zs, conn, isWatertight=False, cell_node_offsets=off)
    for kind in (["reference"] if have_ref else []) + ["port"]:
        rphi, rcp, rn = oracle.SignedDistance(xs, ys, zs, conn, watertight=False, offsets=off, kind=kind).compute(q, True, True)
        for mode in (0, 1):
            gpu.setMode(mode)
            gphi, gcp, gn = gpu.computeDistances(q, True, True)
            assert np.array_equal(rphi, gphi) and np.array_equal(rcp, gcp), (kind, mode)
            assert np.allclose(rn, gn, rtol=0, atol=1e-12)
    from axom_b200._lib import AxbError
    bad = off.copy()
    bad[1] += 2  # a 6-node "cell" followed by a 1-node one
    with pytest.raises(AxbError):
        SignedDistance(xs, ys, zs, conn, isWatertight=False, cell_node_offsets=bad)


def test_host_to_host_query_is_pipelined_in_chunks(oracle, monkeypatch):
    """host inputs + host outputs above 2 x AXB_SD_PIPE_CHUNK points go through the two-stream chunk pipeline
    (upload / kernel / download of neighbouring chunks overlap); results must not depend on the chunking, with
    closest points and normals, AoS and SoA query layouts, and a ragged last chunk"""
    from axom_b200 import SignedDistance
    monkeypatch.setenv("AXB_SD_PIPE_CHUNK", "4096")
    x, y, z, conn = synth.icosphere(12)
    q = synth.random_points(4096 * 5 + 123, seed=4, lo=-1.0, hi=1.0)
    want_phi, want_cp, _ = oracle.SignedDistance(x, y, z, conn).compute(q, True, True)
    sd = SignedDistance(x, y, z, conn)
    phi, cp, nrm = sd.computeDistances(q, True, True)
    assert np.array_equal(phi, want_phi) and np.array_equal(cp, want_cp)
    assert np.allclose(np.linalg.norm(nrm, axis=1), 1.0, atol=1e-12)
    soa = tuple(np.ascontiguousarray(q[:, c]) for c in range(3))
    phi2, _, _ = sd.computeDistances(soa)
    assert np.array_equal(phi2, want_phi)
    monkeypatch.setenv("AXB_SD_PIPE_CHUNK", "0")  # pipeline off: same answer
    phi3, cp3, nrm3 = sd.computeDistances(q, True, True)
    assert np.array_equal(phi3, want_phi) and np.array_equal(cp3, cp) and np.array_equal(nrm3, nrm)


def test_heavy_queries_take_the_warp_cooperative_path(oracle, monkeypatch):
    """Queries at and around the centre of a sphere: hundreds to thousands of triangles within the 1e-6 tie window, the
    remembered-leaf list of phase 1 overflows and the query is finished by sd_solo_kernel (exact minimum, then the
    in-window leaves in the reference's visiting order, one warp per query).  icosphere(20): 4000 triangles, the ordered
    frontier fits the scratch; icosphere(32): 10240 triangles, at the exact centre it does not and the warp falls back to
    the serial ordered walk.  Both must match the oracle bit for bit, with and without the cooperative kernel."""
    from axom_b200 import SignedDistance
    rng = np.random.default_rng(3)
    for freq in (20, 32):
        x, y, z, conn = synth.icosphere(freq)
        q = np.concatenate([np.zeros((1, 3)), rng.normal(0, 1e-9, (8, 3)), rng.normal(0, 1e-6, (16, 3)), rng.normal(0, 1e-3, (40, 3)),
                            rng.normal(0, 0.02, (200, 3)), synth.uniform_grid_points(-0.05, 0.05, 12)])
        q = np.concatenate([q, rng.uniform(-1.2, 1.2, (4096, 3))])  # enough for the Morton-sorted path
        for cs in (True, False):
            rphi, rcp, rn = oracle.SignedDistance(x, y, z, conn, 3, True, cs).compute(q, True, True, nthreads=0)
            for solo_off in (False, True):
                if solo_off:
                    monkeypatch.setenv("AXB_SD_NO_SOLO", "1")
                else:
                    monkeypatch.delenv("AXB_SD_NO_SOLO", raising=False)
                sd = SignedDistance(x, y, z, conn, 3, True, cs)
                phi, cp, nr = sd.computeDistances(q, True, True)
                assert np.array_equal(rphi, phi), (freq, cs, solo_off, int((rphi != phi).sum()))
                assert np.array_equal(rcp, cp), (freq, cs, solo_off)
                assert np.allclose(rn, nr, rtol=0, atol=1e-12), (freq, cs, solo_off)


def test_first_bound_from_a_point_the_reference_arithmetic_cannot_reach(oracle, have_ref):
    """One Morton part (2.5 M triangles, an open patch) of the 20 M-triangle sphere of BASELINE config C5, unsigned
    distance, queries on the far side of the patch so that every closest point lies on the patch boundary.  Triangle
    areas are 1e-7: their square is below the EPS = 1e-12 of the reference's closest_point(Point, Triangle, loc, EPS)
    (primal/operators/closest_point.hpp:162-290), whose fuzzy region tests then return an edge point where a vertex is
    nearer (3e-6 off in the squared distance).  A first bound taken from a neighbouring query's closest point (a vertex,
    reached exactly from there) is then BELOW every value the reference arithmetic yields: the search must notice the
    total miss and start over without a bound.  Truth: the reference-order kernel (mode 0, no bounds from hints), itself
    pinned to the unmodified reference on a sample here."""
    import torch
    from axom_b200 import SignedDistance
    from axom_b200 import dist as D
    x, y, z, conn = synth.icosphere(1000)
    P = np.stack([x, y, z], 1)
    cen = (P[conn[:, 0]] + P[conn[:, 1]] + P[conn[:, 2]]) / 3.0
    c = conn[D.morton_partition(cen, 8)[7]]
    pc = cen[D.morton_partition(cen, 8)[7]].mean(axis=0)
    rng = np.random.default_rng(11)
    q = rng.uniform(-1.0, 1.0, (6_000_000, 3))
    q = q[(q * np.sign(pc)).max(axis=1) < 0.0][:3_000_000]  # the octant opposite to the patch
    assert len(q) > 500_000
    qd = torch.from_numpy(np.ascontiguousarray(q)).cuda()
    sd = SignedDistance(x, y, z, c, 3, False, False)
    sd.setMode(0)
    truth = sd.computeDistances(qd)[0]
    sd.setMode(1)
    got = sd.computeDistances(qd)[0]
    assert int((got > 1e100).sum()) == 0
    assert torch.equal(got, truth), int((got != truth).sum())
    kind = "reference" if have_ref else "port"
    want, _, _ = oracle.SignedDistance(x, y, z, c, 3, False, False, kind=kind).compute(q[:100_000], nthreads=0)
    assert np.array_equal(want, truth[:100_000].cpu().numpy())
    # the quirk itself, on one triangle of this mesh: EPS = 1e-12 picks edge AB, the exact tests pick vertex C
    tri = np.array([[0.42520727074501397, 0.00068895192447202, 0.263055701802531, 0.42494387910357284, 0.00052631214548544, 0.2634813515373215,
                     4.2513491742628706e-01, 2.6301083961343877e-04, 2.6317338925172956e-01]])
    qq = np.array([[0.7702697736345363, -0.0036593517992807856, 0.47534102216064422]])
    from axom_b200 import primal
    for eps, loc_want in ((1e-12, -1), (1e-50, 2)):
        cp_r, loc_r = oracle.closest_point_tri(qq, tri, eps, kind)
        cp_g, loc_g = primal.closest_point(qq, tri.reshape(-1, 3, 3), eps)
        assert loc_r[0] == loc_want and loc_g[0] == loc_want and np.array_equal(cp_r, cp_g)


@pytest.mark.parametrize("scale,shift", [(1e-3, (1.0e5, -3.0e5, 7.0e5)), (1.0e7, (0.0, 0.0, 0.0)), (1e-6, (0.0, 0.0, 0.0)), (1.0, (2.0e9, 2.0e9, -2.0e9))])
def test_compact_records_far_from_the_origin_and_at_extreme_scales(oracle, scale, shift):
    """The 64-byte node records keep the node origin in binary32 and centres / extents on a per-node 16-bit grid.  A tiny mesh
    far from the origin (the rounded origin is then farther from the node than the node is wide), a huge one and a
    microscopic one must still give the reference's distances bit for bit: the quantisation only loosens the bounds."""
    from axom_b200 import SignedDistance
    x, y, z, conn = synth.icosphere(14)
    s = np.array(shift)
    x, y, z = x * scale + s[0], y * scale + s[1], z * scale + s[2]
    rng = np.random.default_rng(21)
    q = rng.uniform(-1.3, 1.3, (6000, 3)) * scale * 0.5 + s
    q = np.concatenate([q, np.stack([x, y, z], 1)[:200], rng.uniform(-40.0, 40.0, (500, 3)) * scale + s])
    for cs in (True, False):
        rphi, rcp, _ = oracle.SignedDistance(x, y, z, conn, 3, True, cs).compute(q, True, False, nthreads=0)
        sd = SignedDistance(x, y, z, conn, 3, True, cs)
        phi, cp, _ = sd.computeDistances(q, True, False)
        assert np.array_equal(rphi, phi), (scale, shift, cs, int((rphi != phi).sum()))
        assert np.array_equal(rcp, cp), (scale, shift, cs)


def test_update_min_distances_over_surface_parts_equals_the_whole_surface(oracle):
    """axb_sd_update_min_distances: the parts of a partitioned surface evaluated in turn on one GPU into a running minimum that
    also bounds the next part's search; the result is the distance to the whole surface, bit for bit (device and host
    arrays, parts taken in two different orders, DBL_MAX entries)."""
    import torch
    from axom_b200 import SignedDistance
    from axom_b200 import dist as D
    x, y, z, conn = synth.icosphere(24)
    P = np.stack([x, y, z], 1)
    cen = (P[conn[:, 0]] + P[conn[:, 1]] + P[conn[:, 2]]) / 3.0
    parts = D.morton_partition(cen, 5)
    rng = np.random.default_rng(31)
    q = np.concatenate([rng.uniform(-1.2, 1.2, (40000, 3)), P[:500], rng.normal(0, 1e-3, (200, 3))])
    want, _, _ = oracle.SignedDistance(x, y, z, conn, 3, False, False).compute(q, nthreads=0)
    sds = [SignedDistance(x, y, z, conn[p], 3, False, False) for p in parts]
    qd = torch.from_numpy(q).cuda()
    for order in ([0, 1, 2, 3, 4], [3, 0, 4, 2, 1]):
        out = torch.full((len(q),), float(np.finfo(np.float64).max), dtype=torch.float64, device="cuda")
        for k in order:
            sds[k].updateMinDistances(qd, out)
        assert np.array_equal(out.cpu().numpy(), want), order
        outh = np.full(len(q), np.finfo(np.float64).max)
        for k in order:
            sds[k].updateMinDistances(q, outh)
        assert np.array_equal(outh, want), order
    small = np.ascontiguousarray(q[:100])  # below the sorted / sampled path
    o = np.full(100, np.finfo(np.float64).max)
    for sd in sds:
        sd.updateMinDistances(small, o)
    assert np.array_equal(o, want[:100])
    from axom_b200._lib import AxbError
    with pytest.raises(AxbError):
        SignedDistance(x, y, z, conn, 3, True, True).updateMinDistances(small, o)
