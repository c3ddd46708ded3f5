"""Known-answer tests taken from the reference's own unit tests, shared by the oracle (CPU) and
the CUDA (GPU) test files.  Each function takes a `make_bvh(boxes, ndims, scale)` factory returning
an object with find_points / find_boxes / find_rays (oracle.Bvh interface)."""
import numpy as np


def unit_cells(ncell, ndims):
    """AABBs of the cells of mint::UniformMesh([0,ncell]^D, ncell+1 nodes/dim), x fastest
    (spin/tests/spin_bvh.cpp:90-122 generate_aabbs)"""
    idx = np.stack(np.meshgrid(*[np.arange(ncell)] * ndims, indexing="ij"), axis=-1).reshape(-1, ndims)[:, ::-1]
    idx = idx.astype(np.float64)
    return np.ascontiguousarray(np.concatenate([idx, idx + 1.0], axis=1))


# sorted Morton codes / permutation / radix-tree children of the 3x3x3 case with scale 1.0, obtained
# from the reference's lbvh::build_radix_tree<SEQ_EXEC>
KAT_MCODES = [14913080, 147000368, 164043889, 279087656, 313174698, 411174944, 428218465, 445261986, 462305507, 543262232,
              611436316, 675349520, 692393041, 743523604, 760567125, 807436808, 841523850, 875610892, 909697934, 939524096,
              956567617, 973611138, 990654659, 1007698180, 1024741701, 1041785222, 1058828743]
KAT_LEAFS = [0, 1, 2, 3, 6, 4, 5, 7, 8, 9, 18, 10, 11, 19, 20, 12, 15, 21, 24, 13, 14, 16, 17, 22, 23, 25, 26]
KAT_LCHILD = [8, 27, 26, 4, 29, 6, 31, 33, 2, 14, 35, 12, 37, 39, 10, 18, 41, 43, 16, 22, 45, 47, 20, 24, 49, 51]
KAT_RCHILD = [9, 28, 1, 5, 30, 7, 32, 34, 3, 15, 36, 13, 38, 40, 11, 19, 42, 44, 17, 23, 46, 48, 21, 25, 50, 52]


def children_to_lr(inner_children, n):
    """decode LinearBVH child ids (2*idx | -(pos+1)) back to RadixTree ids (idx | pos+n-1)"""
    c = np.asarray(inner_children).reshape(-1, 2)
    dec = np.where(c >= 0, c // 2, -c - 1 + (n - 1))
    return dec[:, 0], dec[:, 1]


def hits(off, cnt, cand, i):
    return sorted(np.asarray(cand)[off[i]:off[i] + cnt[i]].tolist())


def check_boxes_3d(bvh):
    # spin_bvh.cpp:279-404: 18 hits = cells 0..17; second box none
    q = np.array([[-1, -1, -1, 2.5, 2.5, 1.5], [-1, -1, -1, -0.5, -0.5, -0.5]], np.float64)
    off, cnt, cand = bvh.find_boxes(q)
    assert cnt[0] == 18 and cnt[1] == 0
    assert hits(off, cnt, cand, 0) == list(range(18))


def check_boxes_2d(bvh):
    # spin_bvh.cpp:408-516: 6 hits, cells 6,7,8 missed
    q = np.array([[-1, -1, 2.5, 1.5], [-1, -1, -0.1, -0.1]], np.float64)
    off, cnt, cand = bvh.find_boxes(q)
    assert cnt[0] == 6 and cnt[1] == 0
    assert hits(off, cnt, cand, 0) == [0, 1, 2, 3, 4, 5]


def check_rays_3d(bvh):
    # spin_bvh.cpp:519-646
    o = np.array([[-1, -1, -1], [-1, -1, -1]], np.float64)
    d = np.array([[1, 1, 1], [-1, -1, -1]], np.float64)
    off, cnt, cand = bvh.find_rays(o, d, True)
    assert cnt[0] == 15 and cnt[1] == 0
    assert hits(off, cnt, cand, 0) == [0, 1, 3, 4, 9, 10, 12, 13, 14, 16, 17, 22, 23, 25, 26]


def check_rays_2d(bvh):
    # spin_bvh.cpp:650-774: 7 hits, cells 2 and 6 missed
    o = np.array([[-1, -1], [-1, -1]], np.float64)
    d = np.array([[1, 1], [-1, -1]], np.float64)
    off, cnt, cand = bvh.find_rays(o, d, True)
    assert cnt[0] == 7 and cnt[1] == 0
    assert hits(off, cnt, cand, 0) == [0, 1, 3, 4, 5, 7, 8]


def check_points(bvh, boxes, ndims):
    # spin_bvh.cpp:792-999: cell centroids -> exactly one candidate (the cell); shifted by +10 -> none
    cen = 0.5 * (boxes[:, :ndims] + boxes[:, ndims:])
    off, cnt, cand = bvh.find_points(cen)
    assert (cnt == 1).all()
    assert np.array_equal(np.asarray(cand)[off], np.arange(len(boxes)))
    off, cnt, cand = bvh.find_points(cen + 10.0)
    assert (cnt == 0).all() and len(cand) == 0


# ---- primal::intersect(Triangle3, Triangle3): primal/tests/primal_intersect.cpp:826-1106 --------------------
# (t1, t2, expected with includeBoundary=True, expected with includeBoundary=False); every case is run over the
# 36 corner permutations of permuteCornersTest (:53-125), EPS = 1e-8.
_TRI_A = [(0, 0, 0), (1, 0, 0), (0, 1.7, 2.3)]
_I152 = [(1, 0, 0.5), (1, 0, -0.5), (0, 0, 0)]
TRI_TRI_KATS = [
    ([(-1, -1, -1), (-2, -5, -5), (-4, -8, -8)], [(-1, -1, -1), (-2, -5, -5), (-4, -8, -8)], True, True),   # identical
    ([(-1, -1, -1), (-2, -5, -5), (-4, -8, -8)], [(1, 1, 1), (5, 5, 5), (8, 7, 92)], False, False),         # disjunct
    (_TRI_A, [(0, 0, 0), (1, 0, 0), (0, -2, 1.2)], True, False),          # sharing a segment
    (_TRI_A, [(-0.2, 0, 0), (0.7, 0, 0), (0, -2, 1.2)], True, False),     # sharing part of a segment
    (_TRI_A, [(-1, 0, 0), (0, 4.3, 6), (0, 1.7, 2.3)], True, False),      # sharing a vertex
    (_TRI_A, [(0, -1, 0), (1, 1, 0), (0, 1.7, -2.3)], True, False),       # edges cross
    (_TRI_A, [(0, -1, -1), (0.5, 0, 0), (1, 1, -1)], True, False),        # B vertex lands on A's edge
    (_TRI_A, [(0.5, -1, 0.1), (0.5, 1, 0.1), (1, 1, -1)], True, True),    # two links in a chain
    (_TRI_A, [(-1, -1, 1), (0, 2, 1), (5, 0, 1)], True, True),            # A pokes through B
    (_TRI_A, [(1, -1, 1), (1, 2, 1), (1, 0, -1)], True, False),           # A vertex tangent on B
    (_TRI_A, [(1.00001, -1, 1), (1, 2, 1), (1, 0, -1)], False, False),    # not quite tangent
    # regression cases :923-1106
    ([(-1.83697e-14, 62.5, 300), (16.17619, 60.37037, 300), (-5.790149e-16, 11.26926, 9.456031)],
     [(-5.790149e-16, 11.26926, 9.456031), (16.17619, 60.37037, 300), (2.916699, 10.88527, 9.456031)], True, False),
    ([(-138.02488708496094, -14398.0908203125, 111881.2421875), (0.067092768847942352, -14407.21875, 111891.078125),
      (-136.77900695800781, -14416.4912109375, 111891.078125)],
     [(1.1611454486846924, -14423.3466796875, 111904.359375), (0.067092768847942352, -14407.21875, 111891.078125),
      (136.91319274902344, -14397.947265625, 111891.078125)], True, False),
    ([(-1, -1, -1), (0, 0, -0.005), (-1, 0, 0)], [(0, 1, -1), (0, 0, -0.005), (1, 0, 0)], True, False),
    ([(76.648, 54.6752, 15.0012), (76.648, 54.6752, 14.5542), (76.582, 54.6752, 14.7879)],
     [(76.6252, 54.6752, 14.892), (76.582, 54.6752, 14.7879), (76.5617, 54.6752, 14.7929)], True, True),
    ([(0.066, 0, 0.2133), (0.066, 0, -0.2337), (0, 0, 0)], [(0.0432, 0, 0.1041), (0, 0, 0), (-0.0203, 0, 0.005)], True, True),
    (_I152, [(0.5, 0, 0.1), (0, 0, 0), (-0.1, 0, -0.2)], True, True),
    (_I152, [(0.5, 0, 0.1), (0, 0, 0), (-0.1, 0, 0.05)], True, True),
    (_I152, [(0.5, 0, 0.1), (0, 0, 0), (-0.1, 0, 0.06)], True, True),
    (_I152, [(0.5, 0, 0.1), (0, 0, 0), (-0.1, 0, 0.04)], True, True),
]


def tri_tri_cases():
    """-> (t1 (n,3,3), t2 (n,3,3), include_boundary (n,) bool, expected (n,) bool) over all corner permutations"""
    def roll(t, i):
        return [t[i % 3], t[(i + 1) % 3], t[(i + 2) % 3]]
    T1, T2, INC, WANT = [], [], [], []
    for a, b, with_bdry, without_bdry in TRI_TRI_KATS:
        ap, bp = [a[0], a[2], a[1]], [b[0], b[2], b[1]]
        for inc, want in ((True, with_bdry), (False, without_bdry)):
            for u, v in ((a, b), (ap, bp), (b, a), (bp, ap)):
                for i in range(3):
                    for j in range(3):
                        T1.append(roll(u, i))
                        T2.append(roll(v, j))
                        INC.append(inc)
                        WANT.append(want)
    return np.array(T1, np.float64), np.array(T2, np.float64), np.array(INC, bool), np.array(WANT, bool)


# ---- leaf math: the reference's own unit tests for the arithmetic on the path ------------------------------------------
def closest_point_tri_cases():
    """primal/tests/primal_closest_point.cpp:239-576: (points, triangles, eps, expected cp, expected loc, cp tolerance).
    The 1e-16-thin triangle and the four degenerate-side triangles, EPS = PRIMAL_TINY."""
    P, T, C, L, TOL = [], [], [], [], []

    def add(tri, q, cp, loc, tol=0.0):
        P.append(q), T.append(np.asarray(tri, np.float64).reshape(9)), C.append(cp), L.append(loc), TOL.append(tol)

    tiny = [[0.0, 0.0, 0.0], [0.0, 0.0, 1.0e-16], [0.0, 1.0, 0.0]]  # :246
    A, B, Cv = tiny
    add(tiny, [0, 0, 0], A, 0), add(tiny, [0, 0, -1e-16], A, 0), add(tiny, [0, 0, 1e-16], B, 1), add(tiny, [0, 0, 1e-15], B, 1)
    add(tiny, [0, 1, 0], Cv, 2), add(tiny, [0, 1.0 + 1e-16, 0], Cv, 2)
    add(tiny, [0, 0, 1e-17], [0, 0, 1e-17], -1, 1e-16), add(tiny, [0, -0.1, 1e-17], [0, 0, 1e-17], -1, 1e-16)
    add(tiny, [0, 0.5, 5e-17], [0, 0.5, 5e-17], -2, 1e-16), add(tiny, [0.5, 0.5, 5e-17], [0, 0.5, 5e-17], -2, 1e-16)
    add(tiny, [0, 0.25, 0], [0, 0.25, 0], -3, 1e-16), add(tiny, [-0.25, 0.75, -0.25], [0, 0.75, 0], -3, 1e-16)
    add(tiny, [0, 1.0 / 3.0, 1e-16 / 3.0], [0, 1.0 / 3.0, 1e-16 / 3.0], 3, 1e-16)
    add(tiny, [-0.5, 1.0 / 3.0, 1e-16 / 3.0], [0, 1.0 / 3.0, 1e-16 / 3.0], 3, 1e-16)
    ab = [[0.0, 0.0, 0.0], [0.0, 0.0, 0.0], [0.0, 1.0, 0.0]]  # degenerate side AB :383
    add(ab, [0, 0, 0], ab[0], 0), add(ab, [0, 0, -1e-16], ab[0], 0), add(ab, [0, 0, 1e-16], ab[0], 0), add(ab, [0, -1e-16, 0], ab[0], 0)
    add(ab, [0, 1, 0], ab[2], 2), add(ab, [0, 1.0 + 1e-16, 0], ab[2], 2)
    add(ab, [0, 0.25, 0], [0, 0.25, 0], -3, 1e-16), add(ab, [-0.25, 0.75, -0.25], [0, 0.75, 0], -3, 1e-16)
    bc = [[2.0, 0.0, 0.0], [2.0, 1.0, 1.0], [2.0, 1.0, 1.0]]  # degenerate side BC :449
    add(bc, [2, 0, 0], bc[0], 0), add(bc, [2, 1, 1], bc[1], 1), add(bc, [3, 2, 2], bc[1], 1)
    add(bc, [2, 0.75, 0.75], [2, 0.75, 0.75], -1, 1e-16), add(bc, [2, 1, 0], [2, 0.5, 0.5], -1, 1e-16)
    ca = [[1.0, 3.0, 1.0], [2.0, 4.0, 2.0], [1.0, 3.0, 1.0]]  # degenerate side CA :506
    add(ca, [1, 3, 1], ca[0], 0), add(ca, [0, 0, 0], ca[0], 0), add(ca, [2, 4, 2], ca[1], 1), add(ca, [2.1, 4.1, 2.1], ca[1], 1)
    add(ca, [4.0 / 3.0, 10.0 / 3.0, 4.0 / 3.0], [4.0 / 3.0, 10.0 / 3.0, 4.0 / 3.0], -1, 1e-16)
    add(ca, [1, 4, 1], [4.0 / 3.0, 10.0 / 3.0, 4.0 / 3.0], -1, 1e-16)
    pt = [[1.0, 3.0, 1.0]] * 3  # all sides degenerate :561
    add(pt, [1, 3, 1], pt[0], 0), add(pt, [2, 4, 2], pt[0], 0)
    return (np.asarray(P, np.float64), np.asarray(T, np.float64), 1e-50, np.asarray(C, np.float64), np.asarray(L, np.int32),
            np.asarray(TOL, np.float64))


def check_closest_point_tri(fn):
    """fn(points, triangles, eps) -> (cp, loc)"""
    P, T, eps, C, L, TOL = closest_point_tri_cases()
    cp, loc = fn(P, T, eps)
    assert np.array_equal(loc, L), (loc, L)
    assert np.all(np.abs(cp - C) <= TOL[:, None]), np.abs(cp - C).max(axis=1)


def check_squared_distance_point_box(fn):
    """primal_squared_distance.cpp:167-201: 27 points around the cube [-1,1]^3, and the empty box; fn(points, boxes) -> d2"""
    pts = np.array([[3.0 * i, 3.0 * j, 3.0 * k] for i in (-1, 0, 1) for j in (-1, 0, 1) for k in (-1, 0, 1)], np.float64)
    cube = np.tile(np.array([-1, -1, -1, 1, 1, 1], np.float64), (27, 1))
    want = np.array([(0 if i == 0 else 4) + (0 if j == 0 else 4) + (0 if k == 0 else 4) for i in (-1, 0, 1) for j in (-1, 0, 1) for k in (-1, 0, 1)],
                    np.float64)
    assert np.allclose(fn(pts, cube), want, rtol=0, atol=1e-12)
    big = np.finfo(np.float64).max
    empty = np.tile(np.array([big, big, big, -big, -big, -big], np.float64), (27, 1))
    assert np.array_equal(fn(pts, empty), np.full(27, big))


def ray_box_cases():
    """primal_ray_intersect.cpp:149-380 (ray_aabb_intersection_3D): rays from outside each face of [0,1]^3 (hit) and their
    reverses (miss), then 60 rays from the box centre (hit).  -> (rays (n,6), boxes (n,6), expected)"""
    x = np.linspace(0.0, 1.0, 3)
    rays, want = [], []
    faces = [((None, None, -1.0), (0, 0, 1)), ((None, None, 2.0), (0, 0, -1)), ((None, -1.0, None), (0, 1, 0)),
             ((None, 2.0, None), (0, -1, 0)), ((-1.0, None, None), (1, 0, 0)), ((2.0, None, None), (-1, 0, 0))]
    for origin, n in faces:
        for a in x:
            for b in x:
                free = iter((a, b))
                o = [next(free) if c is None else c for c in origin]
                rays.append(o + list(map(float, n))), want.append(True)
                rays.append(o + [-float(c) for c in n]), want.append(False)
    for ang in np.linspace(0.0, 360.0, 20):
        t = ang * np.pi / 180.0
        c, s = np.cos(t), np.sin(t)
        # Rx e1 = e1, Ry e2 = e2, Rz e3 = e3 (the reference multiplies each rotation with its own axis)
        for n in ([1.0, 0.0, 0.0], [0.0, 1.0, 0.0], [0.0, 0.0, 1.0], [0.0, c, s], [s, 0.0, c], [c, s, 0.0]):
            rays.append([0.5, 0.5, 0.5] + n), want.append(True)
    rays = np.asarray(rays, np.float64)
    boxes = np.tile(np.array([0, 0, 0, 1, 1, 1], np.float64), (len(rays), 1))
    return rays, boxes, np.asarray(want, bool)


def check_ray_box(fn):
    """fn(rays, boxes, tol) -> bool per ray (the Ray constructor's normalisation applied)"""
    rays, boxes, want = ray_box_cases()
    assert np.array_equal(fn(rays, boxes, 1e-9), want)


def check_box_scale(fn):
    """primal_boundingbox.cpp:523-569 (bb_scale); fn(boxes, scale) -> scaled boxes"""
    b = np.array([[1, 1, 1, 3, 3, 3]], np.float64)
    assert np.array_equal(fn(b, 1.5), [[.5, .5, .5, 3.5, 3.5, 3.5]])
    assert np.array_equal(fn(b, 0.5), [[1.5, 1.5, 1.5, 2.5, 2.5, 2.5]])
    assert np.array_equal(fn(b, 0.0), [[2, 2, 2, 2, 2, 2]])
    assert np.array_equal(fn(b, -1.0), b)
    big = np.finfo(np.float64).max
    inv = np.array([[big, big, big, -big, -big, -big]], np.float64)
    assert np.array_equal(fn(inv, 1.5), inv)


def sliver_triangle_cloud(n, seed=5):
    """random point / triangle pairs that stress closest_point's region tests: slivers (area down to 1e-13), needles,
    coincident vertices, query points on vertices / edges / within 1e-12 of them"""
    rng = np.random.default_rng(seed)
    A = rng.uniform(-1, 1, (n, 3))
    e1 = rng.uniform(-1, 1, (n, 3))
    e2 = rng.uniform(-1, 1, (n, 3))
    kind = rng.integers(0, 8, n)
    s = 10.0 ** rng.uniform(-14, -1, (n, 1))
    B = A + e1
    C = A + e2
    C = np.where((kind == 1)[:, None], A + e1 * rng.uniform(0, 1, (n, 1)) + s * e2, C)     # sliver: C almost on AB
    B = np.where((kind == 2)[:, None], A + s * e1, B)                                       # needle: B almost at A
    C = np.where((kind == 3)[:, None], B, C)                                                # B == C
    B = np.where((kind == 4)[:, None], A, B)                                                # A == B
    C = np.where((kind == 5)[:, None], A, C)                                                # A == C
    tri = np.stack([A, B, C], axis=1)
    w = rng.dirichlet([1, 1, 1], n)
    onplane = (w[:, :, None] * tri).sum(axis=1)
    nrm = np.cross(B - A, C - A)
    pk = rng.integers(0, 6, n)
    q = onplane + rng.uniform(-1, 1, (n, 1)) * nrm                                          # above the face
    q = np.where((pk == 1)[:, None], tri[np.arange(n), rng.integers(0, 3, n)], q)          # exactly a vertex
    t = rng.uniform(0, 1, (n, 1))
    q = np.where((pk == 2)[:, None], A + t * (B - A), q)                                    # on edge AB
    q = np.where((pk == 3)[:, None], B + t * (C - B) + 1e-12 * rng.uniform(-1, 1, (n, 3)), q)  # within 1e-12 of BC
    q = np.where((pk == 4)[:, None], rng.uniform(-3, 3, (n, 3)), q)                         # anywhere
    q = np.where((pk == 5)[:, None], A + 1e-13 * rng.uniform(-1, 1, (n, 3)), q)             # within 1e-13 of A
    return np.ascontiguousarray(q), np.ascontiguousarray(tri.reshape(n, 9))


def vtk_boxes(ndims):
    """a small tree for the writeVtkFile golden: 7 boxes with awkward decimals (exercises the stream formatting)"""
    rng = np.random.default_rng(17 + ndims)
    lo = rng.uniform(-3.0, 3.0, (7, ndims))
    ext = rng.uniform(0.01, 1.5, (7, ndims))
    b = np.concatenate([lo, lo + ext], axis=1)
    b[2] = np.concatenate([np.full(ndims, 1.0 / 3.0), np.full(ndims, 2.0 / 3.0)])
    b[5] = np.concatenate([np.full(ndims, -1.0e-7), np.full(ndims, 1234567.891)])
    return b
