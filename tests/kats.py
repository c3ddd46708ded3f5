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
