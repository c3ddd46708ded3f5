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
