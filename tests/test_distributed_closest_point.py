"""quest::DistributedClosestPoint.

CPU part (no GPU): the oracle's per-rank step against the real reference BVH traversal (ties, threshold, empty ranks,
2-D); the host logic of axom_b200.DistributedClosestPoint -- all-gather, pruning test, MIN / ring-position / payload
all-reduces -- under gloo with world sizes 2 and 3, against the reference's sequential ring.
GPU part: axb_dcp_* through the C ABI, chained round a ring of handles exactly like the reference's ring of ranks."""
import os

import numpy as np
import pytest

_DBL_MAX = float(np.finfo(np.float64).max)


def _cloud_parts(nd, seed=2):
    """object points of 4 ranks: a big cloud, an EMPTY rank, a single point, a cloud sharing 50 exact duplicates with
    rank 0 (cross-rank ties); rank 0 has two domains"""
    rng = np.random.default_rng(seed)
    parts = [rng.random((3000, nd)), rng.random((0, nd)), rng.random((1, nd)), rng.random((2500, nd))]
    parts[3][:50] = parts[0][:50]
    parts[0][100:110] = parts[0][100]  # duplicates inside one rank
    q = rng.random((2000, nd)) * 1.2 - 0.1
    q[:20] = parts[0][:20]
    q[20:30] = parts[0][100]
    return parts, q


def _ring(make_rank, parts, q, owner, sq_th):
    """the reference's ring for one query block: owner first, then owner+1, ... (DistributedClosestPointImpl.hpp:737-880)"""
    n = len(parts)
    st = None
    for k in range(n):
        r = (owner + k) % n
        st = make_rank(r).compute_local(r, q, st, sq_th)
    return st


def _same_state(a, b):
    return all(np.array_equal(np.asarray(a[k]), np.asarray(b[k]), equal_nan=True) for k in ("cp_index", "cp_domain_index", "cp_rank", "cp_coords", "cp_distance"))


@pytest.mark.parametrize("nd", [3, 2])
def test_oracle_local_step_equals_reference_traversal(oracle, have_ref, nd):
    if not have_ref:
        pytest.skip("oracle/_ref/libaxom_ref.so not present (needs /root/reference to build)")
    parts, q = _cloud_parts(nd)
    doms = [np.where(np.arange(len(p)) < len(p) // 2, 10 * i, 10 * i + 1).astype(np.int32) for i, p in enumerate(parts)]
    for th in (_DBL_MAX, 0.05 ** 2, 0.0):
        res = {}
        for kind in ("port", "reference"):
            ranks = [oracle.DistributedClosestPointRank(p, d, nd, kind) for p, d in zip(parts, doms)]
            res[kind] = [_ring(lambda r: ranks[r], parts, q, owner, th) for owner in range(4)]
        for a, b in zip(res["port"], res["reference"]):
            assert _same_state(a, b)
        if th == _DBL_MAX:
            assert (res["port"][0]["cp_rank"] >= 0).all()
            assert (res["port"][0]["cp_rank"][:20] == 0).all() and (res["port"][3]["cp_rank"][:20] == 3).all()  # ring-order ties


class _OracleBackend:
    """CPU stand-in for the GPU handle (tests only): the oracle's per-rank step behind the backend interface"""

    def __init__(self, oracle, nd):
        self.O, self.nd, self.rank_obj, self.sq = oracle, nd, None, _DBL_MAX

    def set_object_points(self, coords, ids):
        self.coords, self.ids = coords, ids

    def generate_bvh_tree(self):
        self.rank_obj = self.O.DistributedClosestPointRank(self.coords, self.ids, self.nd)

    def set_sq_threshold(self, t):
        self.sq = t

    def tensor_device(self):
        import torch
        return torch.device("cpu")

    def object_bounds(self):
        if len(self.coords) == 0:
            return np.full(self.nd, _DBL_MAX), np.full(self.nd, -_DBL_MAX)
        b = self.O.Bvh(np.concatenate([self.coords, self.coords], axis=1), ndims=self.nd).arrays()["bounds"]
        return b[:self.nd], b[self.nd:]

    def compute_bounded(self, rank, q, bound_sq):
        import torch
        st = self.rank_obj.compute_local(rank, q.numpy(), None, self.sq)
        v = st["cp_coords"] - q.numpy()
        sq = np.zeros(len(v))
        for d in range(self.nd):
            sq = sq + v[:, d] * v[:, d]
        drop = ~((st["cp_rank"] >= 0) & (sq <= bound_sq.numpy()))
        snan = np.frombuffer(np.array([0x7ff4000000000000], np.int64).tobytes(), np.float64)[0]
        for k in ("cp_index", "cp_domain_index", "cp_rank"):
            st[k][drop] = -1
        st["cp_coords"][drop] = snan
        st["cp_distance"][drop] = snan
        return {k: torch.from_numpy(v) for k, v in st.items()}

    def compute_local(self, rank, q, state=None):
        import torch
        if state is not None:  # preset from the owner: updated in place, like the xferDom arrays
            state = {k: np.ascontiguousarray(v.numpy()) for k, v in state.items()}
        st = self.rank_obj.compute_local(rank, q.numpy(), state, self.sq)
        return {k: torch.from_numpy(v) for k, v in st.items()}


def _dcp_worker(rank, world, port, nd, threshold, out):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from axom_b200 import DistributedClosestPoint
        from oracle import oracle as O
        parts, q = _cloud_parts(nd)
        parts = parts[:world] if world < 4 else parts
        if world == 2:
            parts = [parts[0], _cloud_parts(nd)[0][3]]
        doms = [np.where(np.arange(len(p)) < len(p) // 2, 10 * i, 10 * i + 1).astype(np.int32) for i, p in enumerate(parts)]
        # query blocks of different sizes, one of them empty
        cuts = np.linspace(0, len(q), world + 1).astype(int)
        if world == 3:
            cuts = np.array([0, 0, 1200, len(q)])
        myq = q[cuts[rank]:cuts[rank + 1]]
        d = DistributedClosestPoint(nd, backend=_OracleBackend(O, nd))
        if threshold is not None:
            d.setDistanceThreshold(threshold)
        half = len(parts[rank]) // 2
        d.setObjectMesh([(parts[rank][:half], 10 * rank), (parts[rank][half:], 10 * rank + 1)])
        d.generateBVHTree()
        got = d.computeClosestPoints(myq)
        ranks = [O.DistributedClosestPointRank(p, dd, nd) for p, dd in zip(parts, doms)]
        want = _ring(lambda r: ranks[r], parts, myq, rank, _DBL_MAX if threshold is None else threshold * threshold)
        ok = _same_state({k: v.numpy() for k, v in got.items()}, want)
        out.put((rank, bool(ok), int((want["cp_rank"] >= 0).sum()), len(myq)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,nd,threshold", [(2, 3, None), (3, 3, 0.04), (3, 2, None)])
def test_host_logic_equals_reference_ring_under_gloo(world, nd, threshold):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = 31500 + (os.getpid() % 2000) + 7 * world + nd
    procs = [ctx.Process(target=_dcp_worker, args=(r, world, port, nd, threshold, out)) for r in range(world)]
    for p in procs:
        p.start()
    res = [out.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok, _, _ in res), res
    assert sum(n for _, _, _, n in res) == 2000
    if threshold is not None:
        assert any(found < n for _, _, found, n in res if n)  # the threshold leaves some queries without a closest point


def test_api_argument_checks():
    from axom_b200 import DistributedClosestPoint

    class _Null:
        def set_object_points(self, c, i):
            pass
    d = DistributedClosestPoint(3, backend=_Null())
    with pytest.raises(ValueError):
        d.setDistanceThreshold(-1.0)
    with pytest.raises(ValueError):
        d.setOutput("cp_nonsense", True)
    with pytest.raises(RuntimeError):
        d.generateBVHTree()
    with pytest.raises(RuntimeError):
        d.computeClosestPoints(np.zeros((1, 3)))


@pytest.mark.gpu
@pytest.mark.parametrize("mode", [1, 0])
@pytest.mark.parametrize("nd", [3, 2])
def test_gpu_ring_of_handles_matches_oracle(oracle, nd, mode):
    """four axb_dcp handles on one GPU play the four ranks; the state arrays go round the ring through the C ABI
    (is_first on the owner, then in-place updates), host memspace and device memspace"""
    import ctypes as C
    import torch
    from axom_b200 import _lib
    from axom_b200._lib import MEM_DEVICE, MEM_HOST, check
    L = _lib.lib()
    parts, q = _cloud_parts(nd)
    q = np.ascontiguousarray(np.concatenate([q, q[:3000] * 0.5 + 0.25]))  # > 4096 queries: the Morton-sorted path
    doms = [np.where(np.arange(len(p)) < len(p) // 2, 10 * i, 10 * i + 1).astype(np.int32) for i, p in enumerate(parts)]
    oranks = [oracle.DistributedClosestPointRank(p, d, nd) for p, d in zip(parts, doms)]
    handles = []
    for p, d in zip(parts, doms):
        h = C.c_void_p()
        check(L.axb_dcp_create(C.byref(h), nd, 0))
        pc = np.ascontiguousarray(p)
        check(L.axb_dcp_set_object_points(h, pc.ctypes.data, d.ctypes.data, len(p), MEM_HOST))
        check(L.axb_dcp_generate_bvh_tree(h))
        check(L.axb_dcp_set_mode(h, mode))
        handles.append(h)
    try:
        for th in (_DBL_MAX, 0.05 ** 2):
            for h in handles:
                check(L.axb_dcp_set_squared_distance_threshold(h, th))
            for owner in range(4):
                want = _ring(lambda r: oranks[r], parts, q, owner, th)
                n = len(q)
                # host arrays
                st = {"cp_index": np.empty(n, np.int32), "cp_domain_index": np.empty(n, np.int32), "cp_rank": np.empty(n, np.int32),
                      "cp_coords": np.empty((n, nd)), "cp_distance": np.empty(n)}
                for k in range(4):
                    r = (owner + k) % 4
                    check(L.axb_dcp_compute_local_closest_points(handles[r], r, q.ctypes.data, n, int(k == 0), st["cp_index"].ctypes.data,
                                                                 st["cp_domain_index"].ctypes.data, st["cp_rank"].ctypes.data,
                                                                 st["cp_coords"].ctypes.data, st["cp_distance"].ctypes.data, MEM_HOST))
                assert _same_state(st, want), (nd, th, owner)
                # device arrays, no cp_distance
                qd = torch.from_numpy(q).cuda()
                sd = {"cp_index": torch.empty(n, dtype=torch.int32, device="cuda"), "cp_domain_index": torch.empty(n, dtype=torch.int32, device="cuda"),
                      "cp_rank": torch.empty(n, dtype=torch.int32, device="cuda"), "cp_coords": torch.empty((n, nd), dtype=torch.float64, device="cuda")}
                for k in range(4):
                    r = (owner + k) % 4
                    check(L.axb_dcp_compute_local_closest_points(handles[r], r, qd.data_ptr(), n, int(k == 0), sd["cp_index"].data_ptr(),
                                                                 sd["cp_domain_index"].data_ptr(), sd["cp_rank"].data_ptr(),
                                                                 sd["cp_coords"].data_ptr(), None, MEM_DEVICE))
                for f in ("cp_index", "cp_domain_index", "cp_rank", "cp_coords"):
                    assert np.array_equal(sd[f].cpu().numpy(), want[f], equal_nan=True), (f, nd, th, owner)
    finally:
        for h in handles:
            L.axb_dcp_destroy(h)


@pytest.mark.gpu
def test_gpu_single_rank_class_and_outputs(oracle):
    from axom_b200 import DistributedClosestPoint
    parts, q = _cloud_parts(3)
    d = DistributedClosestPoint(3)
    d.setObjectMesh([parts[0][:1000], (parts[0][1000:], 42)])
    d.generateBVHTree()
    d.setOutput("cp_domain_index", False)
    got = d.computeClosestPoints(q)
    assert "cp_domain_index" not in got and set(got) == {"cp_rank", "cp_index", "cp_distance", "cp_coords"}
    d.setOutput("cp_domain_index", True)
    d.setDistanceThreshold(0.03)
    got = d.computeClosestPoints(q)
    dom = np.where(np.arange(3000) < 1000, 0, 42).astype(np.int32)
    want = oracle.DistributedClosestPointRank(parts[0], dom, 3).compute_local(0, q, None, 0.03 ** 2)
    assert _same_state({k: v.cpu().numpy() for k, v in got.items()}, want)
    assert (want["cp_rank"] < 0).any() and (want["cp_rank"] >= 0).any()


@pytest.mark.gpu
@pytest.mark.parametrize("nd", [3, 2])
def test_gpu_bounded_search_keeps_exactly_the_points_within_the_bound(oracle, nd):
    """axb_dcp_compute_bounded_closest_points == the unbounded first-visit search, kept only where the squared distance
    is <= the bound (ties with the bound included)"""
    import torch
    from axom_b200.distributed_closest_point import _GpuBackend
    parts, q = _cloud_parts(nd)
    q = np.ascontiguousarray(np.concatenate([q, q[:3000] * 0.5 + 0.25]))
    b = _GpuBackend(nd, 0)
    b.set_object_points(parts[0], np.zeros(len(parts[0]), np.int32))
    b.generate_bvh_tree()
    qd = torch.from_numpy(q).cuda()
    full = b.compute_local(5, qd)
    v = full["cp_coords"] - qd
    sq = torch.zeros(len(q), dtype=torch.float64, device="cuda")
    for d in range(nd):
        sq = sq + v[:, d] * v[:, d]
    rng = np.random.default_rng(8)
    factor = torch.from_numpy(rng.choice([0.5, 1.0, 1.0, 2.0], len(q))).cuda()  # below, AT, and above the true distance
    bound = (sq * factor).contiguous()
    got = b.compute_bounded(5, qd, bound)
    keep = sq <= bound
    assert bool(keep.any()) and bool((~keep).any())
    assert torch.equal(got["cp_rank"], torch.where(keep, full["cp_rank"], torch.full_like(full["cp_rank"], -1)))
    assert torch.equal(got["cp_index"][keep], full["cp_index"][keep]) and torch.equal(got["cp_coords"][keep], full["cp_coords"][keep])
    assert torch.equal(got["cp_distance"][keep], full["cp_distance"][keep]) and bool((got["cp_index"][~keep] == -1).all())
