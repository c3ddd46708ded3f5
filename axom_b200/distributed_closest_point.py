"""Host-side mirror of axom::quest::DistributedClosestPoint (quest/DistributedClosestPoint.hpp:56-215,
quest/detail/DistributedClosestPointImpl.hpp) over the C ABI (axb_dcp_*) and torch.distributed.

What the reference does: every rank owns some object POINTS (a point cloud, possibly several domains) and some query
points.  A query block travels round the ring of ranks (owner first, then owner+1, ...; ranks whose object bounding
box is farther than the distance threshold from the block's bounding box are skipped), and each rank overwrites an
entry of cp_rank / cp_index / cp_domain_index / cp_coords / cp_distance only if it holds a STRICTLY nearer object point
(DistributedClosestPointImpl.hpp:737-880, :905-1079).  The result is, for every query, the nearest object point of
the whole machine; among equidistant points the first rank in ring order from the owner wins, and within a rank the
first point in the BVH's traversal order.

On one NVSwitch box the ring of Conduit messages is replaced by collectives.  All query blocks and all object
bounding boxes are all-gathered.  Phase 1: every query is searched, unbounded, by the rank whose object box centre is
nearest to it; one MIN all-reduce turns the distances found into an upper bound for everybody.  Phase 2: every other rank
searches a query only if its own object box is within that bound (and within the distance threshold), and only for a
point at least as near as the bound (axb_dcp_compute_bounded_closest_points) -- the ring prunes with the same two
tests, per block instead of per query.  Phase 3: three all-reduces over all queries select the winner exactly as the
ring would --
  MIN  over the squared distances (recomputed from cp_coords with the reference's own expression, so equal values
       are bit-equal),
  MIN  over the ring position (rank - home) mod N of the ranks that attain that minimum,
  SUM  of the winner's payload as integer bit patterns (everyone else contributes zeros; exact, keeps -0.0).
Which point a rank reports (its nearest, ties by traversal order) does not depend on the bound, a bound only prunes, and
every rank that holds a point at the global minimum still reports it, so the result is identical to the reference's
ring, ties included.  With world size 1 no collective is issued.

The mint / Conduit blueprint nodes of the reference are reduced to arrays: the object mesh is a list of domains
(coords (n_i, D) interleaved, optional state/domain_id), the query mesh is coords (n, D).
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import MEM_DEVICE, MEM_HOST, check

import os
import time

OUTPUT_FIELDS = ("cp_rank", "cp_index", "cp_distance", "cp_coords", "cp_domain_index")


class _Phases:
    """AXB_DCP_TIMING=1: wall time per phase of computeClosestPoints (synchronising), printed by rank 0"""

    def __init__(self, dev):
        self.on = os.environ.get("AXB_DCP_TIMING") == "1"
        self.dev, self.t, self.acc = dev, None, {}

    def mark(self, name):
        if not self.on:
            return
        import torch
        if self.dev.type == "cuda":
            torch.cuda.synchronize(self.dev)
        now = time.perf_counter()
        if self.t is not None:
            self.acc[self._name] = self.acc.get(self._name, 0.0) + (now - self.t) * 1e3
        self.t, self._name = now, name

    def report(self, rank):
        self.mark("end")
        if self.on and rank == 0:
            print("[dcp phases ms]", {k: round(v, 1) for k, v in self.acc.items()}, flush=True)

_DBL_MAX = float(np.finfo(np.float64).max)


class _GpuBackend:
    """the product path: one axb_dcp handle on this rank's GPU; state lives in torch CUDA tensors"""

    def __init__(self, ndims, device):
        self._L = _lib.lib()
        self.ndims, self.device = ndims, device
        h = C.c_void_p()
        check(self._L.axb_dcp_create(C.byref(h), ndims, device))
        self._h = h

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                self._L.axb_dcp_destroy(self._h)
                self._h = None
        except Exception:
            pass

    def set_object_points(self, coords, domain_ids):
        c = np.ascontiguousarray(coords, np.float64).reshape(-1, self.ndims)
        d = np.ascontiguousarray(domain_ids, np.int32)
        check(self._L.axb_dcp_set_object_points(self._h, c.ctypes.data, d.ctypes.data, c.shape[0], MEM_HOST))

    def generate_bvh_tree(self):
        check(self._L.axb_dcp_generate_bvh_tree(self._h))

    def set_sq_threshold(self, t):
        check(self._L.axb_dcp_set_squared_distance_threshold(self._h, float(t)))

    def set_mode(self, mode):
        check(self._L.axb_dcp_set_mode(self._h, int(mode)))

    def tensor_device(self):
        import torch
        return torch.device("cuda", self.device)

    def object_bounds(self):
        from .bvh import BVH
        b = C.c_void_p()
        check(self._L.axb_dcp_get_bvh(self._h, C.byref(b)))
        bvh = BVH(self.ndims, self.device, _borrowed=b)
        if not bvh.isInitialized():
            return np.full(self.ndims, _DBL_MAX), np.full(self.ndims, -_DBL_MAX)
        return bvh.getBounds()

    def compute_bounded(self, rank, q, bound_sq):
        """first-visit search for points with squared distance <= bound_sq (ties included); q, bound_sq CUDA tensors"""
        import torch
        torch.cuda.current_stream(q.device).synchronize()
        n, dev = q.shape[0], q.device
        st = {"cp_index": torch.empty(n, dtype=torch.int32, device=dev), "cp_domain_index": torch.empty(n, dtype=torch.int32, device=dev),
              "cp_rank": torch.empty(n, dtype=torch.int32, device=dev), "cp_coords": torch.empty((n, self.ndims), dtype=torch.float64, device=dev),
              "cp_distance": torch.empty(n, dtype=torch.float64, device=dev)}
        if n:
            check(self._L.axb_dcp_compute_bounded_closest_points(self._h, int(rank), q.data_ptr(), n, bound_sq.data_ptr(), st["cp_index"].data_ptr(),
                                                                 st["cp_domain_index"].data_ptr(), st["cp_rank"].data_ptr(),
                                                                 st["cp_coords"].data_ptr(), st["cp_distance"].data_ptr()))
        return st

    def compute_all(self, comm, q):
        """computeClosestPoints as a whole, inside the library (axb_dcp_compute_closest_points): q is this rank's (n, D)
        float64 CUDA tensor, comm an axom_b200.comm.Comm or None (one rank)"""
        import torch
        torch.cuda.current_stream(q.device).synchronize()
        n, dev = q.shape[0], q.device
        st = {"cp_index": torch.empty(n, dtype=torch.int32, device=dev), "cp_domain_index": torch.empty(n, dtype=torch.int32, device=dev),
              "cp_rank": torch.empty(n, dtype=torch.int32, device=dev), "cp_coords": torch.empty((n, self.ndims), dtype=torch.float64, device=dev),
              "cp_distance": torch.empty(n, dtype=torch.float64, device=dev)}
        check(self._L.axb_dcp_compute_closest_points(self._h, comm._h if comm is not None else None, q.data_ptr() if n else None, n, MEM_DEVICE,
                                                     st["cp_index"].data_ptr(), st["cp_domain_index"].data_ptr(), st["cp_rank"].data_ptr(),
                                                     st["cp_coords"].data_ptr(), st["cp_distance"].data_ptr()))
        return st

    def phase_ms(self, name):
        b = C.c_void_p()
        check(self._L.axb_dcp_get_bvh(self._h, C.byref(b)))
        v = C.c_double()
        check(self._L.axb_bvh_get_phase_ms(b, name.encode(), C.byref(v)))
        return v.value

    def set_profiling(self, on):
        b = C.c_void_p()
        check(self._L.axb_dcp_get_bvh(self._h, C.byref(b)))
        check(self._L.axb_bvh_set_profiling(b, int(bool(on))))

    def compute_local(self, rank, q, state=None):
        """q: (n, D) float64 CUDA tensor.  state None = is_first.  Returns the state dict (updated in place)."""
        import torch
        # the library works on its own stream: whatever torch still has in flight for q (a gather, a slice copy)
        # must be complete before the kernel reads it
        torch.cuda.current_stream(q.device).synchronize()
        n = q.shape[0]
        first = state is None
        if first:
            dev = q.device
            state = {"cp_index": torch.empty(n, dtype=torch.int32, device=dev), "cp_domain_index": torch.empty(n, dtype=torch.int32, device=dev),
                     "cp_rank": torch.empty(n, dtype=torch.int32, device=dev), "cp_coords": torch.empty((n, self.ndims), dtype=torch.float64, device=dev),
                     "cp_distance": torch.empty(n, dtype=torch.float64, device=dev)}
        if n:
            check(self._L.axb_dcp_compute_local_closest_points(self._h, int(rank), q.data_ptr(), n, int(first), state["cp_index"].data_ptr(),
                                                               state["cp_domain_index"].data_ptr(), state["cp_rank"].data_ptr(),
                                                               state["cp_coords"].data_ptr(), state["cp_distance"].data_ptr(), MEM_DEVICE))
        return state


class DistributedClosestPoint:
    def __init__(self, ndims=3, device=0, backend=None):
        """backend: an object with the _GpuBackend interface (tests inject a CPU one to run the host logic under gloo)"""
        self.ndims = ndims
        self._b = backend if backend is not None else _GpuBackend(ndims, device)
        self._sq_threshold = _DBL_MAX  # DistributedClosestPoint.cpp:35
        self._outputs = {f: True for f in OUTPUT_FIELDS}
        self._tree = False
        self._have_mesh = False
        self._comm = None

    def setComm(self, comm):
        """the axom_b200.comm.Comm the exchange runs on (the reference takes an MPI_Comm, DistributedClosestPoint.hpp:95-100).
        Without one, a communicator is made from the initialised torch.distributed group at the first query."""
        self._comm = comm

    # ---- configuration (DistributedClosestPoint.hpp:66-118) ----
    def setDistanceThreshold(self, threshold):
        if threshold < 0.0:
            raise ValueError("Distance threshold must be non-negative.")
        self._sq_threshold = float(threshold) * float(threshold)  # DistributedClosestPoint.cpp:126-130

    def setOutput(self, field, on):
        if field not in self._outputs:
            raise ValueError("Invalid field '%s' should be one of these: %s" % (field, ", ".join(OUTPUT_FIELDS)))
        self._outputs[field] = bool(on)

    def setObjectMesh(self, domains):
        """domains: list of coords arrays (n_i, D), or of (coords, domain_id) pairs (state/domain_id); a single array is one
        domain.  Points are flattened in domain order, domain d gets id d unless one is given (importObjectPoints :590-649)."""
        if isinstance(domains, np.ndarray):
            domains = [domains]
        coords, ids = [], []
        for d, dom in enumerate(domains):
            c, did = (dom if isinstance(dom, tuple) else (dom, d))
            c = np.ascontiguousarray(c, np.float64).reshape(-1, self.ndims)
            coords.append(c)
            ids.append(np.full(c.shape[0], did, np.int32))
        c = np.concatenate(coords) if coords else np.empty((0, self.ndims))
        i = np.concatenate(ids) if ids else np.empty(0, np.int32)
        self._b.set_object_points(c, i)
        self._have_mesh, self._tree = True, False

    def generateBVHTree(self):
        if not self._have_mesh:
            raise RuntimeError("Users must set the object mesh before generating the BVH tree")
        self._b.generate_bvh_tree()
        self._tree = True
        return True

    # ---- the query (computeClosestPoints, DistributedClosestPointImpl.hpp:737-851) ----
    def computeClosestPoints(self, query_coords):
        """query_coords: (n, D) float64, numpy or torch (this rank's query points; n may be 0).
        Returns {field: tensor} for the enabled outputs, on the backend's device."""
        import torch
        import torch.distributed as dist
        if not self._tree:
            raise RuntimeError("BVH tree must be initialized before calling 'computeClosestPoints")
        self._b.set_sq_threshold(self._sq_threshold)
        dev = self._b.tensor_device()
        q = torch.as_tensor(query_coords, dtype=torch.float64).reshape(-1, self.ndims).to(dev).contiguous()
        world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
        rank = dist.get_rank() if world > 1 else 0
        if isinstance(self._b, _GpuBackend):
            # the product path: the whole exchange runs inside the library, over its own NCCL communicator
            if self._comm is None and world > 1:
                from .comm import Comm
                self._comm = Comm.from_torch(self._b.device)
            return self._select(self._b.compute_all(self._comm, q))
        # An injected backend (the CPU tests run a host per-rank step under gloo): the same protocol written with
        # torch.distributed, kept as the executable description of what csrc/comm.cuh does on the device.
        if world == 1:
            return self._select(self._b.compute_local(rank, q))

        ph = _Phases(dev)
        ph.mark("gather")
        D = self.ndims
        # ---- every rank gets all query blocks and all object bounding boxes ----
        counts = torch.zeros(world, dtype=torch.int64, device=dev)
        counts[rank] = q.shape[0]
        dist.all_reduce(counts, op=dist.ReduceOp.SUM)
        counts = [int(c) for c in counts.tolist()]
        nmax = max(max(counts), 1)
        padded = torch.zeros((nmax, D), dtype=torch.float64, device=dev)
        padded[:q.shape[0]] = q
        blocks = [torch.empty_like(padded) for _ in range(world)]
        dist.all_gather(blocks, padded)
        Q = torch.cat([blocks[o][:counts[o]] for o in range(world)]).contiguous()  # all queries, grouped by home rank
        home = torch.cat([torch.full((counts[o],), o, dtype=torch.int64, device=dev) for o in range(world)])
        ntot = Q.shape[0]
        first = sum(counts[:rank])  # this rank's own queries are Q[first : first + counts[rank]]
        olo, ohi = self._b.object_bounds()
        bb = torch.stack([torch.as_tensor(olo, dtype=torch.float64, device=dev), torch.as_tensor(ohi, dtype=torch.float64, device=dev)])
        bbs = [torch.empty_like(bb) for _ in range(world)]
        dist.all_gather(bbs, bb)

        # squared distance from every query to every rank's object box (primal::squared_distance(Point, BoundingBox);
        # an empty rank is infinitely far).  Only used to order and to prune the searches, never in a result.
        INF = float("inf")
        lo_m, hi_m = bbs[rank][0], bbs[rank][1]
        if bool((lo_m > hi_m).any()):
            mine = torch.full((ntot,), INF, dtype=torch.float64, device=dev)
        else:
            gap = torch.clamp(torch.maximum(lo_m - Q, Q - hi_m), min=0.0)
            mine = (gap * gap).sum(dim=1)
        # who searches a query first: the rank whose object box CENTRE is nearest (box distances tie at 0 wherever boxes
        # overlap, which would pile those queries on the lowest rank).  Any choice is correct; this one is balanced.
        cend = torch.empty((world, ntot), dtype=torch.float64, device=dev)
        for r in range(world):
            lo_r, hi_r = bbs[r][0], bbs[r][1]
            if bool((lo_r > hi_r).any()):
                cend[r] = INF
            else:
                dc = Q - 0.5 * (lo_r + hi_r)
                cend[r] = (dc * dc).sum(dim=1)
        nearest_rank = torch.argmin(cend, dim=0)
        del cend

        def scatter_state(full, idx, st):
            for k in ("cp_index", "cp_domain_index", "cp_rank", "cp_coords", "cp_distance"):
                full[k][idx] = st[k]

        snan = torch.tensor([0x7ff4000000000000], dtype=torch.int64, device=dev).view(torch.float64)[0]
        cand = {"cp_index": torch.full((ntot,), -1, dtype=torch.int32, device=dev),
                "cp_domain_index": torch.full((ntot,), -1, dtype=torch.int32, device=dev),
                "cp_rank": torch.full((ntot,), -1, dtype=torch.int32, device=dev),
                "cp_coords": snan.expand(ntot, D).clone(), "cp_distance": snan.expand(ntot).clone()}

        def sq_of(st, qq):
            v = st["cp_coords"] - qq
            sq = torch.zeros(qq.shape[0], dtype=torch.float64, device=dev)
            for d in range(D):  # squared_distance(qpt, query_pos): += in order, separately rounded
                sq = sq + v[:, d] * v[:, d]
            return torch.where(st["cp_rank"] >= 0, sq, torch.full_like(sq, INF))

        # ---- phase 1: each query is searched, unbounded, by the rank whose object box centre is nearest; the distance
        # found there is an upper bound for everybody else (one MIN all-reduce) ----
        ph.mark("first search")
        idx1 = torch.nonzero((nearest_rank == rank) & (mine <= self._sq_threshold)).reshape(-1)
        bound = torch.full((ntot,), INF, dtype=torch.float64, device=dev)
        if idx1.numel():
            q1 = Q[idx1].contiguous()
            st1 = self._b.compute_local(rank, q1)
            scatter_state(cand, idx1, st1)
            bound[idx1] = sq_of(st1, q1)
        ph.mark("bound all-reduce")
        dist.all_reduce(bound, op=dist.ReduceOp.MIN)

        # ---- phase 2: the other ranks look only for a point at least as near as the bound, and only where their
        # object box is that near at all (the reference's ring prunes with the same two tests, per block instead of
        # per query: DistributedClosestPointImpl.hpp:762-775, :859-878, :1037-1040) ----
        ph.mark("bounded search")
        limit = torch.minimum(bound, torch.full_like(bound, self._sq_threshold))
        idx2 = torch.nonzero((nearest_rank != rank) & (mine <= limit)).reshape(-1)
        if idx2.numel():
            q2 = Q[idx2].contiguous()
            st2 = self._b.compute_bounded(rank, q2, bound[idx2].contiguous())
            scatter_state(cand, idx2, st2)

        # ---- phase 3: the winner of every query, exactly as the ring would pick it ----
        ph.mark("combine")
        valid = cand["cp_rank"] >= 0
        sq = sq_of(cand, Q)
        smin = sq.clone()
        dist.all_reduce(smin, op=dist.ReduceOp.MIN)
        pos = torch.where(valid & (sq == smin), (rank - home) % world, torch.full_like(home, world))
        win = pos.clone()
        dist.all_reduce(win, op=dist.ReduceOp.MIN)
        sel = valid & (pos == win)
        zero = torch.zeros(ntot, dtype=torch.int64, device=dev)
        cols = [torch.where(sel, cand["cp_index"].to(torch.int64), zero), torch.where(sel, cand["cp_domain_index"].to(torch.int64), zero),
                torch.where(sel, cand["cp_rank"].to(torch.int64), zero), torch.where(sel, cand["cp_distance"].view(torch.int64), zero)]
        bits = cand["cp_coords"].contiguous().view(torch.int64)
        cols += [torch.where(sel, bits[:, d], zero) for d in range(D)]
        payload = torch.stack(cols, dim=1).contiguous()
        dist.all_reduce(payload, op=dist.ReduceOp.SUM)
        n = counts[rank]
        pl, found = payload[first:first + n], (win[first:first + n] < world)
        neg = torch.full((n,), -1, dtype=torch.int64, device=dev)
        result = {
            "cp_index": torch.where(found, pl[:, 0], neg).to(torch.int32),
            "cp_domain_index": torch.where(found, pl[:, 1], neg).to(torch.int32),
            "cp_rank": torch.where(found, pl[:, 2], neg).to(torch.int32),
            "cp_distance": torch.where(found, pl[:, 3].contiguous().view(torch.float64), snan.expand(n)),
            "cp_coords": torch.where(found[:, None], pl[:, 4:].contiguous().view(torch.float64), snan.expand(n, D)),
        }
        ph.report(rank)
        return self._select(result)

    def _select(self, st):
        return {f: st[f] for f in OUTPUT_FIELDS if self._outputs[f]}
