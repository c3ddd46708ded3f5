"""Host-side mirror of axom::quest::DistributedClosestPoint (quest/DistributedClosestPoint.hpp:56-215,
quest/detail/DistributedClosestPointImpl.hpp) over the C ABI (axb_dcp_*) and torch.distributed.

What the reference does: every rank owns some object POINTS (a point cloud, possibly several domains) and some query
points.  A query block travels round the ring of ranks (owner first, then owner+1, ...; ranks whose object bounding
box is farther than the distance threshold from the block's bounding box are skipped), and each rank overwrites an
entry of cp_rank / cp_index / cp_domain_index / cp_coords / cp_distance only if it holds a STRICTLY nearer object point
(DistributedClosestPointImpl.hpp:737-880, :905-1079).  The result is, for every query, the nearest object point of
the whole machine; among equidistant points the first rank in ring order from the owner wins, and within a rank the
first point in the BVH's traversal order.

On one NVSwitch box the ring of Conduit messages is replaced by: every rank searches its OWN block first (is_first);
the blocks are all-gathered together with the owners' answers; every other rank searches every block it is not pruned
from with the owner's answer as the preset (so, like a later rank of the ring, it reports only a strictly nearer point,
and usually prunes its whole tree at the root); then three all-reduces per block select the winner exactly as the
ring would --
  MIN  over the squared distances (recomputed from cp_coords with the reference's own expression, so equal values
       are bit-equal),
  MIN  over the ring position (rank - owner) mod N of the ranks that attain that minimum,
  SUM  of the winner's payload as integer bit patterns (everyone else contributes zeros; exact, keeps -0.0).
A rank's strictly-nearer answer does not depend on what other ranks found (a preset only prunes), and a tie with the
owner goes to the owner in the ring as well, so the result is identical to the reference's ring, ties included.  With world size 1 no collective is issued.

The mint / Conduit blueprint nodes of the reference are reduced to arrays: the object mesh is a list of domains
(coords (n_i, D) interleaved, optional state/domain_id), the query mesh is coords (n, D).
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import MEM_DEVICE, MEM_HOST, check

OUTPUT_FIELDS = ("cp_rank", "cp_index", "cp_distance", "cp_coords", "cp_domain_index")
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
        if world == 1:
            return self._select(self._b.compute_local(rank, q))

        # ---- phase A: every rank searches its OWN block first; the result is an upper bound for everyone else ----
        own = self._b.compute_local(rank, q)

        # ---- query blocks (with the owner's closest point so far) and bounding boxes of every rank ----
        D = self.ndims
        counts = torch.zeros(world, dtype=torch.int64, device=dev)
        counts[rank] = q.shape[0]
        dist.all_reduce(counts, op=dist.ReduceOp.SUM)
        counts = [int(c) for c in counts.tolist()]
        nmax = max(max(counts), 1)
        # columns: query coords | owner's cp_coords | owner's cp_rank (as float64, exact for small ints)
        padded = torch.zeros((nmax, 2 * D + 1), dtype=torch.float64, device=dev)
        padded[:q.shape[0], :D] = q
        padded[:q.shape[0], D:2 * D] = own["cp_coords"]
        padded[:q.shape[0], 2 * D] = own["cp_rank"].to(torch.float64)
        blocks = [torch.empty_like(padded) for _ in range(world)]
        dist.all_gather(blocks, padded)
        # [query lo, query hi, object lo, object hi] per rank, for the reference's pruning test (:762-775, :859-878)
        olo, ohi = self._b.object_bounds()
        bb = torch.full((4, D), _DBL_MAX, dtype=torch.float64, device=dev)
        bb[1] = -_DBL_MAX
        if q.shape[0]:
            bb[0], bb[1] = q.min(dim=0).values, q.max(dim=0).values
        bb[2], bb[3] = torch.as_tensor(olo, device=dev), torch.as_tensor(ohi, device=dev)
        bbs = [torch.empty_like(bb) for _ in range(world)]
        dist.all_gather(bbs, bb)
        bbs = [b.cpu().numpy() for b in bbs]

        def box_sqdist(alo, ahi, blo, bhi):  # primal::squared_distance(BoundingBox, BoundingBox): gap per dimension
            if np.any(alo > ahi) or np.any(blo > bhi):
                return _DBL_MAX
            gap = np.maximum(np.maximum(blo - ahi, alo - bhi), 0.0)
            return float(np.sum(gap * gap))

        result = None
        for owner in range(world):
            n = counts[owner]
            if n == 0:
                continue
            qo = blocks[owner][:n, :D].contiguous()
            # ---- phase B: the other ranks search the block with the owner's answer as the preset: like the ring's
            # later ranks they only report a STRICTLY nearer point (a tie goes to the owner, ring position 0), and a
            # good preset prunes their whole tree at the root.  A rank whose object box is farther than the threshold
            # from the block's box sits the block out (every rank evaluates this test on the same gathered boxes, so
            # the collectives below match up).
            if owner == rank:
                st = own
                valid = st["cp_rank"] >= 0
            elif box_sqdist(bbs[owner][0], bbs[owner][1], bbs[rank][2], bbs[rank][3]) <= self._sq_threshold:
                st = {"cp_index": torch.full((n,), -1, dtype=torch.int32, device=dev),
                      "cp_domain_index": torch.full((n,), -1, dtype=torch.int32, device=dev),
                      "cp_rank": blocks[owner][:n, 2 * D].to(torch.int32).contiguous(),
                      "cp_coords": blocks[owner][:n, D:2 * D].contiguous(),
                      "cp_distance": torch.zeros(n, dtype=torch.float64, device=dev)}
                st = self._b.compute_local(rank, qo, st)
                valid = st["cp_rank"] == rank
            else:
                st = None
                valid = torch.zeros(n, dtype=torch.bool, device=dev)
            if st is not None:
                v = st["cp_coords"] - qo
                sq = torch.zeros(n, dtype=torch.float64, device=dev)
                for d in range(D):  # squared_distance(qpt, query_pos): += in order, separately rounded
                    sq = sq + v[:, d] * v[:, d]
                sq = torch.where(valid, sq, torch.full_like(sq, float("inf")))
            else:
                sq = torch.full((n,), float("inf"), dtype=torch.float64, device=dev)
            smin = sq.clone()
            dist.all_reduce(smin, op=dist.ReduceOp.MIN)
            pos = torch.full((n,), world, dtype=torch.int64, device=dev)
            pos = torch.where(valid & (sq == smin), torch.full_like(pos, (rank - owner) % world), pos)
            win = pos.clone()
            dist.all_reduce(win, op=dist.ReduceOp.MIN)
            i_win = valid & (pos == win)
            payload = torch.zeros((n, 4 + D), dtype=torch.int64, device=dev)
            if st is not None:
                sel = i_win
                payload[:, 0] = torch.where(sel, st["cp_index"].to(torch.int64), payload[:, 0])
                payload[:, 1] = torch.where(sel, st["cp_domain_index"].to(torch.int64), payload[:, 1])
                payload[:, 2] = torch.where(sel, st["cp_rank"].to(torch.int64), payload[:, 2])
                payload[:, 3] = torch.where(sel, st["cp_distance"].contiguous().view(torch.int64), payload[:, 3])
                bits = st["cp_coords"].contiguous().view(torch.int64)
                for d in range(D):
                    payload[:, 4 + d] = torch.where(sel, bits[:, d], payload[:, 4 + d])
            dist.all_reduce(payload, op=dist.ReduceOp.SUM)
            if owner == rank:
                found = win < world
                snan = torch.tensor([0x7ff4000000000000], dtype=torch.int64, device=dev).view(torch.float64)[0]
                result = {
                    "cp_index": torch.where(found, payload[:, 0], torch.full_like(win, -1)).to(torch.int32),
                    "cp_domain_index": torch.where(found, payload[:, 1], torch.full_like(win, -1)).to(torch.int32),
                    "cp_rank": torch.where(found, payload[:, 2], torch.full_like(win, -1)).to(torch.int32),
                    "cp_distance": torch.where(found, payload[:, 3].view(torch.float64), snan.expand(n)),
                    "cp_coords": torch.where(found[:, None], payload[:, 4:].contiguous().view(torch.float64), snan.expand(n, D)),
                }
        if result is None:
            result = own  # this rank has no queries: empty arrays of the right types
        return self._select(result)

    def _select(self, st):
        return {f: st[f] for f in OUTPUT_FIELDS if self._outputs[f]}
