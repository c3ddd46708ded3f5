// traverse.cuh -- BVH traversal and the batched candidate queries (sm_100a).
//
// Reference path replaced:
//   lbvh::bvh_traverse                  spin/internal/linear_bvh/bvh_traverse.hpp:66-154
//   LinearBVH::findCandidatesImpl       spin/policy/LinearBVH.hpp:271-402  (count -> scan -> fill)
//   predicates of findPoints/Boxes/Rays spin/BVH.hpp:499-501, :558-560, :529-532
#pragma once
#include "common.cuh"
#include "tritri.cuh"

namespace axb
{
constexpr int32_t kBarrier = -2000000000;  // bvh_traverse.hpp:80
constexpr int kStackSize = 64;             // bvh_traverse.hpp:79

// The reference's traversal, restated over the packed node records.  Visits exactly the
// same nodes in exactly the same order as bvh_traverse (including its "park the first leaf,
// keep descending, then handle up to two leaves" shape), so per-query candidate order and
// the SignedDistance state machine evolve identically.
//   pred(box)       -> bool   "B"     (only called on valid boxes, :95-96)
//   order(L, R)     -> bool   "Comp"  true = visit the right child first
//   leaf(sorted_pos)          "A"
template <typename T, int D, class Pred, class Leaf, class Order>
__device__ __forceinline__ void traverse_reference_order(const Node<T, D>* __restrict__ nodes, Pred&& pred, Leaf&& leaf, Order&& order)
{
  int32_t todo[kStackSize];
  int sp = 0;
  todo[0] = kBarrier;
  int32_t found = 0;
  int32_t cur = 0;
  while(cur != kBarrier)
  {
    while(cur >= 0)
    {
      const Node<T, D>& nd = nodes[cur];
      Box<T, D> L = nd.box[0];
      Box<T, D> R = nd.box[1];
      const int32_t lc = nd.child[0];
      int32_t rc = nd.child[1];
      const bool inL = box_valid(L) ? pred(L) : false;
      const bool inR = box_valid(R) ? pred(R) : false;
      if(!inL && !inR)
      {
        cur = todo[sp--];
      }
      else
      {
        cur = inL ? lc : rc;
        if(inL && inR)
        {
          if(order(L, R))
          {
            const int32_t t = cur;
            cur = rc;
            rc = t;
          }
          todo[++sp] = rc;
        }
      }
      if(cur < 0 && !(found < 0))
      {
        found = cur;
        if(cur != kBarrier) cur = todo[sp--];
      }
    }
    while(found < 0 && found != kBarrier)
    {
      leaf(-found - 1);
      found = cur;
      if(cur < 0 && cur != kBarrier) cur = todo[sp--];
    }
    found = 0;
  }
}

struct NoOrder
{
  template <class B>
  __device__ __forceinline__ bool operator()(const B&, const B&) const
  {
    return false;
  }
};

//------------------------------------------------------------------------------------------
// query primitives
//------------------------------------------------------------------------------------------
// bb.contains(p): closed intervals (primal/geometry/BoundingBox.hpp:390-401)
template <typename T, int D>
struct PointQuery
{
  static constexpr int NCOMP = D;
  T p[D];
  __device__ __forceinline__ void load(const Desc<NCOMP>& d, long long i, T /*tol*/, int /*flags*/)
  {
#pragma unroll
    for(int k = 0; k < D; ++k) p[k] = ld_comp<T>(d, k, i);
  }
  __device__ __forceinline__ void center(T* c) const
  {
#pragma unroll
    for(int k = 0; k < D; ++k) c[k] = p[k];
  }
  __device__ __forceinline__ bool operator()(const Box<T, D>& bb) const
  {
    bool in = true;
#pragma unroll
    for(int k = 0; k < D; ++k) in = in && !(p[k] < bb.lo[k] || p[k] > bb.hi[k]);
    return in;
  }
};

// bb1.intersectsWith(bb2), bb1 = query (primal/operators/detail/intersect_bounding_box_impl.hpp:34-41)
template <typename T, int D>
struct BoxQuery
{
  static constexpr int NCOMP = 2 * D;
  Box<T, D> b;
  __device__ __forceinline__ void load(const Desc<NCOMP>& d, long long i, T /*tol*/, int /*flags*/) { b = load_box<T, D>(d, i); }
  __device__ __forceinline__ void center(T* c) const
  {
#pragma unroll
    for(int k = 0; k < D; ++k) c[k] = static_cast<T>(0.5 * (b.lo[k] + b.hi[k]));
  }
  __device__ __forceinline__ bool operator()(const Box<T, D>& bb) const
  {
    bool hit = true;
#pragma unroll
    for(int k = 0; k < D; ++k) hit = hit && ((b.hi[k] >= bb.lo[k]) && (b.lo[k] <= bb.hi[k]));
    return hit;
  }
};

// primal::detail::intersect_ray(R, bb, ip, TOL): per-dimension slab test with short circuit
// (primal/operators/detail/intersect_ray_impl.hpp:150-187, :321-351)
template <typename T, int D>
struct RayQuery
{
  static constexpr int NCOMP = 2 * D;
  T o[D];
  T dir[D];
  T inv[D];  // 1/dir[k]: a pure function of the ray, so hoisting it out of the box test changes nothing
  T tol;
  __device__ __forceinline__ void load(const Desc<NCOMP>& d, long long i, T tol_, int normalized)
  {
    tol = tol_;
#pragma unroll
    for(int k = 0; k < D; ++k)
    {
      o[k] = ld_comp<T>(d, k, i);
      dir[k] = ld_comp<T>(d, D + k, i);
    }
    if(!normalized)
    {
      // Ray ctor -> Vector::unitVector (primal/geometry/Ray.hpp:122-127, Vector.hpp:477-493)
      T acc = (T)0;  // Vector::dot_product accumulates in T (Vector.hpp:543-552), unitVector continues in double
#pragma unroll
      for(int k = 0; k < D; ++k) acc += dir[k] * dir[k];
      const double len2 = (double)acc;
      if(len2 >= 1e-50)
      {
        const double s = 1. / sqrt(len2);  // NumericArray::operator/= (core/NumericArray.hpp:510-514)
#pragma unroll
        for(int k = 0; k < D; ++k) dir[k] = static_cast<T>(dir[k] * s);
      }
      else
      {
        dir[0] = (T)1;
#pragma unroll
        for(int k = 1; k < D; ++k) dir[k] = (T)0;
      }
    }
#pragma unroll
    for(int k = 0; k < D; ++k) inv[k] = (T)1.0 / dir[k];
  }
  __device__ __forceinline__ void center(T* c) const
  {
#pragma unroll
    for(int k = 0; k < D; ++k) c[k] = o[k];
  }
  // where the ray ENTERS the tree's bounds (its origin if inside): rays that enter near each other walk the same nodes,
  // whereas their origins say nothing (an origin outside the bounds would be clamped to the boundary of the Morton grid
  // -- every origin beyond a corner to the same cell).  false: the ray misses the bounds and has no candidates.
  // Processing order only; the candidates come from operator() below.
  __device__ __forceinline__ bool sort_point(const T* bmin, const T* bmax, T* c) const
  {
    T t0 = (T)0, t1 = Lim<T>::max();
    bool hit = true;
#pragma unroll
    for(int k = 0; k < D; ++k)
    {
      if(dir[k] == (T)0)
        hit = hit && !(o[k] < bmin[k] || o[k] > bmax[k]);
      else
      {
        T a = (bmin[k] - o[k]) * inv[k], b = (bmax[k] - o[k]) * inv[k];
        if(a > b)
        {
          const T t = a;
          a = b;
          b = t;
        }
        t0 = a > t0 ? a : t0;
        t1 = b < t1 ? b : t1;
      }
    }
    hit = hit && t0 <= t1 * ((T)1 + (T)1e-5) + (T)1e-30;  // generous: a grazing ray is merely sorted with the hits
#pragma unroll
    for(int k = 0; k < D; ++k) c[k] = o[k] + t0 * dir[k];
    return hit;
  }
  __device__ __forceinline__ bool operator()(const Box<T, D>& bb) const
  {
    T tmin = Lim<T>::min();
    T tmax = Lim<T>::max();
#pragma unroll
    for(int k = 0; k < D; ++k)
    {
      const T diff = dir[k] - (T)0;
      if((diff < 0 ? -diff : diff) <= tol)
      {
        if(o[k] < bb.lo[k] || o[k] > bb.hi[k]) return false;
      }
      else
      {
        const T invn = inv[k];
        T t1 = (bb.lo[k] - o[k]) * invn;
        T t2 = (bb.hi[k] - o[k]) * invn;
        if(t1 > t2)
        {
          const T t = t1;
          t1 = t2;
          t2 = t;
        }
        tmin = (t1 < tmin) ? tmin : t1;  // utilities::max (core/utilities/Utilities.hpp:80-83)
        tmax = (t2 < tmax) ? t2 : tmax;  // utilities::min (:93-96)
        if(tmin > tmax) return false;
      }
    }
    return true;
  }
};

//------------------------------------------------------------------------------------------
// PASS 1: count (LinearBVH.hpp:302-321).  perm == nullptr: thread t handles query t.
//------------------------------------------------------------------------------------------
template <typename T, int D, class Query, class Filter = NoFilter>
__global__ void __launch_bounds__(256) count_kernel(const Node<T, D>* __restrict__ nodes, Desc<Query::NCOMP> prims, int nq, T tol, int flags,
                                                     const int32_t* __restrict__ perm, int32_t* __restrict__ counts,
                                                     const int32_t* __restrict__ leaf_nodes = nullptr, Filter filt = Filter())
{
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if(t >= nq) return;
  const int qi = perm ? perm[t] : t;
  Query q;
  q.load(prims, qi, tol, flags);
  int c = 0;
  if(std::is_same<Filter, NoFilter>::value)
    traverse_reference_order<T, D>(nodes, q, [&](int) { ++c; }, NoOrder {});
  else
    traverse_reference_order<T, D>(
      nodes, q, [&](int pos) { c += filt(qi, __ldg(leaf_nodes + pos)) ? 1 : 0; }, NoOrder {});
  counts[qi] = c;
}

//------------------------------------------------------------------------------------------
// PASS 2: fill (LinearBVH.hpp:346-364): candidates[offset++] = leaf_nodes[pos]
//------------------------------------------------------------------------------------------
template <typename T, int D, class Query, class Filter = NoFilter>
__global__ void __launch_bounds__(256) fill_kernel(const Node<T, D>* __restrict__ nodes, const int32_t* __restrict__ leaf_nodes,
                                                    Desc<Query::NCOMP> prims, int nq, T tol, int flags, const int32_t* __restrict__ perm,
                                                    const int32_t* __restrict__ offsets, int32_t* __restrict__ candidates,
                                                    int32_t* __restrict__ firsts = nullptr, Filter filt = Filter())
{
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if(t >= nq) return;
  const int qi = perm ? perm[t] : t;
  Query q;
  q.load(prims, qi, tol, flags);
  int off = offsets[qi];
  traverse_reference_order<T, D>(
    nodes, q,
    [&](int pos) {
      const int32_t cand = __ldg(leaf_nodes + pos);
      if(filt(qi, cand))
      {
        if(firsts) firsts[off] = qi;
        candidates[off++] = cand;
      }
    },
    NoOrder {});
}

//------------------------------------------------------------------------------------------
// SINGLE-TRAVERSAL candidate search.
//
// The reference walks the tree twice (count, then fill; LinearBVH.hpp:302-364).  Here one walk counts
// AND records every hit as a (query, rank, candidate) triple in a chunked pair buffer; after the scan a
// flat scatter kernel drops candidate k of query q at candidates[offsets[q] + k].  rank is the hit's
// position in the query's own DFS order, so the per-query ORDER of the reference is preserved.
//   * persistent warps pull queries (optionally in Morton order, `perm`) from a device cursor and a lane
//     that finishes is refilled at once;
//   * hits of the lanes that sit on a leaf in the same step are compacted by ballot + prefix into the
//     warp's current chunk; chunks (kPairChunk triples) are reserved with ONE atomic each, and only a
//     warp's last chunk can be partly filled (its slack is recorded in `unused`);
//   * if the buffer runs out the walk keeps counting and sets `overflow`; the host then falls back to
//     the classic fill traversal (fill_kernel) for this call and sizes the buffer better next time.
//------------------------------------------------------------------------------------------
constexpr int kPairChunk = 1024;
constexpr int kFindQueryChunk = 32;

struct PairBuf
{
  int4* pairs;              // .x query  .y rank  .z candidate (original box id)
  unsigned int* cursor;     // [0] next query chunk  [1] next pair chunk  [2] overflow flag
  unsigned int* unused;     // per pair chunk: slots left empty at the end (0 for full chunks)
  unsigned int max_chunks;  // capacity of `pairs` in chunks (0 = recording disabled)
};

template <typename T, int D>
__device__ __forceinline__ void load_node_boxes(const Node<T, D>* __restrict__ nodes, int32_t i, Box<T, D>& L, Box<T, D>& R, int32_t& lc,
                                                int32_t& rc)
{
  const Node<T, D>& nd = nodes[i];
  L = nd.box[0];
  R = nd.box[1];
  lc = nd.child[0];
  rc = nd.child[1];
}
// 3-D double: the 128-byte record as four 256-bit loads
template <>
__device__ __forceinline__ void load_node_boxes<double, 3>(const Node<double, 3>* __restrict__ nodes, int32_t i, Box<double, 3>& L,
                                                            Box<double, 3>& R, int32_t& lc, int32_t& rc)
{
  const D4* p = reinterpret_cast<const D4*>(nodes + i);
  const D4 a = ldg256(p), b = ldg256(p + 1), c = ldg256(p + 2), d = ldg256(p + 3);
  L.lo[0] = a.x;
  L.lo[1] = a.y;
  L.lo[2] = a.z;
  L.hi[0] = a.w;
  L.hi[1] = b.x;
  L.hi[2] = b.y;
  R.lo[0] = b.z;
  R.lo[1] = b.w;
  R.lo[2] = c.x;
  R.hi[0] = c.y;
  R.hi[1] = c.z;
  R.hi[2] = c.w;
  const long long ids = __double_as_longlong(d.x);
  lc = (int32_t)(ids & 0xffffffffll);
  rc = (int32_t)(ids >> 32);
}

// Compact traversal record for comparison-only predicates (point-in-box, box-box) on the 3-D double tree: both child
// boxes as binary32 rounded OUTWARD (lo down, hi up) + the child ids = 64 bytes = two 32-byte sectors instead of four.
// The walk is bound by L1 sector throughput of divergent node reads (a lane reads a record of its own), so bytes per
// visit are what counts.  An outward-rounded box passes every query the exact box passes (the predicates only compare
// coordinates), so inner nodes are entered conservatively; a LEAF child that passes is re-tested against its exact
// double box in the 128-byte record, so the candidates, and their order, are exactly the reference's.  A box that is
// invalid in double stays invalid unless lo and hi straddle a binary32 rounding gap; then only invalid leaves lie below
// it and the exact leaf test rejects them.  Rays keep the 128-byte records: the slab test does arithmetic on the box.
struct alignas(32) FNode
{
  float lo0[3], hi0[3], lo1[3], hi1[3];
  int32_t child[2];
  int32_t pad_[2];
};
static_assert(sizeof(FNode) == 64, "FNode is 2 x 32 B");

__global__ void __launch_bounds__(256) fnode_kernel(const Node<double, 3>* __restrict__ nodes, int inner, FNode* __restrict__ out)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if(i >= inner) return;
  const Node<double, 3> nd = nodes[i];
  FNode f;
#pragma unroll
  for(int d = 0; d < 3; ++d)
  {
    f.lo0[d] = __double2float_rd(nd.box[0].lo[d]);
    f.hi0[d] = __double2float_ru(nd.box[0].hi[d]);
    f.lo1[d] = __double2float_rd(nd.box[1].lo[d]);
    f.hi1[d] = __double2float_ru(nd.box[1].hi[d]);
  }
  f.child[0] = nd.child[0];
  f.child[1] = nd.child[1];
  f.pad_[0] = f.pad_[1] = 0;
  out[i] = f;
}

__device__ __forceinline__ void load_fnode_boxes(const FNode* __restrict__ fnodes, int32_t i, Box<double, 3>& L, Box<double, 3>& R, int32_t& lc,
                                                 int32_t& rc)
{
  const D4* p = reinterpret_cast<const D4*>(fnodes + i);
  const D4 a = ldg256(p), b = ldg256(p + 1);
  const double w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
  float f[16];
#pragma unroll
  for(int k = 0; k < 8; ++k)
  {
    f[2 * k] = __int_as_float(__double2loint(w[k]));
    f[2 * k + 1] = __int_as_float(__double2hiint(w[k]));
  }
#pragma unroll
  for(int d = 0; d < 3; ++d)
  {
    L.lo[d] = (double)f[d];
    L.hi[d] = (double)f[3 + d];
    R.lo[d] = (double)f[6 + d];
    R.hi[d] = (double)f[9 + d];
  }
  lc = __float_as_int(f[12]);
  rc = __float_as_int(f[13]);
}

// visit of inner node `cur` through the compact record; leaf children that pass are confirmed on the exact box
template <class Query>
__device__ __forceinline__ void compact_visit(const Node<double, 3>* __restrict__ nodes, const FNode* __restrict__ fnodes, int32_t cur,
                                              const Query& q, int32_t& lc, int32_t& rc, bool& inL, bool& inR)
{
  Box<double, 3> L, R;
  load_fnode_boxes(fnodes, cur, L, R, lc, rc);
  inL = box_valid(L) ? q(L) : false;
  inR = box_valid(R) ? q(R) : false;
  if(inL && lc < 0)
  {
    const Box<double, 3> e = nodes[cur].box[0];
    inL = box_valid(e) ? q(e) : false;
  }
  if(inR && rc < 0)
  {
    const Box<double, 3> e = nodes[cur].box[1];
    inR = box_valid(e) ? q(e) : false;
  }
}
template <typename T, int D, class Query>
__device__ __forceinline__ void compact_visit(const Node<T, D>*, const FNode*, int32_t, const Query&, int32_t&, int32_t&, bool&, bool&)
{
}

template <typename T, int D, class Query, class Filter = NoFilter>
__global__ void __launch_bounds__(128) find_walk_kernel(const Node<T, D>* __restrict__ nodes, const int32_t* __restrict__ leaf_nodes,
                                                         Desc<Query::NCOMP> prims, int nq, T tol, int flags,
                                                         const int32_t* __restrict__ perm, int32_t* __restrict__ counts, PairBuf pb,
                                                         Filter filt = Filter(), const FNode* __restrict__ fnodes = nullptr)
{
  constexpr unsigned FULL = 0xffffffffu;
  const unsigned lane = lane_id();
  const unsigned lt_mask = (1u << lane) - 1u;
  int qi = -1;
  Query q;
  int32_t todo[kStackSize];
  int sp = 0;
  int32_t cur = kBarrier;
  int cnt = 0;
  unsigned wbase = 0, wcount = 0;
  bool exhausted = false;
  // the warp's pair chunk: [ppos, pend) are free slots of chunk pchunk
  unsigned pchunk = 0xffffffffu, ppos = 0, pend = 0;
  bool recording = pb.max_chunks != 0u;

  while(true)
  {
    // ---- refill free lanes ----
    const unsigned freem = __ballot_sync(FULL, qi < 0);
    if(freem != 0u && !exhausted)
    {
      if(wcount == 0u)
      {
        unsigned b = 0;
        if(lane == 0) b = atomicAdd(pb.cursor, (unsigned)kFindQueryChunk);
        wbase = __shfl_sync(FULL, b, 0);
        wcount = wbase < (unsigned)nq ? min((unsigned)kFindQueryChunk, (unsigned)nq - wbase) : 0u;
        exhausted = (wcount == 0u);
      }
      if(wcount != 0u)
      {
        const unsigned rank = __popc(freem & lt_mask);
        if(qi < 0 && rank < wcount)
        {
          const unsigned t = wbase + rank;
          qi = perm ? perm[t] : (int)t;
          q.load(prims, qi, tol, flags);
          sp = 0;
          cur = 0;  // root
          cnt = 0;
        }
        const unsigned taken = min((unsigned)__popc(freem), wcount);
        wbase += taken;
        wcount -= taken;
      }
    }
    const bool busy = qi >= 0;
    if(__ballot_sync(FULL, busy) == 0u)
    {
      if(exhausted) break;
      continue;
    }
    // ---- lanes that hold a leaf: record the hit (ballot + prefix into the warp's chunk), then pop ----
    const bool at_leaf = busy && cur < 0 && cur != kBarrier;
    bool keep = at_leaf;  // the hit is recorded unless the narrow-phase filter rejects it
    int32_t cand = 0;
    if(__any_sync(FULL, at_leaf))
    {
      if(at_leaf)
      {
        cand = __ldg(leaf_nodes + (-cur - 1));
        keep = filt(qi, cand);
      }
    }
    const unsigned leafm = __ballot_sync(FULL, keep);
    if(leafm != 0u)
    {
      if(recording)
      {
        unsigned n = __popc(leafm);
        unsigned my = __popc(leafm & lt_mask);  // rank of this lane among the recording lanes
        unsigned done = 0;
        while(done < n)
        {
          if(ppos == pend)
          {
            unsigned c = 0;
            if(lane == 0) c = atomicAdd(pb.cursor + 1, 1u);
            c = __shfl_sync(FULL, c, 0);
            if(c >= pb.max_chunks)
            {
              if(lane == 0) atomicExch(pb.cursor + 2, 1u);
              recording = false;
              pchunk = 0xffffffffu;
              break;
            }
            pchunk = c;
            ppos = c * (unsigned)kPairChunk;
            pend = ppos + (unsigned)kPairChunk;
          }
          const unsigned room = pend - ppos;
          const unsigned take = min(room, n - done);
          if(keep && my >= done && my < done + take) pb.pairs[ppos + (my - done)] = make_int4(qi, cnt, cand, 0);
          ppos += take;
          done += take;
        }
      }
    }
    if(at_leaf)
    {
      if(keep) ++cnt;
      cur = sp > 0 ? todo[--sp] : kBarrier;
    }
    // ---- lanes that hold an inner node: the reference's step (bvh_traverse.hpp:86-123), left child first ----
    else if(busy && cur >= 0)
    {
      Box<T, D> L, R;
      int32_t lc, rc;
      bool inL, inR;
      if(std::is_same<T, double>::value && D == 3 && fnodes != nullptr)
      {
        compact_visit(nodes, fnodes, cur, q, lc, rc, inL, inR);
      }
      else
      {
        load_node_boxes<T, D>(nodes, cur, L, R, lc, rc);
        inL = box_valid(L) ? q(L) : false;
        inR = box_valid(R) ? q(R) : false;
      }
      if(inL && inR)
      {
        todo[sp++] = rc;
        cur = lc;
      }
      else if(inL)
      {
        cur = lc;
      }
      else if(inR)
      {
        cur = rc;
      }
      else
      {
        cur = sp > 0 ? todo[--sp] : kBarrier;
      }
    }
    // ---- finished lanes publish their count and become free ----
    if(busy && cur == kBarrier)
    {
      counts[qi] = cnt;
      qi = -1;
    }
  }
  if(pchunk != 0xffffffffu && lane == 0) pb.unused[pchunk] = pend - ppos;
}

// candidates[offsets[q] + rank] = candidate, one thread per recorded slot
__global__ void __launch_bounds__(256) scatter_pairs_kernel(const int4* __restrict__ pairs, const unsigned int* __restrict__ unused,
                                                             unsigned int nchunks, const int32_t* __restrict__ offsets,
                                                             int32_t* __restrict__ candidates, int32_t* __restrict__ firsts = nullptr)
{
  const unsigned long long slot = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
  const unsigned chunk = (unsigned)(slot / kPairChunk);
  if(chunk >= nchunks) return;
  const unsigned in_chunk = (unsigned)(slot % kPairChunk);
  if(in_chunk >= (unsigned)kPairChunk - unused[chunk]) return;
  const int4 p = pairs[slot];
  const long long o = (long long)offsets[p.x] + p.y;
  candidates[o] = p.z;
  if(firsts) firsts[o] = p.x;
}

template <typename T, int D, class Query>
__device__ __forceinline__ bool query_sort_point(const Query& q, const T*, const T*, T* c)
{
  q.center(c);
  return true;
}
template <typename T, int D>
__device__ __forceinline__ bool query_sort_point(const RayQuery<T, D>& q, const T* bmin, const T* bmax, T* c)
{
  return q.sort_point(bmin, bmax, c);
}
// the top D bits of a ray's key: its direction octant (rays entering at the same place but heading apart share nothing below
// the first few levels); other queries: none
template <typename T, int D, class Query>
__device__ __forceinline__ uint32_t query_sort_prefix(const Query&, uint32_t code)
{
  return code;
}
template <typename T, int D>
__device__ __forceinline__ uint32_t query_sort_prefix(const RayQuery<T, D>& q, uint32_t code)
{
  uint32_t oct = 0;
#pragma unroll
  for(int k = 0; k < D; ++k) oct |= (q.dir[k] < (T)0 ? 1u : 0u) << k;
  return (oct << (32 - D)) | (code >> D);
}

// Morton keys of the queries' reference points over the BVH bounds (centroid of a box, entry point of a ray):
// (code << 32) | index, for radix_sort.cuh.  Only the processing ORDER depends on it.  The digit histograms are
// aggregated per warp first (match.any): clustered queries would otherwise serialise on one shared-memory counter.
template <typename T, int D, class Query, class State>
__global__ void __launch_bounds__(256) find_query_keys_kernel(Desc<Query::NCOMP> prims, int nq, const State* __restrict__ st,
                                                               unsigned long long* __restrict__ keys, uint32_t* __restrict__ ghist)
{
  __shared__ uint32_t sh[rsort::MAX_PASSES * rsort::RADIX];
  for(int i = threadIdx.x; i < rsort::MAX_PASSES * rsort::RADIX; i += blockDim.x) sh[i] = 0;
  __syncthreads();
  T mn[D], mx[D], inv[D];
#pragma unroll
  for(int d = 0; d < D; ++d)
  {
    mn[d] = st->bmin[d];
    mx[d] = st->bmax[d];
    inv[d] = st->inv_extent[d];
  }
  const unsigned lane = lane_id();
  for(long long base = (long long)blockIdx.x * blockDim.x; base < nq; base += (long long)gridDim.x * blockDim.x)
  {
    const long long i = base + threadIdx.x;  // whole warps stay in the loop: match.any needs every lane
    const bool valid = i < nq;
    uint32_t code = 0xffffffffu;
    if(valid)
    {
      Query q;
      q.load(prims, i, (T)0, 1);
      T c[D];
      if(query_sort_point<T, D>(q, mn, mx, c))
      {
        uint32_t qd[D];
        constexpr int bits = 32 / D;
#pragma unroll
        for(int d = 0; d < D; ++d)
        {
          T v = (c[d] - mn[d]) * inv[d] * (T)(1 << bits);
          v = v > (T)0 ? v : (T)0;  // also maps NaN to 0
          v = v < (T)((1 << bits) - 1) ? v : (T)((1 << bits) - 1);
          qd[d] = (uint32_t)(int32_t)v;
        }
        if(D == 2)
          code = spread_bits_2d(qd[0]) | (spread_bits_2d(qd[1]) << 1);
        else
          code = spread_bits_3d(qd[0]) | (spread_bits_3d(qd[1]) << 1) | (spread_bits_3d(qd[D - 1]) << 2);
        code = query_sort_prefix<T, D>(q, code);
      }
      keys[i] = ((unsigned long long)code << 32) | (unsigned long long)(uint32_t)i;
    }
    // lanes with the same code add once, together
    const unsigned grp = __match_any_sync(0xffffffffu, code) & __ballot_sync(0xffffffffu, valid);
    if(valid && (grp & ((1u << lane) - 1u)) == 0u)
    {
      const unsigned n = (unsigned)__popc(grp);
#pragma unroll
      for(int p = 0; p < rsort::MAX_PASSES; ++p) atomicAdd(&sh[p * rsort::RADIX + ((code >> (p * 8)) & 255u)], n);
    }
  }
  __syncthreads();
  for(int i = threadIdx.x; i < rsort::MAX_PASSES * rsort::RADIX; i += blockDim.x)
    if(sh[i]) atomicAdd(&ghist[i], sh[i]);
}

//------------------------------------------------------------------------------------------
// exclusive scan of int32 counts -> int32 offsets, int64 total (RAJA::exclusive_scan,
// LinearBVH.hpp:332-334).  Three small kernels; 12 B/query of traffic.
//------------------------------------------------------------------------------------------
constexpr int SCAN_BLOCK = 256;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_TILE = SCAN_BLOCK * SCAN_ITEMS;

__device__ __forceinline__ long long block_reduce_sum(long long v, long long* sh)
{
#pragma unroll
  for(int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if(lane_id() == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  long long r = 0;
  for(int w = 0; w < (int)(blockDim.x >> 5); ++w) r += sh[w];
  __syncthreads();
  return r;
}

__global__ void __launch_bounds__(SCAN_BLOCK) scan_tile_sums_kernel(const int32_t* __restrict__ counts, int n, long long* __restrict__ tile_sums)
{
  __shared__ long long sh[SCAN_BLOCK / 32];
  const int base = blockIdx.x * SCAN_TILE;
  long long s = 0;
#pragma unroll
  for(int k = 0; k < SCAN_ITEMS; ++k)
  {
    const int i = base + k * SCAN_BLOCK + threadIdx.x;
    if(i < n) s += counts[i];
  }
  s = block_reduce_sum(s, sh);
  if(threadIdx.x == 0) tile_sums[blockIdx.x] = s;
}

// single block: exclusive scan of the tile sums in place, total -> *total
__global__ void __launch_bounds__(1024) scan_spine_kernel(long long* tile_sums, int ntiles, long long* total)
{
  __shared__ long long sh[1024];
  __shared__ long long carry;
  if(threadIdx.x == 0) carry = 0;
  __syncthreads();
  for(int base = 0; base < ntiles; base += 1024)
  {
    const int i = base + threadIdx.x;
    const long long v = (i < ntiles) ? tile_sums[i] : 0;
    sh[threadIdx.x] = v;
    __syncthreads();
    long long incl = v;
    for(int off = 1; off < 1024; off <<= 1)
    {
      const long long u = (threadIdx.x >= (unsigned)off) ? sh[threadIdx.x - off] : 0;
      __syncthreads();
      incl += u;
      sh[threadIdx.x] = incl;
      __syncthreads();
    }
    if(i < ntiles) tile_sums[i] = carry + incl - v;
    __syncthreads();
    if(threadIdx.x == 1023) carry += incl;
    __syncthreads();
  }
  if(threadIdx.x == 0) *total = carry;
}

__global__ void __launch_bounds__(SCAN_BLOCK) scan_apply_kernel(const int32_t* __restrict__ counts, int n, const long long* __restrict__ tile_offsets,
                                                                 int32_t* __restrict__ offsets)
{
  __shared__ int wsum[SCAN_BLOCK / 32];
  const int base = blockIdx.x * SCAN_TILE;
  // blocked arrangement: thread owns SCAN_ITEMS consecutive elements
  const int first = base + threadIdx.x * SCAN_ITEMS;
  int v[SCAN_ITEMS];
  int tsum = 0;
#pragma unroll
  for(int k = 0; k < SCAN_ITEMS; ++k)
  {
    const int i = first + k;
    v[k] = (i < n) ? counts[i] : 0;
    tsum += v[k];
  }
  int incl = tsum;
#pragma unroll
  for(int o = 1; o < 32; o <<= 1)
  {
    const int u = __shfl_up_sync(0xffffffffu, incl, o);
    if((int)lane_id() >= o) incl += u;
  }
  if(lane_id() == 31) wsum[threadIdx.x >> 5] = incl;
  __syncthreads();
  int wbase = 0;
  for(int w = 0; w < (int)(threadIdx.x >> 5); ++w) wbase += wsum[w];
  long long run = tile_offsets[blockIdx.x] + wbase + (incl - tsum);
#pragma unroll
  for(int k = 0; k < SCAN_ITEMS; ++k)
  {
    const int i = first + k;
    if(i < n) offsets[i] = (int32_t)run;
    run += v[k];
  }
}

}  // namespace axb
