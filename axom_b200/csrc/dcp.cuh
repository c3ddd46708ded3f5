// dcp.cuh -- the per-rank step of quest::DistributedClosestPoint (sm_100a).
//
// Reference path replaced (quest/detail/DistributedClosestPointImpl.hpp):
//   generateBVHTreeImpl       :883-903    one zero-size box per object point, spin::BVH::initialize
//   computeLocalClosestPoints :905-1079   per query: preset from the state earlier ranks left, traverse_tree with
//                                         checkMinDist / traversePredicate, write back only what this rank improved
// The object "mesh" is a point cloud.  The traversal is the reference's (traverse_reference_order; the query is a
// PointType, so overload resolution picks LinearBVHTraverser::traverse_tree(const PointType&, ...), policy/LinearBVH.hpp:72-85,
// which enters the child with the nearer box centroid first;
// both child predicates evaluated at the parent), the arithmetic is primal::squared_distance (point-point: sum of
// squared differences in order; point-box: clamp, then the same sum) with separately rounded operations, and the
// leaf test is a strict <, so the nearest point, ties included, is the one the reference reports.
#pragma once
#include "common.cuh"
#include "traverse.cuh"
#include "build.cuh"

namespace axb
{
// boxes[i] = BoxType{pts[i]}
template <int D>
__global__ void __launch_bounds__(256) dcp_point_boxes_kernel(const double* __restrict__ pts, int n, Box<double, D>* __restrict__ boxes)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if(i >= n) return;
  Box<double, D> b;
#pragma unroll
  for(int d = 0; d < D; ++d) b.lo[d] = b.hi[d] = pts[(size_t)i * D + d];
  boxes[i] = b;
}

// primal::squared_distance(Point, BoundingBox) (primal/operators/squared_distance.hpp:77-100); the box is valid here
template <int D>
__device__ __forceinline__ double dcp_sqdist_box(const double* p, const Box<double, D>& b)
{
  bool inside = true;
#pragma unroll
  for(int d = 0; d < D; ++d) inside = inside && !(p[d] < b.lo[d] || p[d] > b.hi[d]);
  if(inside) return 0.0;
  double s = 0.0;
#pragma unroll
  for(int d = 0; d < D; ++d)
  {
    const double c = p[d] < b.lo[d] ? b.lo[d] : (p[d] > b.hi[d] ? b.hi[d] : p[d]);  // clampVal
    const double v = c - p[d];
    s += v * v;
  }
  return s;
}

// traversePref of the point overload (policy/LinearBVH.hpp:75-82): true = the right child's centroid is nearer
template <int D>
struct DcpCentroidOrder
{
  const double* p;
  __device__ __forceinline__ bool operator()(const Box<double, D>& L, const Box<double, D>& R) const
  {
    double dl = 0.0, dr = 0.0;
#pragma unroll
    for(int d = 0; d < D; ++d)
    {
      const double c = 0.5 * (L.lo[d] + L.hi[d]) - p[d];
      dl += c * c;
    }
    if(box_valid(R))
    {
#pragma unroll
      for(int d = 0; d < D; ++d)
      {
        const double c = 0.5 * (R.lo[d] + R.hi[d]) - p[d];
        dr += c * c;
      }
    }
    else
    {
      dr = DBL_MAX;
    }
    return dl > dr;
  }
};

// One thread per query.  State arrays (cp_*) are read and updated in place; is_first initialises them (:971-979).
// nodes == nullptr: the rank has no object points (only the initialisation happens, :911-916).
template <int D>
__global__ void __launch_bounds__(128) dcp_local_kernel(const Node<double, D>* __restrict__ nodes, const int32_t* __restrict__ leaf_nodes,
                                                         const double* __restrict__ obj_pts, const int32_t* __restrict__ obj_dom, int rank,
                                                         double sq_thresh, const double* __restrict__ query, int nq,
                                                         const int32_t* __restrict__ perm, int is_first, int32_t* __restrict__ cp_index,
                                                         int32_t* __restrict__ cp_dom, int32_t* __restrict__ cp_rank,
                                                         double* __restrict__ cp_coords, double* __restrict__ cp_dist)
{
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if(t >= nq) return;
  const int i = perm ? perm[t] : t;
  double p[D];
#pragma unroll
  for(int d = 0; d < D; ++d) p[d] = query[(size_t)i * D + d];
  double cur_sq = DBL_MAX;  // MinCandidate{} (:533-543)
  int cur_idx = -1, cur_dom = -1, cur_rank = -1;
  if(is_first)
  {
    const double snan = __longlong_as_double(0x7ff4000000000000ll);  // numeric_limits<double>::signaling_NaN()
    cp_rank[i] = -1;
    cp_index[i] = -1;
    cp_dom[i] = -1;
#pragma unroll
    for(int d = 0; d < D; ++d) cp_coords[(size_t)i * D + d] = snan;
    if(cp_dist) cp_dist[i] = snan;
  }
  else if(cp_rank[i] >= 0)  // preset with the closest point found so far (:1013-1019)
  {
    double s = 0.0;
#pragma unroll
    for(int d = 0; d < D; ++d)
    {
      const double v = cp_coords[(size_t)i * D + d] - p[d];
      s += v * v;
    }
    cur_sq = s;
    cur_idx = cp_index[i];
    cur_dom = cp_dom[i];
    cur_rank = cp_rank[i];
  }
  if(nodes == nullptr) return;
  traverse_reference_order<double, D>(
    nodes,
    [&](const Box<double, D>& bb) {  // traversePredicate (:1037-1040)
      const double sq = dcp_sqdist_box<D>(p, bb);
      return sq <= cur_sq && sq <= sq_thresh;
    },
    [&](int pos) {  // checkMinDist (:1021-1035)
      const int c = __ldg(leaf_nodes + pos);
      double s = 0.0;
#pragma unroll
      for(int d = 0; d < D; ++d)
      {
        const double v = __ldg(obj_pts + (size_t)c * D + d) - p[d];
        s += v * v;
      }
      if(s < cur_sq)
      {
        cur_sq = s;
        cur_idx = c;
        cur_dom = __ldg(obj_dom + c);
        cur_rank = rank;
      }
    },
    DcpCentroidOrder<D> {p});
  if(cur_rank == rank)  // :1045-1058
  {
    cp_index[i] = cur_idx;
    cp_dom[i] = cur_dom;
    cp_rank[i] = cur_rank;
#pragma unroll
    for(int d = 0; d < D; ++d) cp_coords[(size_t)i * D + d] = __ldg(obj_pts + (size_t)cur_idx * D + d);
    if(cp_dist) cp_dist[i] = sqrt(cur_sq);
  }
}

// MODE 1 (default): the same answer from a nearest-first search.
//
// The reference walks its tree nearer-centroid-first with no initial bound, so a query first descends to leaves that
// may be far away and only then starts pruning: thousands of node visits per query on a 2 M-point cloud.  Its RESULT
// has an order-free description except for exact ties: a leaf replaces the running minimum only on a strict <, and
// every leaf at the minimum distance d* is reached (its zero-size box is never pruned), so the reference reports the
// FIRST leaf of its own visiting order among those at d* (a preset from an earlier rank wins every tie).  This kernel
// searches in best-first order -- nearer child first, the other on a (lower bound, node) stack, subtrees pruned when
// their bound exceeds the running minimum, bounds EQUAL to it still entered (<=) -- which finds d* and notices
// whether a second leaf attains it.  Only then (rare: lattice clouds) the query is replayed in the reference's order
// with the prune radius fixed at d*, stopping at the first leaf at d*: the leaves that replay visits are a subsequence
// of the reference's, in the same relative order (the centroid comparison does not depend on the radius), and it
// contains every leaf at d*.  The arithmetic of every distance is the reference's; the answer is bit-identical
// (tests run both modes, including clouds and queries on a lattice where up to 8 points tie).
template <int D>
__global__ void __launch_bounds__(128) dcp_nearest_kernel(const Node<double, D>* __restrict__ nodes, const int32_t* __restrict__ leaf_nodes,
                                                           const double* __restrict__ obj_pts, const int32_t* __restrict__ obj_dom, int rank,
                                                           double sq_thresh, const double* __restrict__ query, int nq,
                                                           const int32_t* __restrict__ perm, int is_first, int32_t* __restrict__ cp_index,
                                                           int32_t* __restrict__ cp_dom, int32_t* __restrict__ cp_rank,
                                                           double* __restrict__ cp_coords, double* __restrict__ cp_dist,
                                                           const double* __restrict__ bound_sq = nullptr)
{
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if(t >= nq) return;
  const int i = perm ? perm[t] : t;
  double p[D];
#pragma unroll
  for(int d = 0; d < D; ++d) p[d] = query[(size_t)i * D + d];
  // bound_sq (optional, with is_first): only a point with squared distance <= bound_sq[i] is of interest -- an upper
  // bound some other rank already achieved.  Unlike a preset it does NOT win ties: a point AT the bound is reported.
  double cur_sq = bound_sq ? bound_sq[i] : DBL_MAX;
  int cur_pos = 0x7fffffff;  // sorted position of the running minimum; -1 = a preset, which wins every tie
  int cur_idx = -1;
  bool improved = false, tie = false;
  if(is_first)
  {
    const double snan = __longlong_as_double(0x7ff4000000000000ll);
    cp_rank[i] = -1;
    cp_index[i] = -1;
    cp_dom[i] = -1;
#pragma unroll
    for(int d = 0; d < D; ++d) cp_coords[(size_t)i * D + d] = snan;
    if(cp_dist) cp_dist[i] = snan;
  }
  else if(cp_rank[i] >= 0)
  {
    double s = 0.0;
#pragma unroll
    for(int d = 0; d < D; ++d)
    {
      const double v = cp_coords[(size_t)i * D + d] - p[d];
      s += v * v;
    }
    cur_sq = s;
    cur_pos = -1;
  }
  if(nodes == nullptr) return;

  auto leaf = [&](int pos) {
    const int c = __ldg(leaf_nodes + pos);
    double s = 0.0;
#pragma unroll
    for(int d = 0; d < D; ++d)
    {
      const double v = __ldg(obj_pts + (size_t)c * D + d) - p[d];
      s += v * v;
    }
    // reached in the reference only if the leaf's own (zero-size) box passes the predicate: s <= threshold
    if(s <= sq_thresh)
    {
      if(s < cur_sq || (s == cur_sq && !improved && cur_pos != -1))
      {
        cur_sq = s;
        cur_pos = pos;
        cur_idx = c;
        improved = true;
        tie = false;
      }
      else if(s == cur_sq && improved)
      {
        tie = true;  // a second point at the running minimum: which one the reference reports depends on its visiting order
      }
    }
  };

  double st_lb[kStackSize];
  int32_t st_id[kStackSize];
  int sp = 0;
  int32_t cur = 0;  // root
  while(true)
  {
    const Node<double, D>& nd = nodes[cur];
    const Box<double, D> L = nd.box[0], R = nd.box[1];
    const int32_t lc = nd.child[0], rc = nd.child[1];
    double dl = box_valid(L) ? dcp_sqdist_box<D>(p, L) : DBL_MAX;
    double dr = box_valid(R) ? dcp_sqdist_box<D>(p, R) : DBL_MAX;
    bool inl = box_valid(L) && dl <= cur_sq && dl <= sq_thresh;
    bool inr = box_valid(R) && dr <= cur_sq && dr <= sq_thresh;
    // leaves are evaluated on the spot (a leaf's bound IS its distance), nearer one first
    if(inl && lc < 0 && (!(inr && rc < 0) || dl <= dr))
    {
      leaf(-lc - 1);
      inl = false;
      inr = inr && dr <= cur_sq;
    }
    if(inr && rc < 0)
    {
      leaf(-rc - 1);
      inr = false;
      inl = inl && dl <= cur_sq;
    }
    if(inl && lc < 0)
    {
      leaf(-lc - 1);
      inl = false;
    }
    int32_t next = kBarrier;
    if(inl && inr)
    {
      const bool left_first = dl <= dr;
      st_lb[sp] = left_first ? dr : dl;
      st_id[sp] = left_first ? rc : lc;
      ++sp;
      next = left_first ? lc : rc;
    }
    else if(inl)
      next = lc;
    else if(inr)
      next = rc;
    while(next == kBarrier && sp > 0)
    {
      --sp;
      if(st_lb[sp] <= cur_sq) next = st_id[sp];
    }
    if(next == kBarrier) break;
    cur = next;
  }
  if(tie)
  {
    // several points at d* = cur_sq: the reference reports the first one of ITS visiting order
    const double dstar = cur_sq;
    bool found = false;
    traverse_reference_order<double, D>(
      nodes,
      [&](const Box<double, D>& bb) {
        if(found) return false;
        const double sq = dcp_sqdist_box<D>(p, bb);
        return sq <= dstar && sq <= sq_thresh;
      },
      [&](int pos) {
        if(found) return;
        const int c = __ldg(leaf_nodes + pos);
        double sl = 0.0;
#pragma unroll
        for(int d = 0; d < D; ++d)
        {
          const double v = __ldg(obj_pts + (size_t)c * D + d) - p[d];
          sl += v * v;
        }
        if(sl == dstar)
        {
          cur_idx = c;
          found = true;
        }
      },
      DcpCentroidOrder<D> {p});
  }
  if(improved)
  {
    cp_index[i] = cur_idx;
    cp_dom[i] = __ldg(obj_dom + cur_idx);
    cp_rank[i] = rank;
#pragma unroll
    for(int d = 0; d < D; ++d) cp_coords[(size_t)i * D + d] = __ldg(obj_pts + (size_t)cur_idx * D + d);
    if(cp_dist) cp_dist[i] = sqrt(cur_sq);
  }
}

// Morton key of a query point over the object BVH's bounds (processing order only): (code << 32) | index
template <int D, class State>
__global__ void __launch_bounds__(256) dcp_query_keys_kernel(const double* __restrict__ query, int nq, const State* __restrict__ st,
                                                              unsigned long long* __restrict__ keys, uint32_t* __restrict__ ghist)
{
  __shared__ uint32_t sh[rsort::MAX_PASSES * rsort::RADIX];
  for(int i = threadIdx.x; i < rsort::MAX_PASSES * rsort::RADIX; i += blockDim.x) sh[i] = 0;
  __syncthreads();
  for(int i = blockIdx.x * blockDim.x + threadIdx.x; i < nq; i += gridDim.x * blockDim.x)
  {
    double c[D];
#pragma unroll
    for(int d = 0; d < D; ++d) c[d] = (query[(size_t)i * D + d] - st->bmin[d]) * st->inv_extent[d];
    const uint32_t code = morton32<double, D>(c);
    keys[i] = ((unsigned long long)code << 32) | (unsigned long long)(uint32_t)i;
#pragma unroll
    for(int p = 0; p < rsort::MAX_PASSES; ++p) atomicAdd(&sh[p * rsort::RADIX + ((code >> (p * 8)) & 255u)], 1u);
  }
  __syncthreads();
  for(int i = threadIdx.x; i < rsort::MAX_PASSES * rsort::RADIX; i += blockDim.x)
    if(sh[i]) atomicAdd(&ghist[i], sh[i]);
}
}  // namespace axb
