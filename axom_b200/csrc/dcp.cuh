// dcp.cuh -- the per-rank step of quest::DistributedClosestPoint (sm_100a).
//
// Reference path replaced (quest/detail/DistributedClosestPointImpl.hpp):
//   generateBVHTreeImpl       :883-903    one zero-size box per object point, spin::BVH::initialize
//   computeLocalClosestPoints :905-1079   per query: preset from the state earlier ranks left, traverse_tree with
//                                         checkMinDist / traversePredicate, write back only what this rank improved
// The object "mesh" is a point cloud.  The traversal is the reference's (traverse_reference_order, left child first,
// both child predicates evaluated at the parent), the arithmetic is primal::squared_distance (point-point: sum of
// squared differences in order; point-box: clamp, then the same sum) with separately rounded operations, and the
// leaf test is a strict <, so the nearest point, ties included, is the one the reference reports.
#pragma once
#include "common.cuh"
#include "traverse.cuh"

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
    NoOrder {});
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
}  // namespace axb
