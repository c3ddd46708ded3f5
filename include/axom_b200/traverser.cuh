// axom_b200/traverser.cuh -- device-side traverse_tree() for the traverser returned by
// spin::BVH::getTraverser(), i.e. the drop-in for LinearBVHTraverser::traverse_tree
// (spin/policy/LinearBVH.hpp:57-109) used by callers that fuse their own leaf action into the walk
// (quest::SignedDistance, DistributedClosestPoint, mir::TopologyMapper).  Include from CUDA code only.
//
//   traverse_tree(traverser, p, leafAction, predicate)            children entered left first -- EXCEPT when p is a
//                                                                  primal::Point: nearer box centroid first (below)
//   traverse_tree(traverser, p, leafAction, predicate, comp)      comp(leftBox, rightBox, p) == true
//                                                                  enters the right child first
//   leafAction(sorted_pos, leaf_nodes)   predicate(p, box) -> bool
//
// The walk is a restatement of lbvh::bvh_traverse (spin/internal/linear_bvh/bvh_traverse.hpp:66-154)
// over the reference-layout arrays: iterative DFS with a 64-entry stack, invalid child boxes never
// entered, and the reference's habit of parking the first leaf it meets until it has found a second
// one, so leaves are reported in exactly the reference's order.
#ifndef AXOM_B200_TRAVERSER_CUH_
#define AXOM_B200_TRAVERSER_CUH_

#include "BVH.hpp"

namespace axom_b200
{
namespace spin
{
namespace detail
{
template <typename T, int D>
__host__ __device__ inline bool box_is_valid(const primal::BoundingBox<T, D>& b)
{
  for(int d = 0; d < D; ++d)
    if(b.m_min.m_components[d] > b.m_max.m_components[d]) return false;
  return true;
}
}  // namespace detail

template <typename T, int D, typename Prim, typename LeafAction, typename Predicate, typename Comp>
__host__ __device__ inline void traverse_tree(const LinearBVHTraverser<T, D>& tr, const Prim& p, LeafAction&& leafAction,
                                              Predicate&& predicate, Comp&& comp)
{
  constexpr std::int32_t kBarrier = -2000000000;  // bvh_traverse.hpp:80
  std::int32_t todo[64];                          // bvh_traverse.hpp:79
  int sp = 0;
  todo[0] = kBarrier;
  std::int32_t parked = 0;
  std::int32_t cur = 0;
  while(cur != kBarrier)
  {
    while(cur >= 0)
    {
      const primal::BoundingBox<T, D> L = tr.m_inner_nodes[cur];
      const primal::BoundingBox<T, D> R = tr.m_inner_nodes[cur + 1];
      const bool inL = detail::box_is_valid(L) && predicate(p, L);
      const bool inR = detail::box_is_valid(R) && predicate(p, R);
      const std::int32_t lc = tr.m_inner_node_children[cur];
      std::int32_t rc = tr.m_inner_node_children[cur + 1];
      if(!inL && !inR)
      {
        cur = todo[sp--];
      }
      else
      {
        cur = inL ? lc : rc;
        if(inL && inR)
        {
          if(comp(L, R, p))
          {
            const std::int32_t t = cur;
            cur = rc;
            rc = t;
          }
          todo[++sp] = rc;
        }
      }
      if(cur < 0 && parked >= 0)
      {
        parked = cur;
        if(cur != kBarrier) cur = todo[sp--];
      }
    }
    while(parked < 0 && parked != kBarrier)
    {
      leafAction(-parked - 1, tr.m_leaf_nodes);
      parked = cur;
      if(cur < 0 && cur != kBarrier) cur = todo[sp--];
    }
    parked = 0;
  }
}

struct NoTraversePreference
{
  template <typename B, typename P>
  __host__ __device__ bool operator()(const B&, const B&, const P&) const
  {
    return false;
  }
};

template <typename T, int D, typename Prim, typename LeafAction, typename Predicate>
__host__ __device__ inline void traverse_tree(const LinearBVHTraverser<T, D>& tr, const Prim& p, LeafAction&& leafAction,
                                              Predicate&& predicate)
{
  traverse_tree(tr, p, leafAction, predicate, NoTraversePreference {});
}

// The POINT overload (spin/policy/LinearBVH.hpp:72-85): when the primitive is a primal::Point of the tree's own
// type and dimension, overload resolution in the reference picks the version that enters the child whose box
// CENTROID is nearer to the point first (an invalid right box counts as infinitely far).  Nearest-neighbour callers
// (quest::SignedDistance, DistributedClosestPoint) depend on that order: with a strict < in the leaf action it decides
// which of several equidistant items is reported.
template <typename T, int D>
struct NearerCentroidFirst
{
  __host__ __device__ bool operator()(const primal::BoundingBox<T, D>& l, const primal::BoundingBox<T, D>& r, const primal::Point<T, D>& p) const
  {
    // squared_distance(Point, Point) in double (primal/operators/squared_distance.hpp:62-75), centroid = 0.5 * (min + max)
    double dl = 0.0, dr = 0.0;
    for(int d = 0; d < D; ++d)
    {
      const T c = static_cast<T>(0.5) * (l.m_min.m_components[d] + l.m_max.m_components[d]);
      const double v = static_cast<double>(c) - static_cast<double>(p.m_components[d]);
      dl += v * v;
    }
    if(detail::box_is_valid(r))
    {
      for(int d = 0; d < D; ++d)
      {
        const T c = static_cast<T>(0.5) * (r.m_min.m_components[d] + r.m_max.m_components[d]);
        const double v = static_cast<double>(c) - static_cast<double>(p.m_components[d]);
        dr += v * v;
      }
    }
    else
    {
      dr = 1.7976931348623157e308;
    }
    return dl > dr;
  }
};

template <typename T, int D, typename LeafAction, typename Predicate>
__host__ __device__ inline void traverse_tree(const LinearBVHTraverser<T, D>& tr, const primal::Point<T, D>& p, LeafAction&& leafAction,
                                              Predicate&& predicate)
{
  traverse_tree(tr, p, leafAction, predicate, NearerCentroidFirst<T, D> {});
}

}  // namespace spin
}  // namespace axom_b200

#endif  // AXOM_B200_TRAVERSER_CUH_
