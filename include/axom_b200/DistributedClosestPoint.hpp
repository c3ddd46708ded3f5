// axom_b200/DistributedClosestPoint.hpp -- C++ header shim for the per-rank half of quest::DistributedClosestPoint
// (quest/DistributedClosestPoint.hpp:56-215, quest/detail/DistributedClosestPointImpl.hpp:551-1095) on top of the
// C ABI in axb200.h.  Header-only, C++14.
//
// The reference class takes Conduit blueprint nodes and an MPI communicator; here the blueprint nodes are reduced to
// what the implementation extracts from them -- interleaved coordinates per object domain, interleaved query
// coordinates, and the five xferDom arrays -- and the class covers what ONE rank does: setObjectMesh,
// generateBVHTree, and computeLocalClosestPoints on a block of query points.  A host that keeps the reference's MPI ring
// (DistributedClosestPointImpl.hpp:737-851) calls computeLocalClosestPoints where the reference does (:768, :826);
// axom_b200/distributed_closest_point.py is the single-box alternative over NCCL collectives.
#ifndef AXOM_B200_DISTRIBUTED_CLOSEST_POINT_HPP_
#define AXOM_B200_DISTRIBUTED_CLOSEST_POINT_HPP_

#include <limits>
#include <string>
#include <vector>

#include "BVH.hpp"

namespace axom_b200
{
namespace quest
{
template <int DIM = 3, typename ExecSpace = B200_EXEC>
class DistributedClosestPoint
{
  static_assert(DIM == 2 || DIM == 3, "DistributedClosestPoint is 2-D or 3-D");

public:
  using PointType = primal::Point<double, DIM>;

  explicit DistributedClosestPoint(int device = 0, int rank = 0) : m_rank(rank) { check(axb_dcp_create(&m_h, DIM, device)); }
  DistributedClosestPoint(const DistributedClosestPoint&) = delete;
  DistributedClosestPoint& operator=(const DistributedClosestPoint&) = delete;
  ~DistributedClosestPoint()
  {
    if(m_h) axb_dcp_destroy(m_h);
  }

  void setRank(int rank) { m_rank = rank; }  // setMpiCommunicator -> MPI_Comm_rank (:285-290)

  // setDistanceThreshold (DistributedClosestPoint.cpp:126-130): distances above the threshold are ignored
  void setDistanceThreshold(double threshold)
  {
    if(threshold < 0.0)
    {
      error_handler()(AXB_ERR_BAD_ARG, "Distance threshold must be non-negative.");
      return;
    }
    check(axb_dcp_set_squared_distance_threshold(m_h, threshold * threshold));
  }

  // setObjectMesh / importObjectPoints (:590-649): the points of all local domains, flattened in domain order;
  // domain d gets id d unless domainIds names it (state/domain_id)
  void setObjectMesh(const std::vector<std::vector<PointType>>& domains, const std::vector<IndexType>& domainIds = {})
  {
    std::vector<double> coords;
    std::vector<IndexType> ids;
    for(std::size_t d = 0; d < domains.size(); ++d)
    {
      const IndexType id = d < domainIds.size() ? domainIds[d] : (IndexType)d;
      for(const PointType& p : domains[d])
      {
        for(int k = 0; k < DIM; ++k) coords.push_back(p[k]);
        ids.push_back(id);
      }
    }
    check(axb_dcp_set_object_points(m_h, coords.data(), ids.data(), (IndexType)ids.size(), AXB_MEM_HOST));
  }

  // generateBVHTree (:651-668)
  bool generateBVHTree()
  {
    const int st = axb_dcp_generate_bvh_tree(m_h);
    check(st);
    return st == AXB_OK;
  }

  // computeLocalClosestPoints (:905-1079) on one block of query points; the arrays are the xferDom fields and live in
  // the caller's memory space (host or device).  isFirst initialises them (-1 / signalling NaN).
  void computeLocalClosestPoints(const PointType* queryPts, IndexType qPtCount, bool isFirst, IndexType* cp_index,
                                 IndexType* cp_domain_index, IndexType* cp_rank, PointType* cp_coords, double* cp_distance = nullptr) const
  {
    check(axb_dcp_compute_local_closest_points(m_h, m_rank, reinterpret_cast<const double*>(queryPts), qPtCount, isFirst ? 1 : 0, cp_index,
                                               cp_domain_index, cp_rank, reinterpret_cast<double*>(cp_coords), cp_distance, AXB_MEM_AUTO));
  }

  axb_dcp* handle() const { return m_h; }

private:
  axb_dcp* m_h = nullptr;
  int m_rank;
};

}  // namespace quest
}  // namespace axom_b200

#endif  // AXOM_B200_DISTRIBUTED_CLOSEST_POINT_HPP_
