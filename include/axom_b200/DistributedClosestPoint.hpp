// axom_b200/DistributedClosestPoint.hpp -- C++ header shim for the per-rank half of quest::DistributedClosestPoint
// (quest/DistributedClosestPoint.hpp:56-215, quest/detail/DistributedClosestPointImpl.hpp:551-1095) on top of the
// C ABI in axb200.h.  Header-only, C++14.
//
// The reference class takes Conduit blueprint nodes and an MPI communicator; here the blueprint nodes are reduced to
// what the implementation extracts from them -- interleaved coordinates per object domain, interleaved query
// coordinates, and the five xferDom arrays -- and the MPI communicator becomes a Communicator (an NCCL communicator the
// library owns: axb_comm).  computeClosestPoints is the whole distributed query (every rank calls it with its own query
// points); the exchange -- the reference's ring of Conduit messages, :737-851 -- runs inside the library as NCCL
// collectives on the handle's stream (csrc/comm.cuh).  computeLocalClosestPoints is the per-rank step for a host that
// prefers to keep the reference's MPI ring and call it where the reference does (:768, :826).
#ifndef AXOM_B200_DISTRIBUTED_CLOSEST_POINT_HPP_
#define AXOM_B200_DISTRIBUTED_CLOSEST_POINT_HPP_

#include <cstdint>
#include <limits>
#include <string>
#include <vector>

#include "BVH.hpp"

namespace axom_b200
{
namespace quest
{
// What stands where the reference has an MPI_Comm (DistributedClosestPoint.hpp:95-100): one per rank, created
// collectively.  Rank 0 calls Communicator::uniqueId() and gives the 128 bytes to the other ranks by whatever means the
// host has (MPI_Bcast, a file); then every rank constructs its Communicator.
class Communicator
{
public:
  using Id = std::vector<std::uint8_t>;
  static Id uniqueId()
  {
    Id id(AXB_COMM_ID_BYTES);
    check(axb_comm_get_unique_id(id.data()));
    return id;
  }
  Communicator(int nranks, int rank, const Id& id, int device = 0) { check(axb_comm_create(&m_c, nranks, rank, id.data(), device)); }
  Communicator(const Communicator&) = delete;
  Communicator& operator=(const Communicator&) = delete;
  ~Communicator()
  {
    if(m_c) axb_comm_destroy(m_c);
  }
  int rank() const
  {
    int r = 0;
    axb_comm_get_rank(m_c, &r, nullptr);
    return r;
  }
  int size() const
  {
    int n = 1;
    axb_comm_get_rank(m_c, nullptr, &n);
    return n;
  }
  axb_comm* handle() const { return m_c; }

private:
  axb_comm* m_c = nullptr;
};

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
  // setMpiCommunicator (DistributedClosestPoint.hpp:95-100): the communicator computeClosestPoints exchanges over
  void setCommunicator(const Communicator& comm)
  {
    m_comm = comm.handle();
    m_rank = comm.rank();
  }

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

  // computeClosestPoints (DistributedClosestPoint.hpp:157-166, impl :737-851): collective over the communicator; every
  // rank passes its own query points (possibly none) and receives the nearest object point of the whole machine for each:
  // cp_rank / cp_index / cp_domain_index / cp_coords / cp_distance, -1 and signalling NaN where nothing lies within the
  // threshold.  Without a communicator this is the single-rank query.  Output pointers may be null (setOutput(field, false)).
  void computeClosestPoints(const PointType* queryPts, IndexType qPtCount, IndexType* cp_index, IndexType* cp_domain_index, IndexType* cp_rank,
                            PointType* cp_coords, double* cp_distance = nullptr) const
  {
    check(axb_dcp_compute_closest_points(m_h, m_comm, reinterpret_cast<const double*>(queryPts), qPtCount, AXB_MEM_AUTO, cp_index, cp_domain_index,
                                         cp_rank, reinterpret_cast<double*>(cp_coords), cp_distance));
  }

  axb_dcp* handle() const { return m_h; }

private:
  axb_dcp* m_h = nullptr;
  axb_comm* m_comm = nullptr;  // borrowed
  int m_rank;
};

}  // namespace quest
}  // namespace axom_b200

#endif  // AXOM_B200_DISTRIBUTED_CLOSEST_POINT_HPP_
