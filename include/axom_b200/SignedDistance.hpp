// axom_b200/SignedDistance.hpp -- C++ header shim keeping quest::SignedDistance's class and method
// names (quest/SignedDistance.hpp:147-397) on top of the C ABI in axb200.h.  Header-only, C++14.
//
// The reference constructor takes a `const mint::Mesh*`; everything it reads from it is what
// SD_GetUcdMeshData extracts (quest/SignedDistance.cpp:15-45): SoA node coordinates and the int32
// cells-to-nodes array of a single-shape (triangle or quad) or mixed-shape unstructured surface mesh.
// `SurfaceMesh` below is that extract; from a real mint::UnstructuredMesh it is filled with
//   {m->getCoordinateArray(0), (1), (2), m->getNumberOfNodes(), m->getCellNodesArray(), m->getNumberOfCells(),
//    nodes per cell, m->hasMixedCellTypes() ? m->getCellNodesOffsetsArray() : nullptr}
// (see INTEGRATION.md).  Unlike the reference, which keeps the mesh pointer and dereferences it on every
// query (:392,543-551), the mesh is copied to the GPU once in setMesh().
#ifndef AXOM_B200_SIGNED_DISTANCE_HPP_
#define AXOM_B200_SIGNED_DISTANCE_HPP_

#include "BVH.hpp"

namespace axom_b200
{
namespace quest
{
struct SurfaceMesh
{
  const double* x = nullptr;
  const double* y = nullptr;
  const double* z = nullptr;
  IndexType num_nodes = 0;
  const IndexType* cells_to_nodes = nullptr;
  IndexType num_cells = 0;
  int nodes_per_cell = 3;                        // 3 = triangles, 4 = quads (split (0,1,2),(0,2,3), :652-658)
  const IndexType* cell_node_offsets = nullptr;  // non-null selects a mixed triangle/quad mesh (num_cells+1 entries)
};

template <int NDIMS = 3, typename ExecSpace = B200_EXEC>
class SignedDistance
{
  static_assert(NDIMS == 3, "quest::SignedDistance is only ever instantiated in 3-D (quest/interface/signed_distance.cpp:116-120)");

public:
  using PointType = primal::Point<double, NDIMS>;
  using VectorType = primal::Vector<double, NDIMS>;
  using BoxType = primal::BoundingBox<double, NDIMS>;
  using BVHTreeType = spin::BVH<NDIMS, ExecSpace, double>;

  // SignedDistance(surfaceMesh, isWatertight, computeSign, allocatorID) (:410-425); allocatorID -> GPU ordinal
  SignedDistance(const SurfaceMesh* surfaceMesh, bool isWatertight, bool computeSign = true, int device = 0)
    : m_isInputWatertight(isWatertight), m_computeSign(computeSign), m_device(device)
  {
    setMesh(surfaceMesh, device);
  }
  SignedDistance(const SignedDistance&) = delete;
  SignedDistance& operator=(const SignedDistance&) = delete;
  ~SignedDistance()
  {
    if(m_sd) axb_sd_destroy(m_sd);
  }

  // setMesh (:427-504): mesh bounds, per-cell AABBs, BVH build with the default scale factor
  bool setMesh(const SurfaceMesh* m, int device = 0)
  {
    if(m_sd) axb_sd_destroy(m_sd);
    m_sd = nullptr;
    m_device = device;
    if(!m)
    {
      error_handler()(AXB_ERR_BAD_ARG, "surfaceMesh != nullptr");
      return false;
    }
    const int st = axb_sd_create(&m_sd, device, m->x, m->y, m->z, m->num_nodes, m->cells_to_nodes, m->cell_node_offsets, m->num_cells,
                                 m->nodes_per_cell, AXB_MEM_AUTO, m_isInputWatertight ? 1 : 0, m_computeSign ? 1 : 0);
    check(st);
    if(st != AXB_OK) return false;
    axb_bvh* b = nullptr;
    check(axb_sd_get_bvh(m_sd, &b));
    m_bvh.borrow(b);
    return true;
  }

  double computeDistance(double x, double y, double z = 0.0) const { return computeDistance(PointType {x, y, z}); }

  double computeDistance(const PointType& queryPnt) const
  {
    double phi = 0.0;
    computeDistances(1, &queryPnt, &phi);
    return phi;
  }

  double computeDistance(const PointType& queryPnt, PointType& closestPnt, VectorType& surfaceNormal) const
  {
    double phi = 0.0;
    computeDistances(1, &queryPnt, &phi, &closestPnt, &surfaceNormal);
    return phi;
  }

  // computeDistances(npts, queryPts, outSgnDist, outClosestPts, outNormals) (:527-605).  Outputs live in the
  // same memory space as the caller allocated them in (host or device).
  template <typename PointIndexable>
  void computeDistances(int npts, PointIndexable queryPts, double* outSgnDist, PointType* outClosestPts = nullptr,
                        VectorType* outNormals = nullptr) const
  {
    if(npts > 0 && outSgnDist == nullptr)
    {
      error_handler()(AXB_ERR_BAD_ARG, "outSgnDist != nullptr");
      return;
    }
    detail::Resolved<double> r;
    detail::indexable_traits<PointType, PointIndexable>::resolve(queryPts, npts, r);
    check(axb_sd_compute_distances(m_sd, &r.desc, npts, outSgnDist, reinterpret_cast<double*>(outClosestPts),
                                   reinterpret_cast<double*>(outNormals), AXB_MEM_AUTO));
  }

  // ---- a surface PARTITIONED over handles (BASELINE config 5; no counterpart in the reference class, which holds one mesh:
  // its distributed query is quest::DistributedClosestPoint).  Handles built with computeSign = false. ----
  // One GPU, parts evaluated in turn: dist (in/out, the caller's memory space) becomes min(dist, distance to this part);
  // the entry values bound the search.  Start from DBL_MAX.
  template <typename PointIndexable>
  void updateMinDistances(int npts, PointIndexable queryPts, double* dist) const
  {
    detail::Resolved<double> r;
    detail::indexable_traits<PointType, PointIndexable>::resolve(queryPts, npts, r);
    check(axb_sd_update_min_distances(m_sd, &r.desc, npts, dist, AXB_MEM_AUTO));
  }
  // One part per rank, the same query points on every rank: query kernels + ncclAllReduce(MIN) on one stream, inside the
  // library (comm: axom_b200::quest::Communicator::handle(), DistributedClosestPoint.hpp).  Collective.
  template <typename PointIndexable>
  void computeDistancesMinReduce(axb_comm* comm, int npts, PointIndexable queryPts, double* outDist) const
  {
    detail::Resolved<double> r;
    detail::indexable_traits<PointType, PointIndexable>::resolve(queryPts, npts, r);
    check(axb_sd_compute_distances_minreduce(m_sd, comm, &r.desc, npts, outDist, AXB_MEM_AUTO));
  }

  const BVHTreeType& getBVHTree() const { return m_bvh; }

  // bounding box of the mesh nodes (m_boxDomain, :455-487)
  BoxType getMeshBounds() const
  {
    double lo[3], hi[3];
    check(axb_sd_get_mesh_bounds(m_sd, lo, hi));
    return BoxType(PointType(lo), PointType(hi));
  }

  axb_sd* handle() const { return m_sd; }

private:
  bool m_isInputWatertight;
  bool m_computeSign;
  int m_device;
  axb_sd* m_sd = nullptr;
  BVHTreeType m_bvh;
};

}  // namespace quest
}  // namespace axom_b200

#endif  // AXOM_B200_SIGNED_DISTANCE_HPP_
