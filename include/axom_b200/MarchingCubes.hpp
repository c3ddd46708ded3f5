// axom_b200/MarchingCubes.hpp -- C++ header shim keeping the names of quest::MarchingCubes
// (quest/MarchingCubes.hpp:107-306) on top of the C ABI in axb200.h (axb_mc_*).  Header-only, C++14.
//
// The reference reads a multi-domain Conduit Blueprint mesh; Conduit is not a dependency here, so setMesh takes what
// quest::MeshViewUtil extracts from each domain (quest/MeshViewUtil.hpp:454-482,560-607,803-848): the real cell shape,
// the coordinate / function / mask arrays with their Blueprint "strides" and "offsets", and state/domain_id.  A host
// code that holds a conduit::Node fills one StructuredDomain per child with the same paths (INTEGRATION.md shows it).
// Arrays may live in host or device memory (AXB_MEM_AUTO asks the driver); device arrays are used in place.
#ifndef AXOM_B200_MARCHING_CUBES_HPP_
#define AXOM_B200_MARCHING_CUBES_HPP_

#include <string>
#include <vector>

#include "BVH.hpp"

namespace axom_b200
{
namespace quest
{
// quest/MarchingCubes.hpp:49-54.  Both variants produce the same contour; the device path does not distinguish them.
enum class MarchingCubesDataParallelism
{
  byPolicy = 0,
  hybridParallel = 1,
  fullParallel = 2
};

// one Blueprint domain as MeshViewUtil sees it.  Strides are in elements; offsets are the ghost-layer offsets per direction.
struct StructuredDomain
{
  int ndims = 3;
  IndexType cellShape[3] = {0, 0, 0};       // topologies/<t>/elements/dims/{i,j,k}
  const double* coords[3] = {nullptr, nullptr, nullptr};  // coordsets/<c>/values/{x,y,z}
  IndexType coordsStrides[3] = {0, 0, 0};   // elements/dims/strides  (0,0,0: direction 0 fastest, no ghosts)
  IndexType coordsOffsets[3] = {0, 0, 0};   // elements/dims/offsets
  const double* fcn = nullptr;              // fields/<fcn>/values (vertex-associated)
  IndexType fcnStrides[3] = {0, 0, 0};      // fields/<fcn>/strides   (0,0,0: as above)
  IndexType fcnOffsets[3] = {0, 0, 0};
  const int* mask = nullptr;                // fields/<mask>/values (element-associated int32) or nullptr
  IndexType maskStrides[3] = {0, 0, 0};
  IndexType maskOffsets[3] = {0, 0, 0};
  IndexType domainId = -1;                  // state/domain_id; -1: the domain's position in the mesh
};

class MarchingCubes
{
public:
  using DomainIdType = IndexType;
  // (runtimePolicy, allocatorID, dataParallelism) of the reference; the execution space here is the B200 `device`
  explicit MarchingCubes(int device = 0, MarchingCubesDataParallelism dataParallelism = MarchingCubesDataParallelism::byPolicy)
    : m_device(device)
    , m_dataParallelism(dataParallelism)
  { }
  ~MarchingCubes()
  {
    if(m_h) axb_mc_destroy(m_h);
  }
  MarchingCubes(const MarchingCubes&) = delete;
  MarchingCubes& operator=(const MarchingCubes&) = delete;

  // setMesh(bpMesh, topologyName, maskField) + setFunctionField(fcnField) (MarchingCubes.cpp:47-105)
  void setMesh(const StructuredDomain* domains, IndexType numDomains)
  {
    const int nd = numDomains > 0 ? domains[0].ndims : 3;
    if(!m_h || nd != m_ndims)
    {
      if(m_h) axb_mc_destroy(m_h);
      m_h = nullptr;
      check(axb_mc_create(&m_h, nd, m_device));
      if(!m_h) return;
      m_ndims = nd;
    }
    std::vector<axb_mc_domain> v((std::size_t)(numDomains > 0 ? numDomains : 1));
    for(IndexType k = 0; k < numDomains; ++k)
    {
      const StructuredDomain& s = domains[k];
      axb_mc_domain& d = v[(std::size_t)k];
      std::int64_t cs[3], fs[3], ms[3], tn = 1, tc = 1;
      for(int i = 0; i < nd; ++i)  // Blueprint defaults: direction 0 fastest (MeshViewUtil.hpp:575-583,824-834)
      {
        cs[i] = s.coordsStrides[i] ? s.coordsStrides[i] : tn;
        fs[i] = s.fcnStrides[i] ? s.fcnStrides[i] : tn;
        ms[i] = s.maskStrides[i] ? s.maskStrides[i] : tc;
        tn *= s.cellShape[i] + 1;
        tc *= s.cellShape[i];
      }
      std::int64_t co = 0, fo = 0, mo = 0;
      for(int i = 0; i < 3; ++i)
      {
        const bool on = i < nd;
        d.cell_shape[i] = on ? s.cellShape[i] : 1;
        d.coords_strides[i] = on ? cs[i] : 0;
        d.fcn_strides[i] = on ? fs[i] : 0;
        d.mask_strides[i] = on ? ms[i] : 0;
        if(on)
        {
          co += s.coordsOffsets[i] * cs[i];
          fo += s.fcnOffsets[i] * fs[i];
          mo += s.maskOffsets[i] * ms[i];
        }
      }
      for(int i = 0; i < 3; ++i) d.coords[i] = (i < nd && s.coords[i]) ? s.coords[i] + co : nullptr;
      d.fcn = s.fcn ? s.fcn + fo : nullptr;
      d.mask = s.mask ? reinterpret_cast<const int32_t*>(s.mask) + mo : nullptr;
      d.domain_id = s.domainId >= 0 ? s.domainId : k;
    }
    check(axb_mc_set_mesh(m_h, v.data(), numDomains, AXB_MEM_AUTO));
  }
  void setMesh(const std::vector<StructuredDomain>& domains) { setMesh(domains.data(), (IndexType)domains.size()); }

  void setMaskValue(int maskVal)
  {
    m_maskVal = maskVal;
  }

  // adds the contour at contourVal to the contour mesh computed so far (MarchingCubes.cpp:107-147)
  void computeIsocontour(double contourVal = 0.0)
  {
    if(!m_h)
    {
      error_handler()(AXB_ERR_NOT_BUILT, "MarchingCubes::computeIsocontour before setMesh");
      return;
    }
    check(axb_mc_set_mask_value(m_h, m_maskVal));
    check(axb_mc_compute_isocontour(m_h, contourVal));
  }

  IndexType getContourCellCount() const
  {
    std::int64_t n = 0;
    if(m_h) check(axb_mc_get_contour_cell_count(m_h, &n));
    return (IndexType)n;
  }
  IndexType getContourFacetCount() const { return getContourCellCount(); }
  IndexType getContourNodeCount() const
  {
    std::int64_t n = 0;
    if(m_h) check(axb_mc_get_contour_node_count(m_h, &n));
    return (IndexType)n;
  }

  // device views (valid until the next compute / clear): getContourFacetCorners [cells][ndims], getContourNodeCoords
  // [nodes][ndims], getContourFacetParents [cells], getContourFacetDomainIds [cells] (quest/MarchingCubes.hpp:203-250)
  const IndexType* getContourFacetCorners() const { return view(0); }
  const double* getContourNodeCoords() const
  {
    const double* p = nullptr;
    if(m_h) check(axb_mc_get_contour_views(m_h, nullptr, &p, nullptr, nullptr));
    return p;
  }
  const IndexType* getContourFacetParents() const { return view(2); }
  const IndexType* getContourFacetDomainIds() const { return view(3); }

  // populateContourMesh (:169-233) into plain host vectors: what the reference appends to the mint::UnstructuredMesh
  // (nodes, cells) and stores in cellIdField / domainIdField
  void populateContourMesh(std::vector<double>& nodeCoords, std::vector<IndexType>& cellNodeIds, std::vector<IndexType>* cellIds = nullptr,
                           std::vector<DomainIdType>* domainIds = nullptr) const
  {
    const std::size_t n = (std::size_t)getContourCellCount(), d = (std::size_t)m_ndims;
    nodeCoords.resize(n * d * d);
    cellNodeIds.resize(n * d);
    if(cellIds) cellIds->resize(n);
    if(domainIds) domainIds->resize(n);
    if(n)
      check(axb_mc_copy_contour(m_h, AXB_MEM_HOST, cellNodeIds.data(), nodeCoords.data(), cellIds ? cellIds->data() : nullptr,
                                domainIds ? domainIds->data() : nullptr));
  }

  void clearOutput()
  {
    if(m_h) check(axb_mc_clear_output(m_h));
  }
  int spatialDimension() const { return m_ndims; }

private:
  const IndexType* view(int which) const
  {
    const int32_t* p[4] = {nullptr, nullptr, nullptr, nullptr};
    if(m_h) check(axb_mc_get_contour_views(m_h, &p[0], nullptr, &p[2], &p[3]));
    return p[which];
  }
  int m_device;
  MarchingCubesDataParallelism m_dataParallelism;
  axb_mc* m_h = nullptr;
  int m_ndims = 3;
  int m_maskVal = 1;
};
}  // namespace quest
}  // namespace axom_b200

#endif  // AXOM_B200_MARCHING_CUBES_HPP_
