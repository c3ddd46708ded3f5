// ref_mc_driver.cpp -- thin C ABI over the UNMODIFIED reference quest::MarchingCubes (LLNL/axom v0.11.0).
//
// TEST INFRASTRUCTURE ONLY (see oracle/build_ref.py).  No algorithm here: a Blueprint tree is assembled node by node
// from Python (axref_node_*) in the Conduit MOCK of oracle/conduit_stub/, handed to the reference's
// MarchingCubes(RuntimePolicy::seq, ...)::setMesh / setFunctionField / setMaskValue / computeIsocontour, and the four
// output views are copied out.  The reference sources (quest/MarchingCubes.cpp, quest/detail/MarchingCubesSingleDomain.cpp
// and the headers they include) are compiled where they lie with -DAXOM_USE_CONDUIT.
// The table / MDMapping accessors at the bottom read the reference's own headers for the pinning tests.
#include "axom/config.hpp"
#include "axom/core/MDMapping.hpp"
#include "axom/core/execution/runtime_policy.hpp"
#include "axom/core/memory_management.hpp"
#include "axom/quest/MarchingCubes.hpp"
#include "conduit_blueprint.hpp"

#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <string>

// ---- axom::MDMapping as MarchingCubesImpl::markCrossings uses it (:165-169) ------------------------------------------
template <int DIM>
static void ref_mapping(const int64_t* fcn_strides, const int64_t* shape, int32_t* slowest, int64_t* case_strides, int64_t flat, int64_t* idx)
{
  axom::StackArray<axom::IndexType, DIM> st, sh;
  for(int d = 0; d < DIM; ++d)
  {
    st[d] = (axom::IndexType)fcn_strides[d];
    sh[d] = (axom::IndexType)shape[d];
  }
  axom::MDMapping<DIM> fcnMapper(st);
  axom::MDMapping<DIM> caseMapper;
  caseMapper.initializeShape(sh, fcnMapper.slowestDirs());
  const auto mi = caseMapper.toMultiIndex((axom::IndexType)flat);
  for(int d = 0; d < DIM; ++d)
  {
    slowest[d] = fcnMapper.slowestDirs()[d];
    case_strides[d] = caseMapper.strides()[d];
    idx[d] = mi[d];
  }
}
extern "C" {

void* axref_node_new() { return new conduit::Node(); }
void axref_node_free(void* n) { delete static_cast<conduit::Node*>(n); }
void axref_node_set_string(void* n, const char* path, const char* value) { static_cast<conduit::Node*>(n)->fetch(path).set(std::string(value)); }
void axref_node_set_int(void* n, const char* path, int32_t value) { static_cast<conduit::Node*>(n)->fetch(path).set_int32(value); }
// kind: 0 int32, 1 int64, 2 float64; the array stays owned by the caller
void axref_node_set_external(void* n, const char* path, int kind, void* data, int64_t count)
{
  const conduit::DataType t = kind == 0 ? conduit::DataType::int32(count)
                                        : kind == 1 ? conduit::DataType::int64(count) : conduit::DataType::float64(count);
  static_cast<conduit::Node*>(n)->fetch(path).set_external(t, data);
}

// MarchingCubes(seq, host allocator, dataParallelism); setMesh; setFunctionField; setMaskValue; computeIsocontour for each
// contour value in turn (the contour mesh accumulates, MarchingCubes.hpp:160-164).  Returns the facet count; outputs malloc'ed.
int64_t axref_mc_run(void* mesh, const char* topology, const char* fcn_field, const char* mask_field, int mask_val, const double* contour_vals,
                     int num_contours, int data_parallelism, int* ndims_out, int32_t** facet_node_ids, double** node_coords,
                     int32_t** facet_parent_ids, int32_t** facet_domain_ids)
{
  using axom::quest::MarchingCubes;
  const conduit::Node& bp = *static_cast<conduit::Node*>(mesh);
  MarchingCubes mc(MarchingCubes::RuntimePolicy::seq, axom::execution_space<axom::SEQ_EXEC>::allocatorID(),
                   static_cast<axom::quest::MarchingCubesDataParallelism>(data_parallelism));
  mc.setMesh(bp, topology, mask_field ? std::string(mask_field) : std::string());
  mc.setFunctionField(fcn_field);
  mc.setMaskValue(mask_val);
  for(int c = 0; c < num_contours; ++c) mc.computeIsocontour(contour_vals[c]);
  const int64_t n = mc.getContourCellCount();
  const int64_t nn = mc.getContourNodeCount();
  const int dim = n > 0 ? (int)(nn / n) : 0;
  *ndims_out = dim;
  const size_t m = (size_t)(n > 0 ? n : 1), d = (size_t)(dim > 0 ? dim : 1);
  *facet_node_ids = (int32_t*)malloc(sizeof(int32_t) * m * d);
  *node_coords = (double*)malloc(sizeof(double) * m * d * d);
  *facet_parent_ids = (int32_t*)malloc(sizeof(int32_t) * m);
  *facet_domain_ids = (int32_t*)malloc(sizeof(int32_t) * m);
  if(n > 0)
  {
    memcpy(*facet_node_ids, mc.getContourFacetCorners().data(), sizeof(int32_t) * n * dim);
    memcpy(*node_coords, mc.getContourNodeCoords().data(), sizeof(double) * nn * dim);
    memcpy(*facet_parent_ids, mc.getContourFacetParents().data(), sizeof(int32_t) * n);
    memcpy(*facet_domain_ids, mc.getContourFacetDomainIds().data(), sizeof(int32_t) * n);
  }
  return n;
}

// ---- the reference's look-up tables (quest/detail/marching_cubes_lookup.hpp), read where they lie -------------------
int axref_mc_table(int dim, int iCase, int iEdge)
{
#define _MC_LOOKUP_CASES2D
#define _MC_LOOKUP_CASES3D
#include "axom/quest/detail/marching_cubes_lookup.hpp"
#undef _MC_LOOKUP_CASES2D
#undef _MC_LOOKUP_CASES3D
  return dim == 2 ? cases2D[iCase][iEdge] : cases3D[iCase][iEdge];
}
int axref_mc_num_contour_cells(int dim, int iCase)
{
#define _MC_LOOKUP_NUM_SEGMENTS
#define _MC_LOOKUP_NUM_TRIANGLES
#include "axom/quest/detail/marching_cubes_lookup.hpp"
#undef _MC_LOOKUP_NUM_SEGMENTS
#undef _MC_LOOKUP_NUM_TRIANGLES
  return dim == 2 ? num_segments[iCase] : num_triangles[iCase];
}

void axref_mc_mapping(int dim, const int64_t* fcn_strides, const int64_t* shape, int32_t* slowest, int64_t* case_strides, int64_t flat,
                      int64_t* idx)
{
  if(dim == 2)
    ref_mapping<2>(fcn_strides, shape, slowest, case_strides, flat, idx);
  else
    ref_mapping<3>(fcn_strides, shape, slowest, case_strides, flat, idx);
}

}  // extern "C"
