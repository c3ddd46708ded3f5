// mc_oracle.cpp -- CPU restatement of quest::MarchingCubes (LLNL/axom v0.11.0), SEQ_EXEC path.
//
// TEST INFRASTRUCTURE ONLY: the parity checker for axom_b200/csrc/mc.cuh.  Only tests/, __graft_entry__.smoke()
// and the CPU-baseline legs of bench.py / tools/bench_configs.py may load the library built from it.
//
// What is restated (paths relative to /root/reference/src/axom), step by step and array by array, as the
// reference's SEQ_EXEC / hybridParallel path runs it:
//   quest/MarchingCubes.cpp:107-147                 computeIsocontour: per-domain mark + scan, output growth, facets, domain ids
//   quest/detail/MarchingCubesImpl.hpp:157-362      markCrossings / computeCaseId / computeCrossingCase
//   quest/detail/MarchingCubesImpl.hpp:487-573      scanCrossings_hybridParallel
//   quest/detail/MarchingCubesImpl.hpp:575-808      computeFacets / get_corner_coords_and_values / linear_interp
//   core/MDMapping.hpp:182-198,255-275,361-371      initializeShape(slowestDirs), initializeStrides, toMultiIndex
//
// PINNING.  Parity is PINNED to the real reference.  quest::MarchingCubes needs Conduit (MarchingCubesImpl.hpp:8-12), an
// external library that is not in this image; oracle/build_ref.py therefore compiles the UNMODIFIED reference sources
// (quest/MarchingCubes.cpp, quest/detail/MarchingCubesSingleDomain.cpp and the headers they include, MeshViewUtil.hpp among
// them) against a ~200-line MOCK of Conduit's Node (oracle/conduit_stub/: a named tree of external arrays, no algorithm)
// into oracle/_ref/libaxom_ref.so, and oracle/ref_mc_driver.cpp drives the reference's public API
// (setMesh / setFunctionField / setMaskValue / computeIsocontour / getContour*).  tests/test_oracle_golden.py checks
//   * this restatement == the real reference, bit for bit, for both data-parallel variants (hybridParallel, fullParallel),
//     2-D and 3-D, multi-domain, ghost layers, row- and column-major fields, masks, curvilinear coordinates and two
//     accumulated contour values (live when oracle/_ref exists, and through tests/golden/mc_*.npz made by
//     tests/golden/make_golden.py from the real reference);
//   * the packed case tables == cases2D / cases3D / num_segments / num_triangles of marching_cubes_lookup.hpp;
//   * the MDMapping pieces (direction order, case-id strides, toMultiIndex) == axom::MDMapping.
// The three result checks of the reference's own test driver (quest/examples/quest_marching_cubes_example.cpp:
// checkContourSurface :1152, checkContourCellLimits :1250, checkCellsContainingContour :1380) are applied as well
// (tests/test_marching_cubes.py).
//
// Compile with -ffp-contract=off, like the reference's x86-64 Release build (no FMA in p1 + w * (p2 - p1)).
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../axom_b200/csrc/mc_tables.h"  // generated, nibble-packed case tables (pinned by test_mc_tables_match_reference)

namespace
{
// same layout as axb_mc_domain (include/axb200.h)
struct McDomain
{
  int64_t cell_shape[3];
  const double* coords[3];
  int64_t coords_strides[3];
  const double* fcn;
  int64_t fcn_strides[3];
  const int32_t* mask;
  int64_t mask_strides[3];
  int64_t domain_id;
};

const uint64_t kCases3D[256] = {AXB_MC_CASES3D_WORDS};
const uint16_t kCases2D[16] = {AXB_MC_CASES2D_WORDS};

// cases_table(iCase, iEdge) (MarchingCubesImpl.hpp:824-854); -1 = unused
int cases_table(int dim, int iCase, int iEdge)
{
  const uint64_t w = dim == 2 ? kCases2D[iCase] : kCases3D[iCase];
  const int v = (int)((w >> (4 * iEdge)) & 0xF);
  return v == 0xF ? -1 : v;
}
// num_contour_cells (:814-843) == number of used entries / DIM
int num_contour_cells(int dim, int iCase)
{
  int used = 0;
  const int width = dim == 2 ? 4 : 16;
  while(used < width && cases_table(dim, iCase, used) >= 0) ++used;
  return used / dim;
}

struct Mapping  // axom::MDMapping<DIM>
{
  int dim;
  int64_t strides[3];
  int slowest[3];
  // initializeStrides(strides, ROW) (core/MDMapping.hpp:255-275)
  void from_strides(int d_, const int64_t* s)
  {
    dim = d_;
    for(int d = 0; d < dim; ++d)
    {
      strides[d] = s[d];
      slowest[d] = d;
    }
    for(int s0 = 0; s0 < dim; ++s0)
      for(int d = s0; d < dim; ++d)
        if(strides[slowest[s0]] < strides[slowest[d]]) std::swap(slowest[s0], slowest[d]);
  }
  // initializeShape(shape, slowestDirs) (:182-198)
  void from_shape(int d_, const int64_t* shape, const int* slowestDirs)
  {
    dim = d_;
    for(int d = 0; d < dim; ++d) slowest[d] = slowestDirs[d];
    strides[slowest[dim - 1]] = 1;
    for(int d = dim - 2; d >= 0; --d)
    {
      const int dir = slowest[d], faster = slowest[d + 1];
      strides[dir] = strides[faster] * shape[faster];
    }
  }
  // toMultiIndex (:361-371)
  void to_multi(int64_t flat, int64_t* idx) const
  {
    for(int d = 0; d < dim; ++d)
    {
      const int dir = slowest[d];
      idx[dir] = flat / strides[dir];
      flat -= idx[dir] * strides[dir];
    }
  }
};

inline bool is_nearly_equal(double a, double b) { return std::fabs(a - b) <= 1.0e-8; }  // core/utilities/Utilities.hpp:317-321

inline int64_t dot(int dim, const int64_t* idx, const int64_t* strides)
{
  int64_t o = 0;
  for(int d = 0; d < dim; ++d) o += idx[d] * strides[d];
  return o;
}

// corner n of cell (i,j[,k]) as node offsets -- the literal lists of :329-333 (2-D) and :349-357 / :685-701 (3-D)
const int kCorner2D[4][2] = {{0, 0}, {1, 0}, {1, 1}, {0, 1}};
const int kCorner3D[8][3] = {{1, 0, 0}, {1, 1, 0}, {0, 1, 0}, {0, 0, 0}, {1, 0, 1}, {1, 1, 1}, {0, 1, 1}, {0, 0, 1}};

struct Single  // MarchingCubesImpl<DIM, SEQ_EXEC, SEQ_EXEC> for one domain
{
  int dim;
  const McDomain* dom;
  Mapping caseMap;
  int64_t cellCount;
  std::vector<uint16_t> caseIds;
  std::vector<int32_t> crossingParentIds, facetIncrs, firstFacetIds;
  std::vector<int16_t> crossingCases;
  int64_t crossingCount = 0, facetCount = 0;

  void corner_node(const int64_t* c, int n, int64_t* node) const
  {
    for(int d = 0; d < dim; ++d) node[d] = c[d] + (dim == 2 ? kCorner2D[n][d] : kCorner3D[n][d]);
  }

  void markCrossings(double contourVal, int maskVal)
  {
    cellCount = 1;
    for(int d = 0; d < dim; ++d) cellCount *= dom->cell_shape[d];
    caseIds.assign((size_t)cellCount, 0);  // :161-162
    if(cellCount == 0) return;
    Mapping fcnMap;
    fcnMap.from_strides(dim, dom->fcn_strides);  // :165
    caseMap.from_shape(dim, dom->cell_shape, fcnMap.slowest);  // :166
    const int ncorner = dim == 2 ? 4 : 8;
    int64_t c[3] = {0, 0, 0};
    const int64_t nk = dim == 3 ? dom->cell_shape[2] : 1;
    for(c[2] = 0; c[2] < nk; ++c[2])
      for(c[1] = 0; c[1] < dom->cell_shape[1]; ++c[1])
        for(c[0] = 0; c[0] < dom->cell_shape[0]; ++c[0])
        {
          const bool useZone = !dom->mask || dom->mask[dot(dim, c, dom->mask_strides)] == maskVal;  // :325 / :345
          if(!useZone) continue;
          int index = 0;
          for(int n = 0; n < ncorner; ++n)  // computeCrossingCase :307-319
          {
            int64_t node[3];
            corner_node(c, n, node);
            if(dom->fcn[dot(dim, node, dom->fcn_strides)] >= contourVal) index |= (1 << n);
          }
          caseIds[(size_t)dot(dim, c, caseMap.strides)] = (uint16_t)index;
        }
  }

  void scanCrossings()  // scanCrossings_hybridParallel :487-573
  {
    int64_t vsum = 0;
    for(int64_t n = 0; n < cellCount; ++n) vsum += bool(num_contour_cells(dim, caseIds[(size_t)n]));
    crossingCount = vsum;
    crossingParentIds.assign((size_t)crossingCount, 0);
    crossingCases.assign((size_t)crossingCount, 0);
    facetIncrs.assign((size_t)crossingCount, 0);
    firstFacetIds.assign((size_t)crossingCount + 1, 0);
    int64_t crossingId = 0;
    for(int64_t n = 0; n < cellCount; ++n)
    {
      const int caseId = caseIds[(size_t)n];
      const int ccc = num_contour_cells(dim, caseId);
      if(ccc != 0)
      {
        crossingParentIds[(size_t)crossingId] = (int32_t)n;
        crossingCases[(size_t)crossingId] = (int16_t)caseId;
        facetIncrs[(size_t)crossingId] = ccc;
        ++crossingId;
      }
    }
    for(int64_t i = 1; i < 1 + crossingCount; ++i) firstFacetIds[(size_t)i] = firstFacetIds[(size_t)i - 1] + facetIncrs[(size_t)i - 1];
    facetCount = firstFacetIds[(size_t)crossingCount];
  }

  // linear_interp :706-807
  void linear_interp(int edgeIdx, const double (*cornerCoords)[3], const double* nodeValues, double contourVal, double* crossingPt) const
  {
    static const int hex_edge_table[] = {0, 1, 1, 2, 2, 3, 3, 0, 4, 5, 5, 6, 6, 7, 7, 4, 0, 4, 1, 5, 2, 6, 3, 7};
    int n1, n2;
    if(dim == 2)
    {
      n1 = edgeIdx;
      n2 = (edgeIdx == 3) ? 0 : edgeIdx + 1;
    }
    else
    {
      n1 = hex_edge_table[edgeIdx * 2];
      n2 = hex_edge_table[edgeIdx * 2 + 1];
    }
    const double f1 = nodeValues[n1], f2 = nodeValues[n2];
    const double* p1 = cornerCoords[n1];
    const double* p2 = cornerCoords[n2];
    if(is_nearly_equal(contourVal, f1) || is_nearly_equal(f1, f2))
    {
      for(int d = 0; d < dim; ++d) crossingPt[d] = p1[d];
      return;
    }
    if(is_nearly_equal(contourVal, f2))
    {
      for(int d = 0; d < dim; ++d) crossingPt[d] = p2[d];
      return;
    }
    const double ptiny = 1.0e-50;  // primal::PRIMAL_TINY (primal/constants.hpp)
    const double df = f2 - f1 + ptiny;
    const double w = (contourVal - f1) / df;
    for(int d = 0; d < dim; ++d) crossingPt[d] = p1[d] + w * (p2[d] - p1[d]);
  }

  // outBase: global id of the facet stored at element 0 of the three output arrays (0 in the reference, whose arrays hold the whole contour)
  void computeFacets(double contourVal, int64_t facetIndexOffset, int64_t outBase, int32_t* facetNodeIds, double* facetNodeCoords,
                     int32_t* facetParentIds) const
  {
    const int ncorner = dim == 2 ? 4 : 8;
    for(int64_t crossingId = 0; crossingId < crossingCount; ++crossingId)
    {
      const int64_t parentCellId = crossingParentIds[(size_t)crossingId];
      const int caseId = crossingCases[(size_t)crossingId];
      double cornerCoords[8][3], cornerValues[8];
      int64_t c[3];
      caseMap.to_multi(parentCellId, c);  // :653 / :679
      for(int n = 0; n < ncorner; ++n)
      {
        int64_t node[3];
        corner_node(c, n, node);
        const int64_t xo = dot(dim, node, dom->coords_strides);
        for(int d = 0; d < dim; ++d) cornerCoords[n][d] = dom->coords[d][xo];
        cornerValues[n] = dom->fcn[dot(dim, node, dom->fcn_strides)];
      }
      const int64_t additionalFacets = firstFacetIds[(size_t)crossingId + 1] - firstFacetIds[(size_t)crossingId];
      const int64_t firstFacetId = facetIndexOffset + firstFacetIds[(size_t)crossingId];
      for(int64_t fId = 0; fId < additionalFacets; ++fId)
      {
        const int64_t newFacetId = firstFacetId + fId;
        const int64_t firstCornerId = newFacetId * dim;
        facetParentIds[newFacetId - outBase] = (int32_t)parentCellId;
        for(int d = 0; d < dim; ++d)
        {
          const int64_t newCornerId = firstCornerId + d;
          facetNodeIds[(newFacetId - outBase) * dim + d] = (int32_t)newCornerId;
          const int edge = cases_table(dim, caseId, (int)(fId * dim + d));
          linear_interp(edge, cornerCoords, cornerValues, contourVal, facetNodeCoords + (newCornerId - outBase * dim) * dim);
        }
      }
    }
  }
};
}  // namespace

extern "C" {

// MarchingCubes::computeIsocontour (MarchingCubes.cpp:107-147) for `ndom` domains, starting from a contour that already
// holds `first_facet` facets (their storage is the caller's business: outputs here hold only the NEW facets, but node ids
// are numbered globally as the reference does).  Returns the number of new facets; arrays are malloc'ed, release with
// axo_mc_free.  num_threads is unused (the reference's SEQ path is what this restates).
int64_t axo_mc_compute_isocontour(int ndims, const McDomain* doms, int32_t ndom, double contour_val, int mask_val, int64_t first_facet,
                                  int32_t** facet_node_ids, double** node_coords, int32_t** facet_parent_ids, int32_t** facet_domain_ids)
{
  std::vector<Single> singles((size_t)ndom);
  std::vector<int64_t> facetIndexOffsets((size_t)ndom);
  int64_t facetCount = first_facet;
  for(int d = 0; d < ndom; ++d)
  {
    Single& s = singles[(size_t)d];
    s.dim = ndims;
    s.dom = &doms[d];
    s.markCrossings(contour_val, mask_val);
    s.scanCrossings();
    facetIndexOffsets[(size_t)d] = facetCount;
    facetCount += s.facetCount;
  }
  const int64_t added = facetCount - first_facet;
  const size_t n = (size_t)(added > 0 ? added : 1);
  int32_t* ids = (int32_t*)malloc(sizeof(int32_t) * n * ndims);
  double* xyz = (double*)malloc(sizeof(double) * n * ndims * ndims);
  int32_t* par = (int32_t*)malloc(sizeof(int32_t) * n);
  int32_t* did = (int32_t*)malloc(sizeof(int32_t) * n);
  for(int d = 0; d < ndom; ++d)
  {
    singles[(size_t)d].computeFacets(contour_val, facetIndexOffsets[(size_t)d], first_facet, ids, xyz, par);
    const int64_t cnt = (d < ndom - 1 ? facetIndexOffsets[(size_t)d + 1] : facetCount) - facetIndexOffsets[(size_t)d];
    for(int64_t f = 0; f < cnt; ++f) did[facetIndexOffsets[(size_t)d] - first_facet + f] = (int32_t)doms[d].domain_id;
  }
  *facet_node_ids = ids;
  *node_coords = xyz;
  *facet_parent_ids = par;
  *facet_domain_ids = did;
  return added;
}

void axo_mc_free(void* p) { free(p); }

// table access for the pinning test
int axo_mc_table(int dim, int iCase, int iEdge) { return cases_table(dim, iCase, iEdge); }
int axo_mc_num_contour_cells(int dim, int iCase) { return num_contour_cells(dim, iCase); }
// MDMapping pieces for the pinning test: slowestDirs + case strides of `shape` from the function strides; toMultiIndex
void axo_mc_mapping(int dim, const int64_t* fcn_strides, const int64_t* shape, int32_t* slowest, int64_t* case_strides)
{
  Mapping f, c;
  f.from_strides(dim, fcn_strides);
  c.from_shape(dim, shape, f.slowest);
  for(int d = 0; d < dim; ++d)
  {
    slowest[d] = f.slowest[d];
    case_strides[d] = c.strides[d];
  }
}
void axo_mc_to_multi_index(int dim, const int64_t* fcn_strides, const int64_t* shape, int64_t flat, int64_t* idx)
{
  Mapping f, c;
  f.from_strides(dim, fcn_strides);
  c.from_shape(dim, shape, f.slowest);
  c.to_multi(flat, idx);
}
}  // extern "C"
