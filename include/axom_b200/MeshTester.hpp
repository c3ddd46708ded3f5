// axom_b200/MeshTester.hpp -- C++ header shim keeping the names of quest::findTriMeshIntersectionsBVH
// (quest/MeshTester.hpp:67-104) and primal::intersect(Triangle3, Triangle3, includeBoundary, EPS)
// (primal/operators/intersect.hpp:64-71) on top of the C ABI in axb200.h.  Header-only, C++14.
//
// The reference takes a mint::UnstructuredMesh<SINGLE_SHAPE>*; everything the candidate finder reads from it
// (quest/detail/MeshTester_detail.hpp:158-199) is the node coordinates and the triangle connectivity, i.e. the
// quest::SurfaceMesh extract of SignedDistance.hpp with nodes_per_cell == 3.  The BVH walk, the i < j filter and
// the exact triangle-triangle test run in one kernel on the device; pairs come back in the SEQ_EXEC order.
#ifndef AXOM_B200_MESH_TESTER_HPP_
#define AXOM_B200_MESH_TESTER_HPP_

#include <utility>
#include <vector>

#include "SignedDistance.hpp"

namespace axom_b200
{
namespace primal
{
// minimal primal::Triangle<double,3>: three vertices, contiguous (72 bytes), operator[]
template <typename T, int NDIMS>
class Triangle
{
public:
  using PointType = Point<T, NDIMS>;
  Triangle() = default;
  Triangle(const PointType& A, const PointType& B, const PointType& C) : m_points {A, B, C} { }
  PointType& operator[](int i) { return m_points[i]; }
  const PointType& operator[](int i) const { return m_points[i]; }

private:
  PointType m_points[3];
};

// primal::intersect(t1, t2, includeBoundary, EPS) -- one pair, evaluated on the device (host triangles are staged)
inline bool intersect(const Triangle<double, 3>& t1, const Triangle<double, 3>& t2, bool includeBoundary = false, double EPS = 1E-08,
                      int device = 0)
{
  static_assert(sizeof(Triangle<double, 3>) == 72, "Triangle<double,3> must be 9 contiguous doubles");
  uint8_t out = 0;
  check(axb_tri_tri_intersect(device, reinterpret_cast<const double*>(&t1), reinterpret_cast<const double*>(&t2), 1, AXB_MEM_HOST,
                              includeBoundary ? 1 : 0, EPS, &out));
  return out != 0;
}

// batched form: n pairs in host or device memory (AXB_MEM_AUTO), out[i] = 0 / 1 in the same space
inline void intersect(const Triangle<double, 3>* t1, const Triangle<double, 3>* t2, std::int64_t n, uint8_t* out,
                      bool includeBoundary = false, double EPS = 1E-08, int device = 0)
{
  check(axb_tri_tri_intersect(device, reinterpret_cast<const double*>(t1), reinterpret_cast<const double*>(t2), n, AXB_MEM_AUTO,
                              includeBoundary ? 1 : 0, EPS, out));
}
}  // namespace primal

namespace quest
{
// findTriMeshIntersectionsBVH<ExecSpace, FloatType>(surface_mesh, intersections, degenerateIndices, intersectionThreshold)
// (quest/MeshTester.hpp:67-83).  FloatType is the BVH's; like the reference's default it is double here.
template <typename ExecSpace = B200_EXEC, typename FloatType = double>
void findTriMeshIntersectionsBVH(const SurfaceMesh* surface_mesh, std::vector<std::pair<int, int>>& intersections,
                                 std::vector<int>& degenerateIndices, double intersectionThreshold = 1E-8, int device = 0)
{
  static_assert(sizeof(FloatType) == 8, "the mesh tester's BVH is built in double");
  intersections.clear();
  degenerateIndices.clear();
  if(!surface_mesh || surface_mesh->nodes_per_cell != 3 || surface_mesh->cell_node_offsets)
  {
    error_handler()(AXB_ERR_BAD_ARG, "findTriMeshIntersectionsBVH needs a single-shape triangle mesh");
    return;
  }
  axb_meshtester* mt = nullptr;
  const SurfaceMesh& m = *surface_mesh;
  int st = axb_meshtester_create(&mt, device, m.x, m.y, m.z, m.num_nodes, m.cells_to_nodes, m.num_cells, AXB_MEM_AUTO);
  check(st);
  if(st != AXB_OK) return;
  int32_t *first = nullptr, *second = nullptr, *deg = nullptr;
  std::int64_t np = 0, nd = 0;
  st = axb_meshtester_find_intersections(mt, intersectionThreshold, AXB_MEM_HOST, &first, &second, &np);
  if(st == AXB_OK) st = axb_meshtester_get_degenerate(mt, AXB_MEM_HOST, &deg, &nd);
  if(st == AXB_OK)
  {
    intersections.resize((std::size_t)np);
    for(std::int64_t i = 0; i < np; ++i) intersections[(std::size_t)i] = {first[i], second[i]};
    degenerateIndices.assign(deg, deg + nd);
  }
  axb_meshtester_free(mt, first, AXB_MEM_HOST);
  axb_meshtester_free(mt, second, AXB_MEM_HOST);
  axb_meshtester_free(mt, deg, AXB_MEM_HOST);
  axb_meshtester_destroy(mt);
  check(st);
}
}  // namespace quest
}  // namespace axom_b200

#endif  // AXOM_B200_MESH_TESTER_HPP_
