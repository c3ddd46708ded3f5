// tritri.cuh -- exact triangle-triangle intersection test (narrow phase downstream of findBoundingBoxes).
//
// Reference path replaced (paths relative to /root/reference/src/axom):
//   primal::intersect(Triangle<T,3>, Triangle<T,3>, includeBoundary, EPS)   primal/operators/intersect.hpp:64-71
//   detail::intersect_tri3D_tri3D and helpers                                primal/operators/detail/intersect_impl.hpp:156-590,1103-1484
//   fuzzy comparators                                                         primal/operators/detail/fuzzy_comparators.hpp:20-64
// Devillers-Guigue: plane-side classification of each triangle against the other's plane, a
// circular permutation that isolates the lone vertex, two orientation predicates on the plane-plane
// line, and a 2-D test for coplanar pairs.  Every comparison is the reference's fuzzy comparator and
// every product/sum is separately rounded (-fmad=false), so the boolean result is the reference's.
#pragma once
#include "common.cuh"

namespace axb
{
namespace tt
{
struct V3
{
  double x, y, z;
};
struct P2
{
  double x, y;
};
__device__ __forceinline__ V3 sub(const V3& h, const V3& t) { return {h.x - t.x, h.y - t.y, h.z - t.z}; }
// Vector::dot_product -> numerics::dot_product: left-to-right accumulation (primal/geometry/Vector.hpp:543-552)
__device__ __forceinline__ double dot(const V3& a, const V3& b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
// Vector::cross_product via 2x2 determinants (Vector.hpp:564-571, core/numerics/Determinants.hpp:31-38)
__device__ __forceinline__ V3 cross(const V3& u, const V3& v) { return {u.y * v.z - v.y * u.z, v.x * u.z - u.x * v.z, u.x * v.y - v.x * u.y}; }
// Triangle::normal (primal/geometry/Triangle.hpp:98-102)
__device__ __forceinline__ V3 tri_normal(const V3& a, const V3& b, const V3& c) { return cross(sub(b, a), sub(c, a)); }
// Vector::unitVector (Vector.hpp:477-493): v *= 1/sqrt(|v|^2), or e_0 below PRIMAL_TINY
__device__ __forceinline__ V3 unit(const V3& v)
{
  const double len2 = dot(v, v);
  if(len2 >= 1e-50)
  {
    const double s = 1. / sqrt(len2);
    return {v.x * s, v.y * s, v.z * s};
  }
  return {1., 0., 0.};
}
__device__ __forceinline__ bool nearly_eq(double a, double b, double th) { return fabs(a - b) <= th; }  // core/utilities/Utilities.hpp isNearlyEqual
__device__ __forceinline__ bool is_gt(double x, double y, double e) { return (x > y) && !nearly_eq(x, y, e); }
__device__ __forceinline__ bool is_lt(double x, double y, double e) { return (x < y) && !nearly_eq(x, y, e); }
__device__ __forceinline__ bool is_geq(double x, double y, double e) { return !is_lt(x, y, e); }
__device__ __forceinline__ bool is_lpeq(double x, double y, bool inc, double e) { return (inc && nearly_eq(x, y, e)) ? true : is_lt(x, y, e); }
__device__ __forceinline__ bool is_gpeq(double x, double y, bool inc, double e) { return (inc && nearly_eq(x, y, e)) ? true : is_gt(x, y, e); }
// twoDcross (intersect_impl.hpp:539-542)
__device__ __forceinline__ double cross2(const P2& A, const P2& B, const P2& C) { return (A.x - C.x) * (B.y - C.y) - (A.y - C.y) * (B.x - C.x); }
__device__ __forceinline__ int sgn(double x) { return (0 < x) - (x < 0); }
// countZeros / nonzeroSignMatch / oneZeroOthersMatch (:548-590)
__device__ __forceinline__ int count_zeros(double x, double y, double z, double e)
{
  return (int)nearly_eq(x, 0., e) + (int)nearly_eq(y, 0., e) + (int)nearly_eq(z, 0., e);
}
__device__ __forceinline__ bool nonzero_sign_match(double x, double y, double z, double e)
{
  return !nearly_eq(x, 0., e) && !nearly_eq(y, 0., e) && !nearly_eq(z, 0., e) && sgn(x) == sgn(y) && sgn(x) == sgn(z);
}
__device__ __forceinline__ bool one_zero_others_match(double x, double y, double z, double e)
{
  return count_zeros(x, y, z, e) == 1 &&
    ((nearly_eq(x, 0., e) && is_gt(y * z, 0., e)) || (nearly_eq(y, 0., e) && is_gt(z * x, 0., e)) || (nearly_eq(z, 0., e) && is_gt(x * y, 0., e)));
}

// checkEdge (:1352-1390)
__device__ __forceinline__ bool check_edge(const P2& p1, const P2& q1, const P2& r1, const P2& p2, const P2& r2, bool b, double e)
{
  if(is_gpeq(cross2(r2, p2, q1), 0., b, e))
  {
    if(is_gpeq(cross2(r2, p1, q1), 0., b, e))
    {
      if(is_gpeq(cross2(p1, p2, q1), 0., b, e)) return true;
      return is_gpeq(cross2(p1, p2, r1), 0., b, e) && is_gpeq(cross2(q1, r1, p2), 0., b, e);
    }
    return false;
  }
  return is_gpeq(cross2(r2, p2, r1), 0., b, e) && is_gpeq(cross2(q1, r1, r2), 0., b, e) && is_gpeq(cross2(p1, p2, r1), 0., b, e);
}
// checkVertex (:1393-1484)
__device__ __forceinline__ bool check_vertex(const P2& p1, const P2& q1, const P2& r1, const P2& p2, const P2& q2, const P2& r2, bool b, double e)
{
  if(is_gpeq(cross2(r2, p2, q1), 0., b, e))
  {
    if(is_gpeq(cross2(q2, r2, q1), 0., b, e))
    {
      if(is_gpeq(cross2(p1, p2, q1), 0., b, e)) return is_lpeq(cross2(p1, q2, q1), 0., b, e);
      return is_gpeq(cross2(p1, p2, r1), 0., b, e) && is_lpeq(cross2(q1, p2, r1), 0., b, e);
    }
    return is_lpeq(cross2(p1, q2, q1), 0., b, e) && is_gpeq(cross2(q2, r2, r1), 0., b, e) && is_gpeq(cross2(q1, r1, q2), 0., b, e);
  }
  if(is_gpeq(cross2(r2, p2, r1), 0., b, e))
  {
    if(is_gpeq(cross2(q1, r1, r2), 0., b, e)) return is_gpeq(cross2(r1, p1, p2), 0., b, e);
    return is_gpeq(cross2(q1, r1, q2), 0., b, e) && is_gpeq(cross2(q2, r2, r1), 0., b, e);
  }
  return false;
}
// intersectPermuted2DTriangles (:1281-1349)
__device__ __forceinline__ bool permuted_2d(const P2& p1, const P2& q1, const P2& r1, const P2& p2, const P2& q2, const P2& r2, bool b, double e)
{
  if(is_gpeq(cross2(p2, q2, p1), 0., b, e))
  {
    if(is_gpeq(cross2(q2, r2, p1), 0., b, e))
    {
      if(is_gpeq(cross2(r2, p2, p1), 0., b, e)) return true;
      return check_edge(p1, q1, r1, p2, r2, b, e);
    }
    if(is_gpeq(cross2(r2, p2, p1), 0., b, e)) return check_edge(p1, q1, r1, r2, q2, b, e);
    return check_vertex(p1, q1, r1, p2, q2, r2, b, e);
  }
  if(is_gpeq(cross2(q2, r2, p1), 0., b, e))
  {
    if(is_gpeq(cross2(r2, p2, p1), 0., b, e)) return check_edge(p1, q1, r1, q2, p2, b, e);
    return check_vertex(p1, q1, r1, q2, r2, p2, b, e);
  }
  return check_vertex(p1, q1, r1, r2, p2, q2, b, e);
}
// TriangleIntersection2D (:1245-1278): make both triangles counter-clockwise
__device__ __forceinline__ bool tri2d(const P2* t1, const P2* t2, bool b, double e)
{
  const bool f1 = is_lt(cross2(t1[0], t1[1], t1[2]), 0., e);
  const bool f2 = is_lt(cross2(t2[0], t2[1], t2[2]), 0., e);
  return permuted_2d(t1[0], f1 ? t1[2] : t1[1], f1 ? t1[1] : t1[2], t2[0], f2 ? t2[2] : t2[1], f2 ? t2[1] : t2[2], b, e);
}
// intersectCoplanar3DTriangles (:1185-1242): project on the plane of largest normal component.
// Rare path, kept out of line; arguments by value so that the callers' vertices stay in registers.
__device__ __noinline__ bool coplanar(V3 p1, V3 q1, V3 r1, V3 p2, V3 q2, V3 r2, V3 n, bool b, double e)
{
  n.x = fabs(n.x);
  n.y = fabs(n.y);
  n.z = fabs(n.z);
  P2 a[3], c[3];
  if(is_gt(n.x, n.z, e) && is_geq(n.x, n.y, e))
  {
    a[0] = {q1.z, q1.y}; a[1] = {p1.z, p1.y}; a[2] = {r1.z, r1.y};
    c[0] = {q2.z, q2.y}; c[1] = {p2.z, p2.y}; c[2] = {r2.z, r2.y};
  }
  else if(is_gt(n.y, n.z, e) && is_geq(n.y, n.x, e))
  {
    a[0] = {q1.x, q1.z}; a[1] = {p1.x, p1.z}; a[2] = {r1.x, r1.z};
    c[0] = {q2.x, q2.z}; c[1] = {p2.x, p2.z}; c[2] = {r2.x, r2.z};
  }
  else
  {
    a[0] = {p1.x, p1.y}; a[1] = {q1.x, q1.y}; a[2] = {r1.x, r1.y};
    c[0] = {p2.x, p2.y}; c[1] = {q2.x, q2.y}; c[2] = {r2.x, r2.y};
  }
  return tri2d(a, c, b, e);
}
// intersectTwoPermutedTriangles (:457-474)
__device__ __forceinline__ bool two_permuted(const V3& p1, const V3& q1, const V3& r1, const V3& p2, const V3& q2, const V3& r2, bool b, double e)
{
  return is_lpeq(dot(sub(q2, q1), tri_normal(q1, p2, p1)), 0., b, e) && is_lpeq(dot(sub(r2, p1), tri_normal(p1, p2, r1)), 0., b, e);
}

// Which cyclic rotation / swap of a triangle isolates its lone vertex, from the signed plane distances of its
// three vertices -- the decision tree shared by :222-435 (triangle 1, acting on triangle 2's order) and
// :1118-1181 (triangle 2).  Returns rot in {0,1,2} (vertex order (rot, rot+1, rot+2)) and `swap_other`
// (exchange the 2nd and 3rd vertex of the OTHER triangle); coplanar = all three distances ~ 0.
struct Perm
{
  int rot;
  bool swap_other;
  bool coplanar;
};
__device__ __forceinline__ Perm classify(double dp, double dq, double dr, double e)
{
  if(is_gt(dp, 0., e))
  {
    if(is_gt(dq, 0., e)) return {2, true, false};
    if(is_gt(dr, 0., e)) return {1, true, false};
    return {0, false, false};
  }
  if(is_lt(dp, 0., e))
  {
    if(is_lt(dq, 0., e)) return {2, false, false};
    if(is_lt(dr, 0., e)) return {1, false, false};
    return {0, true, false};
  }
  if(is_lt(dq, 0., e))
  {
    if(is_geq(dr, 0., e)) return {1, true, false};
    return {0, false, false};
  }
  if(is_gt(dq, 0., e))
  {
    if(is_gt(dr, 0., e)) return {0, true, false};
    return {1, false, false};
  }
  if(is_gt(dr, 0., e)) return {2, false, false};
  if(is_lt(dr, 0., e)) return {2, true, false};
  return {0, false, true};
}

// element `rot` of the cyclic sequence (a, b, c) without dynamic indexing
__device__ __forceinline__ V3 rot3(const V3& a, const V3& b, const V3& c, int rot) { return rot == 0 ? a : (rot == 1 ? b : c); }

// intersect_tri3D_tri3D (:156-436).  t1, t2: three vertices each.
__device__ __forceinline__ bool tri_tri(const V3* t1, const V3* t2, bool b, double e)
{
  const V3 n2 = unit(tri_normal(t2[0], t2[1], t2[2]));
  const double dp1 = dot(sub(t1[0], t2[2]), n2), dq1 = dot(sub(t1[1], t2[2]), n2), dr1 = dot(sub(t1[2], t2[2]), n2);
  if(nonzero_sign_match(dp1, dq1, dr1, e)) return false;
  if(!b && (count_zeros(dp1, dq1, dr1, e) == 2 || one_zero_others_match(dp1, dq1, dr1, e))) return false;
  const V3 n1 = unit(tri_normal(t1[0], t1[1], t1[2]));
  const double d2[3] = {dot(sub(t2[0], t1[2]), n1), dot(sub(t2[1], t1[2]), n1), dot(sub(t2[2], t1[2]), n1)};
  if(nonzero_sign_match(d2[0], d2[1], d2[2], e)) return false;
  if(!b && (count_zeros(d2[0], d2[1], d2[2], e) == 2 || one_zero_others_match(d2[0], d2[1], d2[2], e))) return false;

  // step 3 (:222-435): rotate triangle 1, possibly swap vertices 1 and 2 of triangle 2 (with their distances)
  const Perm a = classify(dp1, dq1, dr1, e);
  if(a.coplanar) return coplanar(t1[0], t1[1], t1[2], t2[0], t2[1], t2[2], n1, b, e);
  const V3 p1 = rot3(t1[0], t1[1], t1[2], a.rot), q1 = rot3(t1[1], t1[2], t1[0], a.rot), r1 = rot3(t1[2], t1[0], t1[1], a.rot);
  const V3 u0 = t2[0], u1 = a.swap_other ? t2[2] : t2[1], u2 = a.swap_other ? t2[1] : t2[2];
  const double e0 = d2[0], e1 = a.swap_other ? d2[2] : d2[1], e2 = a.swap_other ? d2[1] : d2[2];
  // step 4 (:1118-1181): the same decision table for triangle 2; here "swap" exchanges q1 and r1
  const Perm c = classify(e0, e1, e2, e);
  if(c.coplanar) return coplanar(p1, q1, r1, u0, u1, u2, n1, b, e);
  const V3 p2 = rot3(u0, u1, u2, c.rot), q2 = rot3(u1, u2, u0, c.rot), r2 = rot3(u2, u0, u1, c.rot);
  return c.swap_other ? two_permuted(p1, r1, q1, p2, q2, r2, b, e) : two_permuted(p1, q1, r1, p2, q2, r2, b, e);
}

// Triangle::degenerate (Triangle.hpp:326-330): isNearlyEqual(0.5 * |normal|, 0, 1e-12)
__device__ __forceinline__ bool degenerate(const V3* t)
{
  const V3 n = tri_normal(t[0], t[1], t[2]);
  return nearly_eq(0.5 * sqrt(dot(n, n)), 0.0, 1.0e-12);
}

__device__ __forceinline__ void load_tri(const double* __restrict__ tris, long long i, V3* t)
{
  const double* p = tris + i * 9;
#pragma unroll
  for(int k = 0; k < 3; ++k) t[k] = {__ldg(p + 3 * k), __ldg(p + 3 * k + 1), __ldg(p + 3 * k + 2)};
}
// internal triangle record of the mesh tester: 96 bytes = 3 x 32 B (v0.xyz v1.x | v1.yz v2.xy | v2.z pad pad pad),
// fetched with three LDG.E.256 -- a lane reads a record of its own, so every load instruction is one L1 wavefront
constexpr int kTriRecDoubles = 12;
__device__ __forceinline__ void load_tri_rec(const double* __restrict__ recs, long long i, V3* t)
{
  const double* p = recs + i * kTriRecDoubles;
  const D4 a = ldg256(p), b = ldg256(p + 4), c = ldg256(p + 8);
  t[0] = {a.x, a.y, a.z};
  t[1] = {a.w, b.x, b.y};
  t[2] = {b.z, b.w, c.x};
}
}  // namespace tt

// Candidate filter of the fused broad + narrow phase (quest/detail/MeshTester_detail.hpp:236-283):
// keep candidate c of query i iff (i < c or !upper_only) and primal::intersect(tri_i, tri_c, include_boundary, eps)
struct TriTriFilter
{
  const double* __restrict__ query_tris;  // 96-byte records (tt::kTriRecDoubles), indexed by query id
  const double* __restrict__ tree_tris;   // 96-byte records, indexed by candidate (original box) id
  double eps;
  int upper_only;
  int include_boundary;
  __device__ __forceinline__ bool operator()(int qi, int cand) const
  {
    if(upper_only && !(qi < cand)) return false;
    tt::V3 a[3], b[3];
    tt::load_tri_rec(query_tris, qi, a);
    tt::load_tri_rec(tree_tris, cand, b);
    return tt::tri_tri(a, b, include_boundary != 0, eps);
  }
};

// no narrow phase: plain findPoints / findBoundingBoxes / findRays
struct NoFilter
{
  __device__ __forceinline__ bool operator()(int, int) const { return true; }
};

}  // namespace axb
