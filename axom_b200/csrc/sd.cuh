// sd.cuh -- quest::SignedDistance kernels (3-D, FP64, sm_100a).
//
// Reference path replaced (quest/SignedDistance.hpp):
//   setMesh            :427-504   mesh bounds, per-cell AABBs, BVH build
//   computeDistances   :527-605   fused nearest-surface traversal + sign
//   checkCandidate     :636-737   closest point on triangle + pseudo-normal state machine
//   getSurfaceNormal / computeSign :740-763
// Leaf arithmetic restated from primal (closest_point.hpp:162-290, squared_distance.hpp:62-100,
// Triangle.hpp:98-109,326-330,384-400, Vector.hpp:477-493,543-571).  All of it is compiled with
// -fmad=false: mul and add round separately, as in the reference's x86-64 build.
#pragma once
#include "common.cuh"
#include "traverse.cuh"

namespace axb
{
struct V3
{
  double x, y, z;
};
__device__ __forceinline__ V3 v3sub(const V3& h, const V3& t) { return {h.x - t.x, h.y - t.y, h.z - t.z}; }
__device__ __forceinline__ V3 v3add(const V3& a, const V3& b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
__device__ __forceinline__ V3 v3mul(const V3& a, double s) { return {a.x * s, a.y * s, a.z * s}; }
// Vector::dot_product (Vector.hpp:543-552): res = 0; res += u[d]*v[d] in order
__device__ __forceinline__ double v3dot(const V3& a, const V3& b)
{
  double r = 0.0;
  r += a.x * b.x;
  r += a.y * b.y;
  r += a.z * b.z;
  return r;
}
// Vector::cross_product via 2x2 determinants a00*a11 - a10*a01 (Vector.hpp:564-571, Determinants.hpp:31-38)
__device__ __forceinline__ V3 v3cross(const V3& u, const V3& v)
{
  return {u.y * v.z - v.y * u.z, v.x * u.z - u.x * v.z, u.x * v.y - v.x * u.y};
}
// Vector::unitVector (Vector.hpp:477-493): v *= 1./sqrt(len2) when len2 >= PRIMAL_TINY (1e-50)
__device__ __forceinline__ V3 v3unit(const V3& v)
{
  const double len2 = v3dot(v, v);
  if(len2 >= 1e-50)
  {
    const double s = 1. / sqrt(len2);
    return {v.x * s, v.y * s, v.z * s};
  }
  return {1.0, 0.0, 0.0};
}
__device__ __forceinline__ bool nearly_eq(double a, double b, double th) { return fabs(a - b) <= th; }
// detail::isLeq / isGeq (primal/operators/detail/fuzzy_comparators.hpp:26-45)
__device__ __forceinline__ bool is_leq(double x, double y, double e) { return !((x > y) && !nearly_eq(x, y, e)); }
__device__ __forceinline__ bool is_geq(double x, double y, double e) { return !((x < y) && !nearly_eq(x, y, e)); }

// primal::closest_point(Point, Triangle, int* loc, EPS) -- closest_point.hpp:162-290.
// Region tests in the reference's order: A, B, AB (with the !isNearlyEqual(d1,d3) guard), C, AC, BC, face.
//
// The reference returns from the first region that matches.  In a warp every lane tests a different triangle
// and lands in a different region, so the early returns execute 7 ways divergent (profiles/r1h: 3-5 of 32 lanes
// active on these lines, 30 % of the kernel's instructions).  Here all six dot products, the three sub-determinants
// and the seven region predicates are evaluated unconditionally, the FIRST matching region in the reference's order
// is selected, and the single division any region needs is issued once.  Every value that reaches the result is
// computed by the same expression (same operands, same order, no FMA) as in the reference, so the closest point and
// `loc` are bit-identical; values of regions not taken are discarded (a zero denominator there is harmless).
__device__ __forceinline__ V3 closest_point_tri(const V3& P, const V3& A, const V3& B, const V3& C, int& loc, double EPS)
{
  const V3 ab = v3sub(B, A), ac = v3sub(C, A), ap = v3sub(P, A), bp = v3sub(P, B), cp = v3sub(P, C);
  const double d1 = v3dot(ab, ap), d2 = v3dot(ac, ap);
  const double d3 = v3dot(ab, bp), d4 = v3dot(ac, bp);
  const double d5 = v3dot(ab, cp), d6 = v3dot(ac, cp);
  const double vc = d1 * d4 - d3 * d2;
  const double vb = d5 * d2 - d1 * d6;
  const double va = d3 * d6 - d5 * d4;
  const double d43 = d4 - d3, d56 = d5 - d6;
  const bool rA = is_leq(d1, 0, EPS) && is_leq(d2, 0, EPS);
  const bool rB = is_geq(d3, 0, EPS) && is_leq(d4, d3, EPS);
  const bool rAB = is_leq(vc, 0, EPS) && is_geq(d1, 0, EPS) && is_leq(d3, 0, EPS) && !nearly_eq(d1, d3, EPS);
  const bool rC = is_geq(d6, 0, EPS) && is_leq(d5, d6, EPS);
  const bool rAC = is_leq(vb, 0, EPS) && is_geq(d2, 0, EPS) && is_leq(d6, 0, EPS);
  const bool rBC = is_leq(va, 0, EPS) && is_geq(d43, 0, EPS) && is_geq(d56, 0, EPS);
  // first match wins: 0 A, 1 B, 2 AB, 3 C, 4 AC, 5 BC, 6 face
  const int region = rA ? 0 : (rB ? 1 : (rAB ? 2 : (rC ? 3 : (rAC ? 4 : (rBC ? 5 : 6)))));
  // the one quotient of the region: AB d1/(d1-d3), AC d2/(d2-d6), BC (d4-d3)/((d4-d3)+(d5-d6)), face 1/(va+vb+vc)
  const double num = region == 2 ? d1 : (region == 4 ? d2 : (region == 5 ? d43 : 1.0));
  const double den = region == 2 ? (d1 - d3) : (region == 4 ? (d2 - d6) : (region == 5 ? (d43 + d56) : (region == 6 ? (va + vb + vc) : 1.0)));
  const double q = num / den;
  // edge form  base + dir * q
  const V3 bc = v3sub(C, B);
  const V3 base = region == 5 ? B : A;
  const V3 dir = region == 2 ? ab : (region == 4 ? ac : bc);
  const V3 edge_pt = v3add(base, v3mul(dir, q));
  // face form  A + (ab * (vb * q) + ac * (vc * q))
  const double v = vb * q, w = vc * q;
  const V3 face_pt = v3add(A, v3add(v3mul(ab, v), v3mul(ac, w)));
  const V3 vert_pt = region == 0 ? A : (region == 1 ? B : C);
  loc = region == 0 ? 0 : (region == 1 ? 1 : (region == 2 ? -1 : (region == 3 ? 2 : (region == 4 ? -3 : (region == 5 ? -2 : 3)))));
  const bool is_vertex = region == 0 || region == 1 || region == 3;
  const bool is_face = region == 6;
  V3 r;
  r.x = is_vertex ? vert_pt.x : (is_face ? face_pt.x : edge_pt.x);
  r.y = is_vertex ? vert_pt.y : (is_face ? face_pt.y : edge_pt.y);
  r.z = is_vertex ? vert_pt.z : (is_face ? face_pt.z : edge_pt.z);
  return r;
}

// Triangle::angle (Triangle.hpp:384-400); idx in {0,1,2}, selected without indexing so the
// triangle stays in registers
__device__ __forceinline__ double tri_angle(const V3* t, int idx)
{
  const V3 pt = idx == 0 ? t[0] : (idx == 1 ? t[1] : t[2]);
  const V3 p1 = idx == 0 ? t[1] : (idx == 1 ? t[2] : t[0]);
  const V3 p2 = idx == 0 ? t[2] : (idx == 1 ? t[0] : t[1]);
  const V3 v1 = v3unit(v3sub(p1, pt));
  const V3 v2 = v3unit(v3sub(p2, pt));
  const double dp = v3dot(v1, v2);
  return acos(dp < -1.0 ? -1.0 : (dp > 1.0 ? 1.0 : dp));
}

// MinCandidate (quest/SignedDistance.hpp:159-175).  minTri is kept as (leaf position, sub-triangle)
// and re-read at the end; minElem / minCount are not observable through the API.
struct MinCand
{
  double minSq;
  V3 minPt;
  V3 sumN;
  int minType;  // -1 uninitialised, 0 vertex, 1 edge, 2 face (detail::ClosestPointLocType :105-112)
  int minPos;   // sorted leaf position of the closest element
  int minSub;   // 0 | 1: which triangle of a quad
};

__device__ __forceinline__ int loc_type(int loc) { return loc < 0 ? 1 : (loc <= 2 ? 0 : 2); }  // :115-137

// Leaf geometry, gathered at setMesh time into sorted-leaf order so a leaf visit is one
// contiguous read instead of the reference's leaf_nodes -> connectivity -> 3 coordinate arrays
// pointer chase.  One 96-byte record per leaf (3 x 32 B: 4 vertices x 3 doubles, the 4th unused for
// triangles), read with three 256-bit loads.
constexpr int kLeafDoubles = 12;
template <int NV>
__device__ __forceinline__ void load_leaf(const double* __restrict__ soup, int pos, V3* v)
{
  const D4* p = reinterpret_cast<const D4*>(soup + (size_t)pos * kLeafDoubles);
  const D4 a = ldg256(p), b = ldg256(p + 1);
  v[0] = {a.x, a.y, a.z};
  v[1] = {a.w, b.x, b.y};
  if(NV == 4)
  {
    const D4 c = ldg256(p + 2);
    v[2] = {b.z, b.w, c.x};
    v[NV - 1] = {c.y, c.z, c.w};
  }
  else
  {
    v[2] = {b.z, b.w, __ldg(soup + (size_t)pos * kLeafDoubles + 8)};
  }
}

// In a mixed triangle / quad mesh (UcdMeshData with cell_node_offsets, :44-96) the leaf records are
// the 4-vertex ones and a triangle cell marks its missing 4th vertex with NaN.
__device__ __forceinline__ bool has_fourth(const V3& v) { return v.x == v.x; }

// checkCandidate (:636-737) for one (sub-)triangle T
__device__ __forceinline__ void check_triangle(const V3& q, MinCand& m, const V3* T, int pos, int sub, bool computeNormal)
{
  constexpr double EPS = 1e-12;
  int loc;
  const V3 cp = closest_point_tri(q, T[0], T[1], T[2], loc, EPS);
  const V3 dq = v3sub(cp, q);
  const double sq = v3dot(dq, dq);
  const int type = loc_type(loc);
  const bool shared = (type != 2);
  bool upd;
  if(sq < m.minSq)
  {
    const V3 dm = v3sub(m.minPt, cp);
    const bool clear = !shared || (m.minType != type) || !nearly_eq(v3dot(dm, dm), 0., EPS);
    m.minSq = sq;
    m.minPt = cp;
    m.minType = type;
    m.minPos = pos;
    m.minSub = sub;
    if(computeNormal && clear) m.sumN = {0.0, 0.0, 0.0};
    upd = computeNormal && shared;
  }
  else
  {
    const V3 dm = v3sub(m.minPt, cp);
    upd = computeNormal && shared && (m.minType == type) && nearly_eq(v3dot(dm, dm), 0., EPS);
  }
  if(upd)
  {
    const V3 n = v3cross(v3sub(T[1], T[0]), v3sub(T[2], T[0]));  // Triangle::normal :98-102
    if(type == 1)
    {
      m.sumN = v3add(m.sumN, v3unit(n));
    }
    else
    {
      const double area = 0.5 * sqrt(v3dot(n, n));  // Triangle::area :105-109
      if(!nearly_eq(area, 0.0, 1.0e-12))            // !degenerate() :326-330
      {
        const double alpha = tri_angle(T, loc);
        m.sumN = v3add(m.sumN, v3mul(v3unit(n), alpha));
      }
    }
  }
}

template <int NV>
__device__ __forceinline__ void check_leaf(const double* __restrict__ soup, const V3& q, MinCand& m, int pos, bool computeNormal)
{
  V3 v[NV];
  load_leaf<NV>(soup, pos, v);
  {
    const V3 T[3] = {v[0], v[1], v[2]};
    check_triangle(q, m, T, pos, 0, computeNormal);
  }
  if(NV == 4 && has_fourth(v[NV - 1]))
  {
    const V3 T[3] = {v[0], v[2], v[NV - 1]};  // quads split (0,1,2),(0,2,3) :652-658
    check_triangle(q, m, T, pos, 1, computeNormal);
  }
}

// squared_distance(Point, Box) (squared_distance.hpp:77-100); caller guarantees a valid box
__device__ __forceinline__ double sqdist_point_box(const double* p, const Box<double, 3>& b)
{
  bool inside = true;
#pragma unroll
  for(int d = 0; d < 3; ++d) inside = inside && !(p[d] < b.lo[d] || p[d] > b.hi[d]);
  if(inside) return 0.0;
  double s = 0.0;
#pragma unroll
  for(int d = 0; d < 3; ++d)
  {
    const double c = (p[d] < b.lo[d]) ? b.lo[d] : ((p[d] > b.hi[d]) ? b.hi[d] : p[d]);  // clampVal
    const double v = c - p[d];
    s += v * v;
  }
  return s;
}

struct SdParams
{
  double dom_lo[3], dom_hi[3];  // m_boxDomain: bounds of the mesh nodes (:455-487)
  int watertight;
  int compute_sign;
};

// finish one query: sign (:583-592, :751-763), distance (:594), optional outputs (:595-603)
template <int NV>
__device__ __forceinline__ void sd_finish(const double* __restrict__ soup, const SdParams& prm, const V3& q, const MinCand& m, int qi,
                                           double* __restrict__ phi, double* __restrict__ cps, double* __restrict__ nrms)
{
  V3 nrm = m.sumN;
  if(m.minType == 2 && (prm.compute_sign || nrms))
  {
    V3 v[NV];
    load_leaf<NV>(soup, m.minPos, v);
    const V3 a = v[0], b = (m.minSub == 0) ? v[1] : v[2], c = (m.minSub == 0) ? v[2] : v[NV - 1];
    nrm = v3cross(v3sub(b, a), v3sub(c, a));
  }
  double sgn = 1.0;
  if(prm.compute_sign)
  {
    const double mp[3] = {m.minPt.x, m.minPt.y, m.minPt.z};
    bool in_dom = true;
#pragma unroll
    for(int d = 0; d < 3; ++d) in_dom = in_dom && !(mp[d] < prm.dom_lo[d] || mp[d] > prm.dom_hi[d]);
    if(!(prm.watertight && !in_dom))
    {
      const V3 r = v3sub(q, m.minPt);
      sgn = (v3dot(r, nrm) >= 0.0) ? 1.0 : -1.0;
    }
  }
  phi[qi] = sqrt(m.minSq) * sgn;
  if(cps)
  {
    cps[3 * (size_t)qi + 0] = m.minPt.x;
    cps[3 * (size_t)qi + 1] = m.minPt.y;
    cps[3 * (size_t)qi + 2] = m.minPt.z;
  }
  if(nrms)
  {
    const V3 u = v3unit(nrm);
    nrms[3 * (size_t)qi + 0] = u.x;
    nrms[3 * (size_t)qi + 1] = u.y;
    nrms[3 * (size_t)qi + 2] = u.z;
  }
}

// MODE 0: one thread per query, the reference's own visiting order (nearest-centroid-first,
// LinearBVH.hpp:72-85).  Bit-identical distances, closest points and (up to libm acos) normals.
template <int NV>
__global__ void __launch_bounds__(128) sd_reference_order_kernel(const Node<double, 3>* __restrict__ nodes, const double* __restrict__ soup,
                                                                  SdParams prm, Desc<3> qpts, int npts, const int32_t* __restrict__ perm,
                                                                  double* __restrict__ phi, double* __restrict__ cps,
                                                                  double* __restrict__ nrms, unsigned long long* __restrict__ work)
{
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if(t >= npts) return;
  const int qi = perm ? perm[t] : t;
  const double qp[3] = {ld_comp<double>(qpts, 0, qi), ld_comp<double>(qpts, 1, qi), ld_comp<double>(qpts, 2, qi)};
  const V3 q {qp[0], qp[1], qp[2]};
  MinCand m;
  m.minSq = DBL_MAX;
  m.minPt = {0.0, 0.0, 0.0};
  m.sumN = {0.0, 0.0, 0.0};
  m.minType = -1;
  m.minPos = 0;
  m.minSub = 0;
  const bool cn = prm.compute_sign != 0;
  unsigned nleaf = 0, ninner = 0;
  traverse_reference_order<double, 3>(
    nodes,
    [&](const Box<double, 3>& bb) { return sqdist_point_box(qp, bb) <= m.minSq; },
    [&](int pos) {
      ++nleaf;
      check_leaf<NV>(soup, q, m, pos, cn);
    },
    [&](const Box<double, 3>& L, const Box<double, 3>& R) {
      ++ninner;
      double dl = 0.0, dr = 0.0;
#pragma unroll
      for(int d = 0; d < 3; ++d)
      {
        const double c = 0.5 * (L.lo[d] + L.hi[d]) - qp[d];
        dl += c * c;
      }
      if(box_valid(R))
      {
#pragma unroll
        for(int d = 0; d < 3; ++d)
        {
          const double c = 0.5 * (R.lo[d] + R.hi[d]) - qp[d];
          dr += c * c;
        }
      }
      else
      {
        dr = DBL_MAX;
      }
      return dl > dr;
    });
  sd_finish<NV>(soup, prm, q, m, qi, phi, cps, nrms);
  if(work)
  {
    // profiling only
    atomicAdd(&work[0], (unsigned long long)nleaf);
    atomicAdd(&work[1], (unsigned long long)ninner);
  }
}

//------------------------------------------------------------------------------------------
// setMesh kernels
//------------------------------------------------------------------------------------------
// mesh node bounds (:455-487): exact min/max per dimension
__global__ void __launch_bounds__(256) node_bounds_kernel(const double* __restrict__ x, const double* __restrict__ y, const double* __restrict__ z,
                                                           int nnodes, unsigned long long* __restrict__ obounds /* [6]: min xyz, max xyz */)
{
  double mn[3] = {DBL_MAX, DBL_MAX, DBL_MAX}, mx[3] = {-DBL_MAX, -DBL_MAX, -DBL_MAX};
  for(int i = blockIdx.x * blockDim.x + threadIdx.x; i < nnodes; i += gridDim.x * blockDim.x)
  {
    const double p[3] = {x[i], y[i], z[i]};
#pragma unroll
    for(int d = 0; d < 3; ++d)
    {
      if(p[d] < mn[d]) mn[d] = p[d];
      if(p[d] > mx[d]) mx[d] = p[d];
    }
  }
#pragma unroll
  for(int d = 0; d < 3; ++d)
  {
#pragma unroll
    for(int o = 16; o > 0; o >>= 1)
    {
      const double a = __shfl_xor_sync(0xffffffffu, mn[d], o);
      const double b = __shfl_xor_sync(0xffffffffu, mx[d], o);
      mn[d] = a < mn[d] ? a : mn[d];
      mx[d] = b > mx[d] ? b : mx[d];
    }
    if(lane_id() == 0)
    {
      atomicMin(&obounds[d], f64_to_ordered(mn[d]));
      atomicMax(&obounds[3 + d], f64_to_ordered(mx[d]));
    }
  }
}

// getCellBoundingBox (:608-633): AABB over the cell's nodes (addPoint on an invalid box)
template <int NV>
__global__ void __launch_bounds__(256) cell_boxes_kernel(const double* __restrict__ x, const double* __restrict__ y, const double* __restrict__ z,
                                                          const int32_t* __restrict__ conn, int ncells, Box<double, 3>* __restrict__ boxes)
{
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if(c >= ncells) return;
  Box<double, 3> bb;
  box_clear(bb);
#pragma unroll
  for(int k = 0; k < NV; ++k)
  {
    const int nd = conn[(size_t)c * NV + k];
    const double p[3] = {x[nd], y[nd], z[nd]};
#pragma unroll
    for(int d = 0; d < 3; ++d)
    {
      if(p[d] < bb.lo[d]) bb.lo[d] = p[d];
      if(p[d] > bb.hi[d]) bb.hi[d] = p[d];
    }
  }
  boxes[c] = bb;
}

// getCellBoundingBox for a mixed-shape mesh: the cell's nodes are conn[offsets[c] .. offsets[c+1])
__global__ void __launch_bounds__(256) cell_boxes_mixed_kernel(const double* __restrict__ x, const double* __restrict__ y,
                                                                const double* __restrict__ z, const int32_t* __restrict__ conn,
                                                                const int32_t* __restrict__ offsets, int ncells,
                                                                Box<double, 3>* __restrict__ boxes, int* __restrict__ bad)
{
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if(c >= ncells) return;
  const int b = offsets[c], nn = offsets[c + 1] - b;
  if(nn != 3 && nn != 4) atomicExch(bad, 1);  // SLIC_ASSERT(nnodes <= 4) (:648); only triangles and quads are surfaces
  Box<double, 3> bb;
  box_clear(bb);
  for(int k = 0; k < nn; ++k)
  {
    const int nd = conn[b + k];
    const double p[3] = {x[nd], y[nd], z[nd]};
#pragma unroll
    for(int d = 0; d < 3; ++d)
    {
      if(p[d] < bb.lo[d]) bb.lo[d] = p[d];
      if(p[d] > bb.hi[d]) bb.hi[d] = p[d];
    }
  }
  boxes[c] = bb;
}

__global__ void __launch_bounds__(256) gather_soup_mixed_kernel(const double* __restrict__ x, const double* __restrict__ y,
                                                                 const double* __restrict__ z, const int32_t* __restrict__ conn,
                                                                 const int32_t* __restrict__ offsets, const int32_t* __restrict__ leaf_nodes,
                                                                 int nleaves, int ncells, double* __restrict__ soup)
{
  const int pos = blockIdx.x * blockDim.x + threadIdx.x;
  if(pos >= nleaves) return;
  const int cell = leaf_nodes[pos];
  double* o = soup + (size_t)pos * kLeafDoubles;
  for(int k = 0; k < kLeafDoubles; ++k) o[k] = 0.0;
  if(cell >= ncells) return;
  const int b = offsets[cell], nn = offsets[cell + 1] - b;
  for(int k = 0; k < 4; ++k)
  {
    if(k < nn)
    {
      const int nd = conn[b + k];
      o[3 * k + 0] = x[nd];
      o[3 * k + 1] = y[nd];
      o[3 * k + 2] = z[nd];
    }
    else
    {
      o[3 * k + 0] = o[3 * k + 1] = o[3 * k + 2] = __longlong_as_double(0x7ff8000000000000ll);  // NaN: no 4th vertex
    }
  }
}

// gather leaf geometry into sorted-leaf order (one-time, at setMesh)
template <int NV>
__global__ void __launch_bounds__(256) gather_soup_kernel(const double* __restrict__ x, const double* __restrict__ y, const double* __restrict__ z,
                                                           const int32_t* __restrict__ conn, const int32_t* __restrict__ leaf_nodes,
                                                           int nleaves, int ncells, double* __restrict__ soup)
{
  const int pos = blockIdx.x * blockDim.x + threadIdx.x;
  if(pos >= nleaves) return;
  const int cell = leaf_nodes[pos];
  double* o = soup + (size_t)pos * kLeafDoubles;
  for(int k = 0; k < kLeafDoubles; ++k) o[k] = 0.0;
  // a padding leaf of the N<=1 case is never reached (its box is invalid); it stays zero
  if(cell >= ncells) return;
#pragma unroll
  for(int k = 0; k < NV; ++k)
  {
    const int nd = conn[(size_t)cell * NV + k];
    o[3 * k + 0] = x[nd];
    o[3 * k + 1] = y[nd];
    o[3 * k + 2] = z[nd];
  }
}

}  // namespace axb
