// axb_oracle.cpp -- CPU restatement of the reference hot path (LLNL/axom v0.11.0).
//
// TEST INFRASTRUCTURE ONLY.  This file is the parity checker for the CUDA path in
// axom_b200/csrc.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// --impl reference legs may load the library built from it; the product never does.
//
// It restates, in plain scalar C++ (no templates beyond <D>, no reference headers),
// the algorithm of
//   spin::BVH<D,SEQ_EXEC,double>::initialize / findPoints / findBoundingBoxes / findRays
//   quest::SignedDistance<3,SEQ_EXEC>::setMesh / computeDistances
//   primal::intersect(Triangle3, Triangle3) and quest::findTriMeshIntersectionsBVH<SEQ_EXEC,double>   (8(f) rank 1)
//   the per-rank step of quest::DistributedClosestPoint (generateBVHTreeImpl + computeLocalClosestPoints) (8(f) rank 3)
// Every function cites the reference file:line (relative to /root/reference/src/axom)
// it follows.  Parity is PINNED: tests/test_oracle_golden.py checks this file against
// golden vectors produced by the real reference compiled from /root/reference
// (oracle/build_ref.py -> oracle/_ref/libaxom_ref.so, generator tests/golden/make_golden.py)
// and against the known-answer tests of the reference's own unit tests
// (spin/tests/spin_bvh.cpp, quest/tests/quest_signed_distance*.cpp, primal/tests/primal_intersect.cpp).
// DistributedClosestPoint itself needs Conduit + MPI and cannot be built here: its per-rank step is pinned against the
// real spin::BVH traversal driven with the reference's two lambdas (oracle/ref_driver.cpp: axref_dcp_*).
//
// Compile with -ffp-contract=off (no FMA), as the reference's own Release build does
// on x86-64 baseline: the Morton quantisation and the closest-point region tests are
// sensitive to the last bit.
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>

#include <limits>

#ifdef _OPENMP
  #include <omp.h>
#endif

// The BVH half of this file is written in terms of `real` so that the same restatement serves
// spin::BVH<D,SEQ_EXEC,double> (liboracle.so) and spin::BVH<D,SEQ_EXEC,float>
// (liboracle_f32.so, built with -DAXO_REAL=float -DAXO_FLOAT_BUILD; its entry points are axof_*).
// SignedDistance is double only (quest/SignedDistance.hpp:151-156) and is left out of the float build.
#ifndef AXO_REAL
  #define AXO_REAL double
#endif
typedef AXO_REAL real;
#ifdef AXO_FLOAT_BUILD
  #define AXO_FN(name) axof_##name
#else
  #define AXO_FN(name) axo_##name
#endif

namespace
{
constexpr real kInvalidMin = std::numeric_limits<real>::max();   // primal/geometry/BoundingBox.hpp:72
constexpr real kInvalidMax = std::numeric_limits<real>::lowest();  // primal/geometry/BoundingBox.hpp:73

template <int D>
struct Box
{
  real lo[D];
  real hi[D];
};

// primal/geometry/BoundingBox.hpp:451-461
template <int D>
inline bool box_valid(const Box<D>& b)
{
  for(int d = 0; d < D; ++d)
    if(b.lo[d] > b.hi[d]) return false;
  return true;
}

template <int D>
inline void box_clear(Box<D>& b)
{
  for(int d = 0; d < D; ++d)
  {
    b.lo[d] = kInvalidMin;
    b.hi[d] = kInvalidMax;
  }
}

// primal/geometry/BoundingBox.hpp:575-584 (checkAndFixBounds)
template <int D>
inline void box_fix(Box<D>& b)
{
  for(int d = 0; d < D; ++d)
    if(b.lo[d] > b.hi[d]) std::swap(b.lo[d], b.hi[d]);
}

// primal/geometry/BoundingBox.hpp:548-561 (scale) with Point::midpoint (Point.hpp:279-290)
// and range() = max - min (BoundingBox.hpp:159)
template <int D>
inline void box_scale(Box<D>& b, double s)
{
  if(!box_valid(b)) return;
  const real hs = static_cast<real>(s * 0.5);  // static_cast<T>(scaleFactor * 0.5), scaleFactor is a double (:548)
  for(int d = 0; d < D; ++d)
  {
    const real mid = static_cast<real>(0.5 * (b.lo[d] + b.hi[d]));  // Point::midpoint :279-290
    const real r = hs * (b.hi[d] - b.lo[d]);
    b.lo[d] = mid - r;
    b.hi[d] = mid + r;
  }
  box_fix(b);
}

// primal/geometry/BoundingBox.hpp:487-508 (addBox) + :463-484 (addPoint)
template <int D>
inline void box_add(Box<D>& self, const Box<D>& o)
{
  if(box_valid(self))
  {
    if(box_valid(o))
    {
      for(int d = 0; d < D; ++d)
      {
        if(o.lo[d] < self.lo[d]) self.lo[d] = o.lo[d];
        if(o.lo[d] > self.hi[d]) self.hi[d] = o.lo[d];
        if(o.hi[d] < self.lo[d]) self.lo[d] = o.hi[d];
        if(o.hi[d] > self.hi[d]) self.hi[d] = o.hi[d];
      }
    }
  }
  else
  {
    self = o;
  }
}

// spin/MortonIndex.hpp:151-159 with the int32 magic numbers (:226-246 2-D, :377-393 3-D)
inline uint32_t spread2(uint32_t x)
{
  x &= 0x0000FFFFu;
  x = (x | (x << 8)) & 0x00FF00FFu;
  x = (x | (x << 4)) & 0x0F0F0F0Fu;
  x = (x | (x << 2)) & 0x33333333u;
  x = (x | (x << 1)) & 0x55555555u;
  return x;
}
inline uint32_t spread3(uint32_t x)
{
  x &= 0x0000FFFFu;
  x = (x | (x << 16)) & 0xFF0000FFu;
  x = (x | (x << 8)) & 0x0F00F00Fu;
  x = (x | (x << 4)) & 0xC30C30C3u;
  x = (x | (x << 2)) & 0x49249249u;
  return x;
}

// spin/internal/linear_bvh/build_radix_tree.hpp:48-65 (morton32_encode)
template <int D>
inline uint32_t morton32(const real* c01)
{
  constexpr int bits = 32 / D;
  constexpr real to_int = real(1 << bits);
  constexpr real ceil_v = to_int - real(1);
  int32_t q[D];
  for(int d = 0; d < D; ++d) q[d] = (int32_t)std::fmin(std::fmax(c01[d] * to_int, real(0)), ceil_v);
  if(D == 2) return spread2((uint32_t)q[0]) | (spread2((uint32_t)q[1]) << 1);
  return spread3((uint32_t)q[0]) | (spread3((uint32_t)q[1]) << 1) | (spread3((uint32_t)q[D - 1]) << 2);
}

inline int clz32(uint32_t x) { return x == 0 ? 32 : __builtin_clz(x); }

template <int D>
struct Bvh
{
  int n = 0;  // number of leaves after the N<=1 padding
  double scale = 1.000123;  // BoundingBox::scale takes a double
  real tol = std::numeric_limits<real>::epsilon();
  Box<D> bounds;
  std::vector<uint32_t> mcodes;    // sorted
  std::vector<int32_t> leafs;      // sort permutation == final leaf_nodes
  std::vector<int32_t> lchild, rchild, parent;
  std::vector<Box<D>> leaf_aabbs;  // sorted order, scaled
  std::vector<Box<D>> inner_aabbs;
  std::vector<Box<D>> inner_nodes;     // 2*(n-1), LinearBVH layout
  std::vector<int32_t> inner_children; // 2*(n-1)
};

// build_radix_tree.hpp:265-287
template <int D>
inline int delta(const Bvh<D>& t, int a, int b)
{
  const int inner = t.n - 1;
  const bool oor = (b < 0 || b > inner);
  const int bb = oor ? 0 : b;
  uint32_t x = t.mcodes[a] ^ t.mcodes[bb];
  const bool tie = (x == 0);
  if(tie) x = uint32_t(a) ^ uint32_t(bb);
  int c = clz32(x);
  if(tie) c += 32;
  return oor ? -1 : c;
}

// build_radix_tree.hpp:579-611 (build_radix_tree) followed by
// policy/LinearBVH.hpp:191-269 (buildImpl / emit)
template <int D>
void build(Bvh<D>& t, const Box<D>* in, int n)
{
  t.n = n;
  const int inner = n - 1;
  // transform_boxes :85-99
  t.leaf_aabbs.assign(in, in + n);
  for(int i = 0; i < n; ++i) box_scale(t.leaf_aabbs[i], t.scale);
  // reduce :134-141 (SEQ, no RAJA)
  box_clear(t.bounds);
  for(int i = 0; i < n; ++i) box_add(t.bounds, t.leaf_aabbs[i]);
  // get_mcodes :146-176
  real inv_ext[D], mn[D];
  for(int d = 0; d < D; ++d)
  {
    const real ext = t.bounds.hi[d] - t.bounds.lo[d];
    mn[d] = t.bounds.lo[d];
    // isNearlyEqual<FloatType>(extent, .0f) ? 0.f : 1.f / extent   (core/utilities/Utilities.hpp:317-321)
    inv_ext[d] = (std::fabs(ext - real(0)) <= real(1.0e-8)) ? real(0) : real(1) / ext;
  }
  std::vector<uint32_t> codes(n);
  for(int i = 0; i < n; ++i)
  {
    real c[D];
    for(int d = 0; d < D; ++d)
    {
      const real cen = static_cast<real>(0.5 * (t.leaf_aabbs[i].lo[d] + t.leaf_aabbs[i].hi[d]));
      c[d] = (cen - mn[d]) * inv_ext[d];
    }
    codes[i] = morton32<D>(c);
  }
  // sort_mcodes :243-260 (std::stable_sort fallback) + reorder :199-219
  t.leafs.resize(n);
  for(int i = 0; i < n; ++i) t.leafs[i] = i;
  std::stable_sort(t.leafs.begin(), t.leafs.end(), [&](int32_t a, int32_t b) { return codes[a] < codes[b]; });
  t.mcodes.resize(n);
  std::vector<Box<D>> sorted(n);
  for(int i = 0; i < n; ++i)
  {
    t.mcodes[i] = codes[t.leafs[i]];
    sorted[i] = t.leaf_aabbs[t.leafs[i]];
  }
  t.leaf_aabbs.swap(sorted);
  // build_tree :290-384
  t.lchild.assign(inner, 0);
  t.rchild.assign(inner, 0);
  t.parent.assign(inner + n, 0);
  for(int i = 0; i < inner; ++i)
  {
    const int d = (delta(t, i, i + 1) - delta(t, i, i - 1)) < 0 ? -1 : 1;
    const int dmin = delta(t, i, i - d);
    int lmax = 2;
    while(delta(t, i, i + lmax * d) > dmin) lmax *= 2;
    int l = 0;
    for(int s = lmax / 2; s >= 1; s /= 2)
      if(delta(t, i, i + (l + s) * d) > dmin) l += s;
    const int j = i + l * d;
    const int dnode = delta(t, i, j);
    int s = 0;
    real div = 2.0;  // FloatType div_factor = 2.f  (:336); float32(l) / div_factor is evaluated in FloatType
    for(int step = (int)std::ceil((real)(float)l / div);; div *= 2, step = (int)std::ceil((real)(float)l / div))
    {
      if(delta(t, i, i + (s + step) * d) > dnode) s += step;
      if(step == 1) break;
    }
    const int split = i + s * d + std::min(d, 0);
    if(std::min(i, j) == split)
    {
      t.parent[split + inner] = i;
      t.lchild[i] = split + inner;
    }
    else
    {
      t.parent[split] = i;
      t.lchild[i] = split;
    }
    if(std::max(i, j) == split + 1)
    {
      t.parent[split + inner + 1] = i;
      t.rchild[i] = split + inner + 1;
    }
    else
    {
      t.parent[split + 1] = i;
      t.rchild[i] = split + 1;
    }
    if(i == 0) t.parent[0] = -1;
  }
  // propagate_aabbs :505-576 (sequential order: the second arrival merges)
  t.inner_aabbs.resize(inner);
  for(int i = 0; i < inner; ++i) box_clear(t.inner_aabbs[i]);
  std::vector<int32_t> counters(inner, 0);
  for(int i = 0; i < n; ++i)
  {
    Box<D> aabb = t.leaf_aabbs[i];
    int last = inner + i;
    int cur = t.parent[inner + i];
    while(cur != -1)
    {
      if(counters[cur]++ == 0) break;
      const int lc = t.lchild[cur], rc = t.rchild[cur];
      const int other = (lc == last) ? rc : lc;
      const Box<D>& ob = (other >= inner) ? t.leaf_aabbs[other - inner] : t.inner_aabbs[other];
      box_add(aabb, ob);
      t.inner_aabbs[cur] = aabb;
      last = cur;
      cur = t.parent[cur];
    }
  }
  // emit: policy/LinearBVH.hpp:226-263
  t.inner_nodes.resize(2 * inner);
  t.inner_children.resize(2 * inner);
  for(int node = 0; node < inner; ++node)
  {
    for(int side = 0; side < 2; ++side)
    {
      int c = side == 0 ? t.lchild[node] : t.rchild[node];
      if(c >= inner)
      {
        t.inner_nodes[2 * node + side] = t.leaf_aabbs[c - inner];
        c = -(c - inner + 1);
      }
      else
      {
        t.inner_nodes[2 * node + side] = t.inner_aabbs[c];
        c *= 2;
      }
      t.inner_children[2 * node + side] = c;
    }
  }
}

// spin/BVH.hpp:424-477 (initialize, incl. the N<=1 padding :439-464)
template <int D>
Bvh<D>* create(const real* boxes_aos, int n, double scale, double tol)
{
  Bvh<D>* t = new Bvh<D>();
  if(scale > 0) t->scale = scale;
  if(tol >= 0) t->tol = static_cast<real>(tol);
  const Box<D>* in = reinterpret_cast<const Box<D>*>(boxes_aos);
  if(n <= 1)
  {
    Box<D> two[2];
    box_clear(two[0]);
    box_clear(two[1]);
    if(n == 1) two[0] = in[0];
    build(*t, two, 2);
  }
  else
  {
    build(*t, in, n);
  }
  return t;
}

// spin/internal/linear_bvh/bvh_traverse.hpp:66-154.  pred(box) and order(L,R) are the
// "B" and "Comp" functors, leaf(sorted_pos) is "A".
template <int D, class Pred, class Leaf, class Order>
inline void traverse(const Bvh<D>& t, Pred&& pred, Leaf&& leaf, Order&& order)
{
  constexpr int32_t BARRIER = -2000000000;
  int32_t todo[64];
  int sp = 0;
  todo[0] = BARRIER;
  int32_t found = 0;
  int32_t cur = 0;
  const Box<D>* boxes = t.inner_nodes.data();
  const int32_t* kids = t.inner_children.data();
  while(cur != BARRIER)
  {
    while(cur >= 0)
    {
      const Box<D>& L = boxes[cur];
      const Box<D>& R = boxes[cur + 1];
      const bool inL = box_valid(L) ? pred(L) : false;
      const bool inR = box_valid(R) ? pred(R) : false;
      const int32_t lc = kids[cur];
      int32_t rc = kids[cur + 1];
      const bool swp = order(L, R);
      if(!inL && !inR)
      {
        cur = todo[sp--];
      }
      else
      {
        cur = inL ? lc : rc;
        if(inL && inR)
        {
          if(swp) std::swap(cur, rc);
          todo[++sp] = rc;
        }
      }
      if(cur < 0 && !(found < 0))
      {
        found = cur;
        if(cur != BARRIER) cur = todo[sp--];
      }
    }
    while(found < 0 && found != BARRIER)
    {
      leaf(-found - 1);
      found = cur;
      if(cur < 0 && cur != BARRIER) cur = todo[sp--];
    }
    found = 0;
  }
}

// policy/LinearBVH.hpp:368-401 (single-pass candidate fill; per-query order = DFS visit order)
template <int D, class MakePred>
int64_t find_generic(const Bvh<D>& t, int q, int32_t* offsets, int32_t* counts, std::vector<int32_t>& cand, MakePred&& mk)
{
  int64_t total = 0;
  for(int i = 0; i < q; ++i)
  {
    offsets[i] = (int32_t)total;
    int c = 0;
    auto pred = mk(i);
    traverse(
      t,
      pred,
      [&](int pos) {
        cand.push_back(t.leafs[pos]);
        ++c;
      },
      [](const Box<D>&, const Box<D>&) { return false; });
    counts[i] = c;
    total += c;
  }
  return total;
}

// primal/operators/detail/intersect_ray_impl.hpp:150-187 and :321-351
template <int D>
inline bool ray_hits(const real* o, const real* dir, const Box<D>& bb, real eps)
{
  real tmin = std::numeric_limits<real>::min(), tmax = std::numeric_limits<real>::max();
  for(int d = 0; d < D; ++d)
  {
    if(std::fabs(dir[d] - real(0)) <= eps)
    {
      if(o[d] < bb.lo[d] || o[d] > bb.hi[d]) return false;
    }
    else
    {
      const real inv = static_cast<real>(1.0) / dir[d];
      real t1 = (bb.lo[d] - o[d]) * inv;
      real t2 = (bb.hi[d] - o[d]) * inv;
      if(t1 > t2) std::swap(t1, t2);
      tmin = (t1 < tmin) ? tmin : t1;  // utilities::max(x,y) = (y < x) ? x : y  (core/utilities/Utilities.hpp:80-83)
      tmax = (t2 < tmax) ? t2 : tmax;  // utilities::min(x,y) = (y < x) ? y : x  (:93-96)
      if(tmin > tmax) return false;
    }
  }
  return true;
}

// primal/geometry/Vector.hpp:477-493 (unitVector) via NumericArray::operator/= (core/NumericArray.hpp:510-514)
template <int D>
inline void unit_vector(const real* v, real* out)
{
  real acc = 0;  // Vector::dot_product accumulates in T (Vector.hpp:543-552) ...
  for(int d = 0; d < D; ++d) acc += v[d] * v[d];
  const double len2 = acc;  // ... and unitVector continues in double (:482-486)
  if(len2 >= 1e-50)
  {
    const double s = 1. / std::sqrt(len2);
    for(int d = 0; d < D; ++d) out[d] = static_cast<real>(v[d] * s);  // NumericArray::operator*=(double)
  }
  else
  {
    out[0] = 1.0;
    for(int d = 1; d < D; ++d) out[d] = 0.0;
  }
}

// primal/operators/squared_distance.hpp:77-100 (point, box)
template <int D>
inline real sqdist_point_box(const real* p, const Box<D>& b)
{
  if(!box_valid(b)) return std::numeric_limits<real>::max();
  bool inside = true;
  for(int d = 0; d < D; ++d)
    if(p[d] < b.lo[d] || p[d] > b.hi[d]) inside = false;
  if(inside) return 0;
  real s = 0.0;
  for(int d = 0; d < D; ++d)
  {
    const real c = (p[d] < b.lo[d]) ? b.lo[d] : (p[d] > b.hi[d]) ? b.hi[d] : p[d];
    const real v = c - p[d];
    s += v * v;
  }
  return s;
}

#ifndef AXO_FLOAT_BUILD
//------------------------------------------------------------------------------
// SignedDistance (3-D only, quest/SignedDistance.hpp)
//------------------------------------------------------------------------------
struct V3
{
  double x, y, z;
};
inline V3 sub(const V3& h, const V3& t) { return {h.x - t.x, h.y - t.y, h.z - t.z}; }
inline V3 add(const V3& a, const V3& b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline V3 mul(const V3& a, double s) { return {a.x * s, a.y * s, a.z * s}; }
inline double dot(const V3& a, const V3& b)
{
  double r = 0.0;
  r += a.x * b.x;
  r += a.y * b.y;
  r += a.z * b.z;
  return r;
}
// primal/geometry/Vector.hpp:564-571 with core/numerics/Determinants.hpp:31-38
inline V3 cross(const V3& u, const V3& v)
{
  return {u.y * v.z - v.y * u.z, v.x * u.z - u.x * v.z, u.x * v.y - v.x * u.y};
}
inline V3 unit(const V3& v)
{
  double in[3] = {v.x, v.y, v.z}, o[3];
  unit_vector<3>(in, o);
  return {o[0], o[1], o[2]};
}
inline bool nearly_eq(double a, double b, double th) { return std::fabs(a - b) <= th; }
// primal/operators/detail/fuzzy_comparators.hpp:26-45
inline bool is_leq(double x, double y, double e) { return !((x > y) && !nearly_eq(x, y, e)); }
inline bool is_geq(double x, double y, double e) { return !((x < y) && !nearly_eq(x, y, e)); }

// primal/operators/closest_point.hpp:162-290
inline V3 closest_point_tri(const V3& P, const V3& A, const V3& B, const V3& C, int* loc, double EPS)
{
  const V3 ab = sub(B, A), ac = sub(C, A), ap = sub(P, A);
  const double d1 = dot(ab, ap), d2 = dot(ac, ap);
  if(is_leq(d1, 0, EPS) && is_leq(d2, 0, EPS))
  {
    *loc = 0;
    return A;
  }
  const V3 bp = sub(P, B);
  const double d3 = dot(ab, bp), d4 = dot(ac, bp);
  if(is_geq(d3, 0, EPS) && is_leq(d4, d3, EPS))
  {
    *loc = 1;
    return B;
  }
  const double vc = d1 * d4 - d3 * d2;
  if(is_leq(vc, 0, EPS) && is_geq(d1, 0, EPS) && is_leq(d3, 0, EPS) && !nearly_eq(d1, d3, EPS))
  {
    const double v = d1 / (d1 - d3);
    *loc = -1;
    return add(A, mul(ab, v));
  }
  const V3 cp = sub(P, C);
  const double d5 = dot(ab, cp), d6 = dot(ac, cp);
  if(is_geq(d6, 0, EPS) && is_leq(d5, d6, EPS))
  {
    *loc = 2;
    return C;
  }
  const double vb = d5 * d2 - d1 * d6;
  if(is_leq(vb, 0, EPS) && is_geq(d2, 0, EPS) && is_leq(d6, 0, EPS))
  {
    const double w = d2 / (d2 - d6);
    *loc = -3;
    return add(A, mul(ac, w));
  }
  const double va = d3 * d6 - d5 * d4;
  if(is_leq(va, 0, EPS) && is_geq(d4 - d3, 0, EPS) && is_geq(d5 - d6, 0, EPS))
  {
    const double w = (d4 - d3) / ((d4 - d3) + (d5 - d6));
    *loc = -2;
    return add(B, mul(sub(C, B), w));
  }
  const double denom = 1.0 / (va + vb + vc);
  const double v = vb * denom, w = vc * denom;
  *loc = 3;
  return add(A, add(mul(ab, v), mul(ac, w)));
}

struct Surface
{
  std::vector<double> x, y, z;
  std::vector<int32_t> conn;
  std::vector<int32_t> offsets;  // mixed-shape meshes only (UcdMeshData::cell_node_offsets, :51,84-89)
  int ncells = 0, npc = 3;
  // UcdMeshData::getCellNodeIDs (:76-93)
  const int32_t* cell_nodes(int cell, int& nnodes) const
  {
    if(offsets.empty())
    {
      nnodes = npc;
      return &conn[(size_t)cell * npc];
    }
    nnodes = offsets[cell + 1] - offsets[cell];
    return &conn[offsets[cell]];
  }
  bool watertight = true, compute_sign = true;
  Box<3> domain;  // node bounds, unscaled (quest/SignedDistance.hpp:483-486)
  Bvh<3>* bvh = nullptr;
  ~Surface() { delete bvh; }
};

struct MinCand  // quest/SignedDistance.hpp:159-175
{
  double minSq = DBL_MAX;
  V3 minPt {0, 0, 0};
  int minType = -1;  // -1 uninit, 0 vertex, 1 edge, 2 face
  int minElem = -1;
  V3 tri[3] {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
  V3 sumN {0, 0, 0};
  int minCount = 0;
};

inline int loc_type(int loc) { return loc < 0 ? 1 : (loc <= 2 ? 0 : 2); }  // :115-137

// primal/geometry/Triangle.hpp:384-400
inline double tri_angle(const V3* t, int idx)
{
  const V3& pt = t[idx];
  const V3 v1 = unit(sub(t[(idx + 1) % 3], pt));
  const V3 v2 = unit(sub(t[(idx + 2) % 3], pt));
  const double dp = dot(v1, v2);
  return std::acos(dp < -1.0 ? -1.0 : (dp > 1.0 ? 1.0 : dp));
}

// quest/SignedDistance.hpp:636-737
inline void check_candidate(const Surface& s, const V3& q, MinCand& m, int cell)
{
  int nnodes;
  const int32_t* nd = s.cell_nodes(cell, nnodes);
  auto P = [&](int k) { return V3 {s.x[nd[k]], s.y[nd[k]], s.z[nd[k]]}; };
  V3 elems[2][3] = {{P(0), P(1), P(2)}, {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}}};
  int ncand = 1;
  if(nnodes == 4)
  {
    ncand = 2;
    elems[1][0] = P(0);
    elems[1][1] = P(2);
    elems[1][2] = P(3);
  }
  constexpr double EPS = 1e-12;
  for(int e = 0; e < ncand; ++e)
  {
    const V3* T = elems[e];
    int loc;
    const V3 cp = closest_point_tri(q, T[0], T[1], T[2], &loc, EPS);
    const V3 dq = sub(cp, q);
    const double sq = dot(dq, dq);
    const int type = loc_type(loc);
    const bool shared = (type == 0 || type == 1);
    bool upd = false;
    if(sq < m.minSq)
    {
      const V3 dm = sub(m.minPt, cp);
      const bool clear = !shared || (m.minType != type) || !nearly_eq(dot(dm, dm), 0., EPS);
      m.minSq = sq;
      m.minPt = cp;
      m.minType = type;
      m.minElem = cell;
      m.tri[0] = T[0];
      m.tri[1] = T[1];
      m.tri[2] = T[2];
      if(s.compute_sign && clear)
      {
        m.sumN = {0, 0, 0};
        m.minCount = 0;
      }
      upd = s.compute_sign && shared;
    }
    else
    {
      const V3 dm = sub(m.minPt, cp);
      upd = s.compute_sign && shared && (m.minType == type) && nearly_eq(dot(dm, dm), 0., EPS);
    }
    if(upd)
    {
      ++m.minCount;
      const V3 n = cross(sub(T[1], T[0]), sub(T[2], T[0]));
      if(type == 1)
      {
        m.sumN = add(m.sumN, unit(n));
      }
      else
      {
        const double area = 0.5 * std::sqrt(dot(n, n));  // Triangle::area :105-109, degenerate :326-330
        if(!nearly_eq(area, 0.0, 1.0e-12))
        {
          const double alpha = tri_angle(T, loc);
          m.sumN = add(m.sumN, mul(unit(n), alpha));
        }
      }
    }
  }
}

// quest/SignedDistance.hpp:563-604 for one query point
inline void sd_one(const Surface& s, const double* qp, double* phi, double* cp_out, double* n_out)
{
  const Bvh<3>& t = *s.bvh;
  const V3 q {qp[0], qp[1], qp[2]};
  MinCand m;
  traverse(
    t,
    [&](const Box<3>& bb) { return sqdist_point_box<3>(qp, bb) <= m.minSq; },
    [&](int pos) { check_candidate(s, q, m, t.leafs[pos]); },
    // policy/LinearBVH.hpp:75-82: prefer the child whose centroid is nearer
    [&](const Box<3>& L, const Box<3>& R) {
      double dl = 0.0, dr = 0.0;
      for(int d = 0; d < 3; ++d)
      {
        const double c = 0.5 * (L.lo[d] + L.hi[d]) - qp[d];
        dl += c * c;
      }
      if(box_valid(R))
      {
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
  // getSurfaceNormal :740-748
  const V3 nrm = (m.minType == 2) ? cross(sub(m.tri[1], m.tri[0]), sub(m.tri[2], m.tri[0])) : m.sumN;
  double sgn = 1.0;
  if(s.compute_sign)
  {
    bool inside_dom = true;
    const double mp[3] = {m.minPt.x, m.minPt.y, m.minPt.z};
    for(int d = 0; d < 3; ++d)
      if(mp[d] < s.domain.lo[d] || mp[d] > s.domain.hi[d]) inside_dom = false;
    if(!(s.watertight && !inside_dom))
    {
      const V3 r = sub(q, m.minPt);  // Vector(minPt -> q), :759
      sgn = (dot(r, nrm) >= 0.0) ? 1.0 : -1.0;
    }
  }
  *phi = std::sqrt(m.minSq) * sgn;
  if(cp_out)
  {
    cp_out[0] = m.minPt.x;
    cp_out[1] = m.minPt.y;
    cp_out[2] = m.minPt.z;
  }
  if(n_out)
  {
    const V3 u = unit(nrm);
    n_out[0] = u.x;
    n_out[1] = u.y;
    n_out[2] = u.z;
  }
}

#endif  // !AXO_FLOAT_BUILD

}  // namespace

//------------------------------------------------------------------------------
// C ABI (ctypes).  Handles are opaque; ndims selects the instantiation.
//------------------------------------------------------------------------------
struct AxoBvh
{
  int ndims;
  void* impl;
};

#define DISPATCH(h, expr2, expr3) ((h)->ndims == 2 ? (expr2) : (expr3))

extern "C" {

AxoBvh* AXO_FN(bvh_create)(int ndims, const real* boxes_aos, int n, double scale, double tol)
{
  AxoBvh* h = new AxoBvh {ndims, nullptr};
  h->impl = ndims == 2 ? (void*)create<2>(boxes_aos, n, scale, tol) : (void*)create<3>(boxes_aos, n, scale, tol);
  return h;
}

void AXO_FN(bvh_destroy)(AxoBvh* h)
{
  if(!h) return;
  if(h->ndims == 2)
    delete(Bvh<2>*)h->impl;
  else
    delete(Bvh<3>*)h->impl;
  delete h;
}

int AXO_FN(bvh_num_leaves)(const AxoBvh* h) { return DISPATCH(h, ((Bvh<2>*)h->impl)->n, ((Bvh<3>*)h->impl)->n); }

extern "C++" {
template <int D>
static void get_arrays(const Bvh<D>& t, uint32_t* mcodes, int32_t* leafs, int32_t* lchild, int32_t* rchild, int32_t* parents,
                       real* inner_nodes, int32_t* inner_children, real* bounds)
{
  const int n = t.n, inner = n - 1;
  if(mcodes) memcpy(mcodes, t.mcodes.data(), sizeof(uint32_t) * n);
  if(leafs) memcpy(leafs, t.leafs.data(), sizeof(int32_t) * n);
  if(lchild) memcpy(lchild, t.lchild.data(), sizeof(int32_t) * inner);
  if(rchild) memcpy(rchild, t.rchild.data(), sizeof(int32_t) * inner);
  if(parents) memcpy(parents, t.parent.data(), sizeof(int32_t) * (inner + n));
  if(inner_nodes) memcpy(inner_nodes, t.inner_nodes.data(), sizeof(Box<D>) * 2 * inner);
  if(inner_children) memcpy(inner_children, t.inner_children.data(), sizeof(int32_t) * 2 * inner);
  if(bounds) memcpy(bounds, &t.bounds, sizeof(Box<D>));
}
}  // extern "C++"

void AXO_FN(bvh_get)(const AxoBvh* h, uint32_t* mcodes, int32_t* leafs, int32_t* lchild, int32_t* rchild, int32_t* parents,
                 real* inner_nodes, int32_t* inner_children, real* bounds)
{
  if(h->ndims == 2)
    get_arrays(*(Bvh<2>*)h->impl, mcodes, leafs, lchild, rchild, parents, inner_nodes, inner_children, bounds);
  else
    get_arrays(*(Bvh<3>*)h->impl, mcodes, leafs, lchild, rchild, parents, inner_nodes, inner_children, bounds);
}

void AXO_FN(free)(void* p) { free(p); }

static int32_t* to_malloc(const std::vector<int32_t>& v)
{
  int32_t* p = (int32_t*)malloc(sizeof(int32_t) * (v.size() ? v.size() : 1));
  if(!v.empty()) memcpy(p, v.data(), sizeof(int32_t) * v.size());
  return p;
}

// spin/BVH.hpp:480-505
extern "C++" {
template <int D>
static int64_t find_points(const Bvh<D>& t, const real* pts, int q, int32_t* off, int32_t* cnt, int32_t** cand)
{
  std::vector<int32_t> c;
  int64_t tot = find_generic(t, q, off, cnt, c, [&](int i) {
    const real* p = pts + (size_t)i * D;
    return [p](const Box<D>& bb) {
      for(int d = 0; d < D; ++d)
        if(p[d] < bb.lo[d] || p[d] > bb.hi[d]) return false;  // BoundingBox.hpp:390-401
      return true;
    };
  });
  *cand = to_malloc(c);
  return tot;
}
}  // extern "C++"

int64_t AXO_FN(bvh_find_points)(const AxoBvh* h, const real* pts_aos, int q, int32_t* offsets, int32_t* counts, int32_t** cand)
{
  return DISPATCH(h, find_points(*(Bvh<2>*)h->impl, pts_aos, q, offsets, counts, cand),
                  find_points(*(Bvh<3>*)h->impl, pts_aos, q, offsets, counts, cand));
}

// spin/BVH.hpp:539-568
extern "C++" {
template <int D>
static int64_t find_boxes(const Bvh<D>& t, const real* qb, int q, int32_t* off, int32_t* cnt, int32_t** cand)
{
  std::vector<int32_t> c;
  const Box<D>* qs = reinterpret_cast<const Box<D>*>(qb);
  int64_t tot = find_generic(t, q, off, cnt, c, [&](int i) {
    const Box<D> b = qs[i];
    return [b](const Box<D>& bb) {
      // bb1 = query, bb2 = bin: detail/intersect_bounding_box_impl.hpp:34-41
      for(int d = 0; d < D; ++d)
        if(!((b.hi[d] >= bb.lo[d]) && (b.lo[d] <= bb.hi[d]))) return false;
      return true;
    };
  });
  *cand = to_malloc(c);
  return tot;
}
}  // extern "C++"

int64_t AXO_FN(bvh_find_boxes)(const AxoBvh* h, const real* boxes_aos, int q, int32_t* offsets, int32_t* counts, int32_t** cand)
{
  return DISPATCH(h, find_boxes(*(Bvh<2>*)h->impl, boxes_aos, q, offsets, counts, cand),
                  find_boxes(*(Bvh<3>*)h->impl, boxes_aos, q, offsets, counts, cand));
}

// spin/BVH.hpp:508-536; normalize != 0 reproduces the primal::Ray constructor (Ray.hpp:122-127)
extern "C++" {
template <int D>
static int64_t find_rays(const Bvh<D>& t, const real* orig, const real* dirs, int q, int normalize, int32_t* off, int32_t* cnt,
                         int32_t** cand)
{
  std::vector<int32_t> c;
  const real tol = t.tol;
  int64_t tot = find_generic(t, q, off, cnt, c, [&](int i) {
    struct R
    {
      real o[D], d[D];
    } r;
    for(int k = 0; k < D; ++k) r.o[k] = orig[(size_t)i * D + k];
    if(normalize)
      unit_vector<D>(dirs + (size_t)i * D, r.d);
    else
      for(int k = 0; k < D; ++k) r.d[k] = dirs[(size_t)i * D + k];
    return [r, tol](const Box<D>& bb) { return ray_hits<D>(r.o, r.d, bb, tol); };
  });
  *cand = to_malloc(c);
  return tot;
}
}  // extern "C++"

int64_t AXO_FN(bvh_find_rays)(const AxoBvh* h, const real* origins_aos, const real* dirs_aos, int q, int normalize, int32_t* offsets,
                          int32_t* counts, int32_t** cand)
{
  return DISPATCH(h, find_rays(*(Bvh<2>*)h->impl, origins_aos, dirs_aos, q, normalize, offsets, counts, cand),
                  find_rays(*(Bvh<3>*)h->impl, origins_aos, dirs_aos, q, normalize, offsets, counts, cand));
}

#ifndef AXO_FLOAT_BUILD
// count-only traversal driven by an external OpenMP loop (what RAJA's omp policy does
// with policy/LinearBVH.hpp:302-321); used for the host-core baseline timing.
int64_t axo_bvh_count_points_omp(const AxoBvh* h, const double* pts_aos, int q, int32_t* counts, int nthreads)
{
  if(h->ndims != 3) return -1;
  const Bvh<3>& t = *(Bvh<3>*)h->impl;
  int64_t total = 0;
#ifdef _OPENMP
  if(nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel for schedule(dynamic, 1024) reduction(+ : total)
  for(int i = 0; i < q; ++i)
  {
    const double* p = pts_aos + (size_t)i * 3;
    int c = 0;
    traverse(
      t,
      [p](const Box<3>& bb) {
        for(int d = 0; d < 3; ++d)
          if(p[d] < bb.lo[d] || p[d] > bb.hi[d]) return false;
        return true;
      },
      [&](int) { ++c; },
      [](const Box<3>&, const Box<3>&) { return false; });
    counts[i] = c;
    total += c;
  }
  return total;
}

// the same for the box-box and ray-box predicates (spin/BVH.hpp:558-560, :527-531)
int64_t axo_bvh_count_boxes_omp(const AxoBvh* h, const double* boxes_aos, int q, int32_t* counts, int nthreads)
{
  if(h->ndims != 3) return -1;
  const Bvh<3>& t = *(Bvh<3>*)h->impl;
  int64_t total = 0;
#ifdef _OPENMP
  if(nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel for schedule(dynamic, 1024) reduction(+ : total)
  for(int i = 0; i < q; ++i)
  {
    const double* b = boxes_aos + (size_t)i * 6;
    int c = 0;
    traverse(
      t,
      [b](const Box<3>& bb) {  // bb1.intersectsWith(bb2), bb1 = query
        for(int d = 0; d < 3; ++d)
          if(!(b[3 + d] >= bb.lo[d] && b[d] <= bb.hi[d])) return false;
        return true;
      },
      [&](int) { ++c; },
      [](const Box<3>&, const Box<3>&) { return false; });
    counts[i] = c;
    total += c;
  }
  return total;
}

int64_t axo_bvh_count_rays_omp(const AxoBvh* h, const double* origins_aos, const double* dirs_aos, int q, int32_t* counts, int nthreads)
{
  if(h->ndims != 3) return -1;
  const Bvh<3>& t = *(Bvh<3>*)h->impl;
  const double tol = t.tol;
  int64_t total = 0;
#ifdef _OPENMP
  if(nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel for schedule(dynamic, 1024) reduction(+ : total)
  for(int i = 0; i < q; ++i)
  {
    double o[3], d[3];
    for(int k = 0; k < 3; ++k) o[k] = origins_aos[(size_t)i * 3 + k];
    unit_vector<3>(dirs_aos + (size_t)i * 3, d);  // the Ray constructor normalises (Ray.hpp:122-127)
    int c = 0;
    traverse(
      t, [&](const Box<3>& bb) { return ray_hits<3>(o, d, bb, tol); }, [&](int) { ++c; }, [](const Box<3>&, const Box<3>&) { return false; });
    counts[i] = c;
    total += c;
  }
  return total;
}

// quest/SignedDistance.hpp:427-504 (setMesh)
static void* sd_create_impl(const double* x, const double* y, const double* z, int nnodes, const int32_t* conn, const int32_t* offsets,
                            int ncells, int nodes_per_cell, int watertight, int compute_sign)
{
  Surface* s = new Surface();
  s->x.assign(x, x + nnodes);
  s->y.assign(y, y + nnodes);
  s->z.assign(z, z + nnodes);
  if(offsets)
  {
    s->offsets.assign(offsets, offsets + ncells + 1);
    s->conn.assign(conn, conn + offsets[ncells]);
  }
  else
  {
    s->conn.assign(conn, conn + (size_t)ncells * nodes_per_cell);
  }
  s->ncells = ncells;
  s->npc = nodes_per_cell;
  s->watertight = watertight != 0;
  s->compute_sign = compute_sign != 0;
  box_clear(s->domain);
  for(int i = 0; i < nnodes; ++i)
  {
    const double p[3] = {x[i], y[i], z[i]};
    for(int d = 0; d < 3; ++d)  // addPoint on a default (invalid) box: BoundingBox.hpp:463-484
    {
      if(p[d] < s->domain.lo[d]) s->domain.lo[d] = p[d];
      if(p[d] > s->domain.hi[d]) s->domain.hi[d] = p[d];
    }
  }
  std::vector<Box<3>> boxes(ncells > 0 ? ncells : 1);
  for(int c = 0; c < ncells; ++c)  // getCellBoundingBox :608-633
  {
    Box<3> bb;
    box_clear(bb);
    int cn;
    const int32_t* ids = s->cell_nodes(c, cn);
    for(int k = 0; k < cn; ++k)
    {
      const int nd = ids[k];
      const double p[3] = {x[nd], y[nd], z[nd]};
      for(int d = 0; d < 3; ++d)
      {
        if(p[d] < bb.lo[d]) bb.lo[d] = p[d];
        if(p[d] > bb.hi[d]) bb.hi[d] = p[d];
      }
    }
    boxes[c] = bb;
  }
  s->bvh = create<3>(reinterpret_cast<const double*>(boxes.data()), ncells, -1.0, -1.0);
  return s;
}

void* axo_sd_create(const double* x, const double* y, const double* z, int nnodes, const int32_t* conn, int ncells, int nodes_per_cell,
                    int watertight, int compute_sign)
{
  return sd_create_impl(x, y, z, nnodes, conn, nullptr, ncells, nodes_per_cell, watertight, compute_sign);
}

// mixed triangle / quad surface (mint::UnstructuredMesh<MIXED_SHAPE>): offsets[ncells+1] into conn
void* axo_sd_create_mixed(const double* x, const double* y, const double* z, int nnodes, const int32_t* conn, const int32_t* offsets,
                          int ncells, int watertight, int compute_sign)
{
  return sd_create_impl(x, y, z, nnodes, conn, offsets, ncells, -1, watertight, compute_sign);
}

void axo_sd_destroy(void* h) { delete(Surface*)h; }

void axo_sd_get_bvh(void* h, AxoBvh* out)
{
  out->ndims = 3;
  out->impl = ((Surface*)h)->bvh;
}

// quest/SignedDistance.hpp:527-605; nthreads > 1 runs the per-query loop under OpenMP
// (queries are independent, the method is const) for the host-core baseline.
void axo_sd_compute(void* h, const double* qpts_aos, int npts, double* phi, double* cp, double* normals, int nthreads)
{
  const Surface& s = *(Surface*)h;
#ifdef _OPENMP
  if(nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel for schedule(dynamic, 64) if(nthreads != 1)
  for(int i = 0; i < npts; ++i)
  {
    sd_one(s, qpts_aos + (size_t)i * 3, phi + i, cp ? cp + (size_t)i * 3 : nullptr, normals ? normals + (size_t)i * 3 : nullptr);
  }
}

#endif  // !AXO_FLOAT_BUILD

#ifndef AXO_FLOAT_BUILD
//------------------------------------------------------------------------------
// Narrow phase downstream of findBoundingBoxes:
// primal::intersect(Triangle3, Triangle3, includeBoundary, EPS) and
// quest::findTriMeshIntersectionsBVH (quest/MeshTester.hpp:67-83,
// quest/detail/MeshTester_detail.hpp:158-307).
//------------------------------------------------------------------------------
extern "C++" {
namespace tt
{
struct P2
{
  double x, y;
};
// primal/operators/detail/fuzzy_comparators.hpp:20-64
inline bool is_gt(double x, double y, double e) { return (x > y) && !nearly_eq(x, y, e); }
inline bool is_lt(double x, double y, double e) { return (x < y) && !nearly_eq(x, y, e); }
inline bool is_lpeq(double x, double y, bool inc, double e) { return (inc && nearly_eq(x, y, e)) ? true : is_lt(x, y, e); }
inline bool is_gpeq(double x, double y, bool inc, double e) { return (inc && nearly_eq(x, y, e)) ? true : is_gt(x, y, e); }
// intersect_impl.hpp:539-542
inline double cross2(const P2& A, const P2& B, const P2& C) { return (A.x - C.x) * (B.y - C.y) - (A.y - C.y) * (B.x - C.x); }
// :548-590
inline int sgn(double x) { return (0 < x) - (x < 0); }
inline int count_zeros(double x, double y, double z, double e) { return (int)nearly_eq(x, 0., e) + (int)nearly_eq(y, 0., e) + (int)nearly_eq(z, 0., e); }
inline bool nonzero_sign_match(double x, double y, double z, double e)
{
  return !nearly_eq(x, 0., e) && !nearly_eq(y, 0., e) && !nearly_eq(z, 0., e) && sgn(x) == sgn(y) && sgn(x) == sgn(z);
}
inline bool one_zero_others_match(double x, double y, double z, double e)
{
  return count_zeros(x, y, z, e) == 1 &&
    ((nearly_eq(x, 0., e) && is_gt(y * z, 0., e)) || (nearly_eq(y, 0., e) && is_gt(z * x, 0., e)) || (nearly_eq(z, 0., e) && is_gt(x * y, 0., e)));
}
// Triangle::normal (primal/geometry/Triangle.hpp:98-102)
inline V3 tri_normal(const V3& a, const V3& b, const V3& c) { return cross(sub(b, a), sub(c, a)); }

// :1286-1385 checkEdge
inline bool check_edge(const P2& p1, const P2& q1, const P2& r1, const P2& p2, const P2& r2, bool b, double e)
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
// :1393-1484 checkVertex
inline bool check_vertex(const P2& p1, const P2& q1, const P2& r1, const P2& p2, const P2& q2, const P2& r2, bool b, double e)
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
// :1281-1350 intersectPermuted2DTriangles
inline bool permuted_2d(const P2& p1, const P2& q1, const P2& r1, const P2& p2, const P2& q2, const P2& r2, bool b, double e)
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
// :1245-1278 TriangleIntersection2D
inline bool tri2d(const P2* t1, const P2* t2, bool b, double e)
{
  const bool f1 = is_lt(cross2(t1[0], t1[1], t1[2]), 0., e);
  const bool f2 = is_lt(cross2(t2[0], t2[1], t2[2]), 0., e);
  return permuted_2d(t1[0], f1 ? t1[2] : t1[1], f1 ? t1[1] : t1[2], t2[0], f2 ? t2[2] : t2[1], f2 ? t2[1] : t2[2], b, e);
}
// :1185-1242 intersectCoplanar3DTriangles
inline bool coplanar(const V3& p1, const V3& q1, const V3& r1, const V3& p2, const V3& q2, const V3& r2, V3 n, bool b, double e)
{
  n.x = std::fabs(n.x);
  n.y = std::fabs(n.y);
  n.z = std::fabs(n.z);
  if(is_gt(n.x, n.z, e) && is_geq(n.x, n.y, e))
  {
    const P2 a[3] = {{q1.z, q1.y}, {p1.z, p1.y}, {r1.z, r1.y}};
    const P2 c[3] = {{q2.z, q2.y}, {p2.z, p2.y}, {r2.z, r2.y}};
    return tri2d(a, c, b, e);
  }
  if(is_gt(n.y, n.z, e) && is_geq(n.y, n.x, e))
  {
    const P2 a[3] = {{q1.x, q1.z}, {p1.x, p1.z}, {r1.x, r1.z}};
    const P2 c[3] = {{q2.x, q2.z}, {p2.x, p2.z}, {r2.x, r2.z}};
    return tri2d(a, c, b, e);
  }
  const P2 a[3] = {{p1.x, p1.y}, {q1.x, q1.y}, {r1.x, r1.y}};
  const P2 c[3] = {{p2.x, p2.y}, {q2.x, q2.y}, {r2.x, r2.y}};
  return tri2d(a, c, b, e);
}
// :457-474 intersectTwoPermutedTriangles
inline bool two_permuted(const V3& p1, const V3& q1, const V3& r1, const V3& p2, const V3& q2, const V3& r2, bool b, double e)
{
  return is_lpeq(dot(sub(q2, q1), tri_normal(q1, p2, p1)), 0., b, e) && is_lpeq(dot(sub(r2, p1), tri_normal(p1, p2, r1)), 0., b, e);
}
// :1103-1182 intersectOnePermutedTriangle
inline bool one_permuted(const V3& p1, const V3& q1, const V3& r1, const V3& p2, const V3& q2, const V3& r2, double dp2, double dq2,
                         double dr2, const V3& n, bool b, double e)
{
  if(is_gt(dp2, 0., e))
  {
    if(is_gt(dq2, 0., e)) return two_permuted(p1, r1, q1, r2, p2, q2, b, e);
    if(is_gt(dr2, 0., e)) return two_permuted(p1, r1, q1, q2, r2, p2, b, e);
    return two_permuted(p1, q1, r1, p2, q2, r2, b, e);
  }
  if(is_lt(dp2, 0., e))
  {
    if(is_lt(dq2, 0., e)) return two_permuted(p1, q1, r1, r2, p2, q2, b, e);
    if(is_lt(dr2, 0., e)) return two_permuted(p1, q1, r1, q2, r2, p2, b, e);
    return two_permuted(p1, r1, q1, p2, q2, r2, b, e);
  }
  if(is_lt(dq2, 0., e))
  {
    if(is_geq(dr2, 0., e)) return two_permuted(p1, r1, q1, q2, r2, p2, b, e);
    return two_permuted(p1, q1, r1, p2, q2, r2, b, e);
  }
  if(is_gt(dq2, 0., e))
  {
    if(is_gt(dr2, 0., e)) return two_permuted(p1, r1, q1, p2, q2, r2, b, e);
    return two_permuted(p1, q1, r1, q2, r2, p2, b, e);
  }
  if(is_gt(dr2, 0., e)) return two_permuted(p1, q1, r1, r2, p2, q2, b, e);
  if(is_lt(dr2, 0., e)) return two_permuted(p1, r1, q1, r2, p2, q2, b, e);
  return coplanar(p1, q1, r1, p2, q2, r2, n, b, e);
}
// :156-436 intersect_tri3D_tri3D
inline bool tri_tri(const V3* t1, const V3* t2, bool b, double e)
{
  const V3 n2 = unit(tri_normal(t2[0], t2[1], t2[2]));
  const double dp1 = dot(sub(t1[0], t2[2]), n2), dq1 = dot(sub(t1[1], t2[2]), n2), dr1 = dot(sub(t1[2], t2[2]), n2);
  if(nonzero_sign_match(dp1, dq1, dr1, e)) return false;
  if(!b && (count_zeros(dp1, dq1, dr1, e) == 2 || one_zero_others_match(dp1, dq1, dr1, e))) return false;
  const V3 n1 = unit(tri_normal(t1[0], t1[1], t1[2]));
  const double dp2 = dot(sub(t2[0], t1[2]), n1), dq2 = dot(sub(t2[1], t1[2]), n1), dr2 = dot(sub(t2[2], t1[2]), n1);
  if(nonzero_sign_match(dp2, dq2, dr2, e)) return false;
  if(!b && (count_zeros(dp2, dq2, dr2, e) == 2 || one_zero_others_match(dp2, dq2, dr2, e))) return false;
  // the permutation tables of :222-435: A = (t2 as is, dp2,dq2,dr2), S = (t2[0],t2[2],t2[1], dp2,dr2,dq2)
  auto A = [&](int i0, int i1, int i2) { return one_permuted(t1[i0], t1[i1], t1[i2], t2[0], t2[1], t2[2], dp2, dq2, dr2, n1, b, e); };
  auto S = [&](int i0, int i1, int i2) { return one_permuted(t1[i0], t1[i1], t1[i2], t2[0], t2[2], t2[1], dp2, dr2, dq2, n1, b, e); };
  if(is_gt(dp1, 0., e))
  {
    if(is_gt(dq1, 0., e)) return S(2, 0, 1);
    if(is_gt(dr1, 0., e)) return S(1, 2, 0);
    return A(0, 1, 2);
  }
  if(is_lt(dp1, 0., e))
  {
    if(is_lt(dq1, 0., e)) return A(2, 0, 1);
    if(is_lt(dr1, 0., e)) return A(1, 2, 0);
    return S(0, 1, 2);
  }
  if(is_lt(dq1, 0., e))
  {
    if(is_geq(dr1, 0., e)) return S(1, 2, 0);
    return A(0, 1, 2);
  }
  if(is_gt(dq1, 0., e))
  {
    if(is_gt(dr1, 0., e)) return S(0, 1, 2);
    return A(1, 2, 0);
  }
  if(is_gt(dr1, 0., e)) return A(2, 0, 1);
  if(is_lt(dr1, 0., e)) return S(2, 0, 1);
  return coplanar(t1[0], t1[1], t1[2], t2[0], t2[1], t2[2], n1, b, e);
}
}  // namespace tt
}  // extern "C++"

// primal::intersect(Triangle<double,3>, Triangle<double,3>, includeBoundary, EPS) on n pairs of 9-double triangles
void axo_tri_tri_intersect(const double* tris1, const double* tris2, int n, int include_boundary, double eps, uint8_t* out)
{
  for(int i = 0; i < n; ++i)
  {
    V3 a[3], b[3];
    for(int k = 0; k < 3; ++k)
    {
      a[k] = {tris1[i * 9 + 3 * k], tris1[i * 9 + 3 * k + 1], tris1[i * 9 + 3 * k + 2]};
      b[k] = {tris2[i * 9 + 3 * k], tris2[i * 9 + 3 * k + 1], tris2[i * 9 + 3 * k + 2]};
    }
    out[i] = tt::tri_tri(a, b, include_boundary != 0, eps) ? 1 : 0;
  }
}

// ---- leaf math on n independent items (the reference's own unit tests run through these: tests/test_leaf_math.py) ----
// primal::closest_point(Point, Triangle, int* loc, EPS) (closest_point.hpp:162-290); tris are 9 doubles A,B,C
void axo_closest_point_tri(const double* pts, const double* tris, int n, double eps, double* cp, int32_t* loc)
{
  for(int i = 0; i < n; ++i)
  {
    const V3 P {pts[3 * i], pts[3 * i + 1], pts[3 * i + 2]};
    const double* t = tris + (size_t)i * 9;
    int l = 0;
    const V3 c = closest_point_tri(P, V3 {t[0], t[1], t[2]}, V3 {t[3], t[4], t[5]}, V3 {t[6], t[7], t[8]}, &l, eps);
    cp[3 * i] = c.x;
    cp[3 * i + 1] = c.y;
    cp[3 * i + 2] = c.z;
    loc[i] = l;
  }
}
// primal::squared_distance(Point, BoundingBox) (squared_distance.hpp:77-100); boxes are lo[3], hi[3]
void axo_squared_distance_point_box(const double* pts, const double* boxes, int n, double* out)
{
  for(int i = 0; i < n; ++i)
  {
    Box<3> b;
    for(int d = 0; d < 3; ++d)
    {
      b.lo[d] = boxes[6 * i + d];
      b.hi[d] = boxes[6 * i + 3 + d];
    }
    out[i] = sqdist_point_box<3>(pts + 3 * i, b);
  }
}
// primal::intersect(Ray, BoundingBox) as BVH::findRays evaluates it (BVH.hpp:529-532, intersect_ray_impl.hpp:321-351);
// rays are origin[3], direction[3]; normalize != 0 applies the primal::Ray constructor's unitVector
void axo_intersect_ray_box(const double* rays, const double* boxes, int n, int normalize, double tol, uint8_t* out)
{
  for(int i = 0; i < n; ++i)
  {
    Box<3> b;
    for(int d = 0; d < 3; ++d)
    {
      b.lo[d] = boxes[6 * i + d];
      b.hi[d] = boxes[6 * i + 3 + d];
    }
    double dir[3] = {rays[6 * i + 3], rays[6 * i + 4], rays[6 * i + 5]};
    if(normalize) unit_vector<3>(rays + 6 * i + 3, dir);
    out[i] = ray_hits<3>(rays + 6 * i, dir, b, tol) ? 1 : 0;
  }
}
// BoundingBox::scale (BoundingBox.hpp:548-561), in place on n boxes
void axo_box_scale(double* boxes, int n, double scale)
{
  for(int i = 0; i < n; ++i)
  {
    Box<3> b;
    for(int d = 0; d < 3; ++d)
    {
      b.lo[d] = boxes[6 * i + d];
      b.hi[d] = boxes[6 * i + 3 + d];
    }
    box_scale(b, scale);
    for(int d = 0; d < 3; ++d)
    {
      boxes[6 * i + d] = b.lo[d];
      boxes[6 * i + 3 + d] = b.hi[d];
    }
  }
}

// quest::findTriMeshIntersectionsBVH<SEQ_EXEC,double> (MeshTester.hpp:67-104, MeshTester_detail.hpp:158-307):
// triangle AABBs -> BVH (default scale) -> findBoundingBoxes(own AABBs) -> pairs i < j in candidate order ->
// primal::intersect(tri_i, tri_j, false, threshold).  Returns the pair count; first/second/degenerate are malloc'ed.
int64_t axo_find_tri_mesh_intersections(const double* x, const double* y, const double* z, int nnodes, const int32_t* conn, int ncells,
                                        double threshold, int32_t** first, int32_t** second, int32_t** degenerate, int64_t* ndegenerate)
{
  (void)nnodes;
  std::vector<V3> tri((size_t)ncells * 3);
  std::vector<Box<3>> boxes(ncells > 0 ? ncells : 1);
  std::vector<int32_t> deg;
  for(int c = 0; c < ncells; ++c)
  {
    Box<3> bb;
    box_clear(bb);
    for(int k = 0; k < 3; ++k)
    {
      const int nd = conn[c * 3 + k];
      tri[c * 3 + k] = {x[nd], y[nd], z[nd]};
      const double p[3] = {x[nd], y[nd], z[nd]};
      for(int d = 0; d < 3; ++d)
      {
        if(p[d] < bb.lo[d]) bb.lo[d] = p[d];
        if(p[d] > bb.hi[d]) bb.hi[d] = p[d];
      }
    }
    boxes[c] = bb;
    // Triangle::degenerate (Triangle.hpp:326-330): |0.5 * |normal|| <= 1e-12
    const V3 n = tt::tri_normal(tri[c * 3], tri[c * 3 + 1], tri[c * 3 + 2]);
    if(nearly_eq(0.5 * std::sqrt(dot(n, n)), 0.0, 1.0e-12)) deg.push_back(c);
  }
  Bvh<3>* bvh = create<3>(reinterpret_cast<const double*>(boxes.data()), ncells, -1.0, -1.0);
  std::vector<int32_t> off(ncells > 0 ? ncells : 1), cnt(ncells > 0 ? ncells : 1);
  int32_t* cand = nullptr;
  find_boxes<3>(*bvh, reinterpret_cast<const double*>(boxes.data()), ncells, off.data(), cnt.data(), &cand);
  std::vector<int32_t> f, s;
  for(int i = 0; i < ncells; ++i)
    for(int j = 0; j < cnt[i]; ++j)
    {
      const int c = cand[off[i] + j];
      if(i < c && tt::tri_tri(&tri[(size_t)i * 3], &tri[(size_t)c * 3], false, threshold))
      {
        f.push_back(i);
        s.push_back(c);
      }
    }
  free(cand);
  delete bvh;
  *first = to_malloc(f);
  *second = to_malloc(s);
  *degenerate = to_malloc(deg);
  *ndegenerate = (int64_t)deg.size();
  return (int64_t)f.size();
}
#endif  // !AXO_FLOAT_BUILD

#ifndef AXO_FLOAT_BUILD
//------------------------------------------------------------------------------
// quest::DistributedClosestPoint, the per-rank step:
// DistributedClosestPointImpl<D, ExecSpace>::generateBVHTreeImpl + computeLocalClosestPoints
// (quest/detail/DistributedClosestPointImpl.hpp:883-1079).  The object "mesh" is a point cloud: every object
// point is a zero-size box, the query keeps the nearest OBJECT POINT (strict <, so the first one visited wins a
// tie), and a rank only overwrites an entry when it improves on what earlier ranks of the ring left there.
//------------------------------------------------------------------------------
extern "C++" {
template <int D>
struct Dcp
{
  Bvh<D>* bvh = nullptr;
  std::vector<double> pts;   // flattened object points of all local domains (:560-640)
  std::vector<int32_t> dom;  // domain id of every object point
  ~Dcp() { delete bvh; }
};

template <int D>
static void dcp_local(const Dcp<D>& o, int rank, double sq_thresh, const double* q, int nq, int is_first, int32_t* cp_index,
                      int32_t* cp_dom, int32_t* cp_rank, double* cp_coords, double* cp_dist)
{
  const double snan = std::numeric_limits<double>::signaling_NaN();
  for(int i = 0; i < nq; ++i)
  {
    if(is_first)  // :971-979
    {
      cp_rank[i] = cp_index[i] = cp_dom[i] = -1;
      for(int d = 0; d < D; ++d) cp_coords[(size_t)i * D + d] = snan;
      if(cp_dist) cp_dist[i] = snan;
    }
    if(!o.bvh) continue;  // hasObjectPoints == false (:911-916, :990)
    const double* p = q + (size_t)i * D;
    double cur_sq = std::numeric_limits<double>::max();  // MinCandidate{} (:231-243)
    int cur_idx = -1, cur_dom = -1, cur_rank = -1;
    if(cp_rank[i] >= 0)  // preset with the closest point found so far (:1013-1019)
    {
      double s = 0.0;
      for(int d = 0; d < D; ++d)
      {
        const double v = cp_coords[(size_t)i * D + d] - p[d];  // Vector(A, B) = B - A, squared_norm
        s += v * v;
      }
      cur_sq = s;
      cur_idx = cp_index[i];
      cur_dom = cp_dom[i];
      cur_rank = cp_rank[i];
    }
    traverse(
      *o.bvh,
      [&](const Box<D>& bb) {  // traversePredicate (:1037-1040)
        const double sq = sqdist_point_box<D>(p, bb);
        return sq <= cur_sq && sq <= sq_thresh;
      },
      [&](int pos) {  // checkMinDist (:1021-1035)
        const int c = o.bvh->leafs[pos];
        double s = 0.0;
        for(int d = 0; d < D; ++d)
        {
          const double v = o.pts[(size_t)c * D + d] - p[d];
          s += v * v;
        }
        if(s < cur_sq)
        {
          cur_sq = s;
          cur_idx = c;
          cur_dom = o.dom[c];
          cur_rank = rank;
        }
      },
      // qpt is a PointType, so overload resolution picks LinearBVHTraverser::traverse_tree(const PointType&, ...)
      // (policy/LinearBVH.hpp:72-85): the child whose box centroid is nearer is entered first
      [&](const Box<D>& L, const Box<D>& R) {
        double dl = 0.0, dr = 0.0;
        for(int d = 0; d < D; ++d)
        {
          const double c = 0.5 * (L.lo[d] + L.hi[d]) - p[d];
          dl += c * c;
        }
        if(box_valid(R))
        {
          for(int d = 0; d < D; ++d)
          {
            const double c = 0.5 * (R.lo[d] + R.hi[d]) - p[d];
            dr += c * c;
          }
        }
        else
        {
          dr = std::numeric_limits<double>::max();
        }
        return dl > dr;
      });
    if(cur_rank == rank)  // :1045-1058
    {
      cp_index[i] = cur_idx;
      cp_dom[i] = cur_dom;
      cp_rank[i] = cur_rank;
      for(int d = 0; d < D; ++d) cp_coords[(size_t)i * D + d] = o.pts[(size_t)cur_idx * D + d];
      if(cp_dist) cp_dist[i] = std::sqrt(cur_sq);
    }
  }
}
}  // extern "C++"

// setObjectMesh (flattened points + domain ids) + generateBVHTree: boxes = BoxType{pt}, default scale factor
void* axo_dcp_create(int ndims, const double* pts, const int32_t* domain_ids, int npts)
{
  if(ndims == 2)
  {
    Dcp<2>* o = new Dcp<2>();
    o->pts.assign(pts, pts + (size_t)npts * 2);
    o->dom.assign(domain_ids, domain_ids + npts);
    if(npts > 0)
    {
      std::vector<double> boxes((size_t)npts * 4);
      for(int i = 0; i < npts; ++i)
        for(int d = 0; d < 2; ++d) boxes[(size_t)i * 4 + d] = boxes[(size_t)i * 4 + 2 + d] = pts[(size_t)i * 2 + d];
      o->bvh = create<2>(boxes.data(), npts, -1.0, -1.0);
    }
    return o;
  }
  Dcp<3>* o = new Dcp<3>();
  o->pts.assign(pts, pts + (size_t)npts * 3);
  o->dom.assign(domain_ids, domain_ids + npts);
  if(npts > 0)
  {
    std::vector<double> boxes((size_t)npts * 6);
    for(int i = 0; i < npts; ++i)
      for(int d = 0; d < 3; ++d) boxes[(size_t)i * 6 + d] = boxes[(size_t)i * 6 + 3 + d] = pts[(size_t)i * 3 + d];
    o->bvh = create<3>(boxes.data(), npts, -1.0, -1.0);
  }
  return o;
}
void axo_dcp_destroy(void* h, int ndims)
{
  if(ndims == 2)
    delete(Dcp<2>*)h;
  else
    delete(Dcp<3>*)h;
}
void axo_dcp_compute_local(void* h, int ndims, int rank, double sq_thresh, const double* q, int nq, int is_first, int32_t* cp_index,
                           int32_t* cp_dom, int32_t* cp_rank, double* cp_coords, double* cp_dist)
{
  if(ndims == 2)
    dcp_local<2>(*(Dcp<2>*)h, rank, sq_thresh, q, nq, is_first, cp_index, cp_dom, cp_rank, cp_coords, cp_dist);
  else
    dcp_local<3>(*(Dcp<3>*)h, rank, sq_thresh, q, nq, is_first, cp_index, cp_dom, cp_rank, cp_coords, cp_dist);
}
#endif  // !AXO_FLOAT_BUILD

int AXO_FN(max_threads)()
{
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

}  // extern "C"
