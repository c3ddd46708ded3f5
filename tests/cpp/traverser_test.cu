// traverser_test.cu -- a third-party-style caller of spin::BVH::getTraverser(): builds a BVH through the
// C++ shim, then walks it from ITS OWN kernel with axom_b200::spin::traverse_tree (include/axom_b200/traverser.cuh),
// the drop-in for LinearBVHTraverser::traverse_tree (spin/policy/LinearBVH.hpp:57-109).  The hits of the
// user kernel must equal findPoints' candidate lists, in the same order (bvh_traverse.hpp:66-154).
//
// A second kernel follows the other getTraverser() caller of the reference, mir::TopologyMapper
// (mir/utilities/TopologyMapper.hpp:489-560): one thread per TARGET zone walks the BVH of the SOURCE zones' bounding
// boxes with a box-intersects predicate and, in the leaf action, computes the overlap of the two zones and accumulates
// overlap / targetAmount per source material -- here for axis-aligned hexahedral zones, whose shapeOverlap is the volume
// of the box intersection.  The per-material sums must equal, bit for bit, the same sums taken on the host over
// findBoundingBoxes' candidate lists in their order.
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include <cuda_runtime.h>

#include "axom_b200/BVH.hpp"
#include "axom_b200/traverser.cuh"

namespace ab = axom_b200;
using Box3 = ab::primal::BoundingBox<double, 3>;
using Pt3 = ab::primal::Point<double, 3>;

#define CK(x)                                                                              \
  do                                                                                       \
  {                                                                                        \
    cudaError_t e_ = (x);                                                                  \
    if(e_ != cudaSuccess)                                                                  \
    {                                                                                      \
      std::fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_));                       \
      std::exit(2);                                                                        \
    }                                                                                      \
  } while(0)

__global__ void user_kernel(ab::spin::LinearBVHTraverser<double, 3> tr, const Pt3* pts, int n, const int* offsets, int* out, int* counts)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if(i >= n) return;
  const Pt3 p = pts[i];
  int k = 0;
  const int base = offsets[i];
  auto pred = [](const Pt3& q, const Box3& b) {
    for(int d = 0; d < 3; ++d)
      if(q.m_components[d] < b.m_min.m_components[d] || q.m_components[d] > b.m_max.m_components[d]) return false;
    return true;
  };
  auto leaf = [&](std::int32_t pos, const std::int32_t* leaf_nodes) { out[base + k++] = leaf_nodes[pos]; };
  ab::spin::traverse_tree(tr, p, leaf, pred, ab::spin::NoTraversePreference {});  // findPoints' own order (LinearBVH.hpp:302-364)
  counts[i] = k;
}

// mir::TopologyMapper's loop body (TopologyMapper.hpp:493-560) for axis-aligned zones
constexpr int kMaterials = 4;
__host__ __device__ inline double box_overlap(const Box3& a, const Box3& b)
{
  double v = 1.0;
  for(int d = 0; d < 3; ++d)
  {
    const double lo = a.m_min.m_components[d] > b.m_min.m_components[d] ? a.m_min.m_components[d] : b.m_min.m_components[d];
    const double hi = a.m_max.m_components[d] < b.m_max.m_components[d] ? a.m_max.m_components[d] : b.m_max.m_components[d];
    if(hi <= lo) return 0.0;
    v = v * (hi - lo);
  }
  return v;
}
__host__ __device__ inline double box_volume(const Box3& a)
{
  double v = 1.0;
  for(int d = 0; d < 3; ++d) v = v * (a.m_max.m_components[d] - a.m_min.m_components[d]);
  return v;
}

__global__ void mapper_kernel(ab::spin::LinearBVHTraverser<double, 3> tr, const Box3* src, const int* src_material, const Box3* target, int n,
                              double* vf /* [n][kMaterials] */)
{
  const int zi = blockIdx.x * blockDim.x + threadIdx.x;
  if(zi >= n) return;
  const Box3 targetBBox = target[zi];
  const double targetAmount = box_volume(targetBBox);
  double acc[kMaterials] = {0., 0., 0., 0.};
  auto bbIsect = [](const Box3& q, const Box3& b) {  // BoundingBox::intersectsWith
    for(int d = 0; d < 3; ++d)
      if(q.m_max.m_components[d] < b.m_min.m_components[d] || q.m_min.m_components[d] > b.m_max.m_components[d]) return false;
    return true;
  };
  auto handleIntersection = [&](std::int32_t currentNode, const std::int32_t* leafNodes) {
    const int srcZone = leafNodes[currentNode];
    const double srcOverlapsTarget = box_overlap(src[srcZone], targetBBox);
    if(srcOverlapsTarget > 0.)
    {
      const double f = srcOverlapsTarget / targetAmount;
      const int m = src_material[srcZone];
      for(int k = 0; k < kMaterials; ++k) acc[k] = (k == m) ? acc[k] + f : acc[k];
    }
  };
  ab::spin::traverse_tree(tr, targetBBox, handleIntersection, bbIsect);
  for(int k = 0; k < kMaterials; ++k) vf[zi * kMaterials + k] = acc[k];
}

static int test_topology_mapper_pattern()
{
  // source: 24^3 unit-ish cells of a rectilinear grid with jittered planes; target: 17^3 cells over the same region
  const int NS = 24, NT = 17;
  unsigned s = 777u;
  auto rnd = [&]() {
    s = s * 1664525u + 1013904223u;
    return (s >> 8) * (1.0 / 16777216.0);
  };
  auto planes = [&](int n) {
    std::vector<double> p(n + 1);
    for(int i = 0; i <= n; ++i) p[i] = (i + (i > 0 && i < n ? 0.6 * (rnd() - 0.5) : 0.0)) / n;
    return p;
  };
  auto cells = [&](int n, std::vector<Box3>& out) {
    const std::vector<double> px = planes(n), py = planes(n), pz = planes(n);
    for(int k = 0; k < n; ++k)
      for(int j = 0; j < n; ++j)
        for(int i = 0; i < n; ++i) out.emplace_back(Pt3 {px[i], py[j], pz[k]}, Pt3 {px[i + 1], py[j + 1], pz[k + 1]});
  };
  std::vector<Box3> src, tgt;
  cells(NS, src);
  cells(NT, tgt);
  std::vector<int> mat(src.size());
  for(auto& m : mat) m = (int)(rnd() * kMaterials) % kMaterials;
  const int nS = (int)src.size(), nT = (int)tgt.size();

  ab::spin::BVH<3> bvh;
  bvh.setScaleFactor(1.0);
  bvh.initialize(src.data(), nS);
  std::vector<ab::IndexType> off(nT), cnt(nT);
  ab::Array<ab::IndexType> cand;
  bvh.findBoundingBoxes(ab::ArrayView<ab::IndexType>(off), ab::ArrayView<ab::IndexType>(cnt), cand, nT, tgt.data());
  std::vector<double> want((size_t)nT * kMaterials, 0.0);
  for(int zi = 0; zi < nT; ++zi)
  {
    const double amount = box_volume(tgt[zi]);
    for(int c = 0; c < cnt[zi]; ++c)
    {
      const int z = cand[off[zi] + c];
      const double o = box_overlap(src[z], tgt[zi]);
      if(o > 0.) want[(size_t)zi * kMaterials + mat[z]] += o / amount;
    }
  }

  Box3 *d_src, *d_tgt;
  int* d_mat;
  double* d_vf;
  CK(cudaMalloc(&d_src, sizeof(Box3) * nS));
  CK(cudaMalloc(&d_tgt, sizeof(Box3) * nT));
  CK(cudaMalloc(&d_mat, sizeof(int) * nS));
  CK(cudaMalloc(&d_vf, sizeof(double) * nT * kMaterials));
  CK(cudaMemcpy(d_src, src.data(), sizeof(Box3) * nS, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_tgt, tgt.data(), sizeof(Box3) * nT, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_mat, mat.data(), sizeof(int) * nS, cudaMemcpyHostToDevice));
  mapper_kernel<<<(nT + 127) / 128, 128>>>(bvh.getTraverser(), d_src, d_mat, d_tgt, nT, d_vf);
  CK(cudaGetLastError());
  CK(cudaDeviceSynchronize());
  std::vector<double> got((size_t)nT * kMaterials);
  CK(cudaMemcpy(got.data(), d_vf, sizeof(double) * got.size(), cudaMemcpyDeviceToHost));
  long bad = 0;
  double worst_sum = 0.0;
  for(int zi = 0; zi < nT; ++zi)
  {
    double sum = 0.0;
    for(int k = 0; k < kMaterials; ++k)
    {
      if(got[(size_t)zi * kMaterials + k] != want[(size_t)zi * kMaterials + k]) ++bad;
      sum += got[(size_t)zi * kMaterials + k];
    }
    const double e = sum > 1.0 ? sum - 1.0 : 1.0 - sum;  // the source cells tile the region: fractions add up to 1
    if(e > worst_sum) worst_sum = e;
  }
  std::printf("traverser_test (TopologyMapper pattern): %s (%d target zones, %d source zones, %ld mismatches, |sum vf - 1| <= %.2e)\n",
              bad == 0 && worst_sum < 1e-12 ? "OK" : "FAILED", nT, nS, bad, worst_sum);
  return bad == 0 && worst_sum < 1e-12 ? 0 : 1;
}

// The nearest-neighbour pattern of spin_bvh.cpp:1401-1554 (borrowed there from DistributedClosestPoint): a BVH over
// zero-size boxes of 2-D points, one thread per query point, traverse_tree with the POINT overload, a strict-<
// leaf action and a <= predicate on the running minimum.  Which of several equidistant points is reported depends on
// the visiting order, i.e. on the centroid rule.  Results go to a file; tests/test_cpp_shim.py compares them with the
// unmodified reference's own traverser on the same input.
using Box2 = ab::primal::BoundingBox<double, 2>;
using Pt2 = ab::primal::Point<double, 2>;
__global__ void nearest_point_kernel(ab::spin::LinearBVHTraverser<double, 2> tr, const Pt2* src, const Pt2* qpts, int nq, int* min_elem,
                                     double* min_sq)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if(i >= nq) return;
  const Pt2 q = qpts[i];
  double best = 1.7976931348623157e308;
  int elem = -1;
  auto checkMinDist = [&](std::int32_t pos, const std::int32_t* leaf_nodes) {
    const int c = leaf_nodes[pos];
    double s = 0.0;
    for(int d = 0; d < 2; ++d)
    {
      const double v = src[c].m_components[d] - q.m_components[d];
      s += v * v;
    }
    if(s < best)
    {
      best = s;
      elem = c;
    }
  };
  auto traversePredicate = [&](const Pt2& p, const Box2& bb) -> bool {
    double s = 0.0;  // squared_distance(Point, BoundingBox): clamp, then sum of squares
    for(int d = 0; d < 2; ++d)
    {
      const double x = p.m_components[d], lo = bb.m_min.m_components[d], hi = bb.m_max.m_components[d];
      const double c = x < lo ? lo : (x > hi ? hi : x);
      const double v = c - x;
      s += v * v;
    }
    return s <= best;
  };
  ab::spin::traverse_tree(tr, q, checkMinDist, traversePredicate);
  min_elem[i] = elem;
  min_sq[i] = best;
}

static int run_nearest_2d(const char* src_file, const char* query_file, const char* out_file)
{
  auto slurp = [](const char* f, std::vector<double>& v) {
    FILE* fp = std::fopen(f, "rb");
    if(!fp) return false;
    std::fseek(fp, 0, SEEK_END);
    const long n = std::ftell(fp);
    std::fseek(fp, 0, SEEK_SET);
    v.resize((size_t)n / sizeof(double));
    const bool ok = v.empty() || std::fread(v.data(), 1, (size_t)n, fp) == (size_t)n;
    std::fclose(fp);
    return ok;
  };
  std::vector<double> s, q;
  if(!slurp(src_file, s) || !slurp(query_file, q)) return 3;
  const int ns = (int)(s.size() / 2), nq = (int)(q.size() / 2);
  std::vector<Box2> boxes(ns > 0 ? ns : 1);
  for(int i = 0; i < ns; ++i) boxes[i] = Box2(Pt2 {s[2 * i], s[2 * i + 1]}, Pt2 {s[2 * i], s[2 * i + 1]});  // BoxType(point)
  ab::spin::BVH<2> bvh;
  bvh.initialize(boxes.data(), ns);
  Pt2 *d_src, *d_q;
  int* d_elem;
  double* d_sq;
  CK(cudaMalloc(&d_src, sizeof(Pt2) * (ns > 0 ? ns : 1)));
  CK(cudaMalloc(&d_q, sizeof(Pt2) * nq));
  CK(cudaMalloc(&d_elem, sizeof(int) * nq));
  CK(cudaMalloc(&d_sq, sizeof(double) * nq));
  if(ns) CK(cudaMemcpy(d_src, s.data(), sizeof(double) * 2 * ns, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_q, q.data(), sizeof(double) * 2 * nq, cudaMemcpyHostToDevice));
  nearest_point_kernel<<<(nq + 127) / 128, 128>>>(bvh.getTraverser(), d_src, d_q, nq, d_elem, d_sq);
  CK(cudaGetLastError());
  CK(cudaDeviceSynchronize());
  std::vector<int> elem(nq);
  std::vector<double> sq(nq);
  CK(cudaMemcpy(elem.data(), d_elem, sizeof(int) * nq, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(sq.data(), d_sq, sizeof(double) * nq, cudaMemcpyDeviceToHost));
  FILE* fo = std::fopen(out_file, "wb");
  if(!fo) return 3;
  std::fwrite(elem.data(), sizeof(int), nq, fo);
  std::fwrite(sq.data(), sizeof(double), nq, fo);
  std::fclose(fo);
  std::printf("traverser_test nearest2d: %d source points, %d queries\n", ns, nq);
  return 0;
}

int main(int argc, char** argv)
{
  if(argc == 5 && std::string(argv[1]) == "--nearest2d") return run_nearest_2d(argv[2], argv[3], argv[4]);

  const int N = 20000, Q = 5000;
  std::vector<Box3> boxes;
  std::vector<Pt3> pts;
  unsigned s = 12345u;
  auto rnd = [&]() {
    s = s * 1664525u + 1013904223u;
    return (s >> 8) * (1.0 / 16777216.0);
  };
  const double h = 0.06;
  for(int i = 0; i < N; ++i)
  {
    const double c[3] = {rnd(), rnd(), rnd()};
    boxes.emplace_back(Pt3 {c[0] - h * rnd(), c[1] - h * rnd(), c[2] - h * rnd()}, Pt3 {c[0] + h * rnd(), c[1] + h * rnd(), c[2] + h * rnd()});
  }
  for(int i = 0; i < Q; ++i) pts.push_back(Pt3 {rnd(), rnd(), rnd()});

  ab::spin::BVH<3> bvh;
  bvh.initialize(boxes.data(), N);
  std::vector<ab::IndexType> off(Q), cnt(Q);
  ab::Array<ab::IndexType> cand;
  bvh.findPoints(ab::ArrayView<ab::IndexType>(off), ab::ArrayView<ab::IndexType>(cnt), cand, Q, pts.data());

  Pt3* d_pts;
  int *d_off, *d_out, *d_cnt;
  CK(cudaMalloc(&d_pts, sizeof(Pt3) * Q));
  CK(cudaMalloc(&d_off, sizeof(int) * Q));
  CK(cudaMalloc(&d_cnt, sizeof(int) * Q));
  CK(cudaMalloc(&d_out, sizeof(int) * (cand.size() + 1)));
  CK(cudaMemcpy(d_pts, pts.data(), sizeof(Pt3) * Q, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_off, off.data(), sizeof(int) * Q, cudaMemcpyHostToDevice));
  user_kernel<<<(Q + 127) / 128, 128>>>(bvh.getTraverser(), d_pts, Q, d_off, d_out, d_cnt);
  CK(cudaGetLastError());
  CK(cudaDeviceSynchronize());
  std::vector<int> out(cand.size() + 1), kc(Q);
  CK(cudaMemcpy(out.data(), d_out, sizeof(int) * cand.size(), cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(kc.data(), d_cnt, sizeof(int) * Q, cudaMemcpyDeviceToHost));
  long bad = 0, total = 0;
  for(int i = 0; i < Q; ++i)
  {
    if(kc[i] != cnt[i]) ++bad;
    total += cnt[i];
  }
  for(ab::IndexType i = 0; i < cand.size(); ++i)
    if(out[i] != cand[i]) ++bad;
  const int mapper_rc = test_topology_mapper_pattern();
  std::printf("traverser_test: %s (%d queries, %ld candidates, %ld mismatches)\n", bad == 0 && total > 0 && mapper_rc == 0 ? "OK" : "FAILED", Q,
              total, bad);
  return bad == 0 && total > 0 && mapper_rc == 0 ? 0 : 1;
}
