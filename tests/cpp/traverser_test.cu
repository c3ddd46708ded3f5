// traverser_test.cu -- a third-party-style caller of spin::BVH::getTraverser(): builds a BVH through the
// C++ shim, then walks it from ITS OWN kernel with axom_b200::spin::traverse_tree (include/axom_b200/traverser.cuh),
// the drop-in for LinearBVHTraverser::traverse_tree (spin/policy/LinearBVH.hpp:57-109).  The hits of the
// user kernel must equal findPoints' candidate lists, in the same order (bvh_traverse.hpp:66-154).
#include <cstdio>
#include <cstdlib>
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
  ab::spin::traverse_tree(tr, p, leaf, pred);
  counts[i] = k;
}

int main()
{
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
  std::printf("traverser_test: %s (%d queries, %ld candidates, %ld mismatches)\n", bad == 0 && total > 0 ? "OK" : "FAILED", Q, total, bad);
  return bad == 0 && total > 0 ? 0 : 1;
}
