// meshtester.cuh -- setup kernels of the triangle-mesh self-intersection test (sm_100a).
//
// Reference path replaced: quest::detail::CandidateFinderBase::initialize
// (quest/detail/MeshTester_detail.hpp:158-199): per cell, the Triangle3, its degenerate flag and its AABB
// (primal::compute_bounding_box, primal/operators/compute_bounding_box.hpp:111-114).  The narrow phase itself is
// the TriTriFilter of tritri.cuh, applied inside the BVH walk of traverse.cuh.
#pragma once
#include "common.cuh"
#include "tritri.cuh"

namespace axb
{
// one thread per cell: tris[c] = 96-byte record (tt::load_tri_rec), boxes[c] = AABB of the vertices, degenerate[c] = 0/1
__global__ void __launch_bounds__(256) tri_prepare_kernel(const double* __restrict__ x, const double* __restrict__ y, const double* __restrict__ z,
                                                           const int32_t* __restrict__ conn, int ncells, double* __restrict__ tris,
                                                           Box<double, 3>* __restrict__ boxes, int32_t* __restrict__ degenerate)
{
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if(c >= ncells) return;
  tt::V3 t[3];
  Box<double, 3> bb;
  box_clear(bb);
#pragma unroll
  for(int k = 0; k < 3; ++k)
  {
    const int nd = conn[(size_t)c * 3 + k];
    const double p[3] = {x[nd], y[nd], z[nd]};
    t[k] = {p[0], p[1], p[2]};
#pragma unroll
    for(int d = 0; d < 3; ++d)
    {
      if(p[d] < bb.lo[d]) bb.lo[d] = p[d];
      if(p[d] > bb.hi[d]) bb.hi[d] = p[d];
      tris[(size_t)c * tt::kTriRecDoubles + 3 * k + d] = p[d];
    }
  }
#pragma unroll
  for(int k = 9; k < tt::kTriRecDoubles; ++k) tris[(size_t)c * tt::kTriRecDoubles + k] = 0.0;
  boxes[c] = bb;
  degenerate[c] = tt::degenerate(t) ? 1 : 0;
}

// primal::intersect(tri1[i], tri2[i], include_boundary, eps) on explicit pairs (parity harness of the narrow phase)
__global__ void __launch_bounds__(256) tri_tri_pairs_kernel(const double* __restrict__ tris1, const double* __restrict__ tris2, long long n,
                                                             int include_boundary, double eps, uint8_t* __restrict__ out)
{
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if(i >= n) return;
  tt::V3 a[3], b[3];
  tt::load_tri(tris1, i, a);
  tt::load_tri(tris2, i, b);
  out[i] = tt::tri_tri(a, b, include_boundary != 0, eps) ? 1 : 0;
}

// ids of the flagged cells in ascending order: out[offsets[c]] = c where flag[c]
__global__ void __launch_bounds__(256) compact_flagged_kernel(const int32_t* __restrict__ flag, const int32_t* __restrict__ offsets, int n,
                                                               int32_t* __restrict__ out)
{
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if(c < n && flag[c]) out[offsets[c]] = c;
}
}  // namespace axb
