// leafmath.cuh -- the leaf arithmetic of the hot path on n independent items, so that the reference's own unit
// tests for it (primal/tests/primal_closest_point.cpp:239-576, primal_squared_distance.cpp:167-230,
// primal_ray_intersect.cpp:149-380, primal_boundingbox.cpp:523-569) can be run against the DEVICE functions the
// query kernels use: closest_point_tri / sqdist_point_box (sd.cuh), RayQuery (traverse.cuh), box_scale (common.cuh).
#pragma once
#include "common.cuh"
#include "sd.cuh"
#include "traverse.cuh"

namespace axb
{
// primal::closest_point(Point, Triangle, int* loc, EPS) (closest_point.hpp:162-290)
__global__ void __launch_bounds__(256) closest_point_tri_kernel(const double* __restrict__ pts, const double* __restrict__ tris, long long n,
                                                                 double eps, double* __restrict__ cp, int32_t* __restrict__ loc)
{
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if(i >= n) return;
  const double* t = tris + 9 * i;
  int l = 0;
  const V3 c = closest_point_tri(V3 {pts[3 * i], pts[3 * i + 1], pts[3 * i + 2]}, V3 {t[0], t[1], t[2]}, V3 {t[3], t[4], t[5]},
                                 V3 {t[6], t[7], t[8]}, l, eps);
  cp[3 * i] = c.x;
  cp[3 * i + 1] = c.y;
  cp[3 * i + 2] = c.z;
  loc[i] = l;
}

// primal::squared_distance(Point, BoundingBox) (squared_distance.hpp:77-100): invalid box -> DBL_MAX
__global__ void __launch_bounds__(256) sqdist_point_box_kernel(const double* __restrict__ pts, const double* __restrict__ boxes, long long n,
                                                                double* __restrict__ out)
{
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if(i >= n) return;
  Box<double, 3> b;
#pragma unroll
  for(int d = 0; d < 3; ++d)
  {
    b.lo[d] = boxes[6 * i + d];
    b.hi[d] = boxes[6 * i + 3 + d];
  }
  const double p[3] = {pts[3 * i], pts[3 * i + 1], pts[3 * i + 2]};
  out[i] = box_valid(b) ? sqdist_point_box(p, b) : DBL_MAX;
}

// the findRays predicate (spin/BVH.hpp:529-532 -> intersect_ray_impl.hpp:321-351) with the Ray constructor's normalisation
__global__ void __launch_bounds__(256) ray_box_kernel(const double* __restrict__ rays, const double* __restrict__ boxes, long long n,
                                                       int normalized, double tol, uint8_t* __restrict__ out)
{
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if(i >= n) return;
  Desc<6> d;
#pragma unroll
  for(int c = 0; c < 6; ++c) d.comp[c] = reinterpret_cast<const char*>(rays + c);
  d.stride = 48;
  RayQuery<double, 3> q;
  q.load(d, i, tol, normalized);
  Box<double, 3> b;
#pragma unroll
  for(int k = 0; k < 3; ++k)
  {
    b.lo[k] = boxes[6 * i + k];
    b.hi[k] = boxes[6 * i + 3 + k];
  }
  out[i] = q(b) ? 1 : 0;
}

// BoundingBox::scale (BoundingBox.hpp:548-561) exactly as transform_boxes applies it (build_radix_tree.hpp:85-99)
__global__ void __launch_bounds__(256) box_scale_kernel(const double* __restrict__ in, long long n, double half_scale, double* __restrict__ out)
{
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if(i >= n) return;
  Box<double, 3> b;
#pragma unroll
  for(int d = 0; d < 3; ++d)
  {
    b.lo[d] = in[6 * i + d];
    b.hi[d] = in[6 * i + 3 + d];
  }
  box_scale(b, half_scale);
#pragma unroll
  for(int d = 0; d < 3; ++d)
  {
    out[6 * i + d] = b.lo[d];
    out[6 * i + 3 + d] = b.hi[d];
  }
}
}  // namespace axb
