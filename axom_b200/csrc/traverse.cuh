// traverse.cuh -- BVH traversal and the batched candidate queries (sm_100a).
//
// Reference path replaced:
//   lbvh::bvh_traverse                  spin/internal/linear_bvh/bvh_traverse.hpp:66-154
//   LinearBVH::findCandidatesImpl       spin/policy/LinearBVH.hpp:271-402  (count -> scan -> fill)
//   predicates of findPoints/Boxes/Rays spin/BVH.hpp:499-501, :558-560, :529-532
#pragma once
#include "common.cuh"

namespace axb
{
constexpr int32_t kBarrier = -2000000000;  // bvh_traverse.hpp:80
constexpr int kStackSize = 64;             // bvh_traverse.hpp:79

// The reference's traversal, restated over the packed node records.  Visits exactly the
// same nodes in exactly the same order as bvh_traverse (including its "park the first leaf,
// keep descending, then handle up to two leaves" shape), so per-query candidate order and
// the SignedDistance state machine evolve identically.
//   pred(box)       -> bool   "B"     (only called on valid boxes, :95-96)
//   order(L, R)     -> bool   "Comp"  true = visit the right child first
//   leaf(sorted_pos)          "A"
template <typename T, int D, class Pred, class Leaf, class Order>
__device__ __forceinline__ void traverse_reference_order(const Node<T, D>* __restrict__ nodes, Pred&& pred, Leaf&& leaf, Order&& order)
{
  int32_t todo[kStackSize];
  int sp = 0;
  todo[0] = kBarrier;
  int32_t found = 0;
  int32_t cur = 0;
  while(cur != kBarrier)
  {
    while(cur >= 0)
    {
      const Node<T, D>& nd = nodes[cur];
      Box<T, D> L = nd.box[0];
      Box<T, D> R = nd.box[1];
      const int32_t lc = nd.child[0];
      int32_t rc = nd.child[1];
      const bool inL = box_valid(L) ? pred(L) : false;
      const bool inR = box_valid(R) ? pred(R) : false;
      if(!inL && !inR)
      {
        cur = todo[sp--];
      }
      else
      {
        cur = inL ? lc : rc;
        if(inL && inR)
        {
          if(order(L, R))
          {
            const int32_t t = cur;
            cur = rc;
            rc = t;
          }
          todo[++sp] = rc;
        }
      }
      if(cur < 0 && !(found < 0))
      {
        found = cur;
        if(cur != kBarrier) cur = todo[sp--];
      }
    }
    while(found < 0 && found != kBarrier)
    {
      leaf(-found - 1);
      found = cur;
      if(cur < 0 && cur != kBarrier) cur = todo[sp--];
    }
    found = 0;
  }
}

struct NoOrder
{
  template <class B>
  __device__ __forceinline__ bool operator()(const B&, const B&) const
  {
    return false;
  }
};

//------------------------------------------------------------------------------------------
// query primitives
//------------------------------------------------------------------------------------------
// bb.contains(p): closed intervals (primal/geometry/BoundingBox.hpp:390-401)
template <typename T, int D>
struct PointQuery
{
  static constexpr int NCOMP = D;
  T p[D];
  __device__ __forceinline__ void load(const Desc<NCOMP>& d, long long i, T /*tol*/, int /*flags*/)
  {
#pragma unroll
    for(int k = 0; k < D; ++k) p[k] = ld_comp<T>(d, k, i);
  }
  __device__ __forceinline__ bool operator()(const Box<T, D>& bb) const
  {
    bool in = true;
#pragma unroll
    for(int k = 0; k < D; ++k) in = in && !(p[k] < bb.lo[k] || p[k] > bb.hi[k]);
    return in;
  }
};

// bb1.intersectsWith(bb2), bb1 = query (primal/operators/detail/intersect_bounding_box_impl.hpp:34-41)
template <typename T, int D>
struct BoxQuery
{
  static constexpr int NCOMP = 2 * D;
  Box<T, D> b;
  __device__ __forceinline__ void load(const Desc<NCOMP>& d, long long i, T /*tol*/, int /*flags*/) { b = load_box<T, D>(d, i); }
  __device__ __forceinline__ bool operator()(const Box<T, D>& bb) const
  {
    bool hit = true;
#pragma unroll
    for(int k = 0; k < D; ++k) hit = hit && ((b.hi[k] >= bb.lo[k]) && (b.lo[k] <= bb.hi[k]));
    return hit;
  }
};

// primal::detail::intersect_ray(R, bb, ip, TOL): per-dimension slab test with short circuit
// (primal/operators/detail/intersect_ray_impl.hpp:150-187, :321-351)
template <typename T, int D>
struct RayQuery
{
  static constexpr int NCOMP = 2 * D;
  T o[D];
  T dir[D];
  T tol;
  __device__ __forceinline__ void load(const Desc<NCOMP>& d, long long i, T tol_, int normalized)
  {
    tol = tol_;
#pragma unroll
    for(int k = 0; k < D; ++k)
    {
      o[k] = ld_comp<T>(d, k, i);
      dir[k] = ld_comp<T>(d, D + k, i);
    }
    if(!normalized)
    {
      // Ray ctor -> Vector::unitVector (primal/geometry/Ray.hpp:122-127, Vector.hpp:477-493)
      double len2 = 0.0;
#pragma unroll
      for(int k = 0; k < D; ++k) len2 += (double)(dir[k] * dir[k]);
      if(len2 >= 1e-50)
      {
        const double s = 1. / sqrt(len2);  // NumericArray::operator/= (core/NumericArray.hpp:510-514)
#pragma unroll
        for(int k = 0; k < D; ++k) dir[k] = static_cast<T>(dir[k] * s);
      }
      else
      {
        dir[0] = (T)1;
#pragma unroll
        for(int k = 1; k < D; ++k) dir[k] = (T)0;
      }
    }
  }
  __device__ __forceinline__ bool operator()(const Box<T, D>& bb) const
  {
    T tmin = Lim<T>::min();
    T tmax = Lim<T>::max();
#pragma unroll
    for(int k = 0; k < D; ++k)
    {
      const T diff = dir[k] - (T)0;
      if((diff < 0 ? -diff : diff) <= tol)
      {
        if(o[k] < bb.lo[k] || o[k] > bb.hi[k]) return false;
      }
      else
      {
        const T invn = (T)1.0 / dir[k];
        T t1 = (bb.lo[k] - o[k]) * invn;
        T t2 = (bb.hi[k] - o[k]) * invn;
        if(t1 > t2)
        {
          const T t = t1;
          t1 = t2;
          t2 = t;
        }
        tmin = (t1 < tmin) ? tmin : t1;  // utilities::max (core/utilities/Utilities.hpp:80-83)
        tmax = (t2 < tmax) ? t2 : tmax;  // utilities::min (:93-96)
        if(tmin > tmax) return false;
      }
    }
    return true;
  }
};

//------------------------------------------------------------------------------------------
// PASS 1: count (LinearBVH.hpp:302-321).  perm == nullptr: thread t handles query t.
//------------------------------------------------------------------------------------------
template <typename T, int D, class Query>
__global__ void __launch_bounds__(256) count_kernel(const Node<T, D>* __restrict__ nodes, Desc<Query::NCOMP> prims, int nq, T tol, int flags,
                                                     const int32_t* __restrict__ perm, int32_t* __restrict__ counts)
{
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if(t >= nq) return;
  const int qi = perm ? perm[t] : t;
  Query q;
  q.load(prims, qi, tol, flags);
  int c = 0;
  traverse_reference_order<T, D>(nodes, q, [&](int) { ++c; }, NoOrder {});
  counts[qi] = c;
}

//------------------------------------------------------------------------------------------
// PASS 2: fill (LinearBVH.hpp:346-364): candidates[offset++] = leaf_nodes[pos]
//------------------------------------------------------------------------------------------
template <typename T, int D, class Query>
__global__ void __launch_bounds__(256) fill_kernel(const Node<T, D>* __restrict__ nodes, const int32_t* __restrict__ leaf_nodes,
                                                    Desc<Query::NCOMP> prims, int nq, T tol, int flags, const int32_t* __restrict__ perm,
                                                    const int32_t* __restrict__ offsets, int32_t* __restrict__ candidates)
{
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if(t >= nq) return;
  const int qi = perm ? perm[t] : t;
  Query q;
  q.load(prims, qi, tol, flags);
  int off = offsets[qi];
  traverse_reference_order<T, D>(
    nodes, q, [&](int pos) { candidates[off++] = __ldg(leaf_nodes + pos); }, NoOrder {});
}

//------------------------------------------------------------------------------------------
// exclusive scan of int32 counts -> int32 offsets, int64 total (RAJA::exclusive_scan,
// LinearBVH.hpp:332-334).  Three small kernels; 12 B/query of traffic.
//------------------------------------------------------------------------------------------
constexpr int SCAN_BLOCK = 256;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_TILE = SCAN_BLOCK * SCAN_ITEMS;

__device__ __forceinline__ long long block_reduce_sum(long long v, long long* sh)
{
#pragma unroll
  for(int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if(lane_id() == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  long long r = 0;
  for(int w = 0; w < (int)(blockDim.x >> 5); ++w) r += sh[w];
  __syncthreads();
  return r;
}

__global__ void __launch_bounds__(SCAN_BLOCK) scan_tile_sums_kernel(const int32_t* __restrict__ counts, int n, long long* __restrict__ tile_sums)
{
  __shared__ long long sh[SCAN_BLOCK / 32];
  const int base = blockIdx.x * SCAN_TILE;
  long long s = 0;
#pragma unroll
  for(int k = 0; k < SCAN_ITEMS; ++k)
  {
    const int i = base + k * SCAN_BLOCK + threadIdx.x;
    if(i < n) s += counts[i];
  }
  s = block_reduce_sum(s, sh);
  if(threadIdx.x == 0) tile_sums[blockIdx.x] = s;
}

// single block: exclusive scan of the tile sums in place, total -> *total
__global__ void __launch_bounds__(1024) scan_spine_kernel(long long* tile_sums, int ntiles, long long* total)
{
  __shared__ long long sh[1024];
  __shared__ long long carry;
  if(threadIdx.x == 0) carry = 0;
  __syncthreads();
  for(int base = 0; base < ntiles; base += 1024)
  {
    const int i = base + threadIdx.x;
    const long long v = (i < ntiles) ? tile_sums[i] : 0;
    sh[threadIdx.x] = v;
    __syncthreads();
    long long incl = v;
    for(int off = 1; off < 1024; off <<= 1)
    {
      const long long u = (threadIdx.x >= (unsigned)off) ? sh[threadIdx.x - off] : 0;
      __syncthreads();
      incl += u;
      sh[threadIdx.x] = incl;
      __syncthreads();
    }
    if(i < ntiles) tile_sums[i] = carry + incl - v;
    __syncthreads();
    if(threadIdx.x == 1023) carry += incl;
    __syncthreads();
  }
  if(threadIdx.x == 0) *total = carry;
}

__global__ void __launch_bounds__(SCAN_BLOCK) scan_apply_kernel(const int32_t* __restrict__ counts, int n, const long long* __restrict__ tile_offsets,
                                                                 int32_t* __restrict__ offsets)
{
  __shared__ int wsum[SCAN_BLOCK / 32];
  const int base = blockIdx.x * SCAN_TILE;
  // blocked arrangement: thread owns SCAN_ITEMS consecutive elements
  const int first = base + threadIdx.x * SCAN_ITEMS;
  int v[SCAN_ITEMS];
  int tsum = 0;
#pragma unroll
  for(int k = 0; k < SCAN_ITEMS; ++k)
  {
    const int i = first + k;
    v[k] = (i < n) ? counts[i] : 0;
    tsum += v[k];
  }
  int incl = tsum;
#pragma unroll
  for(int o = 1; o < 32; o <<= 1)
  {
    const int u = __shfl_up_sync(0xffffffffu, incl, o);
    if((int)lane_id() >= o) incl += u;
  }
  if(lane_id() == 31) wsum[threadIdx.x >> 5] = incl;
  __syncthreads();
  int wbase = 0;
  for(int w = 0; w < (int)(threadIdx.x >> 5); ++w) wbase += wsum[w];
  long long run = tile_offsets[blockIdx.x] + wbase + (incl - tsum);
#pragma unroll
  for(int k = 0; k < SCAN_ITEMS; ++k)
  {
    const int i = first + k;
    if(i < n) offsets[i] = (int32_t)run;
    run += v[k];
  }
}

}  // namespace axb
