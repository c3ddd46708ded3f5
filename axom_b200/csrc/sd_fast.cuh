// sd_fast.cuh -- the fast SignedDistance query: oriented-bound overlay + Morton-coherent queries.
//
// Why: the reference prunes with the AABBs of the (fixed, parity-pinned) BVH.  For a smooth surface a
// small tilted patch has an AABB that is fat along the surface normal, so ~10^3 leaves sit within the
// AABB slack of the true minimum and every one of them costs a full closest_point evaluation
// (measured: 1100 leaf tests + 1500 inner nodes per query on the 2M-triangle icosphere).
// The tree TOPOLOGY is kept bit-identical to the reference (getTraverser() exposes it), but for our own
// traversal every child additionally carries an oriented box (OBB) whose first axis is the
// area-weighted mean normal of the triangles below it.  The OBB is a second, much tighter, conservative
// lower bound on the distance to anything in the subtree, so only the few nodes actually under the query
// survive.  Leaves that survive are evaluated with exactly the reference arithmetic (check_triangle in
// sd.cuh), and the running minimum is an exact min over exact per-triangle values, so
//   * distances are bit-identical to the reference,
//   * every candidate the reference's pseudo-normal state machine would have used is still visited:
//     the prune threshold is widened by the machine's own tie window (closest points within 1e-6,
//     quest/SignedDistance.hpp:690,709),
//   * children are entered in the reference's order (nearer AABB centroid first), so the evaluated
//     leaves are a subsequence of the reference's visiting order and every skipped leaf is a no-op of
//     the state machine: strict-< tie-breaks between equidistant triangles (closest point, minElem)
//     and the summation order of the pseudo-normal are the reference's.  (The only inputs on which
//     the two can differ are meshes with DISTINCT features closer than 1e-6 to each other, where the
//     reference's own answer already depends on its visiting order.)
// Queries are processed in Morton order of their position (one thread per query), so the threads of a
// warp walk the same few nodes and their loads coalesce into L1-resident lines.
#pragma once
#include "common.cuh"
#include "sd.cuh"

namespace axb
{
// oriented bound of one child: rows of `axis` are (nearly) orthonormal; the subtree's vertices satisfy
// lo[k] <= axis[k] . x <= hi[k].  128 bytes = one L2 line.
struct alignas(32) Obb
{
  double axis[3][3];
  double lo[3];
  double hi[3];
  double pad_;
};
static_assert(sizeof(Obb) == 128, "Obb must be one 128-byte line");

constexpr int kObbMaxRange = 4096;  // subtrees with more leaves keep only their AABB (top ~9 levels)

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
  for(int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_min(double v)
{
#pragma unroll
  for(int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ double warp_max(double v)
{
#pragma unroll
  for(int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// One warp per tree entity e: e < inner -> inner node e (leaves node_range[e]), else leaf e - inner.
// obb[e] bounds everything below entity e.
template <int NV>
__global__ void __launch_bounds__(256) obb_build_kernel(const double* __restrict__ soup, const int2* __restrict__ node_range, int nleaves,
                                                         Obb* __restrict__ obb)
{
  const int inner = nleaves - 1;
  const int e = (int)((blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5);
  if(e >= inner + nleaves) return;
  const int lane = (int)lane_id();
  int first, last;
  if(e < inner)
  {
    const int2 r = node_range[e];
    first = r.x;
    last = r.y;
  }
  else
  {
    first = last = e - inner;
  }
  Obb o;
  if(last - first + 1 > kObbMaxRange)
  {
    // no oriented bound: gaps evaluate to 0, the AABB decides
    if(lane == 0)
    {
      for(int k = 0; k < 3; ++k)
      {
        for(int c = 0; c < 3; ++c) o.axis[k][c] = 0.0;
        o.lo[k] = -DBL_MAX;
        o.hi[k] = DBL_MAX;
      }
      o.pad_ = 0.0;
      obb[e] = o;
    }
    return;
  }
  // pass 1: area-weighted normal sum
  double nx = 0.0, ny = 0.0, nz = 0.0;
  for(int p = first + lane; p <= last; p += 32)
  {
    V3 v[NV];
    load_leaf<NV>(soup, p, v);
    V3 c = v3cross(v3sub(v[1], v[0]), v3sub(v[2], v[0]));
    if(NV == 4) c = v3add(c, v3cross(v3sub(v[2], v[0]), v3sub(v[NV - 1], v[0])));
    nx += c.x;
    ny += c.y;
    nz += c.z;
  }
  nx = warp_sum(nx);
  ny = warp_sum(ny);
  nz = warp_sum(nz);
  double len2 = nx * nx + ny * ny + nz * nz;
  V3 n;
  if(len2 > 1e-280 && len2 < 1e280)
  {
    const double s = 1.0 / sqrt(len2);
    n = {nx * s, ny * s, nz * s};
  }
  else
  {
    n = {0.0, 0.0, 1.0};
  }
  // tangent frame: t1 = normalise(n x e_k) with e_k the coordinate axis least aligned with n
  const double ax = fabs(n.x), ay = fabs(n.y), az = fabs(n.z);
  V3 ek = (ax <= ay && ax <= az) ? V3 {1.0, 0.0, 0.0} : ((ay <= az) ? V3 {0.0, 1.0, 0.0} : V3 {0.0, 0.0, 1.0});
  V3 t1 = v3cross(n, ek);
  {
    const double s = 1.0 / sqrt(v3dot(t1, t1));
    t1 = v3mul(t1, s);
  }
  const V3 t2 = v3cross(n, t1);
  const V3 A[3] = {n, t1, t2};
  // pass 2: extents of all vertices along the frame
  double lo[3] = {DBL_MAX, DBL_MAX, DBL_MAX}, hi[3] = {-DBL_MAX, -DBL_MAX, -DBL_MAX};
  for(int p = first + lane; p <= last; p += 32)
  {
    V3 v[NV];
    load_leaf<NV>(soup, p, v);
#pragma unroll
    for(int j = 0; j < NV; ++j)
    {
#pragma unroll
      for(int k = 0; k < 3; ++k)
      {
        const double d = A[k].x * v[j].x + A[k].y * v[j].y + A[k].z * v[j].z;
        lo[k] = fmin(lo[k], d);
        hi[k] = fmax(hi[k], d);
      }
    }
  }
#pragma unroll
  for(int k = 0; k < 3; ++k)
  {
    lo[k] = warp_min(lo[k]);
    hi[k] = warp_max(hi[k]);
  }
  if(lane == 0)
  {
#pragma unroll
    for(int k = 0; k < 3; ++k)
    {
      o.axis[k][0] = A[k].x;
      o.axis[k][1] = A[k].y;
      o.axis[k][2] = A[k].z;
      // pad by the rounding of the projections (a few ulp of the coordinate magnitude)
      const double pad = 1e-14 * (fabs(lo[k]) + fabs(hi[k])) + 1e-300;
      o.lo[k] = lo[k] - pad;
      o.hi[k] = hi[k] + pad;
    }
    o.pad_ = 0.0;
    obb[e] = o;
  }
}

// squared distance lower bound from q to the oriented box (FMA is fine here: it only has to be a bound)
__device__ __forceinline__ double obb_sqdist(const Obb* __restrict__ ob, const double* q)
{
  double s = 0.0;
  const double2* p = reinterpret_cast<const double2*>(ob);
  // 15 doubles as 8 x 16-byte loads: axis[0..2][0..2], lo[0..2], hi[0..2]
  const double2 a0 = __ldg(p + 0), a1 = __ldg(p + 1), a2 = __ldg(p + 2), a3 = __ldg(p + 3), a4 = __ldg(p + 4), a5 = __ldg(p + 5),
                a6 = __ldg(p + 6), a7 = __ldg(p + 7);
  const double ax[3][3] = {{a0.x, a0.y, a1.x}, {a1.y, a2.x, a2.y}, {a3.x, a3.y, a4.x}};
  const double lo[3] = {a4.y, a5.x, a5.y};
  const double hi[3] = {a6.x, a6.y, a7.x};
#pragma unroll
  for(int k = 0; k < 3; ++k)
  {
    const double d = fma(ax[k][0], q[0], fma(ax[k][1], q[1], ax[k][2] * q[2]));
    const double g = fmax(fmax(lo[k] - d, d - hi[k]), 0.0);
    s = fma(g, g, s);
  }
  return s;
}

// prune threshold on squared distance: everything that could still improve the minimum or tie with it
// inside the reference's 1e-6 closest-point window (sqrt(EPS), EPS = 1e-12), plus rounding head-room.
__device__ __forceinline__ double prune_threshold(double minSq)
{
  if(minSq >= 1e300) return DBL_MAX;
  const double d = sqrt(minSq) + 1.0000001e-6;
  return d * d * (1.0 + 1e-12) + 1e-300;
}

// MODE 1 kernel: one thread per query, queries taken in Morton order (perm), reference-ordered DFS
// with a (node, lower bound) stack so stale entries are dropped without touching memory.
template <int NV>
__global__ void __launch_bounds__(128) sd_fast_kernel(const Node<double, 3>* __restrict__ nodes, const Obb* __restrict__ obb,
                                                       const double* __restrict__ soup, int nleaves, SdParams prm, Desc<3> qpts, int npts,
                                                       const int32_t* __restrict__ perm, double* __restrict__ phi, double* __restrict__ cps,
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
  const int inner = nleaves - 1;
  double thr = DBL_MAX;
  unsigned nleaf = 0, ninner = 0;

  int32_t st_node[kStackSize];
  float st_d2[kStackSize];
  int sp = 0;
  int32_t cur = 0;  // root
  while(true)
  {
    int32_t next = kBarrier;
    if(cur < 0)
    {
      // ---- leaf: the reference's checkCandidate, then tighten the threshold ----
      ++nleaf;
      check_leaf<NV>(soup, q, m, -cur - 1, cn);
      thr = prune_threshold(m.minSq);
    }
    else
    {
      // ---- inner node ----
      ++ninner;
      const Node<double, 3>& nd = nodes[cur];
      int32_t child[2] = {nd.child[0], nd.child[1]};
      const Box<double, 3> bx[2] = {nd.box[0], nd.box[1]};
      double d2[2];
      bool in[2];
#pragma unroll
      for(int s = 0; s < 2; ++s)
      {
        double v = DBL_MAX;
        bool ok = box_valid(bx[s]);  // bvh_traverse.hpp:95-96: invalid boxes are never entered
        if(ok)
        {
          v = sqdist_point_box(qp, bx[s]);
          ok = v <= thr;
          if(ok)
          {
            const int e = child[s] >= 0 ? child[s] : inner + (-child[s] - 1);
            v = fmax(v, obb_sqdist(obb + e, qp));
            ok = v <= thr;
          }
        }
        d2[s] = v;
        in[s] = ok;
      }
      // Child order = the reference's (LinearBVH.hpp:72-85): when both children are entered, the one whose
      // AABB centroid is nearer goes first and the other waits on the stack -- leaf or not -- until
      // everything below the first is done.  The leaves this kernel evaluates are then a SUBSEQUENCE of
      // the reference's own visiting order, so strict-< tie-breaks between equidistant triangles and the
      // summation order of the pseudo-normal are the reference's.
      if(in[0] && in[1])
      {
        double dl = 0.0, dr = 0.0;
#pragma unroll
        for(int d = 0; d < 3; ++d)
        {
          const double cl = 0.5 * (bx[0].lo[d] + bx[0].hi[d]) - qp[d];
          dl += cl * cl;
          const double cr = 0.5 * (bx[1].lo[d] + bx[1].hi[d]) - qp[d];
          dr += cr * cr;
        }
        const int first = dl > dr ? 1 : 0;
        next = child[first];
        st_node[sp] = child[first ^ 1];
        st_d2[sp] = __double2float_rd(d2[first ^ 1]);
        ++sp;
      }
      else if(in[0])
      {
        next = child[0];
      }
      else if(in[1])
      {
        next = child[1];
      }
    }
    // ---- pop until an entry still beats the threshold (stale entries cost no memory access) ----
    while(next == kBarrier && sp > 0)
    {
      --sp;
      if((double)st_d2[sp] <= thr) next = st_node[sp];
    }
    if(next == kBarrier) break;
    cur = next;
  }
  sd_finish<NV>(soup, prm, q, m, qi, phi, cps, nrms);
  if(work)
  {
    atomicAdd(&work[0], (unsigned long long)nleaf);
    atomicAdd(&work[1], (unsigned long long)ninner);
  }
}

// Morton keys of query points over the BVH bounds: (code30 << 32) | index, for radix_sort.cuh
__global__ void __launch_bounds__(256) query_keys_kernel(Desc<3> qpts, int npts, const unsigned long long* __restrict__ ob /* [6] ordered */,
                                                          unsigned long long* __restrict__ keys, uint32_t* __restrict__ ghist)
{
  __shared__ uint32_t sh[rsort::MAX_PASSES * rsort::RADIX];
  for(int i = threadIdx.x; i < rsort::MAX_PASSES * rsort::RADIX; i += blockDim.x) sh[i] = 0;
  __syncthreads();
  const double lox = ordered_to_f64(ob[0]), loy = ordered_to_f64(ob[1]), loz = ordered_to_f64(ob[2]);
  const double ex = ordered_to_f64(ob[3]) - lox, ey = ordered_to_f64(ob[4]) - loy, ez = ordered_to_f64(ob[5]) - loz;
  const double sx = ex > 0.0 ? 1024.0 / ex : 0.0, sy = ey > 0.0 ? 1024.0 / ey : 0.0, sz = ez > 0.0 ? 1024.0 / ez : 0.0;
  for(int i = blockIdx.x * blockDim.x + threadIdx.x; i < npts; i += gridDim.x * blockDim.x)
  {
    const double x = (ld_comp<double>(qpts, 0, i) - lox) * sx;
    const double y = (ld_comp<double>(qpts, 1, i) - loy) * sy;
    const double z = (ld_comp<double>(qpts, 2, i) - loz) * sz;
    const uint32_t qx = (uint32_t)(int)fmin(fmax(x, 0.0), 1023.0);
    const uint32_t qy = (uint32_t)(int)fmin(fmax(y, 0.0), 1023.0);
    const uint32_t qz = (uint32_t)(int)fmin(fmax(z, 0.0), 1023.0);
    const uint32_t code = spread_bits_3d(qx) | (spread_bits_3d(qy) << 1) | (spread_bits_3d(qz) << 2);
    keys[i] = ((unsigned long long)code << 32) | (unsigned long long)(uint32_t)i;
#pragma unroll
    for(int p = 0; p < rsort::MAX_PASSES; ++p) atomicAdd(&sh[p * rsort::RADIX + ((code >> (p * 8)) & 255u)], 1u);
  }
  __syncthreads();
  for(int i = threadIdx.x; i < rsort::MAX_PASSES * rsort::RADIX; i += blockDim.x)
    if(sh[i]) atomicAdd(&ghist[i], sh[i]);
}

__global__ void __launch_bounds__(256) keys_to_perm_kernel(const unsigned long long* __restrict__ keys, int n, int32_t* __restrict__ perm)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if(i < n) perm[i] = (int32_t)(uint32_t)(keys[i] & 0xffffffffull);
}

// min / max of the query coordinates (the Morton quantisation box of the queries)
__global__ void __launch_bounds__(256) query_bounds_kernel(Desc<3> qpts, int npts, unsigned long long* __restrict__ ob /* [6] */)
{
  double mn[3] = {DBL_MAX, DBL_MAX, DBL_MAX}, mx[3] = {-DBL_MAX, -DBL_MAX, -DBL_MAX};
  for(int i = blockIdx.x * blockDim.x + threadIdx.x; i < npts; i += gridDim.x * blockDim.x)
  {
#pragma unroll
    for(int d = 0; d < 3; ++d)
    {
      const double v = ld_comp<double>(qpts, d, i);
      mn[d] = fmin(mn[d], v);
      mx[d] = fmax(mx[d], v);
    }
  }
#pragma unroll
  for(int d = 0; d < 3; ++d)
  {
    mn[d] = warp_min(mn[d]);
    mx[d] = warp_max(mx[d]);
    if(lane_id() == 0)
    {
      atomicMin(&ob[d], f64_to_ordered(mn[d]));
      atomicMax(&ob[3 + d], f64_to_ordered(mx[d]));
    }
  }
}

}  // namespace axb
