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
#include <cuda_fp16.h>

#include "common.cuh"
#include "sd.cuh"

namespace axb
{
// What the fast traversal keeps per CHILD of an inner node: an oriented bound -- unit axes n, t1
// (t2 = n x t1 is recomputed) and the extent lo[k] <= axis_k . (x - org) <= hi[k] of every vertex x below
// the child, org being a point of the PARENT (the centroid of its AABB) so the numbers are node-sized and
// survive storage in binary32: axes rounded to float, extents rounded OUTWARD to float.  All arithmetic
// on them is done in double on the exactly-converted floats (build and query share obb_axes()), so the
// bound stays conservative; the frame is orthonormal only to ~1e-7, which bound_scale accounts for.
// Large subtrees (no useful orientation) and invalid boxes use the coordinate axes: the bound is the AABB.
//
// Traversal record of the fast path: everything one inner-node visit needs in ONE aligned 128-byte line
// (4 x LDG.256 per lane).  A lane reads its own record, nothing coalesces, every 32-byte sector is one
// L1 wavefront -- the limiter of this kernel (profiles/r1c: l1tex 83 %) -- so bytes per visit are what
// counts.  The centroids of the child AABBs (the reference orders children by them, LinearBVH.hpp:72-85;
// precomputed with the reference's own 0.5*(min+max), kept in double because the ORDER must be exact)
// live in a second 64-byte record that is read only when both children are entered.
// The reference-layout view for getTraverser() and the reference-order kernel keep using Node<>.
struct alignas(32) SdNode
{
  int32_t child[2];  // >= 0 inner node, < 0 leaf -(sorted_pos+1)
  double org[3];
  float cb[2][12];  // per child: n[3], t1[3], c[3] (extent centre), h[3] (half extent, rounded up; -inf = invalid)
};
static_assert(sizeof(SdNode) == 128, "SdNode is 4 x 32 B");
struct alignas(32) SdCen
{
  double cen[2][3];
  double pad_[2];
};
static_assert(sizeof(SdCen) == 64, "SdCen is 2 x 32 B");

// The same bounds in HALF the bytes, for the order-free search of sd_two.cuh, which is bound by the L1 data pipe (every
// lane reads its own record: one wavefront per 32-byte sector per lane).  2 x LDG.256 per visit instead of 4:
//   frame    the first axis n as three 16-bit integers (direction only; decoded = normalised in binary32); the second
//            axis is not stored: t1 = normalise(n x e_k), e_k the coordinate axis least aligned with n, t2 = n x t1, all
//            in binary32 by sd_frame() -- the ONE function the build and the queries both call, so the extents are
//            measured in exactly the frame the query evaluates (SdNode stores the same frame as floats);
//   centres  16-bit integers in units of `step` (>= the node's half diagonal / 16000), decoded c = float(ci) * step;
//   extents  binary16 in units of `step`, rounded UP against the DECODED centre (-inf marks an invalid child);
//   origin   three floats (the node's AABB centre rounded to binary32; SdNode carries the same value as doubles).
// Everything lossy happens in the build, before the extents are measured: the bound stays conservative and merely a
// little looser (centre quantum = node size / 32000).
struct alignas(32) SdNode64
{
  int32_t child[2];
  float org[3];
  float step;
  int16_t nq[2][3];
  int16_t cq[2][3];
  uint16_t hq[2][3];
  uint32_t pad_;
};
static_assert(sizeof(SdNode64) == 64, "SdNode64 is 2 x 32 B");

// MUFU.RSQ: one instruction, 2 ulp, and -- what matters here -- a pure function of its input on a given GPU: the build
// and the queries run on the same device and get the same bits
__device__ __forceinline__ float sd_rsqrt(float x)
{
  float y;
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// frame from the 16-bit code of its first axis: fr[0..2] = n, fr[3..5] = t1 (t2 is obb_axes' binary32 cross product);
// the axes are unit to 3e-7
__device__ __forceinline__ void sd_frame(int cx, int cy, int cz, float* fr)
{
  const float vx = (float)cx, vy = (float)cy, vz = (float)cz;
  const float inv = sd_rsqrt(__fadd_rn(__fadd_rn(__fmul_rn(vx, vx), __fmul_rn(vy, vy)), __fmul_rn(vz, vz)));
  const float nx = __fmul_rn(vx, inv), ny = __fmul_rn(vy, inv), nz = __fmul_rn(vz, inv);
  const float ax = fabsf(nx), ay = fabsf(ny), az = fabsf(nz);
  // n x e_k for the axis k least aligned with n
  const bool kx = ax <= ay && ax <= az, ky = !kx && ay <= az;
  const float ux = kx ? 0.f : (ky ? -nz : ny), uy = kx ? nz : (ky ? 0.f : -nx), uz = kx ? -ny : (ky ? nx : 0.f);
  const float iu = sd_rsqrt(__fadd_rn(__fadd_rn(__fmul_rn(ux, ux), __fmul_rn(uy, uy)), __fmul_rn(uz, uz)));
  fr[0] = nx;
  fr[1] = ny;
  fr[2] = nz;
  fr[3] = __fmul_rn(ux, iu);
  fr[4] = __fmul_rn(uy, iu);
  fr[5] = __fmul_rn(uz, iu);
}
// 16-bit code of a unit vector (any scaling of the direction will do; the largest component uses the full range)
__device__ __forceinline__ void sd_frame_code(const V3& n, int* code)
{
  const double m = fmax(fabs(n.x), fmax(fabs(n.y), fabs(n.z)));
  const double s = m > 0.0 ? 32767.0 / m : 0.0;
  code[0] = (int)rint(n.x * s);
  code[1] = (int)rint(n.y * s);
  code[2] = (int)rint(n.z * s);
  if(code[0] == 0 && code[1] == 0 && code[2] == 0) code[2] = 32767;
}

// |M v|^2 <= (1 + 5e-7) |v|^2 for the rows M of a float-rounded frame: scale the bound down accordingly
constexpr double kBoundScale = 1.0 - 2.0e-6;

// the frame in double from the stored floats; the build and the queries MUST agree bit for bit on t2, which is the
// BINARY32 cross product of the two stored axes (separately rounded products and differences) so that the binary32
// bound of sd_two.cuh sees exactly the frame the extents were measured in
__device__ __forceinline__ void obb_axes(const float* f, V3* A)
{
  A[0] = {(double)f[0], (double)f[1], (double)f[2]};
  A[1] = {(double)f[3], (double)f[4], (double)f[5]};
  A[2] = {(double)__fsub_rn(__fmul_rn(f[1], f[5]), __fmul_rn(f[4], f[2])), (double)__fsub_rn(__fmul_rn(f[3], f[2]), __fmul_rn(f[0], f[5])),
          (double)__fsub_rn(__fmul_rn(f[0], f[4]), __fmul_rn(f[3], f[1]))};
}

// extent [lo, hi] of a child along axis k, stored as centre c (float, nearest) and half extent h (float, rounded UP
// from the larger one-sided distance to the ROUNDED centre), so that  max(lo - d, d - hi, 0) >= max(|d - c| - h, 0):
// one |.|, one subtraction and one clamp per axis in the query instead of three double-precision fmax
__device__ __forceinline__ void store_extent(float* out, int k, double lo, double hi)
{
  const float c = __double2float_rn(0.5 * (lo + hi));
  const double cd = (double)c;
  out[6 + k] = c;
  // the relative pad covers the two binary32 subtractions of obb_sqdist_f32 (2^-23 (|c| + h))
  const double h = fmax(hi - cd, cd - lo);
  out[9 + k] = __double2float_ru(h + 2.4e-7 * (fabs(cd) + h));
}

constexpr int kObbMaxRange = 262144;  // subtrees with more leaves keep only their AABB (an orientation no longer helps)

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
// The oriented bound of entity e is stored in its PARENT's record (slot = which child it is); the warp
// of an inner entity also writes that node's child ids and origin into its own record.
// origin of a node's records: the centre of its AABB rounded to binary32 (both record formats carry the same value),
// and the quantum of the compact record's centres / extents
__device__ __forceinline__ void node_origin(const Node<double, 3>& nd, double* org, float* step = nullptr)
{
  Box<double, 3> u = nd.box[0];
  box_add(u, nd.box[1]);
  const bool ok = box_valid(u);
  double diag2 = 0.0, off = 0.0;
#pragma unroll
  for(int d = 0; d < 3; ++d)
  {
    const double c = ok ? 0.5 * (u.lo[d] + u.hi[d]) : 0.0;
    const float cf = (float)c;
    org[d] = (fabsf(cf) <= 3.0e38f) ? (double)cf : 0.0;
    const double e = ok ? fmax(u.hi[d] - org[d], org[d] - u.lo[d]) : 0.0;  // half extent seen from the rounded origin
    diag2 += e * e;
    off += fabs(c - org[d]);
  }
  if(step)
  {
    const double q = (sqrt(diag2) * (1.0 + 1e-6) + off) / 16000.0;
    *step = (q > 1e-37 && q < 1e37) ? __double2float_ru(q) : (q >= 1e37 ? 3.0e38f : 1e-37f);
  }
}

// one axis of a child's extent in the compact record: centre on the 16-bit grid, half extent in binary16 units of step,
// rounded up against the decoded centre (pads as in store_extent)
__device__ __forceinline__ void store_extent64(int16_t* cq, uint16_t* hq, int k, float step, double lo, double hi)
{
  const double mid = 0.5 * (lo + hi);
  double t = rint(mid / (double)step);
  t = fmin(fmax(t, -32767.0), 32767.0);
  const int ci = (int)t;
  const double cd = (double)__fmul_rn((float)ci, step);  // what the query decodes
  const double h = fmax(hi - cd, cd - lo);
  const double need = (h + 2.4e-7 * (fabs(cd) + h)) * (1.0 + 2.0e-7) / (double)step;  // decoded h = half * step, one rounding
  cq[k] = (int16_t)ci;
  const float nf = need < 6.0e4 ? __double2float_ru(need) : 6.0e4f;
  __half hh = __float2half_ru(nf);
  if(!(need < 6.0e4)) hh = __ushort_as_half((unsigned short)0x7c00);  // +inf: never prunes (cannot happen for |c|, h within the node)
  hq[k] = __half_as_ushort(hh);
}

// Reductions over the threads that share one entity: a warp (small subtrees) or a whole block (big ones)
struct WarpGroup
{
  __device__ __forceinline__ int rank() const { return (int)lane_id(); }
  __device__ __forceinline__ int size() const { return 32; }
  __device__ __forceinline__ double sum(double v) const { return warp_sum(v); }
  __device__ __forceinline__ double min(double v) const { return warp_min(v); }
  __device__ __forceinline__ double max(double v) const { return warp_max(v); }
};
struct SerialGroup  // one thread does the whole entity (subtrees of a few leaves: three quarters of all entities)
{
  __device__ __forceinline__ int rank() const { return 0; }
  __device__ __forceinline__ int size() const { return 1; }
  __device__ __forceinline__ double sum(double v) const { return v; }
  __device__ __forceinline__ double min(double v) const { return v; }
  __device__ __forceinline__ double max(double v) const { return v; }
};
struct BlockGroup
{
  double* sh;  // 32 doubles of shared memory
  __device__ __forceinline__ int rank() const { return (int)threadIdx.x; }
  __device__ __forceinline__ int size() const { return (int)blockDim.x; }
  template <typename F>
  __device__ __forceinline__ double reduce(double v, double identity, F f) const
  {
    v = f(v);
    __syncthreads();
    if(lane_id() == 0) sh[threadIdx.x >> 5] = v;
    __syncthreads();
    double w = (lane_id() < (blockDim.x >> 5)) ? sh[lane_id()] : identity;
    return f(w);
  }
  __device__ __forceinline__ double sum(double v) const { return reduce(v, 0.0, [](double x) { return warp_sum(x); }); }
  __device__ __forceinline__ double min(double v) const { return reduce(v, DBL_MAX, [](double x) { return warp_min(x); }); }
  __device__ __forceinline__ double max(double v) const { return reduce(v, -DBL_MAX, [](double x) { return warp_max(x); }); }
};

template <int NV>
__device__ __forceinline__ V3 leaf_area_normal(const double* __restrict__ soup, int p)
{
  V3 v[NV];
  load_leaf<NV>(soup, p, v);
  V3 c = v3cross(v3sub(v[1], v[0]), v3sub(v[2], v[0]));
  if(NV == 4 && has_fourth(v[NV - 1])) c = v3add(c, v3cross(v3sub(v[2], v[0]), v3sub(v[NV - 1], v[0])));
  return c;
}

// Area-weighted normal sums of every subtree in ONE bottom-up sweep (they are additive): one thread per leaf climbs the
// parent links; at every node the first child to arrive parks its sum and retires, the second adds LEFT + RIGHT (a fixed
// order: the result does not depend on who arrives first) and carries on.  child_sum[2 * node + side] ends up holding the
// sum of the subtree hanging on that side -- what obb_of_range needs for the entity stored in that slot.
template <int NV>
__global__ void __launch_bounds__(256) normal_sums_kernel(const double* __restrict__ soup, const Node<double, 3>* __restrict__ nodes,
                                                           const int32_t* __restrict__ leaf_parent, int nleaves, double* __restrict__ child_sum,
                                                           unsigned int* __restrict__ arrived)
{
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if(p >= nleaves) return;
  V3 c = leaf_area_normal<NV>(soup, p);
  int link = leaf_parent[p];
  while(link >= 0)
  {
    const int node = link >> 1, side = link & 1;
    double* mine = child_sum + 3 * (size_t)link;
    mine[0] = c.x;
    mine[1] = c.y;
    mine[2] = c.z;
    __threadfence();
    if(atomicAdd(arrived + node, 1u) == 0u) return;
    __threadfence();
    const volatile double* other = child_sum + 3 * (size_t)(link ^ 1);
    const V3 o {other[0], other[1], other[2]};
    c = side == 0 ? v3add(c, o) : v3add(o, c);
    link = nodes[node].parent;
  }
}

// oriented bound of the leaves [first, last] written into slot `link` of the parent's record
template <int NV, typename Group>
__device__ __forceinline__ void obb_of_range(const Group& g, const double* __restrict__ soup, int first, int last, const double* org,
                                             float* __restrict__ out, const double* __restrict__ nsum = nullptr)
{
  const int t = g.rank(), nt = g.size();
  const double omag = fabs(org[0]) + fabs(org[1]) + fabs(org[2]);
  // pass 1: area-weighted normal sum -- additive over the tree, so normally it arrives precomputed (normal_sums_kernel)
  double nx = 0.0, ny = 0.0, nz = 0.0;
  if(nsum)
  {
    nx = nsum[0];
    ny = nsum[1];
    nz = nsum[2];
  }
  else
  {
    for(int p = first + t; p <= last; p += nt)
    {
      const V3 c = leaf_area_normal<NV>(soup, p);
      nx += c.x;
      ny += c.y;
      nz += c.z;
    }
    nx = g.sum(nx);
    ny = g.sum(ny);
    nz = g.sum(nz);
  }
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
  // the stored frame: the normal on the 16-bit grid, the tangents derived from it in binary32 (sd_frame): everything
  // below uses exactly what the queries will see, in either record format
  int code[3];
  sd_frame_code(n, code);
  float fr[6];
  sd_frame(code[0], code[1], code[2], fr);
  V3 A[3];
  obb_axes(fr, A);
  // pass 2: extents of all vertices along the frame, relative to the parent's origin
  double lo[3] = {DBL_MAX, DBL_MAX, DBL_MAX}, hi[3] = {-DBL_MAX, -DBL_MAX, -DBL_MAX};
  for(int p = first + t; p <= last; p += nt)
  {
    V3 v[NV];
    load_leaf<NV>(soup, p, v);
#pragma unroll
    for(int j = 0; j < NV; ++j)
    {
      if(j == 3 && !has_fourth(v[j])) continue;  // triangle cell of a mixed mesh
      const double rx = v[j].x - org[0], ry = v[j].y - org[1], rz = v[j].z - org[2];
#pragma unroll
      for(int k = 0; k < 3; ++k)
      {
        const double d = A[k].x * rx + A[k].y * ry + A[k].z * rz;
        lo[k] = fmin(lo[k], d);
        hi[k] = fmax(hi[k], d);
      }
    }
  }
#pragma unroll
  for(int k = 0; k < 3; ++k)
  {
    lo[k] = g.min(lo[k]);
    hi[k] = g.max(hi[k]);
  }
  if(t == 0)
  {
#pragma unroll
    for(int k = 0; k < 6; ++k) out[k] = fr[k];
#pragma unroll
    for(int k = 0; k < 3; ++k)
    {
      // pad by the rounding of the projections (a few ulp of the coordinate magnitude), then round outward
      const double pad = 1e-14 * (fabs(lo[k]) + fabs(hi[k]) + omag) + 1e-300;
      store_extent(out, k, lo[k] - pad, hi[k] + pad);
    }
  }
}

constexpr int kObbWarpRange = 4096;  // subtrees up to this many leaves are bounded by one warp, larger ones by a block

constexpr int kObbSerialRange = 8;  // subtrees up to this many leaves are bounded by ONE THREAD

// One THREAD per tree entity e: e < inner -> inner node e (leaves node_range[e]), else leaf e - inner.  The thread writes the
// node's own header (child ids, origin), its centroid in the parent's SdCen, and -- for invalid boxes, huge subtrees (AABB
// kept) and subtrees of at most kObbSerialRange leaves -- the oriented bound itself.  Larger subtrees are queued: up to
// kObbWarpRange leaves for a warp each (obb_build_mid_kernel), beyond that for a block each (obb_build_big_kernel).
// (Round 1 gave every entity a warp: 4 M warps for 2 M triangles, three quarters of them for one or two leaves.)
template <int NV>
__global__ void __launch_bounds__(256) obb_build_kernel(const double* __restrict__ soup, const Node<double, 3>* __restrict__ nodes,
                                                         const int32_t* __restrict__ leaf_parent, const int2* __restrict__ node_range,
                                                         int nleaves, SdNode* __restrict__ sdn, SdCen* __restrict__ sdc, int obb_max_range,
                                                         int32_t* __restrict__ mid_list, unsigned int* __restrict__ mid_count,
                                                         int32_t* __restrict__ big_list, unsigned int* __restrict__ big_count,
                                                         const double* __restrict__ child_sum)
{
  const int inner = nleaves - 1;
  const long long gid = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if(gid >= (long long)inner + nleaves) return;
  const int e = (int)gid;
  int first, last, link;
  if(e < inner)
  {
    const int2 r = node_range[e];
    first = r.x;
    last = r.y;
    link = nodes[e].parent;
    sdn[e].child[0] = nodes[e].child[0];
    sdn[e].child[1] = nodes[e].child[1];
    double o[3];
    node_origin(nodes[e], o);
#pragma unroll
    for(int d = 0; d < 3; ++d) sdn[e].org[d] = o[d];
  }
  else
  {
    first = last = e - inner;
    link = leaf_parent[first];
  }
  if(link < 0) return;  // the root is nobody's child
  float* out = sdn[link >> 1].cb[link & 1];
  double org[3];
  node_origin(nodes[link >> 1], org);
  const Box<double, 3> bb = nodes[link >> 1].box[link & 1];  // this entity's AABB as the reference has it
  const bool valid = box_valid(bb);
#pragma unroll
  for(int d = 0; d < 3; ++d) sdc[link >> 1].cen[link & 1][d] = 0.5 * (bb.lo[d] + bb.hi[d]);
  const int count = last - first + 1;
  if(!valid || count > obb_max_range)
  {
    // coordinate axes: the bound is the AABB itself (an invalid box is infinitely far).  The frame is the one sd_frame
    // derives from the code (32767, 0, 0): n = e_x, t1 = e_z, t2 = -e_y.
    const double omag = fabs(org[0]) + fabs(org[1]) + fabs(org[2]);
    float fr[6];
    sd_frame(32767, 0, 0, fr);
    V3 A[3];
    obb_axes(fr, A);
#pragma unroll
    for(int k = 0; k < 3; ++k)
    {
      out[k] = fr[k];
      out[3 + k] = fr[3 + k];
      if(valid)
      {
        const double a[3] = {A[k].x, A[k].y, A[k].z};
        double lo = 0.0, hi = 0.0;  // extent of the box's corners along axis k of the frame
#pragma unroll
        for(int d = 0; d < 3; ++d)
        {
          const double u = a[d] * (bb.lo[d] - org[d]), v = a[d] * (bb.hi[d] - org[d]);
          lo += fmin(u, v);
          hi += fmax(u, v);
        }
        const double pad = 1e-14 * (fabs(lo) + fabs(hi) + omag) + 1e-300;
        store_extent(out, k, lo - pad, hi + pad);
      }
      else
      {
        out[6 + k] = 0.f;
        out[9 + k] = __int_as_float(0xff800000);  // half extent -inf: |d - c| - h = +inf, infinitely far
      }
    }
    return;
  }
  if(count > kObbWarpRange)
    big_list[atomicAdd(big_count, 1u)] = e;  // left to obb_build_big_kernel
  else if(count > kObbSerialRange)
    mid_list[atomicAdd(mid_count, 1u)] = e;  // left to obb_build_mid_kernel
  else
    obb_of_range<NV>(SerialGroup {}, soup, first, last, org, out, child_sum + 3 * (size_t)link);
}

// the subtrees of kObbSerialRange < leaves <= kObbWarpRange queued by obb_build_kernel: one warp per entity
template <int NV>
__global__ void __launch_bounds__(256) obb_build_mid_kernel(const double* __restrict__ soup, const Node<double, 3>* __restrict__ nodes,
                                                             const int2* __restrict__ node_range, SdNode* __restrict__ sdn,
                                                             const int32_t* __restrict__ mid_list, const unsigned int* __restrict__ mid_count,
                                                             const double* __restrict__ child_sum)
{
  const unsigned n = *mid_count;
  const unsigned warps = (gridDim.x * blockDim.x) >> 5;
  for(unsigned i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; i < n; i += warps)
  {
    const int e = mid_list[i];
    const int2 r = node_range[e];
    const int link = nodes[e].parent;
    double org[3];
    node_origin(nodes[link >> 1], org);
    obb_of_range<NV>(WarpGroup {}, soup, r.x, r.y, org, sdn[link >> 1].cb[link & 1], child_sum + 3 * (size_t)link);
  }
}

// the big subtrees queued by obb_build_kernel: one block per entity
template <int NV>
__global__ void __launch_bounds__(512) obb_build_big_kernel(const double* __restrict__ soup, const Node<double, 3>* __restrict__ nodes,
                                                             const int2* __restrict__ node_range, SdNode* __restrict__ sdn,
                                                             const int32_t* __restrict__ big_list, const unsigned int* __restrict__ big_count,
                                                             const double* __restrict__ child_sum)
{
  __shared__ double sh[32];
  const unsigned n = *big_count;
  for(unsigned i = blockIdx.x; i < n; i += gridDim.x)
  {
    const int e = big_list[i];
    const int2 r = node_range[e];
    const int link = nodes[e].parent;
    double org[3];
    node_origin(nodes[link >> 1], org);
    obb_of_range<NV>(BlockGroup {sh}, soup, r.x, r.y, org, sdn[link >> 1].cb[link & 1], child_sum + 3 * (size_t)link);
  }
}

// SdNode -> SdNode64, one thread per inner node (a stream: 128 B in, 64 B out).  The frame code is recovered exactly from
// the stored first axis (n = normalise(code): code_k = rint(32767 n_k / max|n|)); the extent [c - h, c + h] of the full
// record (already padded outward) is re-expressed on the compact record's grid, rounded outward again.
__global__ void __launch_bounds__(256) sd64_pack_kernel(const SdNode* __restrict__ sdn, const Node<double, 3>* __restrict__ nodes, int inner,
                                                         SdNode64* __restrict__ out)
{
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if(e >= inner) return;
  const SdNode a = sdn[e];
  SdNode64 r;
  double org[3];
  float step;
  node_origin(nodes[e], org, &step);
  r.child[0] = a.child[0];
  r.child[1] = a.child[1];
#pragma unroll
  for(int d = 0; d < 3; ++d) r.org[d] = (float)a.org[d];  // exact: the origin is a binary32 value
  r.step = step;
  r.pad_ = 0u;
#pragma unroll
  for(int s = 0; s < 2; ++s)
  {
    const float* f = a.cb[s];
    const float m = fmaxf(fabsf(f[0]), fmaxf(fabsf(f[1]), fabsf(f[2])));
#pragma unroll
    for(int k = 0; k < 3; ++k)
    {
      r.nq[s][k] = (int16_t)(int)rintf(32767.0f * (f[k] / m));
      if(f[9 + k] >= 0.f)
        store_extent64(r.cq[s], r.hq[s], k, step, (double)f[6 + k] - (double)f[9 + k], (double)f[6 + k] + (double)f[9 + k]);
      else
      {
        r.cq[s][k] = 0;
        r.hq[s][k] = (uint16_t)0xfc00;  // binary16 -inf: an invalid child is infinitely far
      }
    }
  }
  out[e] = r;
}

// squared distance lower bound from the query to the oriented box of a child: f = the 12 floats of the
// child's bound, r = q - org (FMA is fine here: it only has to be a bound)
__device__ __forceinline__ double obb_sqdist(const float* f, const double* r)
{
  V3 A[3];
  obb_axes(f, A);
  double s = 0.0;
#pragma unroll
  for(int k = 0; k < 3; ++k)
  {
    const double d = fma(A[k].x, r[0], fma(A[k].y, r[1], A[k].z * r[2]));
    const double t = fabs(d - (double)f[6 + k]) - (double)f[9 + k];
    const double g = t > 0.0 ? t : 0.0;
    s = fma(g, g, s);
  }
  return s * kBoundScale;
}

// prune threshold on squared distance: everything that could still improve the minimum or tie with it
// inside the reference's 1e-6 closest-point window (sqrt(EPS), EPS = 1e-12), plus rounding head-room.
__device__ __forceinline__ double prune_threshold(double minSq)
{
  if(minSq >= 1e300) return DBL_MAX;
  const double d = sqrt(minSq) + 1.0000001e-6;
  return d * d * (1.0 + 1e-12) + 1e-300;
}

//------------------------------------------------------------------------------------------
// Lazy pseudo-normal: checkCandidate (:636-737) adds a normal contribution whenever a candidate on a
// shared feature becomes (or ties with) the minimum; every later new minimum usually clears the sum
// again.  The contributions cost a sqrt, a division and (vertices) an acos each, so the traversal only
// RECORDS who contributed since the last clear, in order, and the sum is formed once at the end --
// same terms, same order, same rounding as the reference's running sum.
//------------------------------------------------------------------------------------------
constexpr int kContribCap = 8;

struct Contribs
{
  int32_t pos[kContribCap];
  int8_t code[kContribCap];  // (sub << 3) | (loc + 3)
  int n;
};

template <int NV>
__device__ __noinline__ V3 contrib_flush(const double* __restrict__ soup, Contribs& cl, V3 sumN)
{
  for(int i = 0; i < cl.n; ++i)
  {
    V3 v[NV];
    load_leaf<NV>(soup, cl.pos[i], v);
    const int sub = cl.code[i] >> 3, loc = (cl.code[i] & 7) - 3;
    const V3 T[3] = {v[0], sub == 0 ? v[1] : v[2], sub == 0 ? v[2] : v[NV - 1]};
    const V3 n = v3cross(v3sub(T[1], T[0]), v3sub(T[2], T[0]));  // Triangle::normal :98-102
    if(loc < 0)
    {
      sumN = v3add(sumN, v3unit(n));  // edge: += normal().unitVector() (:703)
    }
    else
    {
      const double area = 0.5 * sqrt(v3dot(n, n));  // Triangle::area :105-109
      if(!nearly_eq(area, 0.0, 1.0e-12))            // !degenerate() :326-330
      {
        const double alpha = tri_angle(T, loc);
        sumN = v3add(sumN, v3mul(v3unit(n), alpha));  // vertex: += angle(loc)*normal().unitVector() (:722-728)
      }
    }
  }
  cl.n = 0;
  return sumN;
}

// checkCandidate (:636-737) for one (sub-)triangle with the normal contribution deferred
template <int NV>
__device__ __forceinline__ void check_triangle_lazy(const double* __restrict__ soup, const V3& q, MinCand& m, Contribs& cl, const V3* T,
                                                     int pos, int sub, bool computeNormal)
{
  constexpr double EPS = 1e-12;
  int loc;
  const V3 cp = closest_point_tri(q, T[0], T[1], T[2], loc, EPS);
  const V3 dq = v3sub(cp, q);
  const double sq = v3dot(dq, dq);
  const int type = loc_type(loc);
  const bool shared = (type != 2);
  const V3 dm = v3sub(m.minPt, cp);
  const bool same_spot = (m.minType == type) && nearly_eq(v3dot(dm, dm), 0., EPS);
  bool upd;
  if(sq < m.minSq)
  {
    const bool clear = !shared || !same_spot;
    m.minSq = sq;
    m.minPt = cp;
    m.minType = type;
    m.minPos = pos;
    m.minSub = sub;
    if(computeNormal && clear)
    {
      m.sumN = {0.0, 0.0, 0.0};
      cl.n = 0;
    }
    upd = computeNormal && shared;
  }
  else
  {
    upd = computeNormal && shared && same_spot;
  }
  if(upd)
  {
    if(cl.n == kContribCap) m.sumN = contrib_flush<NV>(soup, cl, m.sumN);
    cl.pos[cl.n] = pos;
    cl.code[cl.n] = (int8_t)((sub << 3) | (loc + 3));
    ++cl.n;
  }
}

template <int NV>
__device__ __forceinline__ void check_leaf_lazy(const double* __restrict__ soup, const V3& q, MinCand& m, Contribs& cl, int pos,
                                                 bool computeNormal)
{
  V3 v[NV];
  load_leaf<NV>(soup, pos, v);
  {
    const V3 T[3] = {v[0], v[1], v[2]};
    check_triangle_lazy<NV>(soup, q, m, cl, T, pos, 0, computeNormal);
  }
  if(NV == 4 && has_fourth(v[NV - 1]))
  {
    const V3 T[3] = {v[0], v[2], v[NV - 1]};  // quads split (0,1,2),(0,2,3) :652-658
    check_triangle_lazy<NV>(soup, q, m, cl, T, pos, 1, computeNormal);
  }
}

//------------------------------------------------------------------------------------------
// MODE 1 kernel.  One query per LANE, but the warp is scheduled as a unit:
//   * persistent warps pull Morton-ordered queries from a global cursor and a lane that finishes its
//     query is refilled at once, so no lane idles behind the slowest query of its warp;
//   * a lane that reaches a leaf does not evaluate it on the spot: it queues it (FIFO, so the leaf ORDER
//     of the lane is unchanged) and keeps walking inner nodes; the warp switches to "leaf steps" when
//     enough lanes have a leaf waiting, so closest_point() runs with most lanes active instead of 2-3;
//   * finished lanes are finalised (pseudo-normal, sign, stores) a few at a time for the same reason.
// Each lane still walks the tree in the reference's child order with a (node, lower bound) stack, so
// everything said in the header about bit-identical results holds.
//------------------------------------------------------------------------------------------
#ifndef AXB_SD_MIN_BLOCKS
  #define AXB_SD_MIN_BLOCKS 4  // resident blocks per SM the register allocation is sized for
#endif
#ifndef AXB_SD_SMEM_STACK
  #define AXB_SD_SMEM_STACK 32
#endif
constexpr int kSmemStack = AXB_SD_SMEM_STACK;   // stack levels kept in shared memory (the rest, rarely reached, in local memory)
constexpr size_t kSdFastSmem = (size_t)kSmemStack * 128 * sizeof(unsigned long long);  // 128 threads per block
#ifndef AXB_SD_PEND
  #define AXB_SD_PEND 4
#endif
#ifndef AXB_SD_LEAF_VOTE
  #define AXB_SD_LEAF_VOTE 16
#endif
#ifndef AXB_SD_FINISH_VOTE
  #define AXB_SD_FINISH_VOTE 4
#endif
constexpr int kPend = AXB_SD_PEND;                // queued leaves per lane
constexpr int kLeafVote = AXB_SD_LEAF_VOTE;       // lanes with a queued leaf that trigger a leaf step
constexpr int kFinishVote = AXB_SD_FINISH_VOTE;   // finished lanes that trigger a finalisation step
constexpr int kQueryChunk = 128;  // queries a warp takes from the cursor at a time (a run of Morton neighbours)

template <int NV>
__global__ void __launch_bounds__(128, AXB_SD_MIN_BLOCKS) sd_fast_kernel(const SdNode* __restrict__ nodes, const SdCen* __restrict__ cens, const double* __restrict__ soup, SdParams prm,
                                                       Desc<3> qpts, int npts, const int32_t* __restrict__ perm, double* __restrict__ phi,
                                                       double* __restrict__ cps, double* __restrict__ nrms,
                                                       unsigned long long* __restrict__ work, unsigned int* __restrict__ cursor,
                                                       unsigned chunk)
{
  constexpr unsigned FULL = 0xffffffffu;
  const unsigned lane = lane_id();
  const unsigned lt_mask = (1u << lane) - 1u;
  const bool cn = prm.compute_sign != 0;

  // ---- per-lane query state ----
  int qi = -1;  // original index of the lane's query, -1 = lane is free
  double qp[3] = {0.0, 0.0, 0.0};
  MinCand m;
  Contribs cl;
  cl.n = 0;
  double thr = DBL_MAX;
  // A point of the SURFACE near the lane's next query: the closest point of the query the lane just
  // finished (consecutive queries of a lane are Morton neighbours).  Its distance to the new query is an
  // upper bound on the new minimum, so the traversal starts with a finite prune radius instead of
  // walking blind until its first leaf has been evaluated.
  V3 hint = {0.0, 0.0, 0.0};
  bool have_hint = false;
  // Traversal stack, entries (lower bound as float bits) << 32 | node id.  The first kSmemStack levels live
  // in shared memory as [level][thread]: whatever level each lane is at, lane t touches bank pair t, so a
  // push or pop is 2 wavefronts for the warp -- in local memory lanes at different depths hit different
  // lines (up to 64 wavefronts, and the spills went to DRAM: 2.9 GB of writes in profiles/r1c).
  extern __shared__ unsigned long long sstack[];
  unsigned long long st_over[kStackSize - kSmemStack];
  unsigned long long* const my_stack = sstack + threadIdx.x;
  const unsigned stride = blockDim.x;
  int sp = 0;
  auto st_get = [&](int k) -> unsigned long long { return k < kSmemStack ? my_stack[(unsigned)k * stride] : st_over[k - kSmemStack]; };
  auto st_put = [&](int k, unsigned long long e) {
    if(k < kSmemStack)
      my_stack[(unsigned)k * stride] = e;
    else
      st_over[k - kSmemStack] = e;
  };
  int32_t cur = kBarrier;  // node in hand: >= 0 inner, < 0 leaf, kBarrier = traversal finished
  float cur_lb = 0.f;
  int32_t pend_id[kPend];
  float pend_lb[kPend];
  int npend = 0;
  unsigned nleaf = 0, ninner = 0;

  // ---- the warp's share of the query stream ----
  unsigned wbase = 0, wcount = 0;
  bool exhausted = false;

  auto pop = [&]() {
    cur = kBarrier;
    while(sp > 0)
    {
      --sp;
      const unsigned long long e = st_get(sp);
      const float lb = __uint_as_float((unsigned)(e >> 32));
      if((double)lb <= thr)
      {
        cur = (int32_t)(unsigned)(e & 0xffffffffull);
        cur_lb = lb;
        break;
      }
    }
  };

  while(true)
  {
    // ---- refill free lanes ----
    const unsigned freem = __ballot_sync(FULL, qi < 0);
    if(freem != 0u && !exhausted)
    {
      if(wcount == 0u)
      {
        // guided self-scheduling: long runs of Morton neighbours while there is plenty of work, short
        // ones near the end so the last warps finish together
        unsigned b = 0, g = 0;
        if(lane == 0)
        {
          const unsigned seen = *reinterpret_cast<volatile unsigned*>(cursor);
          const unsigned rem = seen < (unsigned)npts ? (unsigned)npts - seen : 0u;
          const unsigned nwarps = gridDim.x * (blockDim.x >> 5);
          g = max(32u, min(chunk, (rem / (2u * nwarps)) & ~31u));
          b = atomicAdd(cursor, g);
        }
        wbase = __shfl_sync(FULL, b, 0);
        g = __shfl_sync(FULL, g, 0);
        wcount = wbase < (unsigned)npts ? min(g, (unsigned)npts - wbase) : 0u;
        exhausted = (wcount == 0u);
      }
      if(wcount != 0u)
      {
        const unsigned rank = __popc(freem & lt_mask);
        if(qi < 0 && rank < wcount)
        {
          const unsigned t = wbase + rank;
          qi = perm ? perm[t] : (int)t;
          qp[0] = ld_comp<double>(qpts, 0, qi);
          qp[1] = ld_comp<double>(qpts, 1, qi);
          qp[2] = ld_comp<double>(qpts, 2, qi);
          m.minSq = DBL_MAX;
          m.minPt = {0.0, 0.0, 0.0};
          m.sumN = {0.0, 0.0, 0.0};
          m.minType = -1;
          m.minPos = 0;
          m.minSub = 0;
          cl.n = 0;
          thr = DBL_MAX;
          if(have_hint)
          {
            const double hx = hint.x - qp[0], hy = hint.y - qp[1], hz = hint.z - qp[2];
            thr = prune_threshold(hx * hx + hy * hy + hz * hz);
          }
          sp = 0;
          cur = 0;  // root
          cur_lb = 0.f;
          npend = 0;
        }
        const unsigned taken = min((unsigned)__popc(freem), wcount);
        wbase += taken;
        wcount -= taken;
      }
    }
    const bool busy = qi >= 0;
    const unsigned busym = __ballot_sync(FULL, busy);
    if(busym == 0u)
    {
      if(exhausted) break;
      continue;
    }
    const bool finished = busy && cur == kBarrier && npend == 0;
    const bool want_inner = busy && cur != kBarrier && !(cur < 0 && npend == kPend);
    const bool want_leaf = busy && npend > 0;
    const unsigned mfin = __ballot_sync(FULL, finished);
    const unsigned minner = __ballot_sync(FULL, want_inner);
    const unsigned mleaf = __ballot_sync(FULL, want_leaf);
    const unsigned mfull = __ballot_sync(FULL, busy && npend == kPend);

    if(__popc(mfin) >= kFinishVote || (mfin != 0u && minner == 0u && mleaf == 0u))
    {
      // ---- finalisation step: pseudo-normal, sign, distance, outputs ----
      if(finished)
      {
        if(cl.n) m.sumN = contrib_flush<NV>(soup, cl, m.sumN);
        const V3 q {qp[0], qp[1], qp[2]};
        sd_finish<NV>(soup, prm, q, m, qi, phi, cps, nrms);
        if(m.minType >= 0)
        {
          hint = m.minPt;
          have_hint = true;
        }
        qi = -1;
      }
      continue;
    }
    if(__popc(mleaf) >= kLeafVote || mfull != 0u || minner == 0u)
    {
      // ---- leaf step: every lane with a queued leaf evaluates its oldest one ----
      if(want_leaf)
      {
        const int32_t id = pend_id[0];
        const float lb = pend_lb[0];
#pragma unroll
        for(int k = 0; k + 1 < kPend; ++k)
        {
          pend_id[k] = pend_id[k + 1];
          pend_lb[k] = pend_lb[k + 1];
        }
        --npend;
        if((double)lb <= thr)
        {
          ++nleaf;
          const V3 q {qp[0], qp[1], qp[2]};
          check_leaf_lazy<NV>(soup, q, m, cl, -id - 1, cn);
          thr = fmin(thr, prune_threshold(m.minSq));
        }
      }
      continue;
    }
    // ---- inner step ----
    if(want_inner)
    {
      bool need_pop = false;
      if(cur >= 0)
      {
        ++ninner;
        const D4* rec = reinterpret_cast<const D4*>(nodes + cur);
        const D4 r0 = ldg256(rec), r1 = ldg256(rec + 1), r2 = ldg256(rec + 2), r3 = ldg256(rec + 3);
        const long long ids = __double_as_longlong(r0.x);
        const int32_t child0 = (int32_t)(ids & 0xffffffffll), child1 = (int32_t)(ids >> 32);
        const double rq[3] = {qp[0] - r0.y, qp[1] - r0.z, qp[2] - r0.w};
        float f[24];
        {
          const double w[12] = {r1.x, r1.y, r1.z, r1.w, r2.x, r2.y, r2.z, r2.w, r3.x, r3.y, r3.z, r3.w};
#pragma unroll
          for(int k = 0; k < 12; ++k)
          {
            f[2 * k] = __int_as_float(__double2loint(w[k]));
            f[2 * k + 1] = __int_as_float(__double2hiint(w[k]));
          }
        }
        // an invalid box is infinitely far (bvh_traverse.hpp:95-96)
        const double d20 = obb_sqdist(f, rq), d21 = obb_sqdist(f + 12, rq);
        const bool in0 = d20 <= thr, in1 = d21 <= thr;
        // Child order = the reference's (LinearBVH.hpp:72-85): when both children are entered, the one whose
        // AABB centroid is nearer goes first and the other waits on the stack -- leaf or not -- until
        // everything below the first is done.  The leaves a lane evaluates are then a SUBSEQUENCE of the
        // reference's own visiting order, so strict-< tie-breaks between equidistant triangles and the
        // summation order of the pseudo-normal are the reference's.
        if(in0 && in1)
        {
          const D4* cr = reinterpret_cast<const D4*>(cens + cur);
          const D4 c0 = ldg256(cr), c1 = ldg256(cr + 1);
          const double cl3[3] = {c0.x, c0.y, c0.z}, cr3[3] = {c0.w, c1.x, c1.y};
          double dl = 0.0, dr = 0.0;
#pragma unroll
          for(int d = 0; d < 3; ++d)
          {
            const double cl_ = cl3[d] - qp[d];
            dl += cl_ * cl_;
            const double cr_ = cr3[d] - qp[d];
            dr += cr_ * cr_;
          }
          const bool right_first = dl > dr;
          const int32_t c_second = right_first ? child0 : child1;
          const double d_second = right_first ? d20 : d21;
          st_put(sp, ((unsigned long long)__float_as_uint(__double2float_rd(d_second)) << 32) | (unsigned)c_second);
          ++sp;
          cur = right_first ? child1 : child0;
          cur_lb = __double2float_rd(right_first ? d21 : d20);
        }
        else if(in0)
        {
          cur = child0;
          cur_lb = __double2float_rd(d20);
        }
        else if(in1)
        {
          cur = child1;
          cur_lb = __double2float_rd(d21);
        }
        else
        {
          need_pop = true;
        }
      }
      // One pop site for the whole step (lanes whose children were both pruned and lanes holding a leaf converge
      // here; with a pop() in each branch they ran it 3 lanes at a time, 11 % of the kernel's instructions).
      // A leaf in hand joins the queue and what follows it on the stack comes into hand, until the lane holds an
      // inner node, the queue is full or the traversal is finished.
      while(true)
      {
        if(need_pop)
        {
          pop();
          need_pop = false;
        }
        if(cur < 0 && cur != kBarrier && npend < kPend)
        {
#pragma unroll
          for(int k = 0; k < kPend; ++k)
            if(k == npend)
            {
              pend_id[k] = cur;
              pend_lb[k] = cur_lb;
            }
          ++npend;
          need_pop = true;
        }
        else
          break;
      }
    }
  }
  if(work)
  {
    unsigned long long a = nleaf, b = ninner;
#pragma unroll
    for(int o = 16; o > 0; o >>= 1)
    {
      a += __shfl_xor_sync(FULL, a, o);
      b += __shfl_xor_sync(FULL, b, o);
    }
    if(lane == 0)
    {
      atomicAdd(&work[0], a);
      atomicAdd(&work[1], b);
    }
  }
}

// empty bounds (ordered-uint64 encoding) for query_bounds_kernel
__global__ void init_query_bounds_kernel(unsigned long long* __restrict__ ob)
{
  if(threadIdx.x < 3)
  {
    ob[threadIdx.x] = f64_to_ordered(DBL_MAX);
    ob[3 + threadIdx.x] = f64_to_ordered(-DBL_MAX);
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
