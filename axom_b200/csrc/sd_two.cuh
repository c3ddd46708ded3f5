// sd_two.cuh -- the SignedDistance query in two phases: an order-free search for the exact minimum, then the
// reference's state machine replayed on the handful of leaves that can matter, in the reference's visiting order.
//
// Why two phases.  checkCandidate (quest/SignedDistance.hpp:636-737) is a state machine over the leaves in the
// order LinearBVHTraverser visits them (nearer AABB centroid first, spin/policy/LinearBVH.hpp:72-85), but only
// the leaves whose squared distance lies within the machine's own 1e-6 tie window of the final minimum d* can
// leave a trace in its final state (see sd_fast.cuh: a leaf outside the window is overwritten by a later strict
// minimum with a cleared normal sum).  sd_fast_kernel walks the tree in the reference's order with that state
// machine in registers (122 registers, 13.5 of 32 lanes active, 260 instructions per node visit: profiles/r2b).
// Here:
//   phase 1  finds d* with whatever is fastest: nearer-BOUND-first depth-first search, oriented bounds evaluated
//            in binary32 (conservatively, see obb_sqdist_f32), a (bound, node) stack in shared memory, leaves
//            evaluated in warp-wide batches with the exact reference arithmetic but WITHOUT the state machine
//            (closest_point + squared_distance only).  Every leaf evaluated within the window of the running
//            minimum is remembered (its sorted position); when a new minimum leaves the old one outside its
//            window the list restarts.  At the end the list holds every in-window leaf (and a few stale ones).
//   phase 2  puts the remembered leaves into the reference's visiting order -- by construction the order of two
//            leaves is decided at their lowest common ancestor by the centroid comparison, so a walk from that
//            ancestor that enters only subtrees containing remembered positions reproduces it -- and feeds
//            them to the unchanged state machine (check_leaf_lazy), then signs and stores (sd_finish).
// A query whose list overflows (kCandCap leaves within 1e-6 of each other: the centre of a sphere), or that phase 1 gives
// up after heavy_visits node visits, is only LISTED by phase 2 and finished by one whole warp of sd_solo_kernel (exact
// minimum over a shared stack, then the in-window leaves by an order-preserving level expansion, then the state machine).
// Launches of one call: sd_min_kernel on a SAMPLE (one query in 32 of the Morton order: its closest point is the first
// bound of its neighbours), [partitioned surface: the ranks MIN-reduce a per-query bound], sd_min_kernel proper,
// sd_resolve_kernel, sd_solo_kernel.  Phase 1 reads the 64-byte records (SdNode64, sd_fast.cuh); the ordered walks read
// the full ones in binary64.
#pragma once
#include "sd_fast.cuh"

namespace axb
{
// what phase 2 needs to climb the tree: parent and leaf range of every inner node
struct alignas(16) SdUp
{
  int32_t parent;  // inner node id, -1 for the root
  int32_t first, last;
  int32_t pad_;
};

__global__ void __launch_bounds__(256) sd_up_kernel(const Node<double, 3>* __restrict__ nodes, const int2* __restrict__ node_range, int inner,
                                                     SdUp* __restrict__ up)
{
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if(e >= inner) return;
  const int link = nodes[e].parent;
  const int2 r = node_range[e];
  SdUp u;
  u.parent = link < 0 ? -1 : (link >> 1);
  u.first = r.x;
  u.last = r.y;
  u.pad_ = 0;
  up[e] = u;
}

#ifndef AXB_SD2_MIN_BLOCKS
  #define AXB_SD2_MIN_BLOCKS 5
#endif
#ifndef AXB_SD2_PREFETCH
  #define AXB_SD2_PREFETCH 0  // prefetch.global.L1 of both child records at every visit: measured 46.6 -> 47.4 ms, off
#endif
#ifndef AXB_SD2_SMEM_STACK
  #define AXB_SD2_SMEM_STACK 16
#endif
#ifndef AXB_SD2_CAND_CAP
  #define AXB_SD2_CAND_CAP 15
#endif
#ifndef AXB_SD2_LEAF_VOTE
  #define AXB_SD2_LEAF_VOTE 16
#endif
#ifndef AXB_SD2_FINISH_VOTE
  #define AXB_SD2_FINISH_VOTE 8
#endif
#ifndef AXB_SD2_BATCH_MIN
  #define AXB_SD2_BATCH_MIN 32  // leaves waiting in the warp's pool that trigger a leaf batch
#endif
#ifndef AXB_SD2_FULL_VOTE
  #define AXB_SD2_FULL_VOTE 4
#endif
constexpr int kSd2Threads = 128;
constexpr int kSd2Stack = AXB_SD2_SMEM_STACK;  // stack levels in shared memory; deeper ones (rare) in local memory
constexpr int kCandCap = AXB_SD2_CAND_CAP;     // remembered leaves per query

// Conservative lower bound (squared) on the distance from the query to a child's oriented box, in binary32.
//   f   the child's 12 floats (n, t1, extent centre c, half extent h), rf = float(q - org), m = 5e-7 * |rf|_1
// Error budget of t = |d - c| - h against exact arithmetic on the stored floats: rounding of rf 2^-24 |r|_1, the
// three-term dot product with FMAs 3 * 2^-24 |r|_1, the two subtractions 2^-24 (2 |r|_1 + 2 |c| + h): at most
// 3.6e-7 |r|_1 (covered by m) + 2.4e-7 (|c| + h) (covered by the pad store_extent adds to h).  The third axis is
// the binary32 cross product of the stored two, bit-identical in the build (obb_axes); the sum of three squares
// loses 4 * 2^-24 = 2.4e-7 relative and the frame is orthonormal to ~2e-7 (|M v|^2 <= (1 + 5e-7) |v|^2): kBoundScaleF.
constexpr float kBoundScaleF = 1.0f - 2.0e-6f;
#ifndef AXB_SD2_NORMAL_F64
  #define AXB_SD2_NORMAL_F64 1
#endif
// the same with the first axis (the patch normal: the thin extent, where a margin of 5e-7 |r| is comparable to the
// extent itself for far queries) evaluated in double: no margin on that axis, the result rounded DOWN to binary32
__device__ __forceinline__ float obb_sqdist_mixed(const float* f, double qrx, double qry, double qrz, float rx, float ry, float rz, float m)
{
  const float t2x = __fsub_rn(__fmul_rn(f[1], f[5]), __fmul_rn(f[4], f[2]));
  const float t2y = __fsub_rn(__fmul_rn(f[3], f[2]), __fmul_rn(f[0], f[5]));
  const float t2z = __fsub_rn(__fmul_rn(f[0], f[4]), __fmul_rn(f[3], f[1]));
  const double d0 = fma((double)f[0], qrx, fma((double)f[1], qry, (double)f[2] * qrz));
  const float d1 = __fmaf_rn(f[3], rx, __fmaf_rn(f[4], ry, __fmul_rn(f[5], rz)));
  const float d2 = __fmaf_rn(t2x, rx, __fmaf_rn(t2y, ry, __fmul_rn(t2z, rz)));
  const double t0 = fabs(d0 - (double)f[6]) - (double)f[9];
  const float g0 = t0 > 0.0 ? __double2float_rd(t0 * (1.0 - 1e-12)) : 0.f;
  const float g1 = fmaxf(__fsub_rn(__fsub_rn(fabsf(__fsub_rn(d1, f[7])), f[10]), m), 0.f);
  const float g2 = fmaxf(__fsub_rn(__fsub_rn(fabsf(__fsub_rn(d2, f[8])), f[11]), m), 0.f);
  return __fmul_rn(__fmaf_rn(g0, g0, __fmaf_rn(g1, g1, __fmul_rn(g2, g2))), kBoundScaleF);
}
__device__ __forceinline__ float obb_sqdist_f32(const float* f, float rx, float ry, float rz, float m)
{
  const float t2x = __fsub_rn(__fmul_rn(f[1], f[5]), __fmul_rn(f[4], f[2]));
  const float t2y = __fsub_rn(__fmul_rn(f[3], f[2]), __fmul_rn(f[0], f[5]));
  const float t2z = __fsub_rn(__fmul_rn(f[0], f[4]), __fmul_rn(f[3], f[1]));
  const float d0 = __fmaf_rn(f[0], rx, __fmaf_rn(f[1], ry, __fmul_rn(f[2], rz)));
  const float d1 = __fmaf_rn(f[3], rx, __fmaf_rn(f[4], ry, __fmul_rn(f[5], rz)));
  const float d2 = __fmaf_rn(t2x, rx, __fmaf_rn(t2y, ry, __fmul_rn(t2z, rz)));
  const float g0 = fmaxf(__fsub_rn(__fsub_rn(fabsf(__fsub_rn(d0, f[6])), f[9]), m), 0.f);
  const float g1 = fmaxf(__fsub_rn(__fsub_rn(fabsf(__fsub_rn(d1, f[7])), f[10]), m), 0.f);
  const float g2 = fmaxf(__fsub_rn(__fsub_rn(fabsf(__fsub_rn(d2, f[8])), f[11]), m), 0.f);
  return __fmul_rn(__fmaf_rn(g0, g0, __fmaf_rn(g1, g1, __fmul_rn(g2, g2))), kBoundScaleF);
}

// squared distance from q to leaf `pos` exactly as checkCandidate computes it (closest_point, then
// squared_distance(q, cp)); a quad is its two triangles (:652-658)
template <int NV>
__device__ __forceinline__ double leaf_min_sq(const double* __restrict__ soup, const V3& q, int pos)
{
  constexpr double EPS = 1e-12;
  V3 v[NV];
  load_leaf<NV>(soup, pos, v);
  int loc;
  const V3 cp = closest_point_tri(q, v[0], v[1], v[2], loc, EPS);
  const V3 dq = v3sub(cp, q);
  double sq = v3dot(dq, dq);
  if(NV == 4 && has_fourth(v[NV - 1]))
  {
    const V3 cp2 = closest_point_tri(q, v[0], v[2], v[NV - 1], loc, EPS);
    const V3 dq2 = v3sub(cp2, q);
    sq = fmin(sq, v3dot(dq2, dq2));
  }
  return sq;
}

// prune_threshold with the tie window as a parameter: sqrt(EPS) + head-room when the pseudo-normal state machine
// runs (computeSign), 0 when it does not -- then only the strict-< minimum and its exact ties can matter
__device__ __forceinline__ double prune_threshold_w(double minSq, double window)
{
  if(minSq >= 1e300) return DBL_MAX;
  const double d = sqrt(minSq) + window;
  return d * d * (1.0 + 1e-12) + 1e-300;
}
constexpr double kTieWindow = 1.0000001e-6;  // prune_threshold's

__device__ __forceinline__ void mincand_reset(MinCand& m)
{
  m.minSq = DBL_MAX;
  m.minPt = {0.0, 0.0, 0.0};
  m.sumN = {0.0, 0.0, 0.0};
  m.minType = -1;
  m.minPos = 0;
  m.minSub = 0;
}

// The reference-order walk of sd_fast_kernel for ONE query (the lane's own), seeded with an upper bound on the
// minimum: children entered nearer-centroid-first, double-precision oriented bounds, the state machine at every
// leaf.  Only queries whose remembered-leaf list overflowed come here, so it is kept out of line.
template <int NV>
__device__ __noinline__ void sd_ordered_query(const SdNode* __restrict__ nodes, const SdCen* __restrict__ cens, const double* __restrict__ soup,
                                              const V3& q, double seed_sq, bool cn, MinCand& m, unsigned& nleaf, unsigned& ninner)
{
  unsigned long long st[kStackSize];
  int sp = 0;
  Contribs cl;
  cl.n = 0;
  mincand_reset(m);
  const double window = cn ? kTieWindow : 0.0;
  double thr = prune_threshold_w(seed_sq, window);
  const double qp[3] = {q.x, q.y, q.z};
  int32_t cur = 0;
  float cur_lb = 0.f;
  while(true)
  {
    if(cur >= 0)
    {
      ++ninner;
      const D4* rec = reinterpret_cast<const D4*>(nodes + cur);
      const D4 r0 = ldg256(rec), r1 = ldg256(rec + 1), r2 = ldg256(rec + 2), r3 = ldg256(rec + 3);
      const long long ids = __double_as_longlong(r0.x);
      const int32_t child0 = (int32_t)(ids & 0xffffffffll), child1 = (int32_t)(ids >> 32);
      const double rq[3] = {qp[0] - r0.y, qp[1] - r0.z, qp[2] - r0.w};
      float f[24];
      const double w[12] = {r1.x, r1.y, r1.z, r1.w, r2.x, r2.y, r2.z, r2.w, r3.x, r3.y, r3.z, r3.w};
#pragma unroll
      for(int k = 0; k < 12; ++k)
      {
        f[2 * k] = __int_as_float(__double2loint(w[k]));
        f[2 * k + 1] = __int_as_float(__double2hiint(w[k]));
      }
      const double d20 = obb_sqdist(f, rq), d21 = obb_sqdist(f + 12, rq);
      const bool in0 = d20 <= thr, in1 = d21 <= thr;
      if(in0 && in1)
      {
        const D4* cr = reinterpret_cast<const D4*>(cens + cur);
        const D4 c0 = ldg256(cr), c1 = ldg256(cr + 1);
        const double cl3[3] = {c0.x, c0.y, c0.z}, cr3[3] = {c0.w, c1.x, c1.y};
        double dl = 0.0, dr = 0.0;
#pragma unroll
        for(int d = 0; d < 3; ++d)
        {
          const double a = cl3[d] - qp[d];
          dl += a * a;
          const double b = cr3[d] - qp[d];
          dr += b * b;
        }
        const bool right_first = dl > dr;
        if(sp < kStackSize)
          st[sp++] = ((unsigned long long)__float_as_uint(__double2float_rd(right_first ? d20 : d21)) << 32) |
                     (unsigned)(right_first ? child0 : child1);
        cur = right_first ? child1 : child0;
        cur_lb = __double2float_rd(right_first ? d21 : d20);
        continue;
      }
      if(in0 || in1)
      {
        cur = in0 ? child0 : child1;
        cur_lb = __double2float_rd(in0 ? d20 : d21);
        continue;
      }
    }
    else if((double)cur_lb <= thr)
    {
      ++nleaf;
      check_leaf_lazy<NV>(soup, q, m, cl, -cur - 1, cn);
      thr = fmin(thr, prune_threshold_w(m.minSq, window));
    }
    // next entry of the stack that can still matter
    bool got = false;
    while(sp > 0)
    {
      const unsigned long long e = st[--sp];
      const float lb = __uint_as_float((unsigned)(e >> 32));
      if((double)lb <= thr)
      {
        cur = (int32_t)(unsigned)(e & 0xffffffffull);
        cur_lb = lb;
        got = true;
        break;
      }
    }
    if(!got) break;
  }
  if(cl.n) m.sumN = contrib_flush<NV>(soup, cl, m.sumN);
}

// checkCandidate (:636-737) for one (sub-)triangle whose closest point, squared distance and location code are
// already known (evaluated by another lane of the warp); the normal contribution is deferred as in check_triangle_lazy
template <int NV>
__device__ __forceinline__ void apply_candidate(const double* __restrict__ soup, MinCand& m, Contribs& cl, const V3& cp, double sq, int loc, int pos,
                                                int sub, bool computeNormal)
{
  constexpr double EPS = 1e-12;
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

// Phase 2, ordering.  Input: k >= 2 entries e[i] = (sorted position << 8) | list index, in the lane's column of shared
// memory (stride es).  Output: the list indices in the reference's visiting order, 4 bits each, first visited in
// bits 0-3.  The order of two leaves is decided at their lowest common ancestor (right child first iff its AABB
// centroid is nearer, LinearBVH.hpp:75-82).  With the entries sorted by position, the ancestors that matter are the
// k-1 lowest common ancestors of NEIGHBOURING entries (found by climbing from the left one until the range reaches
// the right one: short unless the pair straddles a big subtree boundary), and the ancestor of any pair (i, j) is the
// one with the largest range among the gaps between them.  Rank of an entry = how many others come before it.
__device__ __forceinline__ unsigned long long sd_order(const SdCen* __restrict__ cens, const SdUp* __restrict__ up,
                                                       const int32_t* __restrict__ leaf_parent, double qx, double qy, double qz,
                                                       unsigned long long* e, unsigned es, int k)
{
  for(int i = 1; i < k; ++i)  // insertion sort by position
  {
    const unsigned long long v = e[(unsigned)i * es];
    int j = i - 1;
    while(j >= 0 && e[(unsigned)j * es] > v)
    {
      e[(unsigned)(j + 1) * es] = e[(unsigned)j * es];
      --j;
    }
    e[(unsigned)(j + 1) * es] = v;
  }
  // gap g (between entries g and g+1): range length of their lowest common ancestor and who goes first there,
  // kept at level kCandCap + g as (length << 1) | right_first
  for(int g = 0; g + 1 < k; ++g)
  {
    const int32_t pa = (int32_t)(e[(unsigned)g * es] >> 8), pb = (int32_t)(e[(unsigned)(g + 1) * es] >> 8);
    int32_t nd = __ldg(leaf_parent + pa) >> 1;
    int4 u;
    while(true)
    {
      u = __ldg(reinterpret_cast<const int4*>(up + nd));
      if(u.z >= pb || u.x < 0) break;
      nd = u.x;
    }
    const D4* cr = reinterpret_cast<const D4*>(cens + nd);
    const D4 c0 = ldg256(cr), c1 = ldg256(cr + 1);
    const double cl3[3] = {c0.x, c0.y, c0.z}, cr3[3] = {c0.w, c1.x, c1.y};
    const double qp[3] = {qx, qy, qz};
    double dl = 0.0, dr = 0.0;
#pragma unroll
    for(int d = 0; d < 3; ++d)
    {
      const double a = cl3[d] - qp[d];
      dl += a * a;
      const double b = cr3[d] - qp[d];
      dr += b * b;
    }
    e[(unsigned)(kCandCap + g) * es] = ((unsigned long long)(unsigned)(u.z - u.y) << 1) | (dl > dr ? 1ull : 0ull);
  }
  unsigned long long rank = 0;  // 4 bits per sorted entry
  for(int i = 0; i + 1 < k; ++i)
  {
    unsigned long long top = 0;  // the largest gap between i and j so far
    for(int j = i + 1; j < k; ++j)
    {
      const unsigned long long gj = e[(unsigned)(kCandCap + j - 1) * es];
      top = gj > top ? gj : top;
      // right first at their common ancestor: j (on the right) precedes i
      rank += 1ull << (4 * ((top & 1ull) ? i : j));
    }
  }
  unsigned long long perm = 0;
  for(int i = 0; i < k; ++i)
  {
    const unsigned r = (unsigned)(rank >> (4 * i)) & 15u;
    perm |= (e[(unsigned)i * es] & 15ull) << (4 * r);
  }
  return perm;
}

__device__ __forceinline__ double shfl_f64(double v, int src)
{
  return __hiloint2double(__shfl_sync(0xffffffffu, __double2hiint(v), src), __shfl_sync(0xffffffffu, __double2loint(v), src));
}

// distance (not squared) from every query to the closest point the sample pass found for its group of Morton neighbours
// (DBL_MAX where it found nothing): this rank's contribution to the bound the ranks of a partitioned surface exchange
__global__ void __launch_bounds__(256) sd_ext_bound_kernel(Desc<3> qpts, const int32_t* __restrict__ perm, int npts,
                                                            const double* __restrict__ hint_tab, int hint_shift, double* __restrict__ ext)
{
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if(t >= npts) return;
  const int qi = perm ? perm[t] : t;
  const double* hp = hint_tab + 3 * (size_t)(t >> hint_shift);
  const double tx = hp[0], ty = hp[1], tz = hp[2];
  double e = DBL_MAX;
  if(tx == tx)
  {
    const double hx = tx - ld_comp<double>(qpts, 0, qi), hy = ty - ld_comp<double>(qpts, 1, qi), hz = tz - ld_comp<double>(qpts, 2, qi);
    e = sqrt(hx * hx + hy * hy + hz * hz) * (1.0 + 1e-15);
  }
  ext[qi] = e;
}

// the largest triangle diameter of the mesh (quads: all four vertices), as an ordered-encoded running maximum
template <int NV>
__global__ void __launch_bounds__(256) sd_max_diam_kernel(const double* __restrict__ soup, int nleaves, unsigned long long* __restrict__ out)
{
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  double m = 0.0;
  if(p < nleaves)
  {
    V3 v[NV];
    load_leaf<NV>(soup, p, v);
    const int nv = (NV == 4 && !has_fourth(v[NV - 1])) ? 3 : NV;
    for(int a = 0; a < nv; ++a)
      for(int b = a + 1; b < nv; ++b)
      {
        const V3 d = v3sub(v[a], v[b]);
        m = fmax(m, v3dot(d, d));
      }
  }
  m = warp_max(m);
  if(lane_id() == 0 && m > 0.0) atomicMax(out, f64_to_ordered(m));
}

//------------------------------------------------------------------------------------------
// PHASE 1 kernel.  Persistent warps pull Morton-ordered queries from a device cursor, one query per lane, a lane that
// finishes is refilled at once.  A warp iteration is either
//   an inner step   every lane with a walk in progress visits one node (binary32 bounds of its two children, the
//                   nearer one next, the other on the lane's stack in shared memory) or takes one entry off its
//                   stack; a LEAF that passes its bound is not evaluated by the lane: it goes to the warp's leaf
//                   pool (owner lane, sorted position, bound) and the lane walks on;
//   a leaf batch    when the pool holds 32 leaves (or nobody can walk): lane j evaluates pool entry j for ITS owner
//                   (query point and threshold fetched by shuffle) -- closest_point at full width whatever the
//                   lanes' own walks are doing -- and the results go back to the owners (match.any groups the
//                   entries by owner; an owner folds its group: exact minimum, closest point for the next hint,
//                   in-window leaves appended to its list in global memory).
// Output per query slot t (Morton rank): cand_n[t] (kCandOverflow: the list overflowed), cand[t][0..n), seed[t] = d*^2.
//------------------------------------------------------------------------------------------
constexpr int kPoolCap = 64;  // leaf pool entries per warp (a batch runs as soon as 32 are waiting)
constexpr uint8_t kCandOverflow = 255;
constexpr size_t kSd2SmemMin = (size_t)kSd2Threads * (size_t)kSd2Stack * sizeof(unsigned long long) +
                               (size_t)(kSd2Threads / 32) * ((size_t)kPoolCap * sizeof(unsigned long long) + 32 * sizeof(unsigned));

template <int NV>
__global__ void __launch_bounds__(kSd2Threads, AXB_SD2_MIN_BLOCKS)
sd_min_kernel(const SdNode64* __restrict__ nodes, const double* __restrict__ soup, Desc<3> qpts, int npts, const int32_t* __restrict__ perm,
              int32_t* __restrict__ cand, uint8_t* __restrict__ cand_n, double* __restrict__ seed, unsigned long long* __restrict__ work,
              unsigned int* __restrict__ cursor, unsigned chunk, double window, const double* __restrict__ hint_tab, int hint_shift,
              double* __restrict__ hint_out, int nfull, unsigned heavy_visits, const double* __restrict__ ext_bound, double ext_slack)
{
  // Two uses.  hint_out == nullptr: the search proper over the npts query slots.  hint_out != nullptr: the SAMPLE pass
  // that runs first -- work item k is the query of Morton rank k * 2^hint_shift + 2^(hint_shift-1) (of nfull), and all
  // that is kept of it is its closest point, hint_out[k].  The search proper then starts every query with the bound
  // |q - hint_tab[rank >> hint_shift]|^2: a point of the surface found for a query at most 2^(hint_shift-1) ranks away,
  // whatever the cursor hands to this lane before and after (small launches, chunk starts, the first query of a lane).
  const bool sampling = hint_out != nullptr;
  constexpr unsigned FULL = 0xffffffffu;
  const unsigned lane = lane_id();
  const unsigned lt_mask = (1u << lane) - 1u;
  const unsigned warp = threadIdx.x >> 5;

  extern __shared__ unsigned long long sd2_smem[];
  unsigned long long* const my_stack = sd2_smem + threadIdx.x;  // [level][thread]
  constexpr unsigned stride = kSd2Threads;
  unsigned long long* const pool = sd2_smem + (size_t)kSd2Stack * kSd2Threads + (size_t)warp * kPoolCap;
  unsigned* const group_of = reinterpret_cast<unsigned*>(sd2_smem + (size_t)kSd2Stack * kSd2Threads + (size_t)(kSd2Threads / 32) * kPoolCap) + warp * 32;
  group_of[lane] = 0u;
  __syncwarp();
  unsigned long long st_over[kStackSize - kSd2Stack];
  auto st_get = [&](int k) -> unsigned long long { return k < kSd2Stack ? my_stack[(unsigned)k * stride] : st_over[k - kSd2Stack]; };
  auto st_put = [&](int k, unsigned long long e) {
    if(k < kSd2Stack)
      my_stack[(unsigned)k * stride] = e;
    else
      st_over[k - kSd2Stack] = e;
  };

  // ---- per-lane query state ----
  int qt = -1;  // Morton rank of the lane's query, -1 = lane is free
  double qx = 0.0, qy = 0.0, qz = 0.0;
  double minSq = DBL_MAX;  // exact minimum over the leaves evaluated so far
  V3 minPt = {0.0, 0.0, 0.0};
  float thr_f = 0.f;  // prune threshold on squared distance, rounded up to binary32 (+inf: none yet)
  double thr = DBL_MAX;
  int ncand = 0;
  bool overflow = false;
  int sp = 0;
  int32_t cur = kBarrier;  // >= 0 inner node in hand, kBarrier: none
  int pending = 0;         // the lane's leaves waiting in the pool
  bool have_hint = false;
  bool hinted = false;  // the lane's query started with a first bound
  double ext_sq = DBL_MAX;  // the never-retracted bound from outside (see the refill)
  unsigned nleaf = 0, ninner = 0;
  unsigned visits0 = 0;  // ninner when the lane's query started
#ifdef AXB_SD_DEBUG_MISS
  double dbg_h0 = 0, dbg_h1 = 0, dbg_h2 = 0, dbg_thr0 = 0, dbg_px = 0, dbg_py = 0, dbg_pz = 0;
  unsigned dbg_back = 0, dbg_pushed = 0, dbg_minlb = 0x7f800000u;
#endif
  unsigned wbase = 0, wcount = 0;
  bool exhausted = false;
  unsigned pool_head = 0, pool_n = 0;  // warp-uniform
  const float inf_f = __int_as_float(0x7f800000);

  while(true)
  {
    // ---- lanes whose walk is over and whose leaves have all come back: store, free the lane ----
    if(qt >= 0 && cur == kBarrier && sp == 0 && pending == 0 && hinted && !overflow && minSq >= 1e300)
    {
      // Nothing within the first bound.  That bound is the distance to a POINT of the surface, and the search minimises
      // what the reference's closest_point(q, triangle, loc, EPS = 1e-12) returns per triangle -- which is not always the
      // triangle's nearest point: on triangles whose squared area is below EPS (a 20 M-triangle sphere of radius 0.5:
      // 1e-14) its fuzzy region tests pick an edge where a vertex is nearer, 3e-6 off in the squared distance.  A point
      // found through one query can therefore be nearer than anything the reference arithmetic yields for its neighbour.
      // Every leaf was rejected, so the miss is total and detectable: search again with no first bound (the bounds are
      // lower bounds of the TRUE distance, hence of the reference's value too).  About 3 queries per million on C5.
      hinted = false;
      thr = ext_sq < 1e300 ? prune_threshold_w(ext_sq, window) : DBL_MAX;
      thr_f = thr < 3.0e38 ? __double2float_ru(thr) : inf_f;
      cur = 0;
      visits0 = ninner;
    }
    if(qt >= 0 && cur == kBarrier && sp == 0 && pending == 0)
    {
      if(sampling)
      {
        const bool got = minSq < 1e300;
        const double nan = __longlong_as_double(0x7ff8000000000000ll);
        hint_out[3 * (size_t)qt + 0] = got ? minPt.x : nan;
        hint_out[3 * (size_t)qt + 1] = got ? minPt.y : nan;
        hint_out[3 * (size_t)qt + 2] = got ? minPt.z : nan;
      }
      else
      {
        cand_n[qt] = overflow ? kCandOverflow : (uint8_t)ncand;
        seed[qt] = minSq;
#ifdef AXB_SD_DEBUG_MISS
        if(work && minSq >= 1e300 && !overflow)
        {
          const unsigned long long k = atomicAdd(&work[2], 1ull);
          if(k < 4)
          {
            unsigned long long* w = work + 4 + 18 * k;
            w[10] = __double_as_longlong(qx);
            w[11] = __double_as_longlong(qy);
            w[12] = __double_as_longlong(qz);
            w[13] = __double_as_longlong(dbg_px);
            w[14] = __double_as_longlong(dbg_py);
            w[15] = __double_as_longlong(dbg_pz);
            w[0] = (unsigned long long)qt;
            w[1] = ninner - visits0;
            w[2] = ((unsigned long long)dbg_pushed << 32) | dbg_back;
            w[3] = __double_as_longlong(dbg_h0);
            w[4] = __double_as_longlong(dbg_h1);
            w[5] = __double_as_longlong(dbg_h2);
            w[6] = __double_as_longlong(dbg_thr0);
            w[7] = __double_as_longlong(thr);
            w[8] = dbg_minlb;
            w[9] = lane;
          }
        }
#endif
      }
      have_hint = have_hint || minSq < 1e300;
      qt = -1;
    }
    // ---- refill free lanes ----
    const unsigned freem = __ballot_sync(FULL, qt < 0);
    if(freem != 0u && !exhausted)
    {
      if(wcount == 0u)
      {
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
        // A second source for the first bound: the running closest point of the lane that holds the LATEST query with a
        // leaf evaluated.  Right after the warp moves on to a new chunk the lane's own previous query is far away (the
        // chunks of one warp are grid-size * 32 queries apart), but a lane that is already working on the new chunk is
        // a Morton neighbour.  Any point of the surface gives a valid upper bound; the smaller of the two is used.
        const int hkey = (qt >= 0 && minSq < 1e300) ? qt : -1;
        const int hbest = __reduce_max_sync(FULL, hkey);
        const int hsrc = max(__ffs(__ballot_sync(FULL, hkey == hbest)) - 1, 0);
        const V3 wpt {shfl_f64(minPt.x, hsrc), shfl_f64(minPt.y, hsrc), shfl_f64(minPt.z, hsrc)};
        const unsigned rank = __popc(freem & lt_mask);
        if(qt < 0 && rank < wcount)
        {
          qt = (int)(wbase + rank);
          const int qrank = sampling ? min((int)(((unsigned)qt << hint_shift) + ((1u << hint_shift) >> 1)), nfull - 1) : qt;
          const int qi = perm ? perm[qrank] : qrank;
          qx = ld_comp<double>(qpts, 0, qi);
          qy = ld_comp<double>(qpts, 1, qi);
          qz = ld_comp<double>(qpts, 2, qi);
          minSq = DBL_MAX;
          thr = DBL_MAX;
          thr_f = inf_f;
          double hsq = DBL_MAX;
#ifdef AXB_SD_DEBUG_MISS
          dbg_h0 = dbg_h1 = dbg_h2 = -1.0;
          dbg_back = dbg_pushed = 0;
          dbg_minlb = 0x7f800000u;
#endif
          if(have_hint)
          {
            // the closest point of the lane's previous query (a Morton neighbour) is a point of the surface
            const double hx = minPt.x - qx, hy = minPt.y - qy, hz = minPt.z - qz;
            hsq = hx * hx + hy * hy + hz * hz;
#ifdef AXB_SD_DEBUG_MISS
            dbg_h0 = hsq;
            dbg_px = minPt.x;
            dbg_py = minPt.y;
            dbg_pz = minPt.z;
#endif
          }
          if(hbest >= 0)
          {
            const double hx = wpt.x - qx, hy = wpt.y - qy, hz = wpt.z - qz;
            hsq = fmin(hsq, hx * hx + hy * hy + hz * hz);
#ifdef AXB_SD_DEBUG_MISS
            dbg_h1 = hx * hx + hy * hy + hz * hz;
            if(dbg_h1 <= hsq)
            {
              dbg_px = wpt.x;
              dbg_py = wpt.y;
              dbg_pz = wpt.z;
            }
#endif
          }
          if(hint_tab)
          {
            const double* hp = hint_tab + 3 * (size_t)(qt >> hint_shift);
            const double tx = __ldg(hp), ty = __ldg(hp + 1), tz = __ldg(hp + 2);
            if(tx == tx)  // NaN: the sample query found nothing
            {
              const double hx = tx - qx, hy = ty - qy, hz = tz - qz;
              hsq = fmin(hsq, hx * hx + hy * hy + hz * hz);
#ifdef AXB_SD_DEBUG_MISS
              dbg_h2 = hx * hx + hy * hy + hz * hz;
              if(dbg_h2 <= hsq)
              {
                dbg_px = tx;
                dbg_py = ty;
                dbg_pz = tz;
              }
#endif
            }
          }
          // A bound from OUTSIDE (partitioned surface: the distance from q to a point of some OTHER part, MIN over the
          // ranks): the value the reference arithmetic yields for that point's own triangle is at most this distance plus
          // the triangle's diameter (closest_point returns a point OF the triangle, even where its fuzzy region tests pick
          // a poor one), so with ext_slack = the largest triangle diameter of the whole surface nothing that can be the
          // global minimum is cut.  Unlike the lane's own first bounds it is never retracted: a part that has nothing
          // within it reports nothing for this query, and another part owns the answer.
          ext_sq = DBL_MAX;
          if(ext_bound)
          {
            const double e = ext_bound[qi];
            if(e < 1e150)
            {
              const double dd = e + ext_slack;
              ext_sq = dd * dd * (1.0 + 1e-12);
            }
          }
          hinted = hsq < ext_sq;  // a bound of the lane's own that is tighter than the safe one: retried on a total miss
          hsq = fmin(hsq, ext_sq);
          if(hsq < 1e300)
          {
            thr = prune_threshold_w(hsq, window);
            thr_f = thr < 3.0e38 ? __double2float_ru(thr) : inf_f;
          }
#ifdef AXB_SD_DEBUG_MISS
          dbg_thr0 = thr;
#endif
          ncand = 0;
          overflow = false;
          sp = 0;
          cur = 0;
          visits0 = ninner;
        }
        const unsigned taken = min((unsigned)__popc(freem), wcount);
        wbase += taken;
        wcount -= taken;
      }
    }
    const bool busy = qt >= 0;
    const bool want_inner = busy && (cur >= 0 || sp > 0);
    const unsigned minner = __ballot_sync(FULL, want_inner);
    if(minner == 0u && pool_n == 0u)
    {
      if(__ballot_sync(FULL, busy) == 0u && exhausted) break;
      continue;  // lanes just stored / were just refilled
    }

    if(pool_n >= (unsigned)AXB_SD2_BATCH_MIN || minner == 0u)
    {
      // ---- leaf batch ----
      const unsigned nb = min(pool_n, 32u);
      const bool have = lane < nb;
      const unsigned long long en = have ? pool[(pool_head + lane) & (kPoolCap - 1)] : 0ull;
      const int owner = (int)((en >> 27) & 31ull);
      const int pos = (int)(unsigned)(en >> 32);
      const float lb = __uint_as_float((unsigned)(en & 0x7ffffffull) << 5);
      const float othr = __shfl_sync(FULL, thr_f, owner);
      const double oqx = shfl_f64(qx, owner), oqy = shfl_f64(qy, owner), oqz = shfl_f64(qz, owner);
      const bool live = have && lb <= othr;
      double sq = DBL_MAX;
      V3 cp = {0.0, 0.0, 0.0};
      if(live)
      {
        constexpr double EPS = 1e-12;
        ++nleaf;
        const V3 oq {oqx, oqy, oqz};
        V3 v[NV];
        load_leaf<NV>(soup, pos, v);
        int loc;
        cp = closest_point_tri(oq, v[0], v[1], v[2], loc, EPS);
        const V3 dq = v3sub(cp, oq);
        sq = v3dot(dq, dq);
        if(NV == 4 && has_fourth(v[NV - 1]))
        {
          const V3 cp2 = closest_point_tri(oq, v[0], v[2], v[NV - 1], loc, EPS);
          const V3 dq2 = v3sub(cp2, oq);
          const double sq2 = v3dot(dq2, dq2);
          if(sq2 < sq)
          {
            sq = sq2;
            cp = cp2;
          }
        }
      }
      // back to the owners: the entries of one owner form a group (match.any); its first lane tells the owner who they are
      const unsigned peers = __match_any_sync(FULL, have ? owner : 32 + (int)lane);
      if(have && (peers & lt_mask) == 0u) group_of[owner] = peers;
      __syncwarp();
      unsigned mine = group_of[lane];
      group_of[lane] = 0u;
      __syncwarp();
      pending -= __popc(mine);
      const int trips = __reduce_max_sync(FULL, (unsigned)__popc(mine));
      for(int i = 0; i < trips; ++i)
      {
        const bool on = mine != 0u;
        const int src = on ? (__ffs(mine) - 1) : (int)lane;
        mine &= mine - 1u;
        const double rsq = shfl_f64(sq, src);
        const int rpos = __shfl_sync(FULL, pos, src);
#ifdef AXB_SD_DEBUG_MISS
        if(on) ++dbg_back;
#endif
        const V3 rcp {shfl_f64(cp.x, src), shfl_f64(cp.y, src), shfl_f64(cp.z, src)};
        if(on && !overflow && rsq <= thr)
        {
          if(rsq < minSq)
          {
            const double nthr = prune_threshold_w(rsq, window);
            if(minSq > nthr) ncand = 0;  // everything remembered so far is outside the new window
            minSq = rsq;
            minPt = rcp;
            if(nthr < thr)
            {
              thr = nthr;
              thr_f = thr < 3.0e38 ? __double2float_ru(thr) : inf_f;
            }
          }
          if(ncand < kCandCap)
          {
            if(!sampling) cand[(size_t)qt * kCandCap + ncand] = rpos;
            ++ncand;
          }
          else
          {
            // too many leaves within the tie window: phase 2 walks the tree in the reference's order for this query
            overflow = true;
            cur = kBarrier;
            sp = 0;
          }
        }
      }
      pool_head = (pool_head + nb) & (kPoolCap - 1);
      pool_n -= nb;
      continue;
    }

    // ---- inner step ----
    int32_t next = kBarrier;
    float next_lb = 0.f;
    if(want_inner)
    {
      if(cur >= 0 && ninner - visits0 >= heavy_visits)
      {
        // A heavy query (the neighbourhood of a centre of curvature: hundreds of near-equidistant patches).  Its serial
        // chain would outlast everybody else's: the lane gives it up, and one whole warp of sd_solo_kernel redoes it.
        overflow = true;
        cur = kBarrier;
        sp = 0;
      }
      if(cur >= 0)
      {
        ++ninner;
        // the compact record (SdNode64): two sectors, decoded into the layout obb_sqdist_mixed expects
        const D4* rec = reinterpret_cast<const D4*>(nodes + cur);
        const D4 r0 = ldg256(rec), r1 = ldg256(rec + 1);
        const int32_t child0 = __double2loint(r0.x), child1 = __double2hiint(r0.x);
#if AXB_SD2_PREFETCH
        // the step that follows reads one of these two records: start both on their way to L1 while this step's ~300
        // instructions of decode and bound arithmetic run (the load of the next record is the largest single stall)
        if(child0 >= 0) asm volatile("prefetch.global.L1 [%0];" ::"l"(nodes + child0));
        if(child1 >= 0) asm volatile("prefetch.global.L1 [%0];" ::"l"(nodes + child1));
#endif
        const float ox = __int_as_float(__double2loint(r0.y)), oy = __int_as_float(__double2hiint(r0.y));
        const float oz = __int_as_float(__double2loint(r0.z)), step = __int_as_float(__double2hiint(r0.z));
        const unsigned w6 = (unsigned)__double2loint(r0.w), w7 = (unsigned)__double2hiint(r0.w);
        const unsigned w8 = (unsigned)__double2loint(r1.x), w9 = (unsigned)__double2hiint(r1.x);
        const unsigned w10 = (unsigned)__double2loint(r1.y), w11 = (unsigned)__double2hiint(r1.y);
        const unsigned w12 = (unsigned)__double2loint(r1.z), w13 = (unsigned)__double2hiint(r1.z);
        const unsigned w14 = (unsigned)__double2loint(r1.w);
        const double qrx = qx - (double)ox, qry = qy - (double)oy, qrz = qz - (double)oz;
        const float rx = __double2float_rn(qrx), ry = __double2float_rn(qry), rz = __double2float_rn(qrz);
        const float r1n = fabsf(rx) + fabsf(ry) + fabsf(rz);
        const float mg = 5.0e-7f * r1n;
        float f[24];
        {
          auto lo16 = [](unsigned w) { return (int)(short)(w & 0xffffu); };
          auto hi16 = [](unsigned w) { return (int)(short)(w >> 16); };
          auto hlo = [](unsigned w) { return __half2float(__ushort_as_half((unsigned short)(w & 0xffffu))); };
          auto hhi = [](unsigned w) { return __half2float(__ushort_as_half((unsigned short)(w >> 16))); };
          sd_frame(lo16(w6), hi16(w6), lo16(w7), f);
          sd_frame(hi16(w7), lo16(w8), hi16(w8), f + 12);
          f[6] = __fmul_rn((float)lo16(w9), step);
          f[7] = __fmul_rn((float)hi16(w9), step);
          f[8] = __fmul_rn((float)lo16(w10), step);
          f[18] = __fmul_rn((float)hi16(w10), step);
          f[19] = __fmul_rn((float)lo16(w11), step);
          f[20] = __fmul_rn((float)hi16(w11), step);
          f[9] = __fmul_rn(hlo(w12), step);
          f[10] = __fmul_rn(hhi(w12), step);
          f[11] = __fmul_rn(hlo(w13), step);
          f[21] = __fmul_rn(hhi(w13), step);
          f[22] = __fmul_rn(hlo(w14), step);
          f[23] = __fmul_rn(hhi(w14), step);
        }
#if AXB_SD2_NORMAL_F64
        float s0 = obb_sqdist_mixed(f, qrx, qry, qrz, rx, ry, rz, mg), s1 = obb_sqdist_mixed(f + 12, qrx, qry, qrz, rx, ry, rz, mg);
#else
        float s0 = obb_sqdist_f32(f, rx, ry, rz, mg), s1 = obb_sqdist_f32(f + 12, rx, ry, rz, mg);
#endif
        // queries farther than binary32 squares can hold: no pruning by bound (valid children only)
        const bool huge = !(r1n < 1.0e18f);
        s0 = huge ? 0.f : s0;
        s1 = huge ? 0.f : s1;
        const bool in0 = f[9] >= 0.f && s0 <= thr_f, in1 = f[21] >= 0.f && s1 <= thr_f;  // half extent -inf marks an invalid box
        // nearer bound first; the other child waits on the stack with its bound
        const bool swap = in1 && (!in0 || s1 < s0);
        const int32_t c_first = swap ? child1 : child0, c_second = swap ? child0 : child1;
        const float s_first = swap ? s1 : s0, s_second = swap ? s0 : s1;
        if(in0 && in1)
        {
          st_put(sp, ((unsigned long long)__float_as_uint(s_second) << 32) | (unsigned)c_second);
          ++sp;
        }
        if(in0 || in1)
        {
          next = c_first;
          next_lb = s_first;
        }
      }
      // No node from the visit: take ONE entry off the stack (a stale one costs the lane this step, not a loop the
      // whole warp waits for)
      if(next == kBarrier && sp > 0)
      {
        --sp;
        const unsigned long long e = st_get(sp);
        const float lb = __uint_as_float((unsigned)(e >> 32));
        if(lb <= thr_f)
        {
          next = (int32_t)(unsigned)(e & 0xffffffffull);
          next_lb = lb;
        }
      }
    }
    // leaves go to the warp's pool
    const bool is_leaf = next < 0 && next != kBarrier;
    const unsigned mpush = __ballot_sync(FULL, is_leaf);
    if(is_leaf)
    {
      const unsigned slot = (pool_head + pool_n + __popc(mpush & lt_mask)) & (kPoolCap - 1);
      // bound truncated to 27 bits (toward zero: never larger than the bound itself)
      pool[slot] = ((unsigned long long)(unsigned)(-next - 1) << 32) | ((unsigned long long)lane << 27) | (unsigned long long)(__float_as_uint(next_lb) >> 5);
      ++pending;
#ifdef AXB_SD_DEBUG_MISS
      ++dbg_pushed;
      dbg_minlb = min(dbg_minlb, __float_as_uint(next_lb));
#endif
      next = kBarrier;
    }
    pool_n += __popc(mpush);
    __syncwarp();
    if(want_inner) cur = next;
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

//------------------------------------------------------------------------------------------
// PHASE 2 kernel.  One lane per query slot ("owner"); the warp works through its owners' remembered leaves together:
//   (a) the leaves, flattened over the warp in rounds of 32: lane j evaluates the j-th one for its owner (closest
//       point, squared distance, location code): the expensive arithmetic at full width;
//   (b) each owner served in the round looks at its results (by shuffle): the exact minimum, which leaves are inside
//       its window (the others are no-ops of the state machine), and whether their ORDER can matter:
//         * all in-window candidates are the same feature (same location type, closest points within EPS/4 of the
//           first minimum's): then no candidate ever clears the normal sum, the final minimum is the first one that
//           attains it, and the sum holds every candidate's term -- the order only decides (1) which of several
//           EXACT ties gives the closest point (matters only if their closest points differ bitwise, or for a face,
//           whose own normal is used) and (2) the rounding of the normal sum (matters only if unit normals are
//           requested, or if the sign test |r . sumN| is within rounding of zero: checked in (e));
//         * anything else (distinct features within 1e-6 of each other): the order matters.
//       Where it matters the owner establishes the reference's visiting order (sd_order);
//   (c) the state machine over the in-window leaves, fed by shuffle from the evaluating lanes;
//   (d) the deferred normal terms of all owners (unit normals, vertex angles: sqrt, divisions, acos), again
//       flattened over the warp, summed by each owner in its own order;
//   (e) sign, distance, outputs (sd_finish).  An owner whose list overflowed, or whose sign test is too close to call
//       without the exact summation order, walks the tree in the reference's order on its own (sd_ordered_query).
//------------------------------------------------------------------------------------------
struct FlatSlots
{
  bool sel;    // this owner's items are in the current round
  int base;    // flat index of the owner's first item
  int total;   // items in the round
  int owner;   // owner lane of flat index `lane`
  int obase;   // that owner's base
};
// owners with `want` > 0 items, in lane order, while their items fit in 32 lanes
__device__ __forceinline__ FlatSlots flat_slots(int want)
{
  constexpr unsigned FULL = 0xffffffffu;
  const int lane = (int)lane_id();
  int incl = want;
#pragma unroll
  for(int o = 1; o < 32; o <<= 1)
  {
    const int v = __shfl_up_sync(FULL, incl, o);
    if(lane >= o) incl += v;
  }
  FlatSlots f;
  f.sel = want > 0 && incl <= 32;
  f.base = incl - want;
  f.total = __shfl_sync(FULL, f.sel ? incl : 0, 31 - __clz(max(__ballot_sync(FULL, f.sel), 1u)));
  int lo = 0, hi = 31;  // incl is non-decreasing over lanes: the first lane whose inclusive count exceeds `lane`
#pragma unroll
  for(int it = 0; it < 5; ++it)
  {
    const int mid = (lo + hi) >> 1;
    const int v = __shfl_sync(FULL, incl, mid);
    if(v > lane)
      hi = mid;
    else
      lo = mid + 1;
  }
  f.owner = lo;
  f.obase = __shfl_sync(FULL, f.base, f.owner);
  return f;
}

template <int NV>
__global__ void __launch_bounds__(kSd2Threads) sd_resolve_kernel(const SdNode* __restrict__ nodes, const SdCen* __restrict__ cens,
                                                                 const SdUp* __restrict__ up, const int32_t* __restrict__ leaf_parent,
                                                                 const double* __restrict__ soup, SdParams prm, Desc<3> qpts, int npts,
                                                                 const int32_t* __restrict__ perm, const int32_t* __restrict__ cand,
                                                                 const uint8_t* __restrict__ cand_n, double* __restrict__ seed,
                                                                 double* __restrict__ phi, double* __restrict__ cps, double* __restrict__ nrms,
                                                                 unsigned long long* __restrict__ work, unsigned int* __restrict__ solo_ctr,
                                                                 unsigned int solo_cap, int32_t* __restrict__ solo_list)
{
  constexpr unsigned FULL = 0xffffffffu;
  constexpr int NSUB = NV == 4 ? 2 : 1;
  constexpr double EPS = 1e-12;
  const unsigned lane = lane_id();
  const bool cn = prm.compute_sign != 0;
  const bool need_sum = cn;  // computeNormal is m_computeSign (quest/SignedDistance.hpp:555,573)
  __shared__ unsigned long long order_smem[2 * kCandCap * kSd2Threads];  // sd_order's scratch / the owners' deferred terms, [level][thread]
  unsigned long long* const my_col = order_smem + threadIdx.x;
  constexpr unsigned stride = kSd2Threads;
  const unsigned warp_thread0 = threadIdx.x & ~31u;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;  // whole warps stay together: lanes past the end own nothing
  const bool valid = t < npts;
  const int qi = valid ? (perm ? perm[t] : t) : 0;
  double qx = 0.0, qy = 0.0, qz = 0.0;
  int kk = 0;
  bool solo = false;  // this owner walks the tree on its own at the end
  if(valid)
  {
    qx = ld_comp<double>(qpts, 0, qi);
    qy = ld_comp<double>(qpts, 1, qi);
    qz = ld_comp<double>(qpts, 2, qi);
    const uint8_t n = cand_n[t];
    solo = n == kCandOverflow;
    kk = solo ? 0 : (int)n;
  }
  const int32_t* const my_cand = cand + (size_t)(valid ? t : 0) * kCandCap;
  unsigned nleaf = 0, ninner = 0;
  MinCand m;
  mincand_reset(m);
  Contribs cl;
  cl.n = 0;
  bool ordered = true;    // the normal sum was (or will be) formed in the reference's order
  double term_abs = 0.0;  // upper bound on the sum of the terms' lengths
  bool served = kk == 0;
  while(__ballot_sync(FULL, !served) != 0u)
  {
    // (a)
    const FlatSlots fs = flat_slots(served ? 0 : kk);
    const bool sel = fs.sel;
    const int base = fs.base;
    const double oqx = shfl_f64(qx, fs.owner), oqy = shfl_f64(qy, fs.owner), oqz = shfl_f64(qz, fs.owner);
    const unsigned long long ocand = (unsigned long long)__shfl_sync(FULL, (unsigned)((unsigned long long)(uintptr_t)my_cand >> 32), fs.owner) << 32 |
                                     (unsigned long long)__shfl_sync(FULL, (unsigned)((unsigned long long)(uintptr_t)my_cand & 0xffffffffull), fs.owner);
    double e_sq[NSUB];
    V3 e_cp[NSUB];
    int e_loc[NSUB];
    int e_pos = 0;
#pragma unroll
    for(int u = 0; u < NSUB; ++u)
    {
      e_sq[u] = DBL_MAX;  // a missing second triangle is never inside a window
      e_cp[u] = {0.0, 0.0, 0.0};
      e_loc[u] = 3;
    }
    if((int)lane < fs.total)
    {
      e_pos = __ldg(reinterpret_cast<const int32_t*>((uintptr_t)ocand) + ((int)lane - fs.obase));
      const V3 oq {oqx, oqy, oqz};
      V3 v[NV];
      load_leaf<NV>(soup, e_pos, v);
      e_cp[0] = closest_point_tri(oq, v[0], v[1], v[2], e_loc[0], EPS);
      const V3 dq = v3sub(e_cp[0], oq);
      e_sq[0] = v3dot(dq, dq);
      if(NV == 4 && has_fourth(v[NV - 1]))
      {
        e_cp[NSUB - 1] = closest_point_tri(oq, v[0], v[2], v[NV - 1], e_loc[NSUB - 1], EPS);
        const V3 dq2 = v3sub(e_cp[NSUB - 1], oq);
        e_sq[NSUB - 1] = v3dot(dq2, dq2);
      }
    }
    // (b) on the candidate lanes, in parallel; an owner's candidates are the lanes [base, base + kk): reductions over
    // that segment are 4 shuffle steps (kk <= kCandCap < 16), flags travel by ballot
    const int okk = __shfl_sync(FULL, sel ? kk : 0, fs.owner);
    const int li = (int)lane - fs.obase;  // list index of this lane's candidate
    const int seg_end = fs.obase + okk;
    // the exact minimum and the first (sub-)triangle in list order that attains it
    int lu = 0;
    double lsq = e_sq[0];
    if(NSUB == 2 && e_sq[NSUB - 1] < lsq)
    {
      lsq = e_sq[NSUB - 1];
      lu = 1;
    }
    double r_sq = (int)lane < fs.total ? lsq : DBL_MAX;
    int r_li = li;
#pragma unroll
    for(int o = 1; o < 16; o <<= 1)
    {
      const double v2 = shfl_f64(r_sq, min((int)lane + o, 31));
      const int i2 = __shfl_sync(FULL, r_li, min((int)lane + o, 31));
      if((int)lane + o < seg_end && v2 < r_sq)  // ties keep the lower list index
      {
        r_sq = v2;
        r_li = i2;
      }
    }
    const int head = min(max(fs.obase, 0), 31);
    const double c_msq = shfl_f64(r_sq, head);
    const int c_first = __shfl_sync(FULL, r_li, head);
    const int first_lane = min(max(fs.obase + c_first, 0), 31);
    const V3 bcp = lu == 0 ? e_cp[0] : e_cp[NSUB - 1];
    const int btype = loc_type(lu == 0 ? e_loc[0] : e_loc[NSUB - 1]);
    const V3 cp0 {shfl_f64(bcp.x, first_lane), shfl_f64(bcp.y, first_lane), shfl_f64(bcp.z, first_lane)};
    const int type0 = __shfl_sync(FULL, btype, first_lane);
    const int first_u = __shfl_sync(FULL, lu, first_lane);
    // The in-window (sub-)triangles in two classes relative to the first minimum c0: G, the same feature (same
    // location type, closest point within sqrt(EPS)/2), and O, clearly another one (other type, or closest point
    // farther than 2 sqrt(EPS)).  If nothing falls in between and every O member is farther from the query than every
    // G member, an O member can only be the running minimum BEFORE the first G member arrives, which then clears
    // whatever it left: the final state is that of the G members alone, whatever the interleaving.
    bool l_in = false, l_inG = false, l_gap = false, l_tie = false, l_G2 = false;
    double l_maxG = 0.0, l_minO = DBL_MAX;
    if((int)lane < fs.total)
    {
      const double wthr = prune_threshold_w(c_msq, cn ? kTieWindow : 0.0);
      int ng = 0;
#pragma unroll
      for(int u = 0; u < NSUB; ++u)
      {
        const double sq = e_sq[u];
        if(sq <= wthr)
        {
          l_in = true;
          const V3 d = v3sub(e_cp[u], cp0);
          const double d2 = v3dot(d, d);
          const bool same_type = loc_type(e_loc[u]) == type0;
          const bool isG = same_type && d2 <= 0.25 * EPS, isO = !same_type || d2 > 4.0 * EPS;
          l_gap = l_gap || (!isG && !isO);
          if(isG)
          {
            ++ng;
            l_maxG = fmax(l_maxG, sq);
            if(sq == c_msq && !(li == c_first && u == first_u))
              l_tie = l_tie || type0 == 2 || e_cp[u].x != cp0.x || e_cp[u].y != cp0.y || e_cp[u].z != cp0.z;
          }
          else if(isO)
            l_minO = fmin(l_minO, sq);
        }
      }
      l_inG = ng > 0;
      l_G2 = ng > 1;
    }
    const unsigned m_in = __ballot_sync(FULL, l_in), m_G = __ballot_sync(FULL, l_inG), m_G2 = __ballot_sync(FULL, l_G2);
    const unsigned m_gap = __ballot_sync(FULL, l_gap), m_tie = __ballot_sync(FULL, l_tie);
#pragma unroll
    for(int o = 1; o < 16; o <<= 1)
    {
      const double g2 = shfl_f64(l_maxG, min((int)lane + o, 31)), o2 = shfl_f64(l_minO, min((int)lane + o, 31));
      if((int)lane + o < seg_end)
      {
        l_maxG = fmax(l_maxG, g2);
        l_minO = fmin(l_minO, o2);
      }
    }
    // back on the owner lanes
    const int obl = min(max(base, 0), 31);
    const double maxG = shfl_f64(l_maxG, obl), minO = shfl_f64(l_minO, obl);
    const unsigned seg = sel ? (((kk >= 32 ? 0u : (1u << kk)) - 1u) << base) : 0u;
    const bool gap = (m_gap & seg) != 0u, tie_hard = (m_tie & seg) != 0u;
    const bool g_only = !gap && minO > maxG;
    const unsigned bits = ((g_only ? m_G : m_in) & seg) >> base;  // list indices that go through the state machine
    int nord = __popc(bits);
    const int nin = g_only ? __popc(m_G & seg) + __popc(m_G2 & seg) : nord * NSUB;  // (sub-)triangles whose normal terms may be summed
    unsigned long long pm = 0;  // those list indices, 4 bits each
    {
      unsigned bb = bits;
      for(int i = 0; i < nord; ++i)
      {
        pm |= (unsigned long long)(__ffs(bb) - 1) << (4 * i);
        bb &= bb - 1u;
      }
    }
    const bool order_matters =
      nord >= 2 && (!g_only || tie_hard || (nrms != nullptr && cn && nin >= 3) || nin > kContribCap);
    if(sel && order_matters)
    {
      for(int i = 0; i < nord; ++i)
      {
        const unsigned idx = (unsigned)(pm >> (4 * i)) & 15u;
        my_col[(unsigned)i * stride] = ((unsigned long long)(unsigned)my_cand[idx] << 8) | idx;
      }
      pm = sd_order(cens, up, leaf_parent, qx, qy, qz, my_col, stride, nord);
    }
    if(sel) ordered = order_matters || nin < 3;
    // (c)
    const int omax = __reduce_max_sync(FULL, sel ? (unsigned)nord : 0u);
    for(int i = 0; i < omax; ++i)
    {
      const int idx = (int)((pm >> (4 * i)) & 15ull);
      const int src = min(base + idx, 31);
      const int pos = __shfl_sync(FULL, e_pos, src);
#pragma unroll
      for(int u = 0; u < NSUB; ++u)
      {
        const double sq = shfl_f64(e_sq[u], src);
        const V3 cp {shfl_f64(e_cp[u].x, src), shfl_f64(e_cp[u].y, src), shfl_f64(e_cp[u].z, src)};
        const int loc = __shfl_sync(FULL, e_loc[u], src);
        if(sel && i < nord && sq < 1e300) apply_candidate<NV>(soup, m, cl, cp, sq, loc, pos, u, need_sum);
      }
    }
    if(sel)
    {
      nleaf += (unsigned)kk;
      served = true;
    }
  }
  // (d) the deferred normal terms of every owner, flattened over the warp
  __syncwarp();
  for(int i = 0; i < kContribCap; ++i)
    if(i < cl.n) my_col[(unsigned)i * stride] = ((unsigned long long)(unsigned)cl.pos[i] << 8) | (unsigned long long)(uint8_t)cl.code[i];
  __syncwarp();
  bool summed = cl.n == 0;
  while(__ballot_sync(FULL, !summed) != 0u)
  {
    const FlatSlots fs = flat_slots(summed ? 0 : cl.n);
    V3 term = {0.0, 0.0, 0.0};
    bool skip = true;
    if((int)lane < fs.total)
    {
      const unsigned long long en = order_smem[(unsigned)((int)lane - fs.obase) * stride + warp_thread0 + (unsigned)fs.owner];
      const int pos = (int)(unsigned)(en >> 8);
      const int code = (int)(int8_t)(en & 0xffull);
      const int sub = code >> 3, loc = (code & 7) - 3;
      V3 v[NV];
      load_leaf<NV>(soup, pos, v);
      const V3 T[3] = {v[0], sub == 0 ? v[1] : v[2], sub == 0 ? v[2] : v[NV - 1]};
      const V3 n = v3cross(v3sub(T[1], T[0]), v3sub(T[2], T[0]));  // Triangle::normal :98-102
      if(loc < 0)
      {
        term = v3unit(n);  // edge: += normal().unitVector() (:703)
        skip = false;
      }
      else
      {
        const double area = 0.5 * sqrt(v3dot(n, n));  // Triangle::area :105-109
        if(!nearly_eq(area, 0.0, 1.0e-12))            // !degenerate() :326-330
        {
          const double alpha = tri_angle(T, loc);
          term = v3mul(v3unit(n), alpha);  // vertex: += angle(loc)*normal().unitVector() (:722-728)
          skip = false;
        }
      }
    }
    const int nmax = __reduce_max_sync(FULL, fs.sel ? (unsigned)cl.n : 0u);
    for(int i = 0; i < nmax; ++i)
    {
      const int src = min(fs.base + i, 31);
      const V3 tv {shfl_f64(term.x, src), shfl_f64(term.y, src), shfl_f64(term.z, src)};
      const bool sk = __shfl_sync(FULL, (int)skip, src) != 0;
      if(fs.sel && i < cl.n && !sk)
      {
        m.sumN = v3add(m.sumN, tv);
        term_abs += fabs(tv.x) + fabs(tv.y) + fabs(tv.z);
      }
    }
    if(fs.sel) summed = true;
  }
  // (e)
  if(valid)
  {
    const V3 q {qx, qy, qz};
    if(!solo && !ordered && cn && m.minType != 2)
    {
      // the sum was formed in list order: its rounding differs from the reference's by a few ulps of the terms;
      // the sign of r . sumN is certain unless it is that close to zero
      const V3 r = v3sub(q, m.minPt);
      const double dotv = v3dot(r, m.sumN);
      const double slack = 1e-13 * (fabs(r.x) + fabs(r.y) + fabs(r.z)) * term_abs;
      solo = !(fabs(dotv) > slack);
    }
    bool handed = false;
    if(solo && solo_list)
    {
      // a heavy query: listed for sd_solo_kernel (one warp each) instead of a serial walk in this lane
      const unsigned at = atomicAdd(solo_ctr, 1u);
      if(at < solo_cap)
      {
        solo_list[at] = t;
        if(cand_n[t] != kCandOverflow) seed[t] = m.minSq;
        handed = true;
      }
    }
    if(!handed)
    {
      if(solo) sd_ordered_query<NV>(nodes, cens, soup, q, cand_n[t] == kCandOverflow ? seed[t] : m.minSq, cn, m, nleaf, ninner);
      sd_finish<NV>(soup, prm, q, m, qi, phi, cps, nrms);
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

//------------------------------------------------------------------------------------------
// The heavy queries, one WARP each.  A query whose remembered-leaf list overflowed (dozens of triangles within 1e-6 of
// each other: the neighbourhood of a sphere's centre), or whose sign test needs the reference's exact summation order,
// used to walk the tree alone in its resolve lane: ~2500 dependent node visits at ~2 us each, a 5 ms serial chain that
// set the tail of every launch (and 70 % of the 8-GPU efficiency).  Here the resolve kernel only lists such queries;
// this kernel, launched behind it, hands them to warps one by one:
//   stage A  the exact minimum d*: a warp-wide search over a shared stack in global memory (32 nodes per step, children
//            pushed farther-first, leaves evaluated on the spot, the bound shrinking with every batch);
//   stage B  every leaf whose oriented bound is within the window of d*, IN THE REFERENCE'S VISITING ORDER: the
//            frontier is an ordered list, each level replaces every inner node by its surviving children in the order
//            LinearBVHTraverser enters them (nearer AABB centroid first, LinearBVH.hpp:72-85) -- a level-synchronous
//            expansion keeps the depth-first order, leaves are carried along in place;
//   replay   the unchanged state machine over that list (closest points evaluated 32 at a time), then sign and store.
// With the threshold fixed at thr(d*) from the start, the leaves visited are exactly those sd_ordered_query ends up
// visiting once its own bound has converged, in the same order.  A frontier that outgrows the scratch (kSoloCap) falls
// back to sd_ordered_query.
//------------------------------------------------------------------------------------------
constexpr int kSoloCap = 8192;  // scratch entries (8 B) per warp: stage A's stack, then two int32 lists of kSoloCap
constexpr int kSoloThreads = 128;

__device__ __forceinline__ int warp_excl_scan(int v, int& total)
{
  const unsigned lane = lane_id();
  int incl = v;
#pragma unroll
  for(int o = 1; o < 32; o <<= 1)
  {
    const int u = __shfl_up_sync(0xffffffffu, incl, o);
    if((int)lane >= o) incl += u;
  }
  total = __shfl_sync(0xffffffffu, incl, 31);
  return incl - v;
}

// the two children of inner node `cur` as the ordered walk sees them: ids and double-precision oriented bounds
__device__ __forceinline__ void solo_visit(const SdNode* __restrict__ nodes, int32_t cur, const double* qp, int32_t& child0, int32_t& child1,
                                           double& d20, double& d21)
{
  const D4* rec = reinterpret_cast<const D4*>(nodes + cur);
  const D4 r0 = ldg256(rec), r1 = ldg256(rec + 1), r2 = ldg256(rec + 2), r3 = ldg256(rec + 3);
  const long long ids = __double_as_longlong(r0.x);
  child0 = (int32_t)(ids & 0xffffffffll);
  child1 = (int32_t)(ids >> 32);
  const double rq[3] = {qp[0] - r0.y, qp[1] - r0.z, qp[2] - r0.w};
  float f[24];
  const double w[12] = {r1.x, r1.y, r1.z, r1.w, r2.x, r2.y, r2.z, r2.w, r3.x, r3.y, r3.z, r3.w};
#pragma unroll
  for(int k = 0; k < 12; ++k)
  {
    f[2 * k] = __int_as_float(__double2loint(w[k]));
    f[2 * k + 1] = __int_as_float(__double2hiint(w[k]));
  }
  d20 = obb_sqdist(f, rq);
  d21 = obb_sqdist(f + 12, rq);
}

template <int NV>
__global__ void __launch_bounds__(kSoloThreads) sd_solo_kernel(const SdNode* __restrict__ nodes, const SdCen* __restrict__ cens,
                                                               const double* __restrict__ soup, SdParams prm, Desc<3> qpts,
                                                               const int32_t* __restrict__ perm, const double* __restrict__ seed,
                                                               unsigned int* __restrict__ ctr, unsigned int cap, const int32_t* __restrict__ list,
                                                               unsigned long long* __restrict__ scratch, double* __restrict__ phi,
                                                               double* __restrict__ cps, double* __restrict__ nrms, unsigned long long* __restrict__ work)
{
  constexpr unsigned FULL = 0xffffffffu;
  constexpr int NSUB = NV == 4 ? 2 : 1;
  constexpr double EPS = 1e-12;
  const unsigned lane = lane_id();
  const bool cn = prm.compute_sign != 0;
  const double window = cn ? kTieWindow : 0.0;
  unsigned long long* const buf = scratch + (size_t)(blockIdx.x * (kSoloThreads / 32) + (threadIdx.x >> 5)) * kSoloCap;
  const unsigned total = min(ctr[0], cap);
  unsigned nleaf = 0, ninner = 0;
  while(true)
  {
    unsigned item = 0;
    if(lane == 0) item = atomicAdd(ctr + 1, 1u);
    item = __shfl_sync(FULL, item, 0);
    if(item >= total) break;
    const int t = list[item];
    const int qi = perm ? perm[t] : t;
    const double qp[3] = {ld_comp<double>(qpts, 0, qi), ld_comp<double>(qpts, 1, qi), ld_comp<double>(qpts, 2, qi)};
    const V3 q {qp[0], qp[1], qp[2]};
    bool fail = false;
    // ---- stage A: the exact minimum ----
    double minSq = seed[t];  // the squared distance of SOME surface point (exact already for a sign re-run)
    int sp = 1;
    if(lane == 0) buf[0] = 0ull;  // (bound 0, root)
    __syncwarp();
    while(sp > 0)
    {
      const int n = min(sp, 32);
      sp -= n;
      const bool have = (int)lane < n;
      const unsigned long long en = have ? buf[sp + (int)lane] : 0ull;
      __syncwarp();
      int32_t push[2] = {kBarrier, kBarrier};
      float push_lb[2] = {0.f, 0.f};
      int cnt = 0;
      double best = DBL_MAX;
      if(have && (double)__uint_as_float((unsigned)(en >> 32)) <= minSq)
      {
        ++ninner;
        int32_t c0, c1;
        double d0, d1;
        solo_visit(nodes, (int32_t)(unsigned)(en & 0xffffffffull), qp, c0, c1, d0, d1);
        // farther child first: the nearer one ends up on top of the stack
        const bool swap = d0 < d1;
        const int32_t ca = swap ? c1 : c0, cb = swap ? c0 : c1;
        const double da = swap ? d1 : d0, db = swap ? d0 : d1;
        const int32_t cc[2] = {ca, cb};
        const double dd[2] = {da, db};
#pragma unroll
        for(int k = 0; k < 2; ++k)
        {
          if(!(dd[k] <= minSq)) continue;
          if(cc[k] < 0)
          {
            ++nleaf;
            best = fmin(best, leaf_min_sq<NV>(soup, q, -cc[k] - 1));
          }
          else
          {
            push[cnt] = cc[k];
            push_lb[cnt] = __double2float_rd(dd[k]);
            ++cnt;
          }
        }
      }
      int tot = 0;
      const int at = warp_excl_scan(cnt, tot);
      if(sp + tot > kSoloCap)
      {
        fail = true;
        break;
      }
      for(int k = 0; k < cnt; ++k) buf[sp + at + k] = ((unsigned long long)__float_as_uint(push_lb[k]) << 32) | (unsigned)push[k];
      sp += tot;
      minSq = fmin(minSq, warp_min(best));
      __syncwarp();
    }
    // ---- stage B: the leaves within the window of d*, in the reference's visiting order ----
    const double thr = prune_threshold_w(minSq, window);
    int32_t* la = reinterpret_cast<int32_t*>(buf);
    int32_t* lb = la + kSoloCap;
    int na = 1;
    __syncwarp();
    if(lane == 0) la[0] = 0;
    __syncwarp();
    bool any_inner = !fail;
    while(any_inner && !fail)
    {
      any_inner = false;
      int nb = 0;
      for(int base = 0; base < na; base += 32)
      {
        const int32_t e = base + (int)lane < na ? la[base + (int)lane] : kBarrier;
        int32_t o[2] = {kBarrier, kBarrier};
        int cnt = 0;
        if(e != kBarrier)
        {
          if(e < 0)
          {
            o[0] = e;
            cnt = 1;
          }
          else
          {
            ++ninner;
            int32_t c0, c1;
            double d0, d1;
            solo_visit(nodes, e, qp, c0, c1, d0, d1);
            const bool in0 = d0 <= thr, in1 = d1 <= thr;
            if(in0 && in1)
            {
              const D4* cr = reinterpret_cast<const D4*>(cens + e);
              const D4 k0 = ldg256(cr), k1 = ldg256(cr + 1);
              const double cl3[3] = {k0.x, k0.y, k0.z}, cr3[3] = {k0.w, k1.x, k1.y};
              double dl = 0.0, dr = 0.0;
#pragma unroll
              for(int d = 0; d < 3; ++d)
              {
                const double a = cl3[d] - qp[d];
                dl += a * a;
                const double b = cr3[d] - qp[d];
                dr += b * b;
              }
              const bool right_first = dl > dr;
              o[0] = right_first ? c1 : c0;
              o[1] = right_first ? c0 : c1;
              cnt = 2;
            }
            else if(in0 || in1)
            {
              o[0] = in0 ? c0 : c1;
              cnt = 1;
            }
          }
        }
        int tot = 0;
        const int at = warp_excl_scan(cnt, tot);
        if(nb + tot > kSoloCap)
        {
          fail = true;
          break;
        }
        for(int k = 0; k < cnt; ++k) lb[nb + at + k] = o[k];
        any_inner = any_inner || __ballot_sync(FULL, (cnt > 0 && o[0] >= 0) || (cnt > 1 && o[1] >= 0)) != 0u;
        nb += tot;
      }
      __syncwarp();
      int32_t* const tmp = la;
      la = lb;
      lb = tmp;
      na = nb;
    }
    // ---- replay: the state machine over the ordered leaves ----
    MinCand m;
    mincand_reset(m);
    if(fail)
    {
      if(lane == 0)
      {
        sd_ordered_query<NV>(nodes, cens, soup, q, minSq, cn, m, nleaf, ninner);
        sd_finish<NV>(soup, prm, q, m, qi, phi, cps, nrms);
      }
      __syncwarp();
      continue;
    }
    Contribs cl;
    cl.n = 0;
    for(int base = 0; base < na; base += 32)
    {
      const int nin = min(32, na - base);
      double e_sq[NSUB];
      V3 e_cp[NSUB];
      int e_loc[NSUB];
      int e_pos = 0;
#pragma unroll
      for(int u = 0; u < NSUB; ++u)
      {
        e_sq[u] = DBL_MAX;
        e_cp[u] = {0.0, 0.0, 0.0};
        e_loc[u] = 3;
      }
      if((int)lane < nin)
      {
        ++nleaf;
        e_pos = -la[base + (int)lane] - 1;
        V3 v[NV];
        load_leaf<NV>(soup, e_pos, v);
        e_cp[0] = closest_point_tri(q, v[0], v[1], v[2], e_loc[0], EPS);
        const V3 dq = v3sub(e_cp[0], q);
        e_sq[0] = v3dot(dq, dq);
        if(NV == 4 && has_fourth(v[NV - 1]))
        {
          e_cp[NSUB - 1] = closest_point_tri(q, v[0], v[2], v[NV - 1], e_loc[NSUB - 1], EPS);
          const V3 dq2 = v3sub(e_cp[NSUB - 1], q);
          e_sq[NSUB - 1] = v3dot(dq2, dq2);
        }
      }
      for(int i = 0; i < nin; ++i)
      {
        const int pos = __shfl_sync(FULL, e_pos, i);
#pragma unroll
        for(int u = 0; u < NSUB; ++u)
        {
          const double sq = shfl_f64(e_sq[u], i);
          const V3 cp {shfl_f64(e_cp[u].x, i), shfl_f64(e_cp[u].y, i), shfl_f64(e_cp[u].z, i)};
          const int loc = __shfl_sync(FULL, e_loc[u], i);
          if(sq < 1e300) apply_candidate<NV>(soup, m, cl, cp, sq, loc, pos, u, cn);  // warp-uniform: every lane keeps the same state
        }
      }
    }
    if(cl.n) m.sumN = contrib_flush<NV>(soup, cl, m.sumN);
    if(lane == 0) sd_finish<NV>(soup, prm, q, m, qi, phi, cps, nrms);
    __syncwarp();
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

}  // namespace axb
