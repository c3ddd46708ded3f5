// build.cuh -- linear-BVH construction kernels (sm_100a).
//
// Reference path replaced: lbvh::build_radix_tree + LinearBVH::buildImpl
//   spin/internal/linear_bvh/build_radix_tree.hpp:579-611, spin/policy/LinearBVH.hpp:191-269
// Reference "kernels" K1..K8 become five launches:
//   bounds_kernel     K1+K2  scale + global AABB in ONE pass over the input (reference: 1 + 3 passes)
//   morton_kernel     K1+K3+(K4 histogram)  re-scale, 32-bit Morton code, 64-bit sort key, digit histograms
//   onesweep_kernel   K4     x4 digit passes (radix_sort.cuh)
//   tree_kernel       K6+K8  Karras hierarchy, written straight into the packed traversal nodes
//   refit_kernel      K5+K7+K8  gather + re-scale leaf boxes, atomic bottom-up union written straight
//                            into the parent's child-box slot (no inner_aabbs array, no emit pass)
// The scaled leaf boxes are never materialised: BoundingBox::scale is a pure function of the
// input box (IEEE mul/add, no FMA), so recomputing it is bit-identical and saves 2 x 48 B/box.
#pragma once
#include "common.cuh"
#include "radix_sort.cuh"

namespace axb
{
// device-resident build parameters / results
template <typename T, int D>
struct BuildState
{
  unsigned long long omin[D];  // ordered-encoded running min of scaled box mins
  unsigned long long omax[D];  // ordered-encoded running max of scaled box maxs
  T bmin[D];                   // decoded bounds (getBounds())
  T bmax[D];
  T inv_extent[D];             // build_radix_tree.hpp:160-163
  uint32_t agglo_mismatch;     // agglo_kernel: a >2^24-leaf node whose float32 split search differs (see there)
  uint32_t agglo_open_count;   // subtrees the block-local pass left for agglo_upper_kernel
};

template <typename T, int D>
__global__ void init_state_kernel(BuildState<T, D>* st)
{
  if(threadIdx.x == 0 && blockIdx.x == 0)
  {
    for(int d = 0; d < D; ++d)
    {
      st->omin[d] = f64_to_ordered((double)Lim<T>::max());
      st->omax[d] = f64_to_ordered((double)Lim<T>::lowest());
    }
    st->agglo_mismatch = 0u;
    st->agglo_open_count = 0u;
  }
}

// K1+K2: transform_boxes (:85-99) fused with reduce (:102-143).  Only valid boxes contribute
// (the SEQ path's addBox, BoundingBox.hpp:487-508); min/max are exact so order is irrelevant.
template <typename T, int D>
__global__ void __launch_bounds__(256) bounds_kernel(Desc<2 * D> boxes, int n, int n_real, T half_scale, BuildState<T, D>* st)
{
  T mn[D], mx[D];
#pragma unroll
  for(int d = 0; d < D; ++d)
  {
    mn[d] = Lim<T>::max();
    mx[d] = Lim<T>::lowest();
  }
  for(int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
  {
    Box<T, D> b;
    if(i < n_real)
      b = load_box<T, D>(boxes, i);
    else
      box_clear(b);  // N<=1 padding, spin/BVH.hpp:439-464
    if(box_valid(b))
    {
      box_scale(b, half_scale);
#pragma unroll
      for(int d = 0; d < D; ++d)
      {
        mn[d] = b.lo[d] < mn[d] ? b.lo[d] : mn[d];
        mx[d] = b.hi[d] > mx[d] ? b.hi[d] : mx[d];
      }
    }
  }
  // warp shuffle reduction, then one atomic pair per warp and dimension
#pragma unroll
  for(int d = 0; d < D; ++d)
  {
#pragma unroll
    for(int o = 16; o > 0; o >>= 1)
    {
      const T a = __shfl_xor_sync(0xffffffffu, mn[d], o);
      const T b = __shfl_xor_sync(0xffffffffu, mx[d], o);
      mn[d] = a < mn[d] ? a : mn[d];
      mx[d] = b > mx[d] ? b : mx[d];
    }
  }
  __shared__ T smn[8][D], smx[8][D];
  const int warp = threadIdx.x >> 5;
  if(lane_id() == 0)
  {
#pragma unroll
    for(int d = 0; d < D; ++d)
    {
      smn[warp][d] = mn[d];
      smx[warp][d] = mx[d];
    }
  }
  __syncthreads();
  if(threadIdx.x < D)
  {
    const int d = threadIdx.x;
    T a = smn[0][d], b = smx[0][d];
    for(int w = 1; w < (int)(blockDim.x >> 5); ++w)
    {
      a = smn[w][d] < a ? smn[w][d] : a;
      b = smx[w][d] > b ? smx[w][d] : b;
    }
    atomicMin(&st->omin[d], f64_to_ordered((double)a));
    atomicMax(&st->omax[d], f64_to_ordered((double)b));
  }
}

// get_mcodes prologue (:154-163): decode bounds, inv_extent = isNearlyEqual(extent,0) ? 0 : 1/extent
template <typename T, int D>
__global__ void finalize_bounds_kernel(BuildState<T, D>* st)
{
  if(threadIdx.x == 0 && blockIdx.x == 0)
  {
    for(int d = 0; d < D; ++d)
    {
      const T lo = (T)ordered_to_f64(st->omin[d]);
      const T hi = (T)ordered_to_f64(st->omax[d]);
      st->bmin[d] = lo;
      st->bmax[d] = hi;
      const T ext = hi - lo;
      const T diff = ext - (T)0;
      st->inv_extent[d] = ((diff < 0 ? -diff : diff) <= (T)1.0e-8) ? (T)0 : (T)1 / ext;  // IEEE div (no fast-math)
    }
  }
}

// spin/MortonIndex.hpp:151-159 bit spreading with the int32 masks (:226-246, :377-393)
__device__ __forceinline__ uint32_t spread_bits_2d(uint32_t x)
{
  x &= 0x0000FFFFu;
  x = (x | (x << 8)) & 0x00FF00FFu;
  x = (x | (x << 4)) & 0x0F0F0F0Fu;
  x = (x | (x << 2)) & 0x33333333u;
  x = (x | (x << 1)) & 0x55555555u;
  return x;
}
__device__ __forceinline__ uint32_t spread_bits_3d(uint32_t x)
{
  x &= 0x0000FFFFu;
  x = (x | (x << 16)) & 0xFF0000FFu;
  x = (x | (x << 8)) & 0x0F00F00Fu;
  x = (x | (x << 4)) & 0xC30C30C3u;
  x = (x | (x << 2)) & 0x49249249u;
  return x;
}

// morton32_encode (:48-65): q = int32(fmin(fmax(c * 2^bits, 0), 2^bits - 1)), truncation
template <typename T, int D>
__device__ __forceinline__ uint32_t morton32(const T* c01)
{
  constexpr int bits = 32 / D;
  constexpr T to_int = (T)(1 << bits);
  constexpr T ceil_v = to_int - (T)1;
  uint32_t q[D];
#pragma unroll
  for(int d = 0; d < D; ++d)
  {
    const T v = fmin(fmax(c01[d] * to_int, (T)0), ceil_v);
    q[d] = (uint32_t)(int32_t)v;  // cvt.rzi: same truncation as the C cast
  }
  if(D == 2) return spread_bits_2d(q[0]) | (spread_bits_2d(q[1]) << 1);
  return spread_bits_3d(q[0]) | (spread_bits_3d(q[1]) << 1) | (spread_bits_3d(q[D - 1]) << 2);
}

// K3 (+ histogram of K4): one thread per box; key = (code << 32) | i
template <typename T, int D>
__global__ void __launch_bounds__(256) morton_kernel(Desc<2 * D> boxes, int n, int n_real, T half_scale, const BuildState<T, D>* __restrict__ st,
                                                      unsigned long long* __restrict__ keys, uint32_t* __restrict__ ghist)
{
  __shared__ uint32_t sh[rsort::MAX_PASSES * rsort::RADIX];
  for(int i = threadIdx.x; i < rsort::MAX_PASSES * rsort::RADIX; i += blockDim.x) sh[i] = 0;
  __syncthreads();
  T mn[D], inv[D];
#pragma unroll
  for(int d = 0; d < D; ++d)
  {
    mn[d] = st->bmin[d];
    inv[d] = st->inv_extent[d];
  }
  for(int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
  {
    Box<T, D> b;
    if(i < n_real)
      b = load_box<T, D>(boxes, i);
    else
      box_clear(b);
    box_scale(b, half_scale);
    T c[D];
#pragma unroll
    for(int d = 0; d < D; ++d)
    {
      const T cen = static_cast<T>(0.5 * (b.lo[d] + b.hi[d]));  // getCentroid, Point.hpp:279-290
      c[d] = (cen - mn[d]) * inv[d];
    }
    const uint32_t code = morton32<T, D>(c);
    keys[i] = ((unsigned long long)code << 32) | (unsigned long long)(uint32_t)i;
#pragma unroll
    for(int p = 0; p < rsort::MAX_PASSES; ++p) atomicAdd(&sh[p * rsort::RADIX + ((code >> (p * 8)) & 255u)], 1u);
  }
  __syncthreads();
  for(int i = threadIdx.x; i < rsort::MAX_PASSES * rsort::RADIX; i += blockDim.x)
    if(sh[i]) atomicAdd(&ghist[i], sh[i]);
}

// delta() (:265-287) on the sorted 64-bit keys: code in the high word; ties broken by the
// sorted position.
__device__ __forceinline__ int karras_delta(const unsigned long long* __restrict__ keys, int a, uint32_t acode, int b, int inner_size)
{
  if(b < 0 || b > inner_size) return -1;
  const uint32_t bcode = (uint32_t)(__ldg(keys + b) >> 32);
  const uint32_t x = acode ^ bcode;
  if(x == 0) return 32 + __clz((int)((uint32_t)a ^ (uint32_t)b));
  return __clz((int)x);
}

// K6 build_tree (:290-384) fused with the child-id half of the emit step (LinearBVH.hpp:231-262):
// node i gets child ids in traversal encoding; each child records (parent<<1)|side.
template <typename T, int D>
__global__ void __launch_bounds__(256) tree_kernel(const unsigned long long* __restrict__ keys, int n, Node<T, D>* __restrict__ nodes,
                                                    int32_t* __restrict__ leaf_parent, int2* __restrict__ node_range)
{
  const int inner_size = n - 1;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if(i >= inner_size) return;
  const uint32_t icode = (uint32_t)(__ldg(keys + i) >> 32);
  auto dl = [&](int b) { return karras_delta(keys, i, icode, b, inner_size); };

  const int d = (dl(i + 1) - dl(i - 1)) < 0 ? -1 : 1;
  const int min_delta = dl(i - d);
  int lmax = 2;
  while(dl(i + lmax * d) > min_delta) lmax *= 2;
  int l = 0;
  for(int t = lmax / 2; t >= 1; t /= 2)
    if(dl(i + (l + t) * d) > min_delta) l += t;
  const int j = i + l * d;
  const int delta_node = dl(j);
  int s = 0;
  // "t = (int32)ceil(float32(l) / div_factor)" with div_factor a FloatType doubling from 2 (:336-339):
  // l is first rounded to float32, then divided in FloatType.  Reproduced verbatim.
  T div_factor = (T)2;
  const T lf = (T)(float)l;
  for(int t = (int)ceil(lf / div_factor);; div_factor *= 2, t = (int)ceil(lf / div_factor))
  {
    if(dl(i + (s + t) * d) > delta_node) s += t;
    if(t == 1) break;
  }
  const int split = i + s * d + (d < 0 ? d : 0);
  const int lo = i < j ? i : j;
  const int hi = i < j ? j : i;
  int lchild, rchild;
  if(lo == split)
  {
    leaf_parent[split] = (i << 1);
    lchild = -(split + 1);
  }
  else
  {
    nodes[split].parent = (i << 1);
    lchild = split;
  }
  if(hi == split + 1)
  {
    leaf_parent[split + 1] = (i << 1) | 1;
    rchild = -(split + 2);
  }
  else
  {
    nodes[split + 1].parent = (i << 1) | 1;
    rchild = split + 1;
  }
  node_range[i] = make_int2(lo, hi);  // sorted-leaf range covered by inner node i (inclusive)
  nodes[i].child[0] = lchild;
  nodes[i].child[1] = rchild;
  nodes[i].counter = 0u;
  if(i == 0) nodes[0].parent = -1;
}

template <typename T, int D>
__device__ __forceinline__ void store_box_cg(Box<T, D>* dst, const Box<T, D>& b)
{
  // write-through to L2 (st.cg): the sibling thread reads it from another SM
  T* p = reinterpret_cast<T*>(dst);
#pragma unroll
  for(int k = 0; k < D; ++k)
  {
    __stcg(p + k, b.lo[k]);
    __stcg(p + D + k, b.hi[k]);
  }
}
template <typename T, int D>
__device__ __forceinline__ Box<T, D> load_box_cg(const Box<T, D>* src)
{
  Box<T, D> b;
  const T* p = reinterpret_cast<const T*>(src);
#pragma unroll
  for(int k = 0; k < D; ++k)
  {
    b.lo[k] = __ldcg(p + k);
    b.hi[k] = __ldcg(p + D + k);
  }
  return b;
}

// K5+K7+K8: reorder (:199-219) + propagate_aabbs (:505-576) + the box half of emit.
// One thread per leaf in sorted order.  A thread arriving at inner node p stores its box in
// p's child slot (the final traversal layout) and bumps p's counter with an acq_rel atomic; the
// first arrival retires, the second reads the sibling slot, unions and climbs.  The single
// acquire-release RMW replaces the reference's atomicExch-store / volatile-poll workaround
// (:391-484); unions are exact min/max, so the result does not depend on arrival order.
template <typename T, int D>
__global__ void __launch_bounds__(256) refit_kernel(Desc<2 * D> boxes, int n, int n_real, T half_scale,
                                                     const unsigned long long* __restrict__ keys, const int32_t* __restrict__ leaf_parent,
                                                     Node<T, D>* nodes, int32_t* __restrict__ leaf_nodes)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if(i >= n) return;
  const int leaf = (int)(uint32_t)(__ldg(keys + i) & 0xffffffffull);
  leaf_nodes[i] = leaf;
  Box<T, D> aabb;
  if(leaf < n_real)
    aabb = load_box<T, D>(boxes, leaf);
  else
    box_clear(aabb);
  box_scale(aabb, half_scale);
  int link = leaf_parent[i];
  while(link != -1)
  {
    const int p = link >> 1;
    const int side = link & 1;
    store_box_cg(&nodes[p].box[side], aabb);
    // one release RMW per arrival publishes the box just stored; the second arrival reads the sibling
    // slot with L2 loads (ld.cg) that are control-dependent on the RMW result, so no acquire fence is needed
    uint32_t old;
    asm volatile("atom.add.release.gpu.global.u32 %0, [%1], 1;" : "=r"(old) : "l"(&nodes[p].counter) : "memory");
    if(old == 0u) return;  // first arrival retires (:547-551)
    const Box<T, D> other = load_box_cg(&nodes[p].box[side ^ 1]);
    box_add(aabb, other);  // aabb.addBox(other_aabb) (:567)
    link = nodes[p].parent;
  }
}

//------------------------------------------------------------------------------------------
// Fused hierarchy + refit ("agglomerative" bottom-up build, Apetrei 2014), the default path.
//
// The reference emits the Karras tree top-down (build_tree, :290-384: one binary search per inner
// node over the sorted codes) and then refits bottom-up (propagate_aabbs, :505-576).  Both are
// functions of the adjacent-key deltas only: a node covering sorted leaves [l, r] is the left child
// of the node that splits at r when delta(r, r+1) > delta(l-1, l), else the right child of the node
// that splits at l-1 (ties are impossible: keys are distinct once the sorted position breaks code
// ties, :276-279).  Karras numbers a left child by its split (= its range's right end) and a right
// child by split+1 (= its range's left end), so a node's reference index, its children's ids and its
// range are all known the moment its two children have met -- no top-down search is needed, and
// the tree comes out bit-identical (checked against the reference-order search for ranges above 2^24
// leaves, where build_tree's float32(l) rounding could in principle pick another split; a mismatch
// raises BuildState::agglo_mismatch and the host re-runs tree_kernel + refit_kernel).
//
// One thread per sorted leaf, B leaves per block.  Children whose parent's range lies inside the
// block meet through shared memory (a slot per child index, an arrival flag per split); the few
// block-straddling nodes meet through a global slot array with a release/acquire RMW.  Each finished
// node is written once, complete: 96 B of child boxes + child ids + range; parents of the children.
// HBM traffic per leaf: 8 B key + 48 B gathered box in, 128 B node record + 4 B leaf id + 12 B links out.
//------------------------------------------------------------------------------------------
template <typename T, int D>
struct alignas(16) AggloSlot
{
  Box<T, D> box;
  int32_t end;  // far end of the arriving child's leaf range
};

// a finished block-local subtree whose parent covers leaves of another block: input of agglo_upper_kernel
template <typename T, int D>
struct alignas(8) AggloOpen
{
  Box<T, D> box;
  int32_t lo, hi;
};

__device__ __forceinline__ int adj_delta(unsigned long long kj, unsigned long long kj1, int j)
{
  // delta(j, j+1) of :265-287 on (code << 32 | sorted position)
  const unsigned long long a = (kj & 0xffffffff00000000ull) | (unsigned long long)(uint32_t)j;
  const unsigned long long b = (kj1 & 0xffffffff00000000ull) | (unsigned long long)(uint32_t)(j + 1);
  return __clzll((long long)(a ^ b));
}

// build_tree's split search (:333-348) for a node with index i, direction d, length l whose true
// split is `gam`: delta(i, i+(s+t)d) > delta_node holds exactly while i+(s+t)d stays on i's side.
template <typename T>
__device__ __noinline__ bool reference_split_agrees(int lo, int hi, int gam, bool index_is_lo)
{
  const int d = index_is_lo ? 1 : -1;
  const int i = index_is_lo ? lo : hi;
  const int l = hi - lo;
  int s = 0;
  T div_factor = (T)2;
  const T lf = (T)(float)l;
  for(int t = (int)ceil(lf / div_factor);; div_factor *= 2, t = (int)ceil(lf / div_factor))
  {
    const long long x = (long long)i + (long long)(s + t) * d;
    const bool same_side = d > 0 ? (x <= gam) : (x >= gam + 1);
    if(same_side) s += t;
    if(t == 1) break;
  }
  return i + s * d + (d < 0 ? d : 0) == gam;
}

// The two children of the node that splits at `gam` have met: [l, r] is the node's range, dl / dr the deltas just
// outside it, `box` the arriving child's box (the left one iff is_left), `other` its sibling's.  Writes the node's
// record under its reference index, its range, and the parent links of the two children.  Returns true for the root.
template <typename T, int D>
__device__ __forceinline__ bool agglo_finish_node(Node<T, D>* nodes, int32_t* leaf_parent, int2* __restrict__ node_range, uint32_t* mismatch, int n,
                                                  int l, int r, int gam, bool is_left, int dl, int dr, const Box<T, D>& box,
                                                  const Box<T, D>& other)
{
  const bool root = (l == 0 && r == n - 1);
  const bool p_is_left = dr > dl;
  const int P = root ? 0 : (p_is_left ? r : l);
  if(r - l > (1 << 24))
    if(!reference_split_agrees<T>(l, r, gam, root || !p_is_left)) atomicOr(mismatch, 1u);
  const int lc = l == gam ? -(gam + 1) : gam;
  const int rc = r == gam + 1 ? -(gam + 2) : gam + 1;
  Node<T, D>* nd = nodes + P;
  nd->box[is_left ? 0 : 1] = box;
  nd->box[is_left ? 1 : 0] = other;
  *reinterpret_cast<int2*>(nd->child) = make_int2(lc, rc);
  node_range[P] = make_int2(l, r);
  // parent links of the two children (leaf_parent for leaves, Node::parent for inner nodes)
  *(lc < 0 ? leaf_parent + gam : &nodes[gam].parent) = P << 1;
  *(rc < 0 ? leaf_parent + gam + 1 : &nodes[gam + 1].parent) = (P << 1) | 1;
  if(root) nd->parent = -1;
  return root;
}

template <typename T, int D, int B>
__global__ void __launch_bounds__(B, 1024 / B) agglo_kernel(Desc<2 * D> boxes, int n, int n_real, T half_scale,
                                                   const unsigned long long* __restrict__ keys, Node<T, D>* nodes,
                                                   int32_t* __restrict__ leaf_nodes, int32_t* leaf_parent, int2* __restrict__ node_range,
                                                   AggloSlot<T, D>* gslot, uint32_t* gflag, uint32_t* mismatch, AggloOpen<T, D>* open_list,
                                                   unsigned int* open_count)
{
  __shared__ unsigned long long s_key[B + 2];  // keys of positions L-1 .. L+B
  __shared__ int s_delta[B + 1];               // s_delta[k] = delta(L-1+k, L+k)
  __shared__ int s_pre[B + 1];                 // min(s_delta[0..k])
  __shared__ int s_suf[B + 2];                 // min(s_delta[k..B]);  s_suf[B+1] = +inf
  __shared__ int s_wmin[2][B / 32];
  __shared__ Box<T, D> s_box[B];               // hand-over slot of the child with index L+k
  __shared__ int s_end[B];
  __shared__ uint32_t s_flag[B];               // arrivals at split L+k; bit 31 = "parent range inside this block"

  const int tid = threadIdx.x;
  const int L = blockIdx.x * B;
  const int g = L + tid;
  const int lane = tid & 31, warp = tid >> 5;

  s_key[tid + 1] = g < n ? __ldg(keys + g) : 0ull;
  if(tid == 0)
  {
    s_key[0] = L > 0 ? __ldg(keys + L - 1) : 0ull;
    s_key[B + 1] = L + B < n ? __ldg(keys + L + B) : 0ull;
  }
  __syncthreads();
  const int dme = g < n - 1 ? adj_delta(s_key[tid + 1], s_key[tid + 2], g) : -1;
  const int d0 = L > 0 ? adj_delta(s_key[0], s_key[1], L - 1) : -1;
  s_delta[tid + 1] = dme;
  if(tid == 0) s_delta[0] = d0;
  // prefix minima over s_delta[0..B] (thread t owns element t+1), suffix minima (thread t owns element t+1 as well)
  int pv = dme, sv = dme;
#pragma unroll
  for(int o = 1; o < 32; o <<= 1)
  {
    const int a = __shfl_up_sync(0xffffffffu, pv, o);
    const int b = __shfl_down_sync(0xffffffffu, sv, o);
    if(lane >= o) pv = min(pv, a);
    if(lane + o < 32) sv = min(sv, b);
  }
  if(lane == 31) s_wmin[0][warp] = pv;
  if(lane == 0) s_wmin[1][warp] = sv;
  __syncthreads();
  if(warp == 0)
  {
    // exclusive prefix / suffix minima over the per-warp minima (at most 32 warps)
    int a = lane < B / 32 ? s_wmin[0][lane] : 0x7fffffff;
    int b = lane < B / 32 ? s_wmin[1][lane] : 0x7fffffff;
#pragma unroll
    for(int o = 1; o < 32; o <<= 1)
    {
      const int ua = __shfl_up_sync(0xffffffffu, a, o);
      const int ub = __shfl_down_sync(0xffffffffu, b, o);
      if(lane >= o) a = min(a, ua);
      if(lane + o < 32) b = min(b, ub);
    }
    int ea = __shfl_up_sync(0xffffffffu, a, 1);
    int eb = __shfl_down_sync(0xffffffffu, b, 1);
    ea = lane == 0 ? d0 : min(ea, d0);
    if(lane == 31) eb = 0x7fffffff;
    if(lane < B / 32)
    {
      s_wmin[0][lane] = ea;
      s_wmin[1][lane] = eb;
    }
  }
  __syncthreads();
  s_pre[tid + 1] = min(pv, s_wmin[0][warp]);
  s_suf[tid + 1] = min(sv, s_wmin[1][warp]);
  if(tid == 0)
  {
    s_pre[0] = d0;
    s_suf[B + 1] = 0x7fffffff;
  }
  __syncthreads();
  {
    // split k = tid (between leaves L+k and L+k+1): its node's range stays inside the block iff a
    // smaller delta exists on both sides within the block's window
    const bool local = (tid < B - 1) && (g < n - 1) && s_pre[tid] < dme && s_suf[tid + 2] < dme;
    s_flag[tid] = local ? 0x80000000u : 0u;
  }
  __syncthreads();
  if(g >= n) return;

  auto dlt = [&](int j) -> int {
    const int k = j - (L - 1);
    if(k >= 0 && k <= B) return s_delta[k];  // s_delta is -1 at and beyond the array ends
    if(j < 0 || j >= n - 1) return -1;
    return adj_delta(__ldg(keys + j), __ldg(keys + j + 1), j);
  };

  const int leaf = (int)(uint32_t)(s_key[tid + 1] & 0xffffffffull);
  leaf_nodes[g] = leaf;
  Box<T, D> box;
  if(leaf < n_real)
  {
    box = load_box<T, D>(boxes, leaf);
    if(!box_valid(box))
    {
      // the unions below are plain min/max, which equals BoundingBox::addBox (:487-508) for valid boxes and
      // for the canonical invalid box (max(), lowest()); any other invalid input box goes to the legacy path
      bool canonical = true;
#pragma unroll
      for(int d = 0; d < D; ++d) canonical = canonical && box.lo[d] == Lim<T>::max() && box.hi[d] == Lim<T>::lowest();
      if(!canonical) atomicOr(mismatch, 2u);
    }
  }
  else
    box_clear(box);
  box_scale(box, half_scale);

  int l = g, r = g;
  int dl = s_delta[tid], dr = dme;
  for(;;)
  {
    const bool is_left = dr > dl;
    const int gam = is_left ? r : l - 1;
    const int self = is_left ? gam : gam + 1;  // this child's reference index
    const int sib = is_left ? gam + 1 : gam;
    const int k = gam - L;
    Box<T, D> other;
    int oend;
    if(k >= 0 && k < B - 1 && (s_flag[k] & 0x80000000u))
    {
      uint32_t old;
#ifndef AXB_AGGLO_RELAXED_SMEM
      s_box[self - L] = box;
      s_end[self - L] = is_left ? l : r;
      // block-scope acq_rel RMW: publishes this child's slot, acquires the sibling's
      asm volatile("atom.acq_rel.cta.shared.add.u32 %0, [%1], 1;"
                   : "=r"(old)
                   : "r"((uint32_t)__cvta_generic_to_shared(&s_flag[k]))
                   : "memory");
#else
      // -DAXB_AGGLO_RELAXED_SMEM: publish through shared memory WITHOUT a fence (relaxed RMW, volatile accesses; relies
      // on shared memory being one in-order pipeline per SM).  Measured on B200: bit-identical trees at 1-20 M boxes
      // and no gain over the fenced form (the membar stall of profiles/r1o came from the gpu-scope RMW of the
      // cross-block path, now moved to agglo_upper_kernel), so the fenced, formally correct form is the default.
      {
        volatile T* sb = reinterpret_cast<volatile T*>(&s_box[self - L]);
#pragma unroll
        for(int d = 0; d < D; ++d)
        {
          sb[d] = box.lo[d];
          sb[D + d] = box.hi[d];
        }
        *reinterpret_cast<volatile int*>(&s_end[self - L]) = is_left ? l : r;
      }
      asm volatile("atom.relaxed.cta.shared.add.u32 %0, [%1], 1;"
                   : "=r"(old)
                   : "r"((uint32_t)__cvta_generic_to_shared(&s_flag[k]))
                   : "memory");
#endif
      if((old & 1u) == 0u) return;  // first arrival retires (:547-551)
#ifndef AXB_AGGLO_RELAXED_SMEM
      other = s_box[sib - L];
      oend = s_end[sib - L];
#else
      {
        const volatile T* ob = reinterpret_cast<const volatile T*>(&s_box[sib - L]);
#pragma unroll
        for(int d = 0; d < D; ++d)
        {
          other.lo[d] = ob[d];
          other.hi[d] = ob[D + d];
        }
        oend = *reinterpret_cast<const volatile int*>(&s_end[sib - L]);
      }
#endif
    }
    else
    {
#ifdef AXB_AGGLO_SINGLE_PASS
      store_box_cg(&gslot[self].box, box);
      __stcg(&gslot[self].end, is_left ? l : r);
      uint32_t old;
      asm volatile("atom.add.acq_rel.gpu.global.u32 %0, [%1], 1;" : "=r"(old) : "l"(gflag + gam) : "memory");
      if(old == 0u) return;
      other = load_box_cg(&gslot[sib].box);
      oend = __ldcg(&gslot[sib].end);
#else
      // The parent's range leaves the block: park this subtree in the open list and retire.  agglo_upper_kernel
      // finishes the upper tree from the open entries.  (Meeting across blocks needs a gpu-scope release/acquire RMW,
      // i.e. a MEMBAR.ALL.GPU that also waits for the node records this thread has just written; done here it stalled
      // whole warps for one or two lanes -- a third of the kernel's stall cycles in profiles/r1o.)
      const unsigned e = atomicAdd(open_count, 1u);
      AggloOpen<T, D> o;
      o.box = box;
      o.lo = l;
      o.hi = r;
      open_list[e] = o;
      return;
#endif
    }
    // the parent covers [l, oend] or [oend, r]: only one of the two outer deltas changes
    if(is_left)
    {
      r = oend;
      dr = dlt(r);
    }
    else
    {
      l = oend;
      dl = dlt(l - 1);
    }
    if(agglo_finish_node<T, D>(nodes, leaf_parent, node_range, mismatch, n, l, r, gam, is_left, dl, dr, box, other)) return;  // root
#pragma unroll
    for(int d = 0; d < D; ++d)
    {
      box.lo[d] = other.lo[d] < box.lo[d] ? other.lo[d] : box.lo[d];
      box.hi[d] = other.hi[d] > box.hi[d] ? other.hi[d] : box.hi[d];
    }
  }
}

// Second pass of the fused build: the upper tree.  One thread per open entry (grid-stride: the count is only known
// on the device); every meeting goes through the global slots with a gpu-scope acq_rel RMW.  All lanes of a warp are
// in this path, so the fences are amortised over 32 merges instead of stalling a warp for one.
template <typename T, int D>
__global__ void __launch_bounds__(128) agglo_upper_kernel(int n, const unsigned long long* __restrict__ keys, Node<T, D>* nodes,
                                                           int32_t* leaf_parent, int2* __restrict__ node_range, AggloSlot<T, D>* gslot,
                                                           uint32_t* gflag, uint32_t* mismatch, const AggloOpen<T, D>* __restrict__ open_list,
                                                           const unsigned int* __restrict__ open_count)
{
  auto dlt = [&](int j) -> int {
    if(j < 0 || j >= n - 1) return -1;
    return adj_delta(__ldg(keys + j), __ldg(keys + j + 1), j);
  };
  const unsigned count = *open_count;
  for(unsigned e = blockIdx.x * blockDim.x + threadIdx.x; e < count; e += gridDim.x * blockDim.x)
  {
    Box<T, D> box = open_list[e].box;
    int l = open_list[e].lo, r = open_list[e].hi;
    int dl = dlt(l - 1), dr = dlt(r);
    for(;;)
    {
      const bool is_left = dr > dl;
      const int gam = is_left ? r : l - 1;
      const int self = is_left ? gam : gam + 1;
      const int sib = is_left ? gam + 1 : gam;
      store_box_cg(&gslot[self].box, box);
      __stcg(&gslot[self].end, is_left ? l : r);
      uint32_t old;
      asm volatile("atom.add.acq_rel.gpu.global.u32 %0, [%1], 1;" : "=r"(old) : "l"(gflag + gam) : "memory");
      if(old == 0u) break;  // first arrival: on to this thread's next open entry
      const Box<T, D> other = load_box_cg(&gslot[sib].box);
      const int oend = __ldcg(&gslot[sib].end);
      if(is_left)
      {
        r = oend;
        dr = dlt(r);
      }
      else
      {
        l = oend;
        dl = dlt(l - 1);
      }
      if(agglo_finish_node<T, D>(nodes, leaf_parent, node_range, mismatch, n, l, r, gam, is_left, dl, dr, box, other)) break;  // root
#pragma unroll
      for(int d = 0; d < D; ++d)
      {
        box.lo[d] = other.lo[d] < box.lo[d] ? other.lo[d] : box.lo[d];
        box.hi[d] = other.hi[d] > box.hi[d] ? other.hi[d] : box.hi[d];
      }
    }
  }
}

// Reference-layout view for getTraverser() / parity checks (LinearBVH.hpp:226-263):
// inner_nodes[2i+s] = child box, inner_node_children[2i+s] = 2*child | -(pos+1)
template <typename T, int D>
__global__ void __launch_bounds__(256) export_kernel(const Node<T, D>* __restrict__ nodes, int inner_size, Box<T, D>* __restrict__ inner_nodes,
                                                      int32_t* __restrict__ inner_children)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if(i >= inner_size) return;
  const Node<T, D> nd = nodes[i];
#pragma unroll
  for(int s = 0; s < 2; ++s)
  {
    inner_nodes[2 * i + s] = nd.box[s];
    inner_children[2 * i + s] = nd.child[s] >= 0 ? nd.child[s] * 2 : nd.child[s];
  }
}

__global__ void extract_mcodes_kernel(const unsigned long long* __restrict__ keys, int n, uint32_t* __restrict__ mcodes)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if(i < n) mcodes[i] = (uint32_t)(keys[i] >> 32);
}

}  // namespace axb
