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
