// common.cuh -- shared device/host helpers for the sm_100a BVH / SignedDistance kernels.
//
// Arithmetic parity: this translation unit set is compiled with -fmad=false, so every
// a*b+c below is two IEEE-rounded operations exactly as in the reference's x86-64 Release
// build (SURVEY.md "quirks": no FMA contraction).  Places that deliberately use FMA call
// fma() explicitly and are documented as value-insensitive filters.
#pragma once
#include <cuda_runtime.h>
#include <cfloat>
#include <cstdint>
#include <cstdio>
#include <string>

#include "../../include/axb200.h"

namespace axb
{
//------------------------------------------------------------------------------------------
// error handling
//------------------------------------------------------------------------------------------
void set_last_error(const std::string& msg);

#define AXB_CUDA_TRY(expr)                                                                           \
  do                                                                                                 \
  {                                                                                                  \
    cudaError_t _e = (expr);                                                                         \
    if(_e != cudaSuccess)                                                                            \
    {                                                                                                \
      ::axb::set_last_error(std::string(#expr) + ": " + cudaGetErrorString(_e) + " (" + __FILE__ + ":" + \
                            std::to_string(__LINE__) + ")");                                         \
      return AXB_ERR_CUDA;                                                                           \
    }                                                                                                \
  } while(0)

#define AXB_TRY(expr)              \
  do                               \
  {                                \
    int _s = (expr);               \
    if(_s != AXB_OK) return _s;    \
  } while(0)

//------------------------------------------------------------------------------------------
// strided component access ("Indexable" of the reference, spin/BVH.hpp:143-171)
//------------------------------------------------------------------------------------------
template <int NC>
struct Desc
{
  const char* comp[NC];
  long long stride;
};

template <typename T, int NC>
__device__ __forceinline__ T ld_comp(const Desc<NC>& d, int c, long long i)
{
  return __ldg(reinterpret_cast<const T*>(d.comp[c] + i * d.stride));
}

//------------------------------------------------------------------------------------------
// boxes.  Invalid box = (max(), lowest())  -- primal/geometry/BoundingBox.hpp:72-73
//------------------------------------------------------------------------------------------
template <typename T>
struct Lim;
template <>
struct Lim<double>
{
  static __host__ __device__ constexpr double max() { return DBL_MAX; }
  static __host__ __device__ constexpr double lowest() { return -DBL_MAX; }
  static __host__ __device__ constexpr double min() { return DBL_MIN; }
};
template <>
struct Lim<float>
{
  static __host__ __device__ constexpr float max() { return FLT_MAX; }
  static __host__ __device__ constexpr float lowest() { return -FLT_MAX; }
  static __host__ __device__ constexpr float min() { return FLT_MIN; }
};

template <typename T, int D>
struct Box
{
  T lo[D];
  T hi[D];
};

template <typename T, int D>
__device__ __forceinline__ bool box_valid(const Box<T, D>& b)
{
  bool ok = true;
#pragma unroll
  for(int d = 0; d < D; ++d) ok = ok && !(b.lo[d] > b.hi[d]);
  return ok;
}

template <typename T, int D>
__device__ __forceinline__ void box_clear(Box<T, D>& b)
{
#pragma unroll
  for(int d = 0; d < D; ++d)
  {
    b.lo[d] = Lim<T>::max();
    b.hi[d] = Lim<T>::lowest();
  }
}

// BoundingBox::scale (primal/geometry/BoundingBox.hpp:548-561): mid -/+ T(s*0.5)*(max-min),
// then checkAndFixBounds.  half_scale = T(scaleFactor*0.5) is computed once on the host.
template <typename T, int D>
__device__ __forceinline__ void box_scale(Box<T, D>& b, T half_scale)
{
  if(!box_valid(b)) return;
#pragma unroll
  for(int d = 0; d < D; ++d)
  {
    const T mid = static_cast<T>(0.5 * (b.lo[d] + b.hi[d]));
    const T r = half_scale * (b.hi[d] - b.lo[d]);
    T lo = mid - r;
    T hi = mid + r;
    if(lo > hi)
    {
      const T t = lo;
      lo = hi;
      hi = t;
    }
    b.lo[d] = lo;
    b.hi[d] = hi;
  }
}

template <typename T, int D>
__device__ __forceinline__ Box<T, D> load_box(const Desc<2 * D>& d, long long i)
{
  Box<T, D> b;
#pragma unroll
  for(int k = 0; k < D; ++k)
  {
    b.lo[k] = ld_comp<T>(d, k, i);
    b.hi[k] = ld_comp<T>(d, D + k, i);
  }
  return b;
}

// BoundingBox::addBox (primal/geometry/BoundingBox.hpp:487-508) for the union of a valid
// running box with another box: an invalid `o` leaves `self` untouched, an invalid `self`
// takes `o`.
template <typename T, int D>
__device__ __forceinline__ void box_add(Box<T, D>& self, const Box<T, D>& o)
{
  if(box_valid(self))
  {
    if(box_valid(o))
    {
#pragma unroll
      for(int d = 0; d < D; ++d)
      {
        // addPoint(min) then addPoint(max), :463-484
        if(o.lo[d] < self.lo[d]) self.lo[d] = o.lo[d];
        if(o.lo[d] > self.hi[d]) self.hi[d] = o.lo[d];
        if(o.hi[d] < self.lo[d]) self.lo[d] = o.hi[d];
        if(o.hi[d] > self.hi[d]) self.hi[d] = o.hi[d];
      }
    }
  }
  else
  {
    self = o;
  }
}

//------------------------------------------------------------------------------------------
// packed traversal node: both child boxes + both child ids in one aligned record, so one
// inner-node visit is one (3-D double: 128 B = one L2 line) contiguous read.
//   child id >= 0 : inner child, index of its node record
//   child id <  0 : leaf, -(sorted_pos+1)              (policy/LinearBVH.hpp:235,248)
// `parent` and `counter` are build-time fields (refit), unused by traversal.
//------------------------------------------------------------------------------------------
template <typename T, int D>
struct alignas(32) Node
{
  Box<T, D> box[2];
  int32_t child[2];
  int32_t parent;   // (parent_index << 1) | side, -1 for the root
  uint32_t counter; // refit arrival counter
};
static_assert(sizeof(Node<double, 3>) == 128, "3-D double node must be one 128-byte line");
static_assert(sizeof(Node<double, 2>) == 96, "2-D double node");

//------------------------------------------------------------------------------------------
// order-preserving double <-> uint64 for atomicMin/atomicMax on doubles
//------------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ unsigned long long f64_to_ordered(double v)
{
#ifdef __CUDA_ARCH__
  unsigned long long b = (unsigned long long)__double_as_longlong(v);
#else
  unsigned long long b;
  memcpy(&b, &v, 8);
#endif
  return (b & 0x8000000000000000ull) ? ~b : (b | 0x8000000000000000ull);
}
__host__ __device__ __forceinline__ double ordered_to_f64(unsigned long long o)
{
  unsigned long long b = (o & 0x8000000000000000ull) ? (o & 0x7FFFFFFFFFFFFFFFull) : ~o;
#ifdef __CUDA_ARCH__
  return __longlong_as_double((long long)b);
#else
  double v;
  memcpy(&v, &b, 8);
  return v;
#endif
}

__device__ __forceinline__ unsigned lane_id() { return threadIdx.x & 31u; }

// 256-bit read-only global load (sm_100: LDG.E.256).  A lane that reads a record of its own
// (nothing to coalesce with) pays one L1 wavefront per load instruction, so the traversal records are
// laid out in 32-byte units and fetched with the widest load the machine has.
struct alignas(32) D4
{
  double x, y, z, w;
};
__device__ __forceinline__ D4 ldg256(const void* p)
{
  D4 r;
  asm("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(r.x), "=d"(r.y), "=d"(r.z), "=d"(r.w) : "l"(p));
  return r;
}

constexpr int kNumSMsB200 = 148;

}  // namespace axb
