// axom_b200/BVH.hpp -- C++ header shim that keeps the reference's spin::BVH class and method names
// (spin/BVH.hpp:129-419) on top of the C ABI in axb200.h.  Header-only, plain C++14, no CUDA, no torch:
// link with -laxb200.
//
//   reference                                              here
//   axom::spin::BVH<NDIMS, ExecSpace, FloatType>           axom_b200::spin::BVH<NDIMS, axom_b200::B200_EXEC, FloatType>
//   axom::ArrayView<IndexType> / axom::Array<IndexType>    axom_b200::ArrayView / axom_b200::Array (pointer+size; owning buffer)
//   "Indexable" arguments (spin/BVH.hpp:143-171)           resolved to an axb_array_desc by indexable_traits below
//
// Defining AXOM_B200_ALIAS_AXOM before including this header adds `namespace axom = axom_b200;` so that
// call sites written against the reference compile unchanged apart from the execution-space tag.
//
// Memory spaces: as in the reference (spin/BVH.hpp:198-203) the caller passes pointers valid in the
// execution space; device and managed pointers are used in place, host pointers are staged by the
// library (AXB_MEM_AUTO).  The candidates Array is allocated in the same space as `offsets`.
//
// Errors: the reference reports misuse through SLIC_ERROR, which aborts by default
// (slic/core/Logger.cpp:37,51).  Here every non-zero axb_status is routed to axom_b200::error_handler(),
// whose default prints the message and calls std::abort(); tests install a throwing handler.
#ifndef AXOM_B200_BVH_HPP_
#define AXOM_B200_BVH_HPP_

#include <cstdio>
#include <cstdlib>
#include <string>
#include <type_traits>
#include <utility>
#include <vector>

#include "../axb200.h"
#include "primal.hpp"

namespace axom_b200
{
// execution-space tag of this implementation (core/execution/execution_space.hpp has SEQ_EXEC,
// OMP_EXEC, CUDA_EXEC<N>, HIP_EXEC<N>)
struct B200_EXEC
{
  static constexpr bool valid() { return true; }
  static constexpr bool async() { return false; }
  static constexpr bool onDevice() { return true; }
  static constexpr const char* name() { return "[B200_EXEC]"; }
};

using ErrorHandler = void (*)(int status, const char* message);
inline void default_error_handler(int status, const char* message)
{
  std::fprintf(stderr, "[axom_b200 ERROR] %s: %s\n", axb_status_string(status), message);
  std::abort();
}
inline ErrorHandler& error_handler()
{
  static ErrorHandler h = default_error_handler;
  return h;
}
inline void check(int status)
{
  if(status != AXB_OK) error_handler()(status, axb_last_error());
}

// axom::ArrayView<T> (core/ArrayView.hpp): non-owning pointer + size
template <typename T>
class ArrayView
{
public:
  ArrayView() = default;
  ArrayView(T* data, IndexType n) : m_data(data), m_size(n) { }
  template <typename A>
  ArrayView(std::vector<A>& v) : m_data(v.data()), m_size(static_cast<IndexType>(v.size()))
  { }
  T* data() const { return m_data; }
  IndexType size() const { return m_size; }
  T& operator[](IndexType i) const { return m_data[i]; }

private:
  T* m_data = nullptr;
  IndexType m_size = 0;
};

// axom::Array<IndexType> as far as find*() needs it: an owning, movable buffer allocated by the
// library in the caller's memory space (policy/LinearBVH.hpp:323-330 allocates with the BVH's allocator
// and move-assigns into the caller's array)
template <typename T>
class Array
{
public:
  Array() = default;
  Array(const Array&) = delete;
  Array& operator=(const Array&) = delete;
  Array(Array&& o) noexcept { swap(o); }
  Array& operator=(Array&& o) noexcept
  {
    if(this != &o)
    {
      clear();
      swap(o);
    }
    return *this;
  }
  ~Array() { clear(); }
  void adopt(axb_bvh* owner, T* data, IndexType n, int memspace)
  {
    clear();
    m_owner = owner;
    m_data = data;
    m_size = n;
    m_space = memspace;
  }
  void clear()
  {
    if(m_data) axb_bvh_free_candidates(m_owner, reinterpret_cast<int32_t*>(m_data), m_space);
    m_data = nullptr;
    m_size = 0;
  }
  T* data() const { return m_data; }
  IndexType size() const { return m_size; }
  bool empty() const { return m_size == 0; }
  T& operator[](IndexType i) const { return m_data[i]; }  // host-space arrays only
  ArrayView<T> view() const { return ArrayView<T>(m_data, m_size); }

private:
  void swap(Array& o)
  {
    std::swap(m_owner, o.m_owner);
    std::swap(m_data, o.m_data);
    std::swap(m_size, o.m_size);
    std::swap(m_space, o.m_space);
  }
  axb_bvh* m_owner = nullptr;
  T* m_data = nullptr;
  IndexType m_size = 0;
  int m_space = AXB_MEM_AUTO;
};

namespace detail
{
// An Indexable resolved for the C ABI.  `normalized` only matters for rays: AoS arrays hold constructed
// primal::Ray objects (direction already unit length), SoA / gathered rays still need the constructor's
// normalisation, which the library applies on the device.
template <typename T>
struct Resolved
{
  axb_array_desc desc;
  std::vector<T> gathered;  // only for generic Indexables
  int normalized = 1;
};

template <typename T, int NCOMP>
inline void aos_desc(axb_array_desc& d, const void* base, int memspace)
{
  for(int c = 0; c < 6; ++c) d.comp[c] = nullptr;
  for(int c = 0; c < NCOMP; ++c) d.comp[c] = static_cast<const char*>(base) + c * sizeof(T);
  d.stride_bytes = NCOMP * sizeof(T);
  d.ncomp = NCOMP;
  d.memspace = memspace;
}

// component c of a geometry object (works for our PODs and for anything layout-compatible)
template <typename T, int D>
inline void put(const primal::Point<T, D>& p, T* out)
{
  for(int d = 0; d < D; ++d) out[d] = p[d];
}
template <typename T, int D>
inline void put(const primal::BoundingBox<T, D>& b, T* out)
{
  for(int d = 0; d < D; ++d)
  {
    out[d] = b.getMin()[d];
    out[D + d] = b.getMax()[d];
  }
}
template <typename T, int D>
inline void put(const primal::Ray<T, D>& r, T* out)
{
  for(int d = 0; d < D; ++d)
  {
    out[d] = r.origin()[d];
    out[D + d] = r.direction()[d];
  }
}

template <typename Geom>
struct geom_info;
template <typename T, int D>
struct geom_info<primal::Point<T, D>>
{
  using Float = T;
  static constexpr int ncomp = D;
};
template <typename T, int D>
struct geom_info<primal::BoundingBox<T, D>>
{
  using Float = T;
  static constexpr int ncomp = 2 * D;
};
template <typename T, int D>
struct geom_info<primal::Ray<T, D>>
{
  using Float = T;
  static constexpr int ncomp = 2 * D;
};

// generic Indexable: anything with a host-callable operator[](int) convertible to Geom
// (the reference's own requirement, spin/BVH.hpp:143-171).  Gathered once on the host.
template <typename Geom, typename Indexable>
struct indexable_traits
{
  using T = typename geom_info<Geom>::Float;
  static void resolve(const Indexable& it, IndexType n, Resolved<T>& r)
  {
    constexpr int NC = geom_info<Geom>::ncomp;
    r.gathered.resize(static_cast<std::size_t>(n > 0 ? n : 0) * NC);
    for(IndexType i = 0; i < n; ++i)
    {
      const Geom g = it[i];
      put(g, r.gathered.data() + static_cast<std::size_t>(i) * NC);
    }
    aos_desc<T, NC>(r.desc, r.gathered.data(), AXB_MEM_HOST);
    r.normalized = 1;  // Geom was constructed on the host: rays are already unit length
  }
};

// raw pointers and ArrayViews of layout-compatible objects: used in place, wherever they live
template <typename Geom, typename Obj>
struct indexable_traits<Geom, Obj*>
{
  using T = typename geom_info<Geom>::Float;
  static_assert(sizeof(Obj) == geom_info<Geom>::ncomp * sizeof(T), "array element is not layout-compatible with the primal type");
  static void resolve(Obj* const& p, IndexType, Resolved<T>& r)
  {
    aos_desc<T, geom_info<Geom>::ncomp>(r.desc, p, AXB_MEM_AUTO);
    r.normalized = 1;
  }
};
template <typename Geom, typename Obj>
struct indexable_traits<Geom, ArrayView<Obj>>
{
  using T = typename geom_info<Geom>::Float;
  static void resolve(const ArrayView<Obj>& v, IndexType n, Resolved<T>& r) { indexable_traits<Geom, Obj*>::resolve(v.data(), n, r); }
};

// SoA views: component pointers go to the library as they are
template <typename T_, int D>
struct indexable_traits<primal::Point<T_, D>, primal::ZipIndexable<primal::Point<T_, D>>>
{
  using T = T_;
  static void resolve(const primal::ZipIndexable<primal::Point<T, D>>& z, IndexType, Resolved<T>& r)
  {
    for(int c = 0; c < 6; ++c) r.desc.comp[c] = nullptr;
    for(int d = 0; d < D; ++d) r.desc.comp[d] = z.pts_arrays[d];
    r.desc.stride_bytes = sizeof(T);
    r.desc.ncomp = D;
    r.desc.memspace = AXB_MEM_AUTO;
  }
};
template <typename T_, int D>
struct indexable_traits<primal::BoundingBox<T_, D>, primal::ZipIndexable<primal::BoundingBox<T_, D>>>
{
  using T = T_;
  static void resolve(const primal::ZipIndexable<primal::BoundingBox<T, D>>& z, IndexType, Resolved<T>& r)
  {
    for(int c = 0; c < 6; ++c) r.desc.comp[c] = nullptr;
    for(int d = 0; d < D; ++d)
    {
      r.desc.comp[d] = z.bb_min_arrays[d];
      r.desc.comp[D + d] = z.bb_max_arrays[d];
    }
    r.desc.stride_bytes = sizeof(T);
    r.desc.ncomp = 2 * D;
    r.desc.memspace = AXB_MEM_AUTO;
  }
};
template <typename T_, int D>
struct indexable_traits<primal::Ray<T_, D>, primal::ZipIndexable<primal::Ray<T_, D>>>
{
  using T = T_;
  static void resolve(const primal::ZipIndexable<primal::Ray<T, D>>& z, IndexType, Resolved<T>& r)
  {
    for(int c = 0; c < 6; ++c) r.desc.comp[c] = nullptr;
    for(int d = 0; d < D; ++d)
    {
      r.desc.comp[d] = z.ray_origs[d];
      r.desc.comp[D + d] = z.ray_dirs[d];
    }
    r.desc.stride_bytes = sizeof(T);
    r.desc.ncomp = 2 * D;
    r.desc.memspace = AXB_MEM_AUTO;
    r.normalized = 0;  // ZipRay::operator[] runs the Ray constructor (ZipRay.hpp:62-77): normalise on the device
  }
};
}  // namespace detail

namespace spin
{
constexpr int BVH_BUILD_OK = AXB_BVH_BUILD_OK;  // spin/BVH.hpp:39-43

// LinearBVHTraverser (spin/policy/LinearBVH.hpp:57-109): the three device arrays in the reference's
// layout.  traverse_tree() is a device-side template: include axom_b200/traverser.cuh from CUDA code.
template <typename FloatType, int NDIMS>
struct LinearBVHTraverser
{
  using BoxType = primal::BoundingBox<FloatType, NDIMS>;
  const BoxType* m_inner_nodes = nullptr;            // [2*(N-1)]  child boxes, pair per inner node
  const std::int32_t* m_inner_node_children = nullptr;  // [2*(N-1)]
  const std::int32_t* m_leaf_nodes = nullptr;           // [N] original ids in sorted order
  IndexType m_num_leaves = 0;
};

template <int NDIMS, typename ExecSpace = B200_EXEC, typename FloatType = double>
class BVH
{
  static_assert(NDIMS == 2 || NDIMS == 3, "The BVH class may be used only in 2D or 3D.");
  static_assert(std::is_floating_point<FloatType>::value, "A valid FloatingType must be used for the BVH.");
  static_assert(std::is_same<ExecSpace, B200_EXEC>::value, "axom_b200::spin::BVH runs on B200_EXEC only (no CPU fallback).");

public:
  using BoxType = primal::BoundingBox<FloatType, NDIMS>;
  using PointType = primal::Point<FloatType, NDIMS>;
  using RayType = primal::Ray<FloatType, NDIMS>;
  using TraverserType = LinearBVHTraverser<FloatType, NDIMS>;
  using ExecSpaceType = ExecSpace;

  explicit BVH(int device = 0) : m_device(device) { }

  // BVH(boxes, numItems, allocatorID, tolerance, scaleFactor)  (:210-221); allocatorID selects the device here
  template <typename BoxIndexable>
  BVH(const BoxIndexable boxes, IndexType numItems, int device = 0, FloatType tolerance = DEFAULT_TOLERANCE,
      FloatType scaleFactor = DEFAULT_SCALE_FACTOR)
    : m_device(device), m_tolerance(tolerance), m_scaleFactor(scaleFactor)
  {
    initialize(boxes, numItems);
  }

  BVH(const BVH&) = delete;
  BVH& operator=(const BVH&) = delete;
  ~BVH()
  {
    if(m_owned && m_bvh) axb_bvh_destroy(m_bvh);
  }

  // initialize(boxes, numItems) (:424-477): copies the boxes, may be called again
  template <typename BoxIndexable>
  int initialize(const BoxIndexable boxes, IndexType numItems)
  {
    ensure();
    detail::Resolved<FloatType> r;
    detail::indexable_traits<BoxType, BoxIndexable>::resolve(boxes, numItems, r);
    check(axb_bvh_set_scale_factor(m_bvh, static_cast<double>(m_scaleFactor)));
    check(axb_bvh_set_tolerance(m_bvh, static_cast<double>(m_tolerance)));
    const int st = axb_bvh_initialize(m_bvh, &r.desc, numItems);
    check(st);
    return st;
  }

  bool isInitialized() const { return m_bvh && axb_bvh_is_initialized(m_bvh) == 1; }

  void setAllocatorID(int device) { m_device = device; }  // an allocator id is a memory space; here: the GPU ordinal
  int getAllocatorID() const { return m_device; }
  void setScaleFactor(FloatType s) { m_scaleFactor = s; }
  FloatType getScaleFactor() const { return m_scaleFactor; }
  void setTolerance(FloatType eps)
  {
    m_tolerance = eps;
    if(m_bvh) check(axb_bvh_set_tolerance(m_bvh, static_cast<double>(eps)));
  }
  FloatType getTolerance() const { return m_tolerance; }

  // getBounds() (:297-308): bounds of the scaled boxes; an invalid box before initialize()
  BoxType getBounds() const
  {
    BoxType b;
    if(isInitialized())
    {
      double lo[3], hi[3];
      check(axb_bvh_get_bounds(m_bvh, lo, hi));
      PointType pl, ph;
      for(int d = 0; d < NDIMS; ++d)
      {
        pl[d] = static_cast<FloatType>(lo[d]);
        ph[d] = static_cast<FloatType>(hi[d]);
      }
      b = BoxType(pl, ph);
    }
    return b;
  }

  // writeVtkFile(fileName) (spin/BVH.hpp:405): the tree's boxes as an ASCII VTK file, for debugging
  void writeVtkFile(const std::string& fileName) const { check(axb_bvh_write_vtk_file(m_bvh, fileName.c_str())); }

  TraverserType getTraverser() const
  {
    axb_traverser t;
    check(axb_bvh_get_traverser(m_bvh, &t));
    TraverserType out;
    out.m_inner_nodes = static_cast<const BoxType*>(t.inner_nodes);
    out.m_inner_node_children = t.inner_node_children;
    out.m_leaf_nodes = t.leaf_nodes;
    out.m_num_leaves = t.num_leaves;
    return out;
  }

  template <typename PointIndexable>
  void findPoints(ArrayView<IndexType> offsets, ArrayView<IndexType> counts, Array<IndexType>& candidates, IndexType numPts,
                  PointIndexable points) const
  {
    detail::Resolved<FloatType> r;
    detail::indexable_traits<PointType, PointIndexable>::resolve(points, numPts, r);
    if(!sizes_ok(offsets, counts, numPts)) return;
    std::int32_t* cand = nullptr;
    std::int64_t total = 0;
    check(axb_bvh_find_points(m_bvh, &r.desc, numPts, offsets.data(), counts.data(), AXB_MEM_AUTO, &cand, &total));
    candidates.adopt(m_bvh, cand, static_cast<IndexType>(total), AXB_MEM_AUTO);
  }

  template <typename RayIndexable>
  void findRays(ArrayView<IndexType> offsets, ArrayView<IndexType> counts, Array<IndexType>& candidates, IndexType numRays,
                RayIndexable rays) const
  {
    detail::Resolved<FloatType> r;
    detail::indexable_traits<RayType, RayIndexable>::resolve(rays, numRays, r);
    if(!sizes_ok(offsets, counts, numRays)) return;
    std::int32_t* cand = nullptr;
    std::int64_t total = 0;
    check(axb_bvh_find_rays(m_bvh, &r.desc, r.normalized, numRays, offsets.data(), counts.data(), AXB_MEM_AUTO, &cand, &total));
    candidates.adopt(m_bvh, cand, static_cast<IndexType>(total), AXB_MEM_AUTO);
  }

  template <typename BoxIndexable>
  void findBoundingBoxes(ArrayView<IndexType> offsets, ArrayView<IndexType> counts, Array<IndexType>& candidates, IndexType numBoxes,
                         BoxIndexable boxes) const
  {
    detail::Resolved<FloatType> r;
    detail::indexable_traits<BoxType, BoxIndexable>::resolve(boxes, numBoxes, r);
    if(!sizes_ok(offsets, counts, numBoxes)) return;
    std::int32_t* cand = nullptr;
    std::int64_t total = 0;
    check(axb_bvh_find_boxes(m_bvh, &r.desc, numBoxes, offsets.data(), counts.data(), AXB_MEM_AUTO, &cand, &total));
    candidates.adopt(m_bvh, cand, static_cast<IndexType>(total), AXB_MEM_AUTO);
  }

  // the underlying C handle (e.g. to set a stream or read phase timers)
  axb_bvh* handle() const { return m_bvh; }
  // wrap a handle owned by someone else (quest::SignedDistance::getBVHTree)
  void borrow(axb_bvh* h)
  {
    if(m_owned && m_bvh) axb_bvh_destroy(m_bvh);
    m_bvh = h;
    m_owned = false;
    double s = 0, t = 0;
    if(h && axb_bvh_get_scale_factor(h, &s) == AXB_OK) m_scaleFactor = static_cast<FloatType>(s);
    if(h && axb_bvh_get_tolerance(h, &t) == AXB_OK) m_tolerance = static_cast<FloatType>(t);
  }

  static constexpr FloatType DEFAULT_SCALE_FACTOR = static_cast<FloatType>(1.000123);  // :410
  static constexpr FloatType DEFAULT_TOLERANCE = std::numeric_limits<FloatType>::epsilon();

private:
  void ensure()
  {
    if(!m_bvh)
    {
      check(axb_bvh_create(&m_bvh, NDIMS, static_cast<int>(sizeof(FloatType)), m_device));
      m_owned = true;
    }
  }
  // policy/LinearBVH.hpp:284-285: SLIC_ERROR_IF(offsets.size() != numObjs) ...
  bool sizes_ok(const ArrayView<IndexType>& offsets, const ArrayView<IndexType>& counts, IndexType n) const
  {
    if(!m_bvh)
    {
      error_handler()(AXB_ERR_NOT_BUILT, "BVH query before initialize()");
      return false;
    }
    if(offsets.size() != n)
    {
      error_handler()(AXB_ERR_BAD_ARG, "offsets length not equal to numObjs");
      return false;
    }
    if(counts.size() != n)
    {
      error_handler()(AXB_ERR_BAD_ARG, "counts length not equal to numObjs");
      return false;
    }
    return true;
  }

  int m_device = 0;
  FloatType m_tolerance {DEFAULT_TOLERANCE};
  FloatType m_scaleFactor {DEFAULT_SCALE_FACTOR};
  mutable axb_bvh* m_bvh = nullptr;
  bool m_owned = false;
};

template <int NDIMS, typename ExecSpace, typename FloatType>
constexpr FloatType BVH<NDIMS, ExecSpace, FloatType>::DEFAULT_SCALE_FACTOR;
template <int NDIMS, typename ExecSpace, typename FloatType>
constexpr FloatType BVH<NDIMS, ExecSpace, FloatType>::DEFAULT_TOLERANCE;

}  // namespace spin
}  // namespace axom_b200

#ifdef AXOM_B200_ALIAS_AXOM
namespace axom = axom_b200;
#endif

#endif  // AXOM_B200_BVH_HPP_
