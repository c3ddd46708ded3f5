// axom_b200/primal.hpp -- the handful of POD geometry types the spin::BVH / quest::SignedDistance
// interfaces are expressed in, layout-compatible with the reference's:
//   primal::Point<T,D>        T[D]                      primal/geometry/Point.hpp
//   primal::Vector<T,D>       T[D]                      primal/geometry/Vector.hpp
//   primal::BoundingBox<T,D>  Point min, Point max      primal/geometry/BoundingBox.hpp:63-366  (48 B for double/3-D)
//   primal::Ray<T,D>          Point origin, Vector dir  primal/geometry/Ray.hpp  (direction normalised by the ctor, :122-127)
//   primal::ZipIndexable<G>   SoA view, one array per component  primal/utils/Zip{Point,BoundingBox,Ray}.hpp
// Only construction and element access live here: every geometric computation of the hot path
// happens in libaxb200 on the GPU.  An array of the reference's own primal objects can be handed to
// the shims unchanged (same bytes); these types exist so the shims compile without the reference.
#ifndef AXOM_B200_PRIMAL_HPP_
#define AXOM_B200_PRIMAL_HPP_

#include <cmath>
#include <cstddef>
#include <initializer_list>
#include <cstdint>
#include <limits>

namespace axom_b200
{
using IndexType = std::int32_t;  // core/Types.hpp:63-65 (default build, no AXOM_USE_64BIT_INDEXTYPE)

namespace primal
{
template <typename T, int D>
struct Point
{
  T m_components[D];
  Point() : m_components {} { }
  explicit Point(T v)
  {
    for(int d = 0; d < D; ++d) m_components[d] = v;
  }
  Point(const T* v)
  {
    for(int d = 0; d < D; ++d) m_components[d] = v[d];
  }
  Point(std::initializer_list<T> v) : m_components {}
  {
    int d = 0;
    for(T x : v)
      if(d < D) m_components[d++] = x;
  }
  static Point make_point(T x, T y, T z = T())
  {
    T v[3] = {x, y, z};
    return Point(v);
  }
  T& operator[](int i) { return m_components[i]; }
  const T& operator[](int i) const { return m_components[i]; }
  const T* data() const { return m_components; }
  T* data() { return m_components; }
};

template <typename T, int D>
struct Vector
{
  T m_components[D];
  Vector() : m_components {} { }
  Vector(const T* v)
  {
    for(int d = 0; d < D; ++d) m_components[d] = v[d];
  }
  Vector(std::initializer_list<T> v) : m_components {}
  {
    int d = 0;
    for(T x : v)
      if(d < D) m_components[d++] = x;
  }
  T& operator[](int i) { return m_components[i]; }
  const T& operator[](int i) const { return m_components[i]; }
  const T* data() const { return m_components; }
  T squared_norm() const
  {
    T r = T();
    for(int d = 0; d < D; ++d) r += m_components[d] * m_components[d];
    return r;
  }
  T norm() const { return std::sqrt(squared_norm()); }
  // Vector::unitVector (Vector.hpp:477-493)
  Vector unitVector() const
  {
    Vector u;
    const T len2 = squared_norm();
    if(len2 >= static_cast<T>(1e-50))
    {
      const T s = static_cast<T>(1.) / std::sqrt(len2);
      for(int d = 0; d < D; ++d) u[d] = static_cast<T>(m_components[d] * s);
    }
    else
    {
      u[0] = static_cast<T>(1);
    }
    return u;
  }
};

template <typename T, int D>
struct BoundingBox
{
  using PointType = Point<T, D>;
  PointType m_min, m_max;
  // invalid box: (max, lowest)  -- BoundingBox.hpp:72-73,83
  BoundingBox() { clear(); }
  BoundingBox(const PointType& lo, const PointType& hi) : m_min(lo), m_max(hi)
  {
    // checkAndFixBounds (:575-584)
    for(int d = 0; d < D; ++d)
      if(m_min[d] > m_max[d])
      {
        const T t = m_min[d];
        m_min[d] = m_max[d];
        m_max[d] = t;
      }
  }
  void clear()
  {
    for(int d = 0; d < D; ++d)
    {
      m_min[d] = std::numeric_limits<T>::max();
      m_max[d] = std::numeric_limits<T>::lowest();
    }
  }
  const PointType& getMin() const { return m_min; }
  const PointType& getMax() const { return m_max; }
  bool isValid() const
  {
    for(int d = 0; d < D; ++d)
      if(m_min[d] > m_max[d]) return false;
    return true;
  }
  void addPoint(const PointType& p)
  {
    for(int d = 0; d < D; ++d)
    {
      if(p[d] < m_min[d]) m_min[d] = p[d];
      if(p[d] > m_max[d]) m_max[d] = p[d];
    }
  }
  bool contains(const PointType& p) const
  {
    for(int d = 0; d < D; ++d)
      if(p[d] < m_min[d] || p[d] > m_max[d]) return false;
    return true;
  }
};

template <typename T, int D>
struct Ray
{
  Point<T, D> m_origin;
  Vector<T, D> m_direction;
  Ray() { m_direction[0] = static_cast<T>(1); }
  Ray(const Point<T, D>& o, const Vector<T, D>& dir) : m_origin(o), m_direction(dir.unitVector()) { }
  const Point<T, D>& origin() const { return m_origin; }
  const Vector<T, D>& direction() const { return m_direction; }
};

static_assert(sizeof(BoundingBox<double, 3>) == 48, "BoundingBox<double,3> must match the reference's 48-byte layout");
static_assert(sizeof(Ray<double, 3>) == 48, "Ray<double,3> layout");
static_assert(sizeof(Point<double, 3>) == 24, "Point<double,3> layout");

// ZipIndexable (primal/utils/ZipIndexable.hpp:58-95): SoA storage presented as an Indexable of geometry
// objects.  Unlike the reference's, the component pointers are public: the shims pass them straight to
// the C ABI as an SoA descriptor instead of gathering element by element.
template <typename Geom>
struct ZipIndexable;

template <typename T, int D>
struct ZipIndexable<Point<T, D>>
{
  using GeomType = Point<T, D>;
  const T* pts_arrays[D];
  ZipIndexable() : pts_arrays {} { }
  template <std::size_t N>
  ZipIndexable(const T* const (&a)[N])
  {
    static_assert(N >= std::size_t(D), "Must provide at least NDIMS arrays");
    for(int d = 0; d < D; ++d) pts_arrays[d] = a[d];
  }
  GeomType operator[](int i) const
  {
    GeomType p;
    for(int d = 0; d < D; ++d) p[d] = pts_arrays[d][i];
    return p;
  }
};

template <typename T, int D>
struct ZipIndexable<BoundingBox<T, D>>
{
  using GeomType = BoundingBox<T, D>;
  const T* bb_min_arrays[D];
  const T* bb_max_arrays[D];
  ZipIndexable() : bb_min_arrays {}, bb_max_arrays {} { }
  template <std::size_t N>
  ZipIndexable(const T* const (&mn)[N], const T* const (&mx)[N])
  {
    static_assert(N >= std::size_t(D), "Must provide at least NDIMS arrays");
    for(int d = 0; d < D; ++d)
    {
      bb_min_arrays[d] = mn[d];
      bb_max_arrays[d] = mx[d];
    }
  }
  GeomType operator[](int i) const
  {
    Point<T, D> lo, hi;
    for(int d = 0; d < D; ++d)
    {
      lo[d] = bb_min_arrays[d][i];
      hi[d] = bb_max_arrays[d][i];
    }
    return GeomType(lo, hi);
  }
};

template <typename T, int D>
struct ZipIndexable<Ray<T, D>>
{
  using GeomType = Ray<T, D>;
  const T* ray_origs[D];
  const T* ray_dirs[D];
  ZipIndexable() : ray_origs {}, ray_dirs {} { }
  template <std::size_t N>
  ZipIndexable(const T* const (&o)[N], const T* const (&dir)[N])
  {
    static_assert(N >= std::size_t(D), "Must provide at least NDIMS arrays");
    for(int d = 0; d < D; ++d)
    {
      ray_origs[d] = o[d];
      ray_dirs[d] = dir[d];
    }
  }
  GeomType operator[](int i) const
  {
    Point<T, D> o;
    Vector<T, D> v;
    for(int d = 0; d < D; ++d)
    {
      o[d] = ray_origs[d][i];
      v[d] = ray_dirs[d][i];
    }
    return GeomType(o, v);
  }
};

}  // namespace primal
}  // namespace axom_b200

#endif  // AXOM_B200_PRIMAL_HPP_
