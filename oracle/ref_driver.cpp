// ref_driver.cpp -- thin C ABI over the UNMODIFIED reference (LLNL/axom v0.11.0) so that
// tests and bench.py can run the reference's own SEQ_EXEC path from Python/ctypes.
//
// TEST INFRASTRUCTURE ONLY (see oracle/build_ref.py).  This file contains no algorithm:
// every entry point forwards to the reference's public API
//   spin::BVH<D,SEQ_EXEC,double>           (spin/BVH.hpp:129)
//   lbvh::build_radix_tree<SEQ_EXEC>       (spin/internal/linear_bvh/build_radix_tree.hpp:579)
//   quest::SignedDistance<3,SEQ_EXEC>      (quest/SignedDistance.hpp:147)
//   primal::intersect(Triangle3,Triangle3) (primal/operators/intersect.hpp:64), quest::findTriMeshIntersectionsBVH
//   (quest/MeshTester.hpp:67), quest::STLReader, quest::weldTriMeshVertices (quest/MeshTester.cpp:218),
//   quest::signed_distance_init / evaluate / finalize (quest/interface/signed_distance.hpp:117-319)
// One exception, marked below: the two lambdas of DistributedClosestPointImpl::computeLocalClosestPoints are restated
// here around the real BVH traverser, because the class itself needs Conduit + MPI.
// The entry points mirror oracle/axb_oracle.cpp one to one (axref_* vs axo_*), so the same
// Python wrapper drives both.
#include "axom/config.hpp"
#include "axom/core/Array.hpp"
#include "axom/core/ArrayView.hpp"
#include "axom/core/execution/execution_space.hpp"
#include "axom/slic/interface/slic.hpp"
#include "axom/slic/core/SimpleLogger.hpp"
#include "axom/primal/geometry/BoundingBox.hpp"
#include "axom/primal/geometry/Point.hpp"
#include "axom/primal/geometry/Ray.hpp"
#include "axom/primal/geometry/Vector.hpp"
#include "axom/spin/BVH.hpp"
#include "axom/mint/mesh/UnstructuredMesh.hpp"
#include "axom/quest/SignedDistance.hpp"
#include "axom/primal/geometry/Triangle.hpp"
#include "axom/primal/operators/intersect.hpp"
#include "axom/quest/MeshTester.hpp"
#include "axom/quest/readers/STLReader.hpp"
#include "axom/quest/interface/signed_distance.hpp"
#include "axom/primal/operators/squared_distance.hpp"
#include "axom/primal/operators/closest_point.hpp"
#include "axom/primal/operators/detail/intersect_ray_impl.hpp"
#include <limits>
#include <memory>

#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>

#ifdef _OPENMP
  #include <omp.h>
#endif

namespace
{
using axom::SEQ_EXEC;
namespace lbvh = axom::spin::internal::linear_bvh;

void ensure_slic()
{
  if(!axom::slic::isInitialized())
  {
    axom::slic::initialize();
    axom::slic::setLoggingMsgLevel(axom::slic::message::Warning);
    axom::slic::addStreamToAllMsgLevels(new axom::slic::GenericOutputStream(&std::cerr));
  }
}

// LinearBVHTraverser (spin/policy/LinearBVH.hpp:57-109) keeps its three ArrayViews private;
// this mirror has the same members in the same order so the arrays can be read out.
template <typename F, int D>
struct TraverserMirror
{
  axom::ArrayView<const axom::primal::BoundingBox<F, D>> inner_nodes;
  axom::ArrayView<const std::int32_t> inner_children;
  axom::ArrayView<const std::int32_t> leaf_nodes;
};

template <typename F, int D>
struct RefBvh
{
  using BVHType = axom::spin::BVH<D, SEQ_EXEC, F>;
  using BoxType = axom::primal::BoundingBox<F, D>;
  BVHType bvh;
  int n_in = 0;
  // radix-tree internals, from a second (identical) call of build_radix_tree
  lbvh::RadixTree<F, D> radix;
};

template <typename F, int D>
RefBvh<F, D>* create(const F* boxes_aos, int n, double scale, double tol)
{
  using BoxType = typename RefBvh<F, D>::BoxType;
  static_assert(sizeof(BoxType) == sizeof(F) * 2 * D, "BoundingBox is min[D],max[D]");
  ensure_slic();
  RefBvh<F, D>* r = new RefBvh<F, D>();
  r->n_in = n;
  if(scale > 0) r->bvh.setScaleFactor(scale);
  if(tol >= 0) r->bvh.setTolerance(tol);
  const BoxType* boxes = reinterpret_cast<const BoxType*>(boxes_aos);
  r->bvh.initialize(boxes, n);
  // internals: same padding as BVH::initialize (spin/BVH.hpp:439-464)
  std::vector<BoxType> tmp;
  int m = n;
  if(n <= 1)
  {
    tmp.resize(2);
    tmp[0].clear();
    tmp[1].clear();
    if(n == 1) tmp[0] = boxes[0];
    boxes = tmp.data();
    m = 2;
  }
  BoxType bounds;
  lbvh::build_radix_tree<SEQ_EXEC>(boxes, m, bounds, r->radix, r->bvh.getScaleFactor(), r->bvh.getAllocatorID());
  return r;
}

template <typename F, int D>
void get_arrays(const RefBvh<F, D>& r, uint32_t* mcodes, int32_t* leafs, int32_t* lchild, int32_t* rchild, int32_t* parents,
                F* inner_nodes, int32_t* inner_children, F* bounds)
{
  using BoxType = typename RefBvh<F, D>::BoxType;
  const int n = r.radix.m_size, inner = n - 1;
  if(mcodes) memcpy(mcodes, r.radix.m_mcodes.data(), sizeof(uint32_t) * n);
  if(leafs) memcpy(leafs, r.radix.m_leafs.data(), sizeof(int32_t) * n);
  if(lchild) memcpy(lchild, r.radix.m_left_children.data(), sizeof(int32_t) * inner);
  if(rchild) memcpy(rchild, r.radix.m_right_children.data(), sizeof(int32_t) * inner);
  if(parents) memcpy(parents, r.radix.m_parents.data(), sizeof(int32_t) * (inner + n));
  auto trav = r.bvh.getTraverser();
  static_assert(sizeof(trav) == sizeof(TraverserMirror<F, D>), "traverser layout changed");
  TraverserMirror<F, D> tm;
  memcpy((void*)&tm, (const void*)&trav, sizeof(tm));
  if(inner_nodes) memcpy(inner_nodes, tm.inner_nodes.data(), sizeof(BoxType) * 2 * inner);
  if(inner_children) memcpy(inner_children, tm.inner_children.data(), sizeof(int32_t) * 2 * inner);
  if(leafs)
  {
    // the final leaf_nodes array must equal the radix tree's permutation
    if(memcmp(leafs, tm.leaf_nodes.data(), sizeof(int32_t) * n) != 0) abort();
  }
  if(bounds)
  {
    BoxType b = r.bvh.getBounds();
    memcpy(bounds, &b, sizeof(BoxType));
  }
}

int32_t* to_malloc(const axom::Array<axom::IndexType>& a)
{
  const size_t n = a.size();
  int32_t* p = (int32_t*)malloc(sizeof(int32_t) * (n ? n : 1));
  if(n) memcpy(p, a.data(), sizeof(int32_t) * n);
  return p;
}

template <typename F, int D>
int64_t find_points(const RefBvh<F, D>& r, const F* pts, int q, int32_t* off, int32_t* cnt, int32_t** cand)
{
  using PointType = axom::primal::Point<F, D>;
  axom::Array<axom::IndexType> c;
  r.bvh.findPoints(axom::ArrayView<axom::IndexType>(off, q), axom::ArrayView<axom::IndexType>(cnt, q), c, q,
                   reinterpret_cast<const PointType*>(pts));
  *cand = to_malloc(c);
  return c.size();
}

template <typename F, int D>
int64_t find_boxes(const RefBvh<F, D>& r, const F* bx, int q, int32_t* off, int32_t* cnt, int32_t** cand)
{
  using BoxType = typename RefBvh<F, D>::BoxType;
  axom::Array<axom::IndexType> c;
  r.bvh.findBoundingBoxes(axom::ArrayView<axom::IndexType>(off, q), axom::ArrayView<axom::IndexType>(cnt, q), c, q,
                          reinterpret_cast<const BoxType*>(bx));
  *cand = to_malloc(c);
  return c.size();
}

template <typename F, int D>
int64_t find_rays(const RefBvh<F, D>& r, const F* orig, const F* dirs, int q, int normalize, int32_t* off, int32_t* cnt,
                  int32_t** cand)
{
  using RayType = axom::primal::Ray<F, D>;
  using PointType = axom::primal::Point<F, D>;
  using VectorType = axom::primal::Vector<F, D>;
  std::vector<RayType> rays;
  rays.reserve(q);
  for(int i = 0; i < q; ++i)
  {
    PointType o(orig + (size_t)i * D, D);
    VectorType d(dirs + (size_t)i * D, D);
    RayType ray(o, d);  // normalises (Ray.hpp:122-127)
    if(!normalize)
    {
      // store the direction verbatim, as a caller holding already-built Ray objects would
      struct Raw
      {
        PointType o;
        VectorType d;
      } raw {o, d};
      static_assert(sizeof(Raw) == sizeof(RayType), "Ray is origin,direction");
      memcpy((void*)&ray, &raw, sizeof(ray));
    }
    rays.push_back(ray);
  }
  axom::Array<axom::IndexType> c;
  r.bvh.findRays(axom::ArrayView<axom::IndexType>(off, q), axom::ArrayView<axom::IndexType>(cnt, q), c, q, rays.data());
  *cand = to_malloc(c);
  return c.size();
}

struct RefSurface
{
  using Mesh = axom::mint::UnstructuredMesh<axom::mint::SINGLE_SHAPE>;
  using MixedMesh = axom::mint::UnstructuredMesh<axom::mint::MIXED_SHAPE>;
  using SD = axom::quest::SignedDistance<3, SEQ_EXEC>;
  axom::mint::Mesh* mesh = nullptr;
  SD* sd = nullptr;
  ~RefSurface()
  {
    delete sd;
    delete mesh;
  }
};

}  // namespace

struct AxrefBvh
{
  int ndims;
  void* impl;
};

extern "C" {

AxrefBvh* axref_bvh_create(int ndims, const double* boxes_aos, int n, double scale, double tol)
{
  AxrefBvh* h = new AxrefBvh {ndims, nullptr};
  h->impl = ndims == 2 ? (void*)create<double, 2>(boxes_aos, n, scale, tol) : (void*)create<double, 3>(boxes_aos, n, scale, tol);
  return h;
}

void axref_bvh_destroy(AxrefBvh* h)
{
  if(!h) return;
  if(h->ndims == 2)
    delete(RefBvh<double, 2>*)h->impl;
  else
    delete(RefBvh<double, 3>*)h->impl;
  delete h;
}

int axref_bvh_num_leaves(const AxrefBvh* h)
{
  return h->ndims == 2 ? ((RefBvh<double, 2>*)h->impl)->radix.m_size : ((RefBvh<double, 3>*)h->impl)->radix.m_size;
}

void axref_bvh_get(const AxrefBvh* h, uint32_t* mcodes, int32_t* leafs, int32_t* lchild, int32_t* rchild, int32_t* parents,
                   double* inner_nodes, int32_t* inner_children, double* bounds)
{
  if(h->ndims == 2)
    get_arrays(*(RefBvh<double, 2>*)h->impl, mcodes, leafs, lchild, rchild, parents, inner_nodes, inner_children, bounds);
  else
    get_arrays(*(RefBvh<double, 3>*)h->impl, mcodes, leafs, lchild, rchild, parents, inner_nodes, inner_children, bounds);
}

// BVH::writeVtkFile (spin/BVH.hpp:405) of the unmodified reference
void axref_bvh_write_vtk(const AxrefBvh* h, const char* file_name)
{
  if(h->ndims == 2)
    ((RefBvh<double, 2>*)h->impl)->bvh.writeVtkFile(file_name);
  else
    ((RefBvh<double, 3>*)h->impl)->bvh.writeVtkFile(file_name);
}
void axreff_bvh_write_vtk(const AxrefBvh* h, const char* file_name)
{
  if(h->ndims == 2)
    ((RefBvh<float, 2>*)h->impl)->bvh.writeVtkFile(file_name);
  else
    ((RefBvh<float, 3>*)h->impl)->bvh.writeVtkFile(file_name);
}

void axref_free(void* p) { free(p); }

int64_t axref_bvh_find_points(const AxrefBvh* h, const double* pts_aos, int q, int32_t* offsets, int32_t* counts, int32_t** cand)
{
  return h->ndims == 2 ? find_points(*(RefBvh<double, 2>*)h->impl, pts_aos, q, offsets, counts, cand)
                       : find_points(*(RefBvh<double, 3>*)h->impl, pts_aos, q, offsets, counts, cand);
}

int64_t axref_bvh_find_boxes(const AxrefBvh* h, const double* boxes_aos, int q, int32_t* offsets, int32_t* counts, int32_t** cand)
{
  return h->ndims == 2 ? find_boxes(*(RefBvh<double, 2>*)h->impl, boxes_aos, q, offsets, counts, cand)
                       : find_boxes(*(RefBvh<double, 3>*)h->impl, boxes_aos, q, offsets, counts, cand);
}

int64_t axref_bvh_find_rays(const AxrefBvh* h, const double* origins_aos, const double* dirs_aos, int q, int normalize,
                            int32_t* offsets, int32_t* counts, int32_t** cand)
{
  return h->ndims == 2 ? find_rays(*(RefBvh<double, 2>*)h->impl, origins_aos, dirs_aos, q, normalize, offsets, counts, cand)
                       : find_rays(*(RefBvh<double, 3>*)h->impl, origins_aos, dirs_aos, q, normalize, offsets, counts, cand);
}

// ---- FloatType = float: spin::BVH<D,SEQ_EXEC,float> (spin/tests/spin_bvh.cpp:1563-1669 instantiates it) ----
AxrefBvh* axreff_bvh_create(int ndims, const float* boxes_aos, int n, double scale, double tol)
{
  AxrefBvh* h = new AxrefBvh {ndims, nullptr};
  h->impl = ndims == 2 ? (void*)create<float, 2>(boxes_aos, n, scale, tol) : (void*)create<float, 3>(boxes_aos, n, scale, tol);
  return h;
}

void axreff_bvh_destroy(AxrefBvh* h)
{
  if(!h) return;
  if(h->ndims == 2)
    delete(RefBvh<float, 2>*)h->impl;
  else
    delete(RefBvh<float, 3>*)h->impl;
  delete h;
}

int axreff_bvh_num_leaves(const AxrefBvh* h)
{
  return h->ndims == 2 ? ((RefBvh<float, 2>*)h->impl)->radix.m_size : ((RefBvh<float, 3>*)h->impl)->radix.m_size;
}

void axreff_bvh_get(const AxrefBvh* h, uint32_t* mcodes, int32_t* leafs, int32_t* lchild, int32_t* rchild, int32_t* parents,
                   float* inner_nodes, int32_t* inner_children, float* bounds)
{
  if(h->ndims == 2)
    get_arrays(*(RefBvh<float, 2>*)h->impl, mcodes, leafs, lchild, rchild, parents, inner_nodes, inner_children, bounds);
  else
    get_arrays(*(RefBvh<float, 3>*)h->impl, mcodes, leafs, lchild, rchild, parents, inner_nodes, inner_children, bounds);
}

void axreff_free(void* p) { free(p); }

int axreff_max_threads() { return 1; }

int64_t axreff_bvh_find_points(const AxrefBvh* h, const float* pts_aos, int q, int32_t* offsets, int32_t* counts, int32_t** cand)
{
  return h->ndims == 2 ? find_points(*(RefBvh<float, 2>*)h->impl, pts_aos, q, offsets, counts, cand)
                       : find_points(*(RefBvh<float, 3>*)h->impl, pts_aos, q, offsets, counts, cand);
}

int64_t axreff_bvh_find_boxes(const AxrefBvh* h, const float* boxes_aos, int q, int32_t* offsets, int32_t* counts, int32_t** cand)
{
  return h->ndims == 2 ? find_boxes(*(RefBvh<float, 2>*)h->impl, boxes_aos, q, offsets, counts, cand)
                       : find_boxes(*(RefBvh<float, 3>*)h->impl, boxes_aos, q, offsets, counts, cand);
}

int64_t axreff_bvh_find_rays(const AxrefBvh* h, const float* origins_aos, const float* dirs_aos, int q, int normalize,
                            int32_t* offsets, int32_t* counts, int32_t** cand)
{
  return h->ndims == 2 ? find_rays(*(RefBvh<float, 2>*)h->impl, origins_aos, dirs_aos, q, normalize, offsets, counts, cand)
                       : find_rays(*(RefBvh<float, 3>*)h->impl, origins_aos, dirs_aos, q, normalize, offsets, counts, cand);
}

// Reference traversal (getTraverser().traverse_tree, policy/LinearBVH.hpp:92-103) driven by an
// external OpenMP loop: RAJA is absent here so OMP_EXEC cannot be instantiated.
int64_t axref_bvh_count_points_omp(const AxrefBvh* h, const double* pts_aos, int q, int32_t* counts, int nthreads)
{
  if(h->ndims != 3) return -1;
  using PointType = axom::primal::Point<double, 3>;
  using BoxType = axom::primal::BoundingBox<double, 3>;
  const auto trav = ((RefBvh<double, 3>*)h->impl)->bvh.getTraverser();
  const PointType* pts = reinterpret_cast<const PointType*>(pts_aos);
  int64_t total = 0;
#ifdef _OPENMP
  if(nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel for schedule(dynamic, 1024) reduction(+ : total)
  for(int i = 0; i < q; ++i)
  {
    int c = 0;
    const PointType p = pts[i];
    auto leafAction = [&c](std::int32_t, const std::int32_t*) { ++c; };
    auto pred = [](const PointType& pp, const BoxType& bb) { return bb.contains(pp); };
    // generic (non-point) overload has no child ordering, like findCandidatesImpl
    struct Wrap
    {
      PointType p;
    } w {p};
    auto pred2 = [](const Wrap& ww, const BoxType& bb) { return bb.contains(ww.p); };
    (void)pred;
    trav.traverse_tree(w, leafAction, pred2);
    counts[i] = c;
    total += c;
  }
  return total;
}

// the same for findBoundingBoxes' and findRays' predicates (spin/BVH.hpp:558-560, :527-531)
int64_t axref_bvh_count_boxes_omp(const AxrefBvh* h, const double* boxes_aos, int q, int32_t* counts, int nthreads)
{
  if(h->ndims != 3) return -1;
  using BoxType = axom::primal::BoundingBox<double, 3>;
  const auto trav = ((RefBvh<double, 3>*)h->impl)->bvh.getTraverser();
  const BoxType* qb = reinterpret_cast<const BoxType*>(boxes_aos);
  int64_t total = 0;
#ifdef _OPENMP
  if(nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel for schedule(dynamic, 1024) reduction(+ : total)
  for(int i = 0; i < q; ++i)
  {
    int c = 0;
    auto leafAction = [&c](std::int32_t, const std::int32_t*) { ++c; };
    auto pred = [](const BoxType& bb1, const BoxType& bb2) -> bool { return bb1.intersectsWith(bb2); };
    trav.traverse_tree(qb[i], leafAction, pred);
    counts[i] = c;
    total += c;
  }
  return total;
}

int64_t axref_bvh_count_rays_omp(const AxrefBvh* h, const double* origins_aos, const double* dirs_aos, int q, int32_t* counts, int nthreads)
{
  if(h->ndims != 3) return -1;
  using PointType = axom::primal::Point<double, 3>;
  using VectorType = axom::primal::Vector<double, 3>;
  using RayType = axom::primal::Ray<double, 3>;
  using BoxType = axom::primal::BoundingBox<double, 3>;
  const RefBvh<double, 3>& r = *(RefBvh<double, 3>*)h->impl;
  const auto trav = r.bvh.getTraverser();
  const double TOL = r.bvh.getTolerance();
  int64_t total = 0;
#ifdef _OPENMP
  if(nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel for schedule(dynamic, 1024) reduction(+ : total)
  for(int i = 0; i < q; ++i)
  {
    int c = 0;
    const RayType ray(PointType(origins_aos + (size_t)i * 3, 3), VectorType(dirs_aos + (size_t)i * 3, 3));  // normalises
    auto leafAction = [&c](std::int32_t, const std::int32_t*) { ++c; };
    auto pred = [TOL](const RayType& rr, const BoxType& bb) -> bool {
      PointType tmp;
      return axom::primal::detail::intersect_ray(rr, bb, tmp, TOL);
    };
    trav.traverse_tree(ray, leafAction, pred);
    counts[i] = c;
    total += c;
  }
  return total;
}

void* axref_sd_create(const double* x, const double* y, const double* z, int nnodes, const int32_t* conn, int ncells,
                      int nodes_per_cell, int watertight, int compute_sign)
{
  ensure_slic();
  RefSurface* s = new RefSurface();
  const axom::mint::CellType ct = nodes_per_cell == 3 ? axom::mint::TRIANGLE : axom::mint::QUAD;
  RefSurface::Mesh* m = new RefSurface::Mesh(3, ct, nnodes, ncells);
  s->mesh = m;
  for(int i = 0; i < nnodes; ++i) m->appendNode(x[i], y[i], z[i]);
  for(int c = 0; c < ncells; ++c)
  {
    axom::IndexType ids[4];
    for(int k = 0; k < nodes_per_cell; ++k) ids[k] = conn[(size_t)c * nodes_per_cell + k];
    m->appendCell(ids);
  }
  s->sd = new RefSurface::SD(s->mesh, watertight != 0, compute_sign != 0);
  return s;
}

// mixed triangle / quad surface: mint::UnstructuredMesh<MIXED_SHAPE> (SD_GetUcdMeshData, quest/SignedDistance.cpp:29-37)
void* axref_sd_create_mixed(const double* x, const double* y, const double* z, int nnodes, const int32_t* conn, const int32_t* offsets,
                            int ncells, int watertight, int compute_sign)
{
  ensure_slic();
  RefSurface* s = new RefSurface();
  RefSurface::MixedMesh* m = new RefSurface::MixedMesh(3, nnodes, ncells);
  s->mesh = m;
  for(int i = 0; i < nnodes; ++i) m->appendNode(x[i], y[i], z[i]);
  for(int c = 0; c < ncells; ++c)
  {
    const int nn = offsets[c + 1] - offsets[c];
    axom::IndexType ids[4];
    for(int k = 0; k < nn; ++k) ids[k] = conn[offsets[c] + k];
    m->appendCell(ids, nn == 3 ? axom::mint::TRIANGLE : axom::mint::QUAD);
  }
  s->sd = new RefSurface::SD(s->mesh, watertight != 0, compute_sign != 0);
  return s;
}

void axref_sd_destroy(void* h) { delete(RefSurface*)h; }

void axref_sd_compute(void* h, const double* qpts_aos, int npts, double* phi, double* cp, double* normals, int nthreads)
{
  using PointType = axom::primal::Point<double, 3>;
  using VectorType = axom::primal::Vector<double, 3>;
  const RefSurface& s = *(RefSurface*)h;
  const PointType* q = reinterpret_cast<const PointType*>(qpts_aos);
  PointType* cps = reinterpret_cast<PointType*>(cp);
  VectorType* nrm = reinterpret_cast<VectorType*>(normals);
  if(nthreads == 1)
  {
    s.sd->computeDistances(npts, q, phi, cps, nrm);
    return;
  }
  // chunked computeDistances from an external OpenMP loop (method is const)
#ifdef _OPENMP
  if(nthreads > 0) omp_set_num_threads(nthreads);
#endif
  const int chunk = 64;
  const int nchunks = (npts + chunk - 1) / chunk;
#pragma omp parallel for schedule(dynamic, 1)
  for(int c = 0; c < nchunks; ++c)
  {
    const int b = c * chunk;
    const int m = (b + chunk <= npts) ? chunk : npts - b;
    s.sd->computeDistances(m, q + b, phi + b, cps ? cps + b : nullptr, nrm ? nrm + b : nullptr);
  }
}

// primal::intersect(Triangle3, Triangle3, includeBoundary, EPS) (primal/operators/intersect.hpp:64-71) on n pairs
void axref_tri_tri_intersect(const double* tris1, const double* tris2, int n, int include_boundary, double eps, uint8_t* out)
{
  using PointType = axom::primal::Point<double, 3>;
  using Tri = axom::primal::Triangle<double, 3>;
  for(int i = 0; i < n; ++i)
  {
    const double* a = tris1 + (size_t)i * 9;
    const double* b = tris2 + (size_t)i * 9;
    const Tri t1(PointType {a[0], a[1], a[2]}, PointType {a[3], a[4], a[5]}, PointType {a[6], a[7], a[8]});
    const Tri t2(PointType {b[0], b[1], b[2]}, PointType {b[3], b[4], b[5]}, PointType {b[6], b[7], b[8]});
    out[i] = axom::primal::intersect(t1, t2, include_boundary != 0, eps) ? 1 : 0;
  }
}

// ---- leaf math on n independent items, forwarded to primal (tests/test_leaf_math.py) ----
void axref_closest_point_tri(const double* pts, const double* tris, int n, double eps, double* cp, int32_t* loc)
{
  using PointType = axom::primal::Point<double, 3>;
  using Tri = axom::primal::Triangle<double, 3>;
  for(int i = 0; i < n; ++i)
  {
    const double* t = tris + (size_t)i * 9;
    const Tri tri(PointType {t[0], t[1], t[2]}, PointType {t[3], t[4], t[5]}, PointType {t[6], t[7], t[8]});
    int l = 0;
    const PointType c = axom::primal::closest_point(PointType {pts[3 * i], pts[3 * i + 1], pts[3 * i + 2]}, tri, &l, eps);
    for(int d = 0; d < 3; ++d) cp[3 * i + d] = c[d];
    loc[i] = l;
  }
}
void axref_squared_distance_point_box(const double* pts, const double* boxes, int n, double* out)
{
  using PointType = axom::primal::Point<double, 3>;
  using BoxType = axom::primal::BoundingBox<double, 3>;
  for(int i = 0; i < n; ++i)
  {
    const double* b = boxes + (size_t)i * 6;
    BoxType bb;  // invalid unless lo <= hi: keep an invalid input invalid (the two-point constructor would swap it)
    if(b[0] <= b[3] && b[1] <= b[4] && b[2] <= b[5]) bb = BoxType(PointType {b[0], b[1], b[2]}, PointType {b[3], b[4], b[5]});
    out[i] = axom::primal::squared_distance(PointType {pts[3 * i], pts[3 * i + 1], pts[3 * i + 2]}, bb);
  }
}
void axref_intersect_ray_box(const double* rays, const double* boxes, int n, int normalize, double tol, uint8_t* out)
{
  using PointType = axom::primal::Point<double, 3>;
  using VectorType = axom::primal::Vector<double, 3>;
  using BoxType = axom::primal::BoundingBox<double, 3>;
  using RayType = axom::primal::Ray<double, 3>;
  for(int i = 0; i < n; ++i)
  {
    const double* r = rays + (size_t)i * 6;
    const double* b = boxes + (size_t)i * 6;
    const BoxType bb(PointType {b[0], b[1], b[2]}, PointType {b[3], b[4], b[5]});
    (void)normalize;  // the primal::Ray constructor always normalises (Ray.hpp:122-127)
    const RayType ray(PointType {r[0], r[1], r[2]}, VectorType {r[3], r[4], r[5]});
    PointType ip;
    out[i] = axom::primal::detail::intersect_ray(ray, bb, ip, tol) ? 1 : 0;  // as BVH::findRays does (spin/BVH.hpp:529-532)
  }
}
void axref_box_scale(double* boxes, int n, double scale)
{
  using PointType = axom::primal::Point<double, 3>;
  using BoxType = axom::primal::BoundingBox<double, 3>;
  for(int i = 0; i < n; ++i)
  {
    double* b = boxes + (size_t)i * 6;
    BoxType bb;
    if(b[0] <= b[3] && b[1] <= b[4] && b[2] <= b[5]) bb = BoxType(PointType {b[0], b[1], b[2]}, PointType {b[3], b[4], b[5]});
    bb.scale(scale);
    for(int d = 0; d < 3; ++d)
    {
      b[d] = bb.getMin()[d];
      b[3 + d] = bb.getMax()[d];
    }
  }
}

// quest::findTriMeshIntersectionsBVH<SEQ_EXEC, double> (quest/MeshTester.hpp:67-104)
int64_t axref_find_tri_mesh_intersections(const double* x, const double* y, const double* z, int nnodes, const int32_t* conn, int ncells,
                                          double threshold, int32_t** first, int32_t** second, int32_t** degenerate,
                                          int64_t* ndegenerate)
{
  ensure_slic();
  axom::slic::setLoggingMsgLevel(axom::slic::message::Warning);
  using UMesh = axom::mint::UnstructuredMesh<axom::mint::SINGLE_SHAPE>;
  UMesh mesh(3, axom::mint::TRIANGLE, nnodes, ncells);
  for(int i = 0; i < nnodes; ++i) mesh.appendNode(x[i], y[i], z[i]);
  for(int c = 0; c < ncells; ++c)
  {
    const axom::IndexType ids[3] = {conn[3 * c], conn[3 * c + 1], conn[3 * c + 2]};
    mesh.appendCell(ids);
  }
  std::vector<std::pair<int, int>> isect;
  std::vector<int> deg;
  axom::quest::findTriMeshIntersectionsBVH<SEQ_EXEC, double>(&mesh, isect, deg, threshold);
  int32_t* f = (int32_t*)malloc(sizeof(int32_t) * (isect.size() + 1));
  int32_t* s = (int32_t*)malloc(sizeof(int32_t) * (isect.size() + 1));
  int32_t* d = (int32_t*)malloc(sizeof(int32_t) * (deg.size() + 1));
  for(size_t i = 0; i < isect.size(); ++i)
  {
    f[i] = isect[i].first;
    s[i] = isect[i].second;
  }
  for(size_t i = 0; i < deg.size(); ++i) d[i] = deg[i];
  *first = f;
  *second = s;
  *degenerate = d;
  *ndegenerate = (int64_t)deg.size();
  return (int64_t)isect.size();
}

// the legacy process-global interface, end to end: quest::signed_distance_init(file) -> evaluate(x[],y[],z[],n,phi[])
// -> get_mesh_bounds -> finalize (quest/interface/signed_distance.hpp:117-319)
int axref_legacy_signed_distance(const char* stl_file, int closed_surface, int compute_sign, const double* x, const double* y,
                                 const double* z, int n, double* phi, double* lo, double* hi)
{
  namespace q = axom::quest;
  q::signed_distance_set_closed_surface(closed_surface != 0);
  q::signed_distance_set_compute_signs(compute_sign != 0);
  q::signed_distance_set_execution_space(q::SignedDistExec::CPU);
  const int rc = q::signed_distance_init(std::string(stl_file));
  if(rc != 0) return rc;
  q::signed_distance_evaluate(x, y, z, n, phi);
  q::signed_distance_get_mesh_bounds(lo, hi);
  q::signed_distance_finalize();
  return 0;
}

// quest::STLReader::read + getMesh, then optionally quest::weldTriMeshVertices(&mesh, eps) (eps > 0)
int axref_stl_read_weld(const char* stl_file, double eps, double** x, double** y, double** z, int32_t* num_nodes, int32_t** conn,
                        int32_t* num_cells)
{
  ensure_slic();
  using UMesh = axom::mint::UnstructuredMesh<axom::mint::SINGLE_SHAPE>;
  axom::quest::STLReader reader;
  reader.setFileName(stl_file);
  if(reader.read() != 0) return -1;
  UMesh* mesh = new UMesh(3, axom::mint::TRIANGLE);
  reader.getMesh(mesh);
  if(eps > 0.) axom::quest::weldTriMeshVertices(&mesh, eps);
  const int nn = mesh->getNumberOfNodes(), nc = mesh->getNumberOfCells();
  *x = (double*)malloc(sizeof(double) * (nn + 1));
  *y = (double*)malloc(sizeof(double) * (nn + 1));
  *z = (double*)malloc(sizeof(double) * (nn + 1));
  *conn = (int32_t*)malloc(sizeof(int32_t) * (3 * nc + 1));
  memcpy(*x, mesh->getCoordinateArray(0), sizeof(double) * nn);
  memcpy(*y, mesh->getCoordinateArray(1), sizeof(double) * nn);
  memcpy(*z, mesh->getCoordinateArray(2), sizeof(double) * nn);
  for(int c = 0; c < nc; ++c)
  {
    const axom::IndexType* ids = mesh->getCellNodeIDs(c);
    for(int k = 0; k < 3; ++k) (*conn)[3 * c + k] = ids[k];
  }
  *num_nodes = nn;
  *num_cells = nc;
  delete mesh;
  return 0;
}

// quest::DistributedClosestPoint needs Conduit + MPI, which are not in this image, so the class itself cannot be
// built.  Its per-rank step is a traversal of the real spin::BVH with two small lambdas; this entry point drives the
// REAL BVH / traverse_tree / primal::squared_distance with those lambdas restated from
// quest/detail/DistributedClosestPointImpl.hpp:883-903 (boxes = BoxType{pt}) and :1008-1060 (preset, checkMinDist,
// traversePredicate, write-back).  It pins the oracle's traversal order, pruning and tie-breaking.
extern "C++" {
template <int D>
struct RefDcp
{
  using PointType = axom::primal::Point<double, D>;
  using BoxType = axom::primal::BoundingBox<double, D>;
  axom::Array<PointType> pts;
  axom::Array<axom::IndexType> dom;
  std::unique_ptr<axom::spin::BVH<D, SEQ_EXEC, double>> bvh;
};

template <int D>
static void* ref_dcp_create(const double* p, const int32_t* dom, int n)
{
  using R = RefDcp<D>;
  R* r = new R();
  r->pts.resize(n);
  r->dom.resize(n);
  axom::Array<typename R::BoxType> boxes(n, n);
  for(int i = 0; i < n; ++i)
  {
    r->pts[i] = typename R::PointType(p + (size_t)i * D);
    r->dom[i] = dom[i];
    boxes[i] = typename R::BoxType {r->pts[i]};
  }
  if(n > 0)
  {
    r->bvh.reset(new axom::spin::BVH<D, SEQ_EXEC, double>());
    r->bvh->initialize(boxes.view(), n);
  }
  return r;
}

template <int D>
static void ref_dcp_local(const RefDcp<D>& o, int rank, double sqThresh, const double* q, int nq, int is_first, int32_t* cp_index,
                          int32_t* cp_dom, int32_t* cp_rank, double* cp_coords, double* cp_dist)
{
  using PointType = typename RefDcp<D>::PointType;
  using BoxType = typename RefDcp<D>::BoxType;
  using axom::primal::squared_distance;
  const double snan = std::numeric_limits<double>::signaling_NaN();
  for(int i = 0; i < nq; ++i)
  {
    if(is_first)
    {
      cp_rank[i] = cp_index[i] = cp_dom[i] = -1;
      for(int d = 0; d < D; ++d) cp_coords[(size_t)i * D + d] = snan;
      if(cp_dist) cp_dist[i] = snan;
    }
  }
  if(!o.bvh) return;
  auto it = o.bvh->getTraverser();
  for(int idx = 0; idx < nq; ++idx)
  {
    const PointType qpt(q + (size_t)idx * D);
    double sqDist = axom::numerics::floating_point_limits<double>::max();
    int pointIdx = -1, domainIdx = -1, minRank = -1;
    if(cp_rank[idx] >= 0)
    {
      sqDist = squared_distance(qpt, PointType(cp_coords + (size_t)idx * D));
      pointIdx = cp_index[idx];
      domainIdx = cp_dom[idx];
      minRank = cp_rank[idx];
    }
    auto checkMinDist = [&](std::int32_t current_node, const std::int32_t* leaf_nodes) {
      const int c = leaf_nodes[current_node];
      const double sq = squared_distance(qpt, o.pts[c]);
      if(sq < sqDist)
      {
        sqDist = sq;
        pointIdx = c;
        domainIdx = o.dom[c];
        minRank = rank;
      }
    };
    auto traversePredicate = [&](const PointType& p, const BoxType& bb) -> bool {
      auto sq = squared_distance(p, bb);
      return sq <= sqDist && sq <= sqThresh;
    };
    it.traverse_tree(qpt, checkMinDist, traversePredicate);
    if(minRank == rank)
    {
      cp_index[idx] = pointIdx;
      cp_dom[idx] = domainIdx;
      cp_rank[idx] = minRank;
      for(int d = 0; d < D; ++d) cp_coords[(size_t)idx * D + d] = o.pts[pointIdx][d];
      if(cp_dist) cp_dist[idx] = sqrt(sqDist);
    }
  }
}

}  // extern "C++"

void* axref_dcp_create(int ndims, const double* pts, const int32_t* domain_ids, int npts)
{
  ensure_slic();
  return ndims == 2 ? ref_dcp_create<2>(pts, domain_ids, npts) : ref_dcp_create<3>(pts, domain_ids, npts);
}
void axref_dcp_destroy(void* h, int ndims)
{
  if(ndims == 2)
    delete(RefDcp<2>*)h;
  else
    delete(RefDcp<3>*)h;
}
void axref_dcp_compute_local(void* h, int ndims, int rank, double sq_thresh, const double* q, int nq, int is_first, int32_t* cp_index,
                             int32_t* cp_dom, int32_t* cp_rank, double* cp_coords, double* cp_dist)
{
  if(ndims == 2)
    ref_dcp_local<2>(*(RefDcp<2>*)h, rank, sq_thresh, q, nq, is_first, cp_index, cp_dom, cp_rank, cp_coords, cp_dist);
  else
    ref_dcp_local<3>(*(RefDcp<3>*)h, rank, sq_thresh, q, nq, is_first, cp_index, cp_dom, cp_rank, cp_coords, cp_dist);
}

int axref_max_threads()
{
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

}  // extern "C"
