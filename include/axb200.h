/* axb200.h -- C ABI of the B200-native spin::BVH / quest::SignedDistance engine.
 *
 * Plain C: opaque handles, raw pointers and sizes, int status codes.  No torch / CUDA
 * types appear in any signature (a stream is passed as void*).  This is the boundary a
 * host code binds against; include/axom_b200/BVH.hpp and SignedDistance.hpp are the C++
 * header shims that keep the reference's class / method names on top of it, and
 * axom_b200/bvh.py is the ctypes mirror the tests use.
 *
 * Reference interfaces replaced (paths relative to /root/reference/src/axom):
 *   spin/BVH.hpp:129-419                    class BVH<NDIMS,ExecSpace,FloatType>
 *   spin/policy/LinearBVH.hpp:57-109        LinearBVHTraverser (device arrays)
 *   quest/SignedDistance.hpp:147-397        class SignedDistance<NDIMS,ExecSpace>
 *   quest/interface/signed_distance.hpp:117-319  (process-global C-style API)
 *   quest/MeshTester.hpp:67-104             findTriMeshIntersectionsBVH (broad + narrow phase)
 *   primal/operators/intersect.hpp:64-71    intersect(Triangle3, Triangle3, includeBoundary, EPS)
 *   quest/MarchingCubes.hpp:107-306         class MarchingCubes (iso-contour of a nodal field)
 *
 * Conventions
 *   - every function returns AXB_OK (0) or a negative axb_status; nothing throws or exits.
 *     axb_last_error() returns a thread-local message for the last failure.
 *   - IndexType is int32_t (reference default, core/Types.hpp:63-65); totals are int64_t.
 *   - calls are synchronous (results complete on return), like CUDA_EXEC<256>
 *     (core/execution/internal/cuda_exec.hpp:38-76), unless axb_*_set_async(h,1) was
 *     called: then work is only enqueued on the handle's stream.
 *   - "Indexable" inputs (raw AoS pointer or primal::ZipIndexable SoA) are described by an
 *     axb_array_desc: component c of item i lives at (char*)comp[c] + i*stride_bytes.
 *   - there is NO CPU fallback: without a CUDA device every compute entry point fails with
 *     AXB_ERR_NO_DEVICE.
 */
#ifndef AXB200_H_
#define AXB200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AXB_VERSION_STRING "0.1.0"

typedef enum axb_status
{
  AXB_OK = 0,
  AXB_ERR_BAD_ARG = -1,     /* null pointer, wrong size, wrong dimension ...               */
  AXB_ERR_CUDA = -2,        /* a CUDA runtime call or kernel failed (see axb_last_error)    */
  AXB_ERR_OVERFLOW = -3,    /* candidate total does not fit the reference's int32 offsets   */
  AXB_ERR_NOT_BUILT = -4,   /* query before initialize() / setMesh()                        */
  AXB_ERR_NO_DEVICE = -5,   /* no CUDA device: there is no CPU fallback                     */
  AXB_ERR_UNSUPPORTED = -6  /* valid request this build does not implement                  */
} axb_status;

typedef enum axb_memspace
{
  AXB_MEM_HOST = 0,  /* pageable or pinned host memory: staged through the handle's stream */
  AXB_MEM_DEVICE = 1, /* device (or managed) memory on the handle's device: used in place  */
  AXB_MEM_AUTO = 2    /* ask the driver (cudaPointerGetAttributes): device / managed pointers are
                         used in place, everything else is treated as host memory.  This is what
                         the C++ shims pass, so that -- like the reference (spin/BVH.hpp:198-203) --
                         the caller simply hands over pointers valid in its execution space.  */
} axb_memspace;

/* BVH_BUILD_OK of spin/BVH.hpp:39-43 */
#define AXB_BVH_BUILD_OK 0

/* Layout descriptor for an Indexable of primitives.
 *   boxes : ncomp = 2*D  -> min[0..D), max[0..D)          (primal::BoundingBox / ZipBoundingBox)
 *   points: ncomp = D                                      (primal::Point / ZipPoint)
 *   rays  : ncomp = 2*D  -> origin[0..D), direction[0..D)  (primal::Ray / ZipRay)
 * AoS array of T[ncomp]:  comp[c] = base + c*sizeof(T), stride_bytes = ncomp*sizeof(T).
 * SoA (ZipIndexable):     comp[c] = array c,            stride_bytes = sizeof(T).       */
typedef struct axb_array_desc
{
  const void* comp[6];
  int64_t stride_bytes;
  int32_t ncomp;
  int32_t memspace; /* axb_memspace */
} axb_array_desc;

/* The three arrays of LinearBVHTraverser (spin/policy/LinearBVH.hpp:105-108), device
 * pointers in the reference's exact layout (policy/LinearBVH.hpp:226-263):
 *   inner_nodes[2i+{0,1}]          = AABB of the left/right child of inner node i, as
 *                                    FloatType min[D],max[D] (primal::BoundingBox layout)
 *   inner_node_children[2i+{0,1}]  = 2*child (inner) or -(sorted_pos+1) (leaf)
 *   leaf_nodes[sorted_pos]         = original box id                                      */
typedef struct axb_traverser
{
  const void* inner_nodes;
  const int32_t* inner_node_children;
  const int32_t* leaf_nodes;
  int32_t num_leaves; /* after the N<=1 padding of spin/BVH.hpp:439-464 */
  int32_t ndims;
  int32_t fp_bytes;
} axb_traverser;

typedef struct axb_bvh axb_bvh;
typedef struct axb_sd axb_sd;
typedef struct axb_meshtester axb_meshtester;
typedef struct axb_dcp axb_dcp;

/* ---- library ------------------------------------------------------------------------ */
const char* axb_version(void);
/* Device memory the library frees stays cached in its OWN stream-ordered pool (not the device's default pool), up to
 * AXB_POOL_KEEP_MB (environment, default 16384); this returns everything above keep_bytes to the driver. */
int axb_trim_pool(int device, uint64_t keep_bytes);
const char* axb_last_error(void);
int axb_device_count(void);
const char* axb_status_string(int status);

/* ---- spin::BVH ---------------------------------------------------------------------- */
/* BVH() (spin/BVH.hpp:185).  ndims 2|3, fp_bytes 8 (FloatType = double) or 4 (float: every box,
 * point and ray component passed to this handle is then a 4-byte float, and all arithmetic is
 * done in float exactly as spin::BVH<D,ExecSpace,float> does), device = CUDA ordinal.
 * Defaults: scale 1.000123, tolerance = machine epsilon of FloatType (spin/BVH.hpp:410-412). */
int axb_bvh_create(axb_bvh** out, int ndims, int fp_bytes, int device);
int axb_bvh_destroy(axb_bvh* bvh);
int axb_bvh_set_stream(axb_bvh* bvh, void* cuda_stream); /* optional: run on the caller's stream */
int axb_bvh_set_async(axb_bvh* bvh, int enabled);
int axb_bvh_synchronize(axb_bvh* bvh);
int axb_bvh_set_scale_factor(axb_bvh* bvh, double scale); /* setScaleFactor :269 */
int axb_bvh_get_scale_factor(const axb_bvh* bvh, double* scale);
int axb_bvh_set_tolerance(axb_bvh* bvh, double tol); /* setTolerance :283 */
int axb_bvh_get_tolerance(const axb_bvh* bvh, double* tol);
/* initialize(boxes, numItems) :424-477.  Input boxes are copied, never reordered; may be
 * called again to rebuild.  Returns AXB_BVH_BUILD_OK. */
int axb_bvh_initialize(axb_bvh* bvh, const axb_array_desc* boxes, int32_t num_boxes);
int axb_bvh_is_initialized(const axb_bvh* bvh); /* 1 / 0 */
/* getBounds() :297-308: bounds of the *scaled* boxes; invalid box (DBL_MAX,-DBL_MAX) if unbuilt */
int axb_bvh_get_bounds(const axb_bvh* bvh, double* lo, double* hi);
/* getTraverser() :319 */
int axb_bvh_get_traverser(axb_bvh* bvh, axb_traverser* out);

/* findPoints / findBoundingBoxes / findRays (:341-398).  offsets and counts are caller
 * arrays of length num_queries in `out_memspace`; *candidates is allocated by the library in
 * `out_memspace` (release with axb_bvh_free_candidates) and *total receives its length.
 * Per-query candidate order is the reference's DFS visit order (bvh_traverse.hpp:66-154).
 * rays_normalized = 0 applies the primal::Ray constructor's normalisation (Ray.hpp:122-127),
 * 1 uses the directions verbatim (an array of already constructed Ray objects). */
int axb_bvh_find_points(axb_bvh* bvh, const axb_array_desc* points, int32_t num_queries, int32_t* offsets, int32_t* counts,
                        int out_memspace, int32_t** candidates, int64_t* total);
int axb_bvh_find_boxes(axb_bvh* bvh, const axb_array_desc* boxes, int32_t num_queries, int32_t* offsets, int32_t* counts,
                       int out_memspace, int32_t** candidates, int64_t* total);
int axb_bvh_find_rays(axb_bvh* bvh, const axb_array_desc* rays, int rays_normalized, int32_t num_queries, int32_t* offsets,
                      int32_t* counts, int out_memspace, int32_t** candidates, int64_t* total);
int axb_bvh_free_candidates(axb_bvh* bvh, int32_t* candidates, int memspace);
/* How find* produces the candidate lists (results are identical, including per-query order):
 *   0 (default) one traversal that counts and records (query, rank, candidate) hits in a chunked
 *               buffer, exclusive scan, flat scatter; queries are processed in Morton order;
 *   1           the reference's shape, count -> scan -> fill with two traversals, one thread per query
 *               (policy/LinearBVH.hpp:302-364);
 *   2           as 0 with a deliberately tiny hit buffer, which forces the overflow path (second
 *               traversal) -- a test hook. */
int axb_bvh_set_find_strategy(axb_bvh* bvh, int strategy);

/* Parity / debugging: copy the build artefacts to HOST buffers (any may be NULL).
 *   mcodes[n]          sorted 32-bit Morton codes        (RadixTree::m_mcodes)
 *   leaf_nodes[n]      sort permutation                  (RadixTree::m_leafs)
 *   inner_nodes[2(n-1)*2D] (FloatType), inner_children[2(n-1)]   (LinearBVH arrays, reference layout) */
int axb_bvh_num_leaves(const axb_bvh* bvh, int32_t* n);
int axb_bvh_copy_arrays(axb_bvh* bvh, uint32_t* mcodes, int32_t* leaf_nodes, void* inner_nodes /* FloatType */, int32_t* inner_children);

/* writeVtkFile(fileName) :405 -- the tree's boxes as an ASCII VTK unstructured grid with a "level" cell field; the file is
 * the reference's byte for byte (policy/LinearBVH.hpp:404-458).  A debugging aid: host code, O(N) text. */
int axb_bvh_write_vtk_file(axb_bvh* bvh, const char* file_name);

/* Device time (ms, CUDA events on the handle's stream) of the phases of the calls made since
 * profiling was (re-)enabled: the MEAN over those calls.  Enabling profiling resets the record.
 * names: "build.total" "build.bounds" "build.morton" "build.sort" "build.tree" "build.refit"
 *        "find.total" "find.sortq" "find.count" (the traversal) "find.scan" "find.fill" (scatter)                   */
int axb_bvh_set_profiling(axb_bvh* bvh, int enabled);
int axb_bvh_get_phase_ms(const axb_bvh* bvh, const char* name, double* ms);
/* number of kernels launched by this handle since creation (bench.py's gpu_launches) */
int axb_bvh_launch_count(const axb_bvh* bvh, int64_t* n);

/* ---- quest::SignedDistance ---------------------------------------------------------- */
/* SignedDistance(mesh, isWatertight, computeSign) + setMesh (quest/SignedDistance.hpp:410-504).
 * The mint::Mesh is reduced to what SD_GetUcdMeshData extracts (quest/SignedDistance.cpp:15-45):
 * SoA node coordinates and int32 connectivity.  nodes_per_cell 3 (triangles) or 4 (quads,
 * split (0,1,2),(0,2,3)); cell_node_offsets != NULL (num_cells+1 entries into cells_to_nodes)
 * selects a mixed triangle/quad mesh (mint::UnstructuredMesh<MIXED_SHAPE>, UcdMeshData :44-96) and
 * nodes_per_cell is ignored.  The mesh is copied and re-laid-out on the device. */
int axb_sd_create(axb_sd** out, int device, const double* x, const double* y, const double* z, int32_t num_nodes,
                  const int32_t* cells_to_nodes, const int32_t* cell_node_offsets, int32_t num_cells, int32_t nodes_per_cell,
                  int mesh_memspace, int is_watertight, int compute_sign);
int axb_sd_destroy(axb_sd* sd);
int axb_sd_set_stream(axb_sd* sd, void* cuda_stream);
int axb_sd_set_async(axb_sd* sd, int enabled);
int axb_sd_synchronize(axb_sd* sd);
/* computeDistances(npts, queryPts, outSgnDist, outClosestPts, outNormals) :527-605.
 * phi[npts]; closest_pts / normals are AoS double[3*npts] or NULL. */
int axb_sd_compute_distances(axb_sd* sd, const axb_array_desc* query_pts, int32_t npts, double* phi, double* closest_pts,
                             double* normals, int out_memspace);
/* getBVHTree() :337 -- borrowed handle, owned by the axb_sd */
int axb_sd_get_bvh(axb_sd* sd, axb_bvh** bvh);
/* bounding box of the mesh nodes (m_boxDomain; quest::signed_distance_get_mesh_bounds) */
int axb_sd_get_mesh_bounds(const axb_sd* sd, double* lo, double* hi);
/* traversal strategy: 0 = reference visiting order, one thread per query (bit-identical
 * closest points and normals); 1 = warp-cooperative packets (default; distances and closest
 * points bit-identical, normals may differ in the last ulp by summation order) */
int axb_sd_set_mode(axb_sd* sd, int mode);
int axb_sd_set_profiling(axb_sd* sd, int level); /* 0 off, 1 phase timers, 2 + work counters (adds a sync per call) */
/* query phases (profiling on): "query.total" "query.sortq" "query.kernel" (= "query.min" + "query.resolve") "query.minreduce";
 * setMesh phases (always recorded at creation): "setmesh.total" "setmesh.upload" "setmesh.cell_boxes" "setmesh.gather_soup"
 * "setmesh.obb_build" and the BVH build inside it "setmesh_build.total|bounds|morton|sort|agglo" */
int axb_sd_get_phase_ms(const axb_sd* sd, const char* name, double* ms);
int axb_sd_launch_count(const axb_sd* sd, int64_t* n);
/* work counters of the last query when profiling is enabled (device-side atomics in a
 * profiling build of the kernel): leaf tests and inner nodes visited */
int axb_sd_get_work_counters(const axb_sd* sd, int64_t* leaf_tests, int64_t* inner_visits);

/* ---- quest::findTriMeshIntersectionsBVH (the narrow phase downstream of findBoundingBoxes) ------ */
/* CandidateFinder<BVH>(surface_mesh, threshold) + initialize() (quest/detail/MeshTester_detail.hpp:158-199,
 * :313-340): the triangle mesh (SoA node coordinates, int32 connectivity, 3 nodes per cell) is reduced to
 * per-cell Triangle3 / AABB / degenerate flag on the device and a spin::BVH<3> is built over the AABBs with
 * the default scale factor. */
int axb_meshtester_create(axb_meshtester** out, int device, const double* x, const double* y, const double* z, int32_t num_nodes,
                          const int32_t* cells_to_nodes, int32_t num_cells, int mesh_memspace);
int axb_meshtester_destroy(axb_meshtester* mt);
/* findTriMeshIntersections (:201-307): pairs (first[k], second[k]), first < second, of triangles whose AABBs
 * overlap (findBoundingBoxes with the mesh's own AABBs) AND for which
 * primal::intersect(tri_first, tri_second, includeBoundary = false, intersection_threshold) holds.  The BVH walk,
 * the i < j filter and the exact test run in ONE kernel; the candidate list is never materialised.  Pairs are
 * returned in the reference's SEQ_EXEC order (first ascending, then the candidate's DFS order).  *first and
 * *second are allocated in `out_memspace` (HOST or DEVICE); release with axb_meshtester_free. */
int axb_meshtester_find_intersections(axb_meshtester* mt, double intersection_threshold, int out_memspace, int32_t** first,
                                      int32_t** second, int64_t* num_pairs);
/* indices of the degenerate triangles (Triangle::degenerate(), primal/geometry/Triangle.hpp:326-330), ascending */
int axb_meshtester_get_degenerate(axb_meshtester* mt, int out_memspace, int32_t** indices, int64_t* n);
int axb_meshtester_free(axb_meshtester* mt, int32_t* p, int memspace);
/* the BVH over the triangle AABBs -- borrowed handle (profiling: "find.count" is the fused walk) */
int axb_meshtester_get_bvh(axb_meshtester* mt, axb_bvh** bvh);
/* primal::intersect(Triangle<double,3>, Triangle<double,3>, includeBoundary, EPS) (primal/operators/intersect.hpp:64-71)
 * on n explicit pairs: tris1 / tris2 are AoS double[9] per triangle, out[i] = 0 / 1; all three live in `memspace`. */
int axb_tri_tri_intersect(int device, const double* tris1, const double* tris2, int64_t n, int memspace, int include_boundary, double eps,
                          uint8_t* out);

/* ---- leaf arithmetic of the path on n independent items (so the reference's own unit tests can be run against the
 * device functions the query kernels use).  All arrays live in `memspace` (host buffers are staged); synchronous. ---- */
/* primal::closest_point(Point, Triangle, int* loc, EPS) (primal/operators/closest_point.hpp:162-290):
 * pts double[3] per item, tris double[9] (A, B, C); cp double[3], loc int32 (0/1/2 vertex, -1/-2/-3 edge AB/BC/CA, 3 face) */
int axb_closest_point_tri(int device, const double* pts, const double* tris, int64_t n, int memspace, double eps, double* cp, int32_t* loc);
/* primal::squared_distance(Point, BoundingBox) (primal/operators/squared_distance.hpp:77-100): boxes double[6] = min, max;
 * an invalid box gives DBL_MAX */
int axb_squared_distance_point_box(int device, const double* pts, const double* boxes, int64_t n, int memspace, double* out);
/* the findRays predicate primal::detail::intersect_ray(Ray, BoundingBox, ip, tol) (spin/BVH.hpp:529-532,
 * primal/operators/detail/intersect_ray_impl.hpp:321-351): rays double[6] = origin, direction; rays_normalized = 0
 * applies the primal::Ray constructor's normalisation (primal/geometry/Ray.hpp:122-127) */
int axb_intersect_ray_box(int device, const double* rays, const double* boxes, int64_t n, int memspace, int rays_normalized, double tol,
                          uint8_t* out);
/* BoundingBox::scale(scale_factor) (primal/geometry/BoundingBox.hpp:548-561) as transform_boxes applies it
 * (spin/internal/linear_bvh/build_radix_tree.hpp:85-99): boxes_in / boxes_out double[6] per item */
int axb_box_scale(int device, const double* boxes_in, int64_t n, int memspace, double scale_factor, double* boxes_out);

/* ---- quest::DistributedClosestPoint, the per-rank step ---------------------------------------------- */
/* DistributedClosestPointExec<DIM, ExecSpace> (quest/detail/DistributedClosestPointImpl.hpp:551-1095) without its
 * Conduit / MPI plumbing: the object "mesh" is a point cloud (all local domains flattened, one domain id per point),
 * the BVH is built over one zero-size box per point (:883-903), and computeLocalClosestPoints (:905-1079) updates
 * the query block's state arrays -- the xferDom fields cp_index (flattened local point index), cp_domain_index,
 * cp_rank, cp_coords (interleaved) and cp_distance -- only where this rank improves on what earlier ranks of the
 * ring left there (strict <: the first point visited wins a tie).  The ring itself (or its NVSwitch replacement:
 * all ranks search all queries, then MIN-reduce with a ring-order tie-break) is host logic,
 * axom_b200/distributed_closest_point.py. */
int axb_dcp_create(axb_dcp** out, int ndims, int device);
int axb_dcp_destroy(axb_dcp* dcp);
int axb_dcp_set_object_points(axb_dcp* dcp, const double* coords_interleaved, const int32_t* domain_ids, int32_t num_points,
                              int memspace);                                              /* importObjectPoints :590-649 */
int axb_dcp_generate_bvh_tree(axb_dcp* dcp);                                               /* generateBVHTree :651-668    */
int axb_dcp_set_squared_distance_threshold(axb_dcp* dcp, double sq_threshold);             /* :297-301, default DBL_MAX   */
/* search strategy: 1 (default) nearest-first with an explicit sorted-position tie-break, 0 the reference's left-first
 * traversal (thousands of node visits per query on large clouds); results are bit-identical */
int axb_dcp_set_mode(axb_dcp* dcp, int mode);
int axb_dcp_get_bvh(axb_dcp* dcp, axb_bvh** bvh); /* borrowed; its getBounds() is what gatherBVHRoots exchanges (:671-677) */
int axb_dcp_compute_local_closest_points(axb_dcp* dcp, int rank, const double* query_coords_interleaved, int32_t num_queries,
                                         int is_first, int32_t* cp_index, int32_t* cp_domain_index, int32_t* cp_rank,
                                         double* cp_coords, double* cp_distance /* may be NULL */, int memspace);

/* The first-visit search restricted to points within a per-query squared-distance bound another rank already achieved
 * (device arrays only): state arrays are initialised and an entry is filled only if this rank holds a point with squared
 * distance <= bound_sq[i], ties included.  A building block of the collective form of the ring. */
int axb_dcp_compute_bounded_closest_points(axb_dcp* dcp, int rank, const double* query_coords_interleaved, int32_t num_queries,
                                           const double* bound_sq, int32_t* cp_index, int32_t* cp_domain_index, int32_t* cp_rank,
                                           double* cp_coords, double* cp_distance);

/* ---- the exchange steps of the distributed cases: NCCL over NVLink / NVSwitch, inside the library ---------------- */
/* One communicator per rank (= per process = per GPU).  Rank 0 obtains an id and hands it to the other ranks by any
 * means the host code has (MPI_Bcast, a file, a socket); every rank then calls axb_comm_create -- collectively, like
 * ncclCommInitRank, which it wraps.  NCCL is bound at run time: the copy already mapped into the process (a torch
 * process has one) or libnccl.so.2 on the loader path, or the file AXB_NCCL_LIB names.  What this replaces: the
 * MPI_Comm the reference's DistributedClosestPoint is given (quest/DistributedClosestPoint.hpp:95-100) and the
 * MPI_Allgather / Isend / Irecv traffic of quest/detail/DistributedClosestPointImpl.hpp:687-693, :737-851. */
#define AXB_COMM_ID_BYTES 128
typedef struct axb_comm axb_comm;
int axb_comm_get_unique_id(uint8_t id[AXB_COMM_ID_BYTES]);
int axb_comm_create(axb_comm** out, int nranks, int rank, const uint8_t id[AXB_COMM_ID_BYTES], int device);
int axb_comm_destroy(axb_comm* comm);
int axb_comm_get_rank(const axb_comm* comm, int* rank, int* nranks);
/* payload bytes this rank has contributed to collectives and the number of collectives issued, since creation */
int axb_comm_get_traffic(const axb_comm* comm, int64_t* bytes, int64_t* collectives);
const char* axb_comm_library(void); /* which NCCL was bound, and its version ("" if none) */
/* in-place elementwise reduction of a DEVICE array over the ranks on `cuda_stream`; op 0 = MIN, 1 = MAX, 2 = SUM */
int axb_comm_allreduce_f64(axb_comm* comm, double* device_buf, int64_t n, int op, void* cuda_stream);

/* BASELINE config C5 -- the surface is PARTITIONED over the ranks (each rank's axb_sd holds one part, created with
 * compute_sign = 0), every rank passes the SAME query points, and every rank receives the distance to the whole surface:
 * the query kernels write the partial (unsigned) distances into the buffer an ncclAllReduce(MIN, double) then reduces in
 * place on the handle's stream.  Between the sample pass and the search proper the ranks also MIN-reduce a per-query bound
 * (the distance to each part's sample point, + the largest triangle diameter), so a part far from a query prunes it at
 * once; the result is unchanged.  Every rank must pass the same points.  Phase timers: "query.kernel",
 * "query.bound_exchange", "query.minreduce". */
int axb_sd_compute_distances_minreduce(axb_sd* sd, axb_comm* comm, const axb_array_desc* query_pts, int32_t npts, double* dist,
                                       int out_memspace);

/* The same partitioned-surface query on ONE GPU, parts evaluated in turn: `dist` is in/out -- on entry the distance to the
 * parts evaluated so far (DBL_MAX where none), on exit min(entry, distance to this handle's part).  The entry value bounds the
 * search, so a part that is farther than what is already known costs a few node visits per query.  compute_sign = 0 handles. */
int axb_sd_update_min_distances(axb_sd* sd, const axb_array_desc* query_pts, int32_t npts, double* dist_inout, int memspace);

/* DistributedClosestPoint::computeClosestPoints (quest/DistributedClosestPoint.hpp:157-166, DistributedClosestPointImpl.hpp
 * :737-851), the whole of it: every rank passes ITS OWN query points (num_queries may be 0 and differ per rank) and gets,
 * for each, the nearest object point of the whole machine -- the five xferDom fields, with the reference's tie rule (first
 * rank in ring order from the query's owner; inside a rank the first point in traversal order) and -1 / signalling NaN where
 * no rank holds a point within the distance threshold.  Collective: every rank of `comm` must call it.  comm == NULL is the
 * single-rank case (no exchange).  Any output pointer may be NULL; arrays live in `memspace`.
 * Phase timers (axb_bvh_get_phase_ms on axb_dcp_get_bvh's handle): "dcpx.total" "dcpx.gather" "dcpx.search1"
 * "dcpx.bound_allreduce" "dcpx.search2" "dcpx.combine" "dcpx.exchange". */
int axb_dcp_compute_closest_points(axb_dcp* dcp, axb_comm* comm, const double* query_coords_interleaved, int32_t num_queries, int memspace,
                                   int32_t* cp_index, int32_t* cp_domain_index, int32_t* cp_rank, double* cp_coords, double* cp_distance);

/* ---- quest::MarchingCubes (the consumer of the distance field: iso-contour of a nodal function) ------------- */
/* One structured domain, as MarchingCubesImpl::setDomain / setFunctionField see it through MeshViewUtil
 * (quest/detail/MarchingCubesImpl.hpp:100-145, quest/MeshViewUtil.hpp:454-482,560-607): ghost-free views, i.e. every
 * pointer addresses node (0,0[,0]) / cell (0,0[,0]) of the REAL mesh and strides are in elements (any ghost layers
 * are skipped by the caller through the pointer, exactly what ArrayView::subspan(offsets, realShape) does).
 * Entries [2] of the 3-vectors are ignored for ndims == 2. */
typedef struct axb_mc_domain
{
  int64_t cell_shape[3];     /* topologies/<t>/elements/dims/{i,j,k}: real cells per direction            */
  const double* coords[3];   /* coordsets/<c>/values/{x,y,z}: one array per direction (not interleaved)   */
  int64_t coords_strides[3]; /* elements/dims/strides, default (1, ni+1, (ni+1)(nj+1))                    */
  const double* fcn;         /* fields/<fcn>/values: the nodal function (double)                          */
  int64_t fcn_strides[3];    /* fields/<fcn>/strides; their order fixes the parent-cell numbering         */
  const int32_t* mask;       /* fields/<mask>/values (cell-centred int32) or NULL                         */
  int64_t mask_strides[3];
  int64_t domain_id;         /* state/domain_id, else the domain's position (MarchingCubesSingleDomain.cpp:168-176) */
} axb_mc_domain;

typedef struct axb_mc axb_mc;
/* MarchingCubes(runtimePolicy, allocatorID, dataParallelism) (quest/MarchingCubes.hpp:112-114): ndims 2 | 3.  The two
 * data-parallel variants of the reference give the same output; there is one device path. */
int axb_mc_create(axb_mc** out, int ndims, int device);
int axb_mc_destroy(axb_mc* mc);
int axb_mc_set_stream(axb_mc* mc, void* cuda_stream);
/* setMesh(bpMesh, topologyName, maskField) + setFunctionField(fcnField) (MarchingCubes.cpp:47-105).  Arrays in
 * `memspace` HOST are staged to the device here (once); DEVICE arrays are used in place and must outlive the
 * compute calls, like the Blueprint node the reference caches views of.  The function's strides must be unique
 * (core/MDMapping.hpp:221-243 aborts otherwise) and the cell count must fit IndexType (int32). */
int axb_mc_set_mesh(axb_mc* mc, const axb_mc_domain* domains, int32_t num_domains, int memspace);
int axb_mc_set_mask_value(axb_mc* mc, int mask_val); /* setMaskValue, default 1 (quest/MarchingCubes.hpp:148) */
/* computeIsocontour(contourVal) (MarchingCubes.cpp:107-147): ADDS the contour to what earlier calls produced. */
int axb_mc_compute_isocontour(axb_mc* mc, double contour_val);
int axb_mc_get_contour_cell_count(const axb_mc* mc, int64_t* n); /* getContourCellCount / getContourFacetCount */
int axb_mc_get_contour_node_count(const axb_mc* mc, int64_t* n); /* getContourNodeCount = cells * ndims          */
/* getContourFacetCorners [cells][ndims], getContourNodeCoords [nodes][ndims], getContourFacetParents [cells],
 * getContourFacetDomainIds [cells] (quest/MarchingCubes.hpp:203-250): borrowed DEVICE views, valid until the next
 * compute / clear / destroy.  Any output pointer may be NULL. */
int axb_mc_get_contour_views(axb_mc* mc, const int32_t** facet_node_ids, const double** node_coords, const int32_t** facet_parent_ids,
                             const int32_t** facet_domain_ids);
/* the same four arrays copied into caller buffers in `memspace` (what populateContourMesh :169-233 appends to the
 * mint::UnstructuredMesh: nodes, cells, cellIdField, domainIdField).  Any pointer may be NULL. */
int axb_mc_copy_contour(axb_mc* mc, int memspace, int32_t* facet_node_ids, double* node_coords, int32_t* facet_parent_ids,
                        int32_t* facet_domain_ids);
int axb_mc_clear_output(axb_mc* mc); /* clearOutput (MarchingCubes.cpp:156-163) */
int axb_mc_set_profiling(axb_mc* mc, int enabled); /* phases: "mc.mark", "mc.count" (tile counts + scan), "mc.emit" */
int axb_mc_get_phase_ms(const axb_mc* mc, const char* phase, double* ms);
int axb_mc_launch_count(const axb_mc* mc, int64_t* n);

#ifdef __cplusplus
}
#endif
#endif /* AXB200_H_ */
