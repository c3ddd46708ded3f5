// api.cu -- host side of libaxb200: handles, device memory, staging, kernel launches and
// the extern "C" entry points declared in include/axb200.h.
//
// There is no CPU fallback anywhere in this file: every compute entry point needs a CUDA
// device and fails with AXB_ERR_NO_DEVICE / AXB_ERR_CUDA otherwise.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <map>
#include <sstream>
#include <mutex>
#include <string>
#include <vector>

#include "build.cuh"
#include "common.cuh"
#include "radix_sort.cuh"
#include "sd.cuh"
#include "sd_fast.cuh"
#include "sd_two.cuh"
#include "meshtester.cuh"
#include "leafmath.cuh"
#include "dcp.cuh"
#include "mc.cuh"
#include "traverse.cuh"

namespace axb
{
static thread_local std::string g_last_error;
void set_last_error(const std::string& msg) { g_last_error = msg; }

static int fail(int status, const std::string& msg)
{
  set_last_error(msg);
  return status;
}

//------------------------------------------------------------------------------------------
// The library's own stream-ordered memory pool, one per device.  Freed blocks stay in it (the per-call candidate
// arrays are hundreds of MB: handing them back to the driver at every synchronisation, the default, turns each call
// into a fresh allocation) up to a bounded release threshold (AXB_POOL_KEEP_MB, default 16384) -- the device's DEFAULT
// pool, which other allocators in the process share, is left alone.  axb_trim_pool() gives the cached blocks back.
//------------------------------------------------------------------------------------------
constexpr int kMaxDevices = 64;
static cudaMemPool_t g_pool[kMaxDevices] = {};
static std::mutex g_pool_mutex;

static cudaError_t pool_for_current_device(cudaMemPool_t* out)
{
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if(e != cudaSuccess) return e;
  if(dev < 0 || dev >= kMaxDevices) return cudaErrorInvalidDevice;
  std::lock_guard<std::mutex> lock(g_pool_mutex);
  if(!g_pool[dev])
  {
    cudaMemPoolProps props;
    memset(&props, 0, sizeof(props));
    props.allocType = cudaMemAllocationTypePinned;
    props.handleTypes = cudaMemHandleTypeNone;
    props.location.type = cudaMemLocationTypeDevice;
    props.location.id = dev;
    e = cudaMemPoolCreate(&g_pool[dev], &props);
    if(e != cudaSuccess)
    {
      g_pool[dev] = nullptr;
      return e;
    }
    unsigned long long keep = 16384ull << 20;
    if(const char* env = getenv("AXB_POOL_KEEP_MB")) keep = (unsigned long long)std::max(0ll, atoll(env)) << 20;
    cudaMemPoolSetAttribute(g_pool[dev], cudaMemPoolAttrReleaseThreshold, &keep);
  }
  *out = g_pool[dev];
  return cudaSuccess;
}

static cudaError_t axb_malloc_async(void** p, size_t bytes, cudaStream_t s)
{
  cudaMemPool_t pool = nullptr;
  cudaError_t e = pool_for_current_device(&pool);
  if(e != cudaSuccess) return e;
  return cudaMallocFromPoolAsync(p, bytes, pool, s);
}

//------------------------------------------------------------------------------------------
// stream-ordered device buffer (grow-only, reused across calls)
//------------------------------------------------------------------------------------------
struct DevBuf
{
  void* p = nullptr;
  size_t cap = 0;
  int reserve(size_t bytes, cudaStream_t s)
  {
    if(bytes <= cap && p) return AXB_OK;
    if(p) AXB_CUDA_TRY(cudaFreeAsync(p, s));
    p = nullptr;
    cap = 0;
    const size_t want = bytes ? bytes : 16;
    AXB_CUDA_TRY(axb_malloc_async(&p, want, s));
    cap = want;
    return AXB_OK;
  }
  void release(cudaStream_t s)
  {
    if(p) cudaFreeAsync(p, s);
    p = nullptr;
    cap = 0;
  }
  template <typename U>
  U* as() const
  {
    return reinterpret_cast<U*>(p);
  }
};

//------------------------------------------------------------------------------------------
// per-handle execution context: stream, profiling events, launch counter
//------------------------------------------------------------------------------------------
struct Phase
{
  std::string name;
  cudaEvent_t a = nullptr, b = nullptr;
};

struct Ctx
{
  int device = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  bool async = false;
  bool profiling = false;
  int64_t launches = 0;
  std::vector<Phase> phases;  // recorded, not yet resolved (events pending)
  struct Acc
  {
    double sum = 0.0;
    long long calls = 0;
    double last = 0.0;
  };
  std::map<std::string, Acc> acc;  // resolved times since profiling was (re-)enabled

  int init(int dev)
  {
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if(e != cudaSuccess || count <= 0)
    {
      cudaGetLastError();
      return fail(AXB_ERR_NO_DEVICE, "no CUDA device available (libaxb200 has no CPU fallback)");
    }
    if(dev < 0 || dev >= count) return fail(AXB_ERR_BAD_ARG, "device ordinal out of range");
    device = dev;
    AXB_CUDA_TRY(cudaSetDevice(device));
    AXB_CUDA_TRY(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
    own_stream = true;
    cudaMemPool_t pool = nullptr;
    AXB_CUDA_TRY(pool_for_current_device(&pool));  // the library's own pool (see above); the default pool is not touched
    return AXB_OK;
  }
  void drop_phases()
  {
    for(auto& ph : phases)
    {
      if(ph.a) cudaEventDestroy(ph.a);
      if(ph.b) cudaEventDestroy(ph.b);
    }
    phases.clear();
  }
  void destroy()
  {
    drop_phases();
    if(own_stream && stream) cudaStreamDestroy(stream);
    stream = nullptr;
  }
  int bind() { AXB_CUDA_TRY(cudaSetDevice(device)); return AXB_OK; }
  void set_profiling(bool on)
  {
    drop_phases();
    acc.clear();
    profiling = on;
  }
  void begin_call() { }
  int phase_begin(const char* name)
  {
    if(!profiling) return -1;
    Phase ph;
    ph.name = name;
    cudaEventCreate(&ph.a);
    cudaEventCreate(&ph.b);
    cudaEventRecord(ph.a, stream);
    phases.push_back(ph);
    return (int)phases.size() - 1;
  }
  void phase_end(int id)
  {
    if(id >= 0) cudaEventRecord(phases[id].b, stream);
  }
  // turn pending event pairs into milliseconds (synchronises the stream)
  void resolve()
  {
    if(phases.empty()) return;
    cudaStreamSynchronize(stream);
    for(auto& ph : phases)
    {
      float t = 0.f;
      if(ph.a && ph.b && cudaEventElapsedTime(&t, ph.a, ph.b) == cudaSuccess)
      {
        Acc& a = acc[ph.name];
        a.sum += t;
        a.calls += 1;
        a.last = t;
      }
    }
    cudaGetLastError();
    drop_phases();
  }
  int sync() { AXB_CUDA_TRY(cudaStreamSynchronize(stream)); return AXB_OK; }
  int finish_call()
  {
    if(!async)
    {
      AXB_TRY(sync());
      if(profiling) resolve();
    }
    return AXB_OK;
  }
};

struct ScopedPhase
{
  Ctx& c;
  int id;
  ScopedPhase(Ctx& ctx, const char* name) : c(ctx), id(ctx.phase_begin(name)) { }
  ~ScopedPhase() { c.phase_end(id); }
};

#define AXB_LAUNCH(ctx, kernel, grid, block, ...)                    \
  do                                                                 \
  {                                                                  \
    kernel<<<(grid), (block), 0, (ctx).stream>>>(__VA_ARGS__);       \
    ++(ctx).launches;                                                \
    AXB_CUDA_TRY(cudaGetLastError());                                \
  } while(0)

#define AXB_LAUNCH_SMEM(ctx, kernel, grid, block, smem, ...)             \
  do                                                                    \
  {                                                                     \
    kernel<<<(grid), (block), (smem), (ctx).stream>>>(__VA_ARGS__);     \
    ++(ctx).launches;                                                   \
    AXB_CUDA_TRY(cudaGetLastError());                                   \
  } while(0)

static inline int blocks_for(long long n, int block) { return (int)std::max<long long>(1, (n + block - 1) / block); }
static inline int capped_grid(long long n, int block, int per_sm = 8)
{
  return (int)std::min<long long>(blocks_for(n, block), (long long)kNumSMsB200 * per_sm);
}

//------------------------------------------------------------------------------------------
// staging of "Indexable" inputs: host arrays are copied to the device (AoS block or one
// array per component), device arrays are used in place.
//------------------------------------------------------------------------------------------
// AXB_MEM_AUTO -> AXB_MEM_HOST | AXB_MEM_DEVICE by asking the driver about the pointer
static int resolve_memspace(int memspace, const void* p)
{
  if(memspace != AXB_MEM_AUTO) return memspace;
  if(!p) return AXB_MEM_HOST;
  cudaPointerAttributes a;
  if(cudaPointerGetAttributes(&a, p) != cudaSuccess)
  {
    cudaGetLastError();
    return AXB_MEM_HOST;
  }
  return (a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged) ? AXB_MEM_DEVICE : AXB_MEM_HOST;
}

template <int NC>
static int stage_desc(Ctx& ctx, const axb_array_desc* in, long long count, size_t elem, DevBuf& stage, Desc<NC>& out)
{
  if(!in) return fail(AXB_ERR_BAD_ARG, "null array descriptor");
  if(in->ncomp != NC) return fail(AXB_ERR_BAD_ARG, "array descriptor has the wrong number of components for this dimension");
  for(int c = 0; c < NC; ++c)
    if(count > 0 && in->comp[c] == nullptr) return fail(AXB_ERR_BAD_ARG, "null component pointer in array descriptor");
  if(count > 0 && in->stride_bytes < (int64_t)elem) return fail(AXB_ERR_BAD_ARG, "array descriptor stride smaller than the element size");
  const int space = resolve_memspace(in->memspace, in->comp[0]);
  if(space == AXB_MEM_DEVICE || count == 0)
  {
    for(int c = 0; c < NC; ++c) out.comp[c] = reinterpret_cast<const char*>(in->comp[c]);
    out.stride = in->stride_bytes;
    return AXB_OK;
  }
  if(space != AXB_MEM_HOST) return fail(AXB_ERR_BAD_ARG, "unknown memspace");
  const char* base = reinterpret_cast<const char*>(in->comp[0]);
  bool aos = (in->stride_bytes == (int64_t)(NC * elem));
  for(int c = 0; c < NC && aos; ++c) aos = (reinterpret_cast<const char*>(in->comp[c]) == base + c * elem);
  if(aos)
  {
    const size_t bytes = (size_t)count * NC * elem;
    AXB_TRY(stage.reserve(bytes, ctx.stream));
    AXB_CUDA_TRY(cudaMemcpyAsync(stage.p, base, bytes, cudaMemcpyHostToDevice, ctx.stream));
    for(int c = 0; c < NC; ++c) out.comp[c] = stage.as<char>() + c * elem;
    out.stride = (long long)(NC * elem);
    return AXB_OK;
  }
  // component-wise (ZipIndexable SoA or a general strided view): one dense device array each
  const size_t comp_bytes = (size_t)count * elem;
  const size_t comp_pitch = (comp_bytes + 255) & ~(size_t)255;
  AXB_TRY(stage.reserve(comp_pitch * NC, ctx.stream));
  for(int c = 0; c < NC; ++c)
  {
    char* dst = stage.as<char>() + c * comp_pitch;
    if(in->stride_bytes == (int64_t)elem)
      AXB_CUDA_TRY(cudaMemcpyAsync(dst, in->comp[c], comp_bytes, cudaMemcpyHostToDevice, ctx.stream));
    else
      AXB_CUDA_TRY(cudaMemcpy2DAsync(dst, elem, in->comp[c], (size_t)in->stride_bytes, elem, (size_t)count, cudaMemcpyHostToDevice,
                                     ctx.stream));
    out.comp[c] = dst;
  }
  out.stride = (long long)elem;
  return AXB_OK;
}

}  // namespace axb

#include "comm.cuh"

using namespace axb;

//==========================================================================================
// spin::BVH
//==========================================================================================
struct axb_bvh
{
  Ctx ctx;
  int ndims = 3;
  int fp_bytes = 8;
  double scale = 1.000123;      // DEFAULT_SCALE_FACTOR, spin/BVH.hpp:410
  double tol = DBL_EPSILON;     // DEFAULT_TOLERANCE, spin/BVH.hpp:411-412
  bool built = false;
  int n = 0;       // leaves after padding
  int n_in = 0;    // boxes supplied by the caller
  double bounds_lo[3] = {DBL_MAX, DBL_MAX, DBL_MAX};
  double bounds_hi[3] = {-DBL_MAX, -DBL_MAX, -DBL_MAX};

  DevBuf nodes, leaf_nodes, leaf_parent, node_range, keys_a, keys_b, state, sort_scratch, stage_in, agglo_slots, agglo_flags, agglo_open;
  int agglo_block = 256;      // leaves per block of agglo_kernel (AXB_AGGLO_BLOCK = 128 | 256 | 512)
  bool legacy_build = false;  // AXB_BUILD_LEGACY=1: tree_kernel + refit_kernel instead of agglo_kernel
  unsigned long long* sorted_keys = nullptr;  // points into keys_a or keys_b
  DevBuf ref_inner_nodes, ref_children;       // reference-layout view, built lazily
  bool ref_view_valid = false;
  DevBuf fnodes;                              // compact 64-byte records for point / box walks (3-D double), built lazily
  bool fnodes_valid = false;
  bool find_compact = true;                   // AXB_FIND_COMPACT=0: walk the 128-byte records
  DevBuf q_stage, q_counts, q_offsets, q_tiles, q_total;
  DevBuf f_keys_a, f_keys_b, f_scratch, f_perm, f_pairs, f_unused, f_cursor;
  long long pair_hint[4] = {0, 0, 0, 0};  // candidates found by the last find* call of each kind (sizes the pair buffer)
  int walk_blocks_per_sm[4] = {0, 0, 0, 0};  // resident blocks/SM of the walk kernel of each kind
  int find_strategy = 0;  // 0 = single traversal + scatter (default), 1 = the reference's count / fill double traversal

  void release_all()
  {
    cudaStream_t s = ctx.stream;
    for(DevBuf* b : {&nodes, &leaf_nodes, &leaf_parent, &node_range, &keys_a, &keys_b, &state, &sort_scratch, &stage_in, &agglo_slots, &agglo_flags, &agglo_open, &fnodes, &ref_inner_nodes,
                     &ref_children, &q_stage, &q_counts, &q_offsets, &q_tiles, &q_total, &f_keys_a, &f_keys_b, &f_scratch, &f_perm, &f_pairs,
                     &f_unused, &f_cursor})
      b->release(s);
  }
};

namespace
{
// LSD radix sort of the 64-bit keys on bits [32, 64): 4 onesweep passes.
// first_pass = 1 skips the lowest digit: the order of QUERIES only decides who is processed next to whom, and 24 bits of
// Morton code (8 per dimension) are as coherent as 32 -- queries that share them stay in index order (the sort is stable).
int sort_keys_generic(Ctx& ctx, unsigned long long* src, unsigned long long* dst, int n, uint32_t* ghist /* 4*256, filled */,
                      uint32_t* tile_counters, uint32_t* lookback, unsigned long long** sorted, int first_pass = 0)
{
  const int tiles = rsort::num_tiles(n);
  if(first_pass > 0)
    if(const char* e = getenv("AXB_QSORT_FIRST_PASS")) first_pass = std::max(0, std::min(atoi(e), rsort::MAX_PASSES - 1));
  for(int p = first_pass; p < rsort::MAX_PASSES; ++p)
  {
    AXB_LAUNCH(ctx, rsort::onesweep_kernel, tiles, rsort::BLOCK, src, dst, (long long)n, 32 + p * rsort::RADIX_BITS,
               ghist + p * rsort::RADIX, lookback + (size_t)p * tiles * rsort::RADIX, tile_counters + p);
    std::swap(src, dst);
  }
  *sorted = src;
  return AXB_OK;
}

int sort_keys(axb_bvh* h, int n, uint32_t* ghist, uint32_t* tile_counters, uint32_t* lookback)
{
  return sort_keys_generic(h->ctx, h->keys_a.as<unsigned long long>(), h->keys_b.as<unsigned long long>(), n, ghist, tile_counters,
                           lookback, &h->sorted_keys);
}

template <typename T, int D>
int build_impl(axb_bvh* h, const axb_array_desc* boxes, int32_t num_boxes)
{
  Ctx& ctx = h->ctx;
  AXB_TRY(ctx.bind());
  ctx.begin_call();
  const int tot = ctx.phase_begin("build.total");
  h->built = false;
  h->ref_view_valid = false;
  h->fnodes_valid = false;
  h->n_in = num_boxes;
  const int n = num_boxes <= 1 ? 2 : num_boxes;  // spin/BVH.hpp:439-464
  h->n = n;
  const int inner = n - 1;
  const T half_scale = static_cast<T>(h->scale * 0.5);

  Desc<2 * D> in;
  AXB_TRY(stage_desc<2 * D>(ctx, boxes, num_boxes, sizeof(T), h->stage_in, in));
  if(num_boxes == 0)
    for(int c = 0; c < 2 * D; ++c) in.comp[c] = nullptr;

  AXB_TRY(h->nodes.reserve(sizeof(Node<T, D>) * (size_t)inner, ctx.stream));
  AXB_TRY(h->leaf_nodes.reserve(sizeof(int32_t) * (size_t)n, ctx.stream));
  AXB_TRY(h->leaf_parent.reserve(sizeof(int32_t) * (size_t)n, ctx.stream));
  AXB_TRY(h->node_range.reserve(sizeof(int2) * (size_t)inner, ctx.stream));
  AXB_TRY(h->keys_a.reserve(sizeof(unsigned long long) * (size_t)n, ctx.stream));
  AXB_TRY(h->keys_b.reserve(sizeof(unsigned long long) * (size_t)n, ctx.stream));
  AXB_TRY(h->state.reserve(sizeof(BuildState<T, D>), ctx.stream));
  const size_t scratch = rsort::scratch_bytes(n);
  AXB_TRY(h->sort_scratch.reserve(scratch, ctx.stream));
  AXB_CUDA_TRY(cudaMemsetAsync(h->sort_scratch.p, 0, scratch, ctx.stream));
  uint32_t* ghist = h->sort_scratch.as<uint32_t>();
  uint32_t* tile_counters = ghist + rsort::MAX_PASSES * rsort::RADIX;
  uint32_t* lookback = tile_counters + 64;
  auto* st = h->state.as<BuildState<T, D>>();

  {
    ScopedPhase ph(ctx, "build.bounds");
    AXB_LAUNCH(ctx, (init_state_kernel<T, D>), 1, 32, st);
    AXB_LAUNCH(ctx, (bounds_kernel<T, D>), capped_grid(n, 256), 256, in, n, num_boxes, half_scale, st);
    AXB_LAUNCH(ctx, (finalize_bounds_kernel<T, D>), 1, 32, st);
  }
  {
    ScopedPhase ph(ctx, "build.morton");
    AXB_LAUNCH(ctx, (morton_kernel<T, D>), capped_grid(n, 256), 256, in, n, num_boxes, half_scale, st,
               h->keys_a.as<unsigned long long>(), ghist);
  }
  {
    ScopedPhase ph(ctx, "build.sort");
    AXB_TRY(sort_keys(h, n, ghist, tile_counters, lookback));
  }
  const bool legacy = h->legacy_build;
  auto legacy_tree_refit = [&]() -> int {
    {
      ScopedPhase ph(ctx, "build.tree");
      AXB_LAUNCH(ctx, (tree_kernel<T, D>), blocks_for(inner, 256), 256, h->sorted_keys, n, h->nodes.as<Node<T, D>>(),
                 h->leaf_parent.as<int32_t>(), h->node_range.as<int2>());
    }
    {
      ScopedPhase ph(ctx, "build.refit");
      AXB_LAUNCH(ctx, (refit_kernel<T, D>), blocks_for(n, 256), 256, in, n, num_boxes, half_scale, h->sorted_keys,
                 h->leaf_parent.as<int32_t>(), h->nodes.as<Node<T, D>>(), h->leaf_nodes.as<int32_t>());
    }
    return AXB_OK;
  };
  if(legacy)
  {
    AXB_TRY(legacy_tree_refit());
  }
  else
  {
    // fused bottom-up hierarchy + refit (build.cuh: agglo_kernel)
    ScopedPhase ph(ctx, "build.agglo");
    AXB_TRY(h->agglo_slots.reserve(sizeof(AggloSlot<T, D>) * (size_t)n, ctx.stream));
    AXB_TRY(h->agglo_flags.reserve(sizeof(uint32_t) * (size_t)inner, ctx.stream));
    AXB_CUDA_TRY(cudaMemsetAsync(h->agglo_flags.p, 0, sizeof(uint32_t) * (size_t)inner, ctx.stream));
    AXB_TRY(h->agglo_open.reserve(sizeof(AggloOpen<T, D>) * (size_t)n, ctx.stream));  // worst case: every leaf is open
#define AXB_AGGLO(AB)                                                                                                          \
  AXB_LAUNCH(ctx, (agglo_kernel<T, D, AB>), blocks_for(n, AB), AB, in, n, num_boxes, half_scale, h->sorted_keys,             \
             h->nodes.as<Node<T, D>>(), h->leaf_nodes.as<int32_t>(), h->leaf_parent.as<int32_t>(), h->node_range.as<int2>(), \
             h->agglo_slots.as<AggloSlot<T, D>>(), h->agglo_flags.as<uint32_t>(), &st->agglo_mismatch,              \
             h->agglo_open.as<AggloOpen<T, D>>(), &st->agglo_open_count)
    if(h->agglo_block == 128)
      AXB_AGGLO(128);
    else if(h->agglo_block == 512)
      AXB_AGGLO(512);
    else
      AXB_AGGLO(256);
#undef AXB_AGGLO
    // the upper tree, from the subtrees the blocks left open (none when the whole tree fitted one block)
    AXB_LAUNCH(ctx, (agglo_upper_kernel<T, D>), capped_grid(std::max(n / 8, 128), 128, 16), 128, n, h->sorted_keys, h->nodes.as<Node<T, D>>(),
               h->leaf_parent.as<int32_t>(), h->node_range.as<int2>(), h->agglo_slots.as<AggloSlot<T, D>>(), h->agglo_flags.as<uint32_t>(),
               &st->agglo_mismatch, h->agglo_open.as<AggloOpen<T, D>>(), &st->agglo_open_count);
  }
  ctx.phase_end(tot);
  // bounds come back to the host (getBounds() is a host query); this is also the build's sync point
  BuildState<T, D> hst;
  AXB_CUDA_TRY(cudaMemcpyAsync(&hst, st, sizeof(hst), cudaMemcpyDeviceToHost, ctx.stream));
  AXB_TRY(ctx.sync());
  if(!legacy && hst.agglo_mismatch)
  {
    // a >2^24-leaf node where the reference's float32 split search leaves the exact Karras split, or a
    // non-canonical invalid input box: rebuild the hierarchy with the reference-order kernels
    AXB_TRY(legacy_tree_refit());
    AXB_TRY(ctx.sync());
  }
  for(int d = 0; d < 3; ++d)
  {
    h->bounds_lo[d] = d < D ? (double)hst.bmin[d] : 0.0;
    h->bounds_hi[d] = d < D ? (double)hst.bmax[d] : 0.0;
  }
  h->built = true;
  if(ctx.profiling && !ctx.async) ctx.resolve();
  return AXB_BVH_BUILD_OK;
}

template <typename T, int D>
int ensure_ref_view(axb_bvh* h)
{
  if(h->ref_view_valid) return AXB_OK;
  Ctx& ctx = h->ctx;
  const int inner = h->n - 1;
  AXB_TRY(h->ref_inner_nodes.reserve(sizeof(Box<T, D>) * 2 * (size_t)inner, ctx.stream));
  AXB_TRY(h->ref_children.reserve(sizeof(int32_t) * 2 * (size_t)inner, ctx.stream));
  AXB_LAUNCH(ctx, (export_kernel<T, D>), blocks_for(inner, 256), 256, h->nodes.as<Node<T, D>>(), inner,
             h->ref_inner_nodes.as<Box<T, D>>(), h->ref_children.as<int32_t>());
  AXB_TRY(ctx.sync());
  h->ref_view_valid = true;
  return AXB_OK;
}

int exclusive_scan(axb_bvh* h, const int32_t* counts, int nq, int32_t* offsets, long long* d_total)
{
  Ctx& ctx = h->ctx;
  const int tiles = blocks_for(nq, SCAN_TILE);
  AXB_TRY(h->q_tiles.reserve(sizeof(long long) * (size_t)tiles, ctx.stream));
  long long* tsum = h->q_tiles.as<long long>();
  AXB_LAUNCH(ctx, scan_tile_sums_kernel, tiles, SCAN_BLOCK, counts, nq, tsum);
  AXB_LAUNCH(ctx, scan_spine_kernel, 1, 1024, tsum, tiles, d_total);
  AXB_LAUNCH(ctx, scan_apply_kernel, tiles, SCAN_BLOCK, counts, nq, tsum, offsets);
  return AXB_OK;
}

// LinearBVH::findCandidatesImpl (policy/LinearBVH.hpp:271-402): count -> scan -> allocate -> fill
// Filter (traverse.cuh / tritri.cuh) is the optional narrow phase applied to every candidate before it is
// counted or recorded; `firsts` (optional) receives the query id of every kept candidate (pair form).
template <typename T, int D, class Query, class Filter = NoFilter>
int find_impl(axb_bvh* h, int kind, const axb_array_desc* prims, int flags, int32_t nq, int32_t* offsets, int32_t* counts, int out_memspace,
              int32_t** candidates, int64_t* total, Filter filt = Filter(), int32_t** firsts = nullptr)
{
  if(!h->built) return fail(AXB_ERR_NOT_BUILT, "BVH query before initialize()");
  if(nq < 0) return fail(AXB_ERR_BAD_ARG, "negative query count");
  if(!candidates || !total) return fail(AXB_ERR_BAD_ARG, "null output pointer");
  if(nq > 0 && (!offsets || !counts)) return fail(AXB_ERR_BAD_ARG, "offsets/counts must hold num_queries entries");
  out_memspace = resolve_memspace(out_memspace, offsets);
  if(out_memspace != AXB_MEM_HOST && out_memspace != AXB_MEM_DEVICE) return fail(AXB_ERR_BAD_ARG, "unknown output memspace");
  Ctx& ctx = h->ctx;
  AXB_TRY(ctx.bind());
  ctx.begin_call();
  *candidates = nullptr;
  if(firsts) *firsts = nullptr;
  *total = 0;
  if(nq == 0) return AXB_OK;
  const int tot = ctx.phase_begin("find.total");

  Desc<Query::NCOMP> q;
  AXB_TRY(stage_desc<Query::NCOMP>(ctx, prims, nq, sizeof(T), h->q_stage, q));
  int32_t* d_counts = counts;
  int32_t* d_offsets = offsets;
  if(out_memspace == AXB_MEM_HOST)
  {
    AXB_TRY(h->q_counts.reserve(sizeof(int32_t) * (size_t)nq, ctx.stream));
    AXB_TRY(h->q_offsets.reserve(sizeof(int32_t) * (size_t)nq, ctx.stream));
    d_counts = h->q_counts.as<int32_t>();
    d_offsets = h->q_offsets.as<int32_t>();
  }
  AXB_TRY(h->q_total.reserve(sizeof(long long), ctx.stream));
  long long* d_total = h->q_total.as<long long>();
  const Node<T, D>* nodes = h->nodes.as<Node<T, D>>();
  const T tol = (T)h->tol;
  long long htotal = 0;
  int32_t* d_cand = nullptr;
  int32_t* d_first = nullptr;
  const int32_t* leaf_nodes = h->leaf_nodes.as<int32_t>();
  if(h->find_strategy == 1)
  {
    // the reference's shape: count -> scan -> fill, one thread per query (LinearBVH.hpp:302-364)
    {
      ScopedPhase ph(ctx, "find.count");
      AXB_LAUNCH(ctx, (count_kernel<T, D, Query, Filter>), blocks_for(nq, 256), 256, nodes, q, nq, tol, flags, (const int32_t*)nullptr,
                 d_counts, leaf_nodes, filt);
    }
    {
      ScopedPhase ph(ctx, "find.scan");
      AXB_TRY(exclusive_scan(h, d_counts, nq, d_offsets, d_total));
    }
    AXB_CUDA_TRY(cudaMemcpyAsync(&htotal, d_total, sizeof(long long), cudaMemcpyDeviceToHost, ctx.stream));
    AXB_TRY(ctx.sync());
    if(htotal > 2147483647LL)
      return fail(AXB_ERR_OVERFLOW, "candidate total " + std::to_string(htotal) + " overflows int32 offsets: split the query batch");
    AXB_CUDA_TRY(axb_malloc_async((void**)&d_cand, sizeof(int32_t) * (size_t)std::max<long long>(htotal, 1), ctx.stream));
    if(firsts) AXB_CUDA_TRY(axb_malloc_async((void**)&d_first, sizeof(int32_t) * (size_t)std::max<long long>(htotal, 1), ctx.stream));
    {
      ScopedPhase ph(ctx, "find.fill");
      AXB_LAUNCH(ctx, (fill_kernel<T, D, Query, Filter>), blocks_for(nq, 256), 256, nodes, leaf_nodes, q, nq, tol, flags,
                 (const int32_t*)nullptr, d_offsets, d_cand, d_first, filt);
    }
  }
  else
  {
    // ---- Morton order of the queries (processing order only) ----
    const int32_t* perm = nullptr;
    if(nq >= 8192)
    {
      ScopedPhase ph(ctx, "find.sortq");
      const size_t kb = sizeof(unsigned long long) * (size_t)nq;
      AXB_TRY(h->f_keys_a.reserve(kb, ctx.stream));
      AXB_TRY(h->f_keys_b.reserve(kb, ctx.stream));
      AXB_TRY(h->f_perm.reserve(sizeof(int32_t) * (size_t)nq, ctx.stream));
      const size_t scratch = rsort::scratch_bytes(nq);
      AXB_TRY(h->f_scratch.reserve(scratch, ctx.stream));
      AXB_CUDA_TRY(cudaMemsetAsync(h->f_scratch.p, 0, scratch, ctx.stream));
      uint32_t* ghist = h->f_scratch.as<uint32_t>();
      uint32_t* tile_counters = ghist + rsort::MAX_PASSES * rsort::RADIX;
      uint32_t* lookback = tile_counters + 64;
      AXB_LAUNCH(ctx, (find_query_keys_kernel<T, D, Query, BuildState<T, D>>), capped_grid(nq, 256), 256, q, nq,
                 h->state.as<BuildState<T, D>>(), h->f_keys_a.as<unsigned long long>(), ghist);
      unsigned long long* sorted = nullptr;
      // (two digits = 16 bits of Morton code are enough here: the walk only wants neighbours to share the upper tree;
      //  C1 1.77 -> 1.90 G points/s.  The SignedDistance search keeps three: its first bounds come from rank neighbours.)
      AXB_TRY(sort_keys_generic(ctx, h->f_keys_a.as<unsigned long long>(), h->f_keys_b.as<unsigned long long>(), nq, ghist, tile_counters,
                                lookback, &sorted, 2));
      AXB_LAUNCH(ctx, keys_to_perm_kernel, blocks_for(nq, 256), 256, sorted, nq, h->f_perm.as<int32_t>());
      perm = h->f_perm.as<int32_t>();
    }
    // ---- pair buffer: sized from the last call of this kind, at least 4 hits per query ----
    const long long want_pairs = std::max<long long>(4LL * nq + 65536, h->pair_hint[kind] + h->pair_hint[kind] / 4 + 65536);
    // every resident warp can strand one partly filled chunk
    if(h->walk_blocks_per_sm[kind] == 0)
    {
      int bps = 0;
      AXB_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, find_walk_kernel<T, D, Query, Filter>, 128, 0));
      int cap = 16;  // as many resident blocks as the registers allow (10-11 for the point / box walks): C1 and C3 +6 % over 8
      if(const char* e = getenv("AXB_FIND_BLOCKS")) cap = std::max(1, atoi(e));
      h->walk_blocks_per_sm[kind] = std::max(1, std::min(bps, cap));
    }
    int sms = kNumSMsB200;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, ctx.device);
    const int grid = (int)std::min<long long>(blocks_for(nq, 128), (long long)sms * h->walk_blocks_per_sm[kind]);
    unsigned max_chunks = (unsigned)std::min<long long>((want_pairs + kPairChunk - 1) / kPairChunk + 4LL * grid, 0x7fffffffLL / kPairChunk);
    if(h->find_strategy == 2) max_chunks = 2;  // test hook: force the overflow fallback
    AXB_TRY(h->f_pairs.reserve(sizeof(int4) * (size_t)max_chunks * kPairChunk, ctx.stream));
    AXB_TRY(h->f_unused.reserve(sizeof(unsigned int) * (size_t)max_chunks, ctx.stream));
    AXB_TRY(h->f_cursor.reserve(sizeof(unsigned int) * 4, ctx.stream));
    AXB_CUDA_TRY(cudaMemsetAsync(h->f_unused.p, 0, sizeof(unsigned int) * (size_t)max_chunks, ctx.stream));
    AXB_CUDA_TRY(cudaMemsetAsync(h->f_cursor.p, 0, sizeof(unsigned int) * 4, ctx.stream));
    PairBuf pb;
    pb.pairs = h->f_pairs.as<int4>();
    pb.cursor = h->f_cursor.as<unsigned int>();
    pb.unused = h->f_unused.as<unsigned int>();
    pb.max_chunks = max_chunks;
    {
      ScopedPhase ph(ctx, "find.count");  // the one traversal: counts + recorded hits
      // comparison-only predicates on the 3-D double tree walk the compact records (traverse.cuh: FNode)
      const FNode* fn = nullptr;
      // (not the fused narrow phase, kind 3: with ~13 overlapping neighbours per triangle most visits end in exact
      //  leaf re-tests and the walk gets slower, 14.9 -> 17.9 ms on the 10 M-triangle case)
      if(std::is_same<T, double>::value && D == 3 && (kind == 0 || kind == 1) && h->find_compact)
      {
        if(!h->fnodes_valid)
        {
          AXB_TRY(h->fnodes.reserve(sizeof(FNode) * (size_t)(h->n - 1), ctx.stream));
          AXB_LAUNCH(ctx, fnode_kernel, blocks_for(h->n - 1, 256), 256, h->nodes.as<Node<double, 3>>(), h->n - 1, h->fnodes.as<FNode>());
          h->fnodes_valid = true;
        }
        fn = h->fnodes.as<FNode>();
      }
      AXB_LAUNCH(ctx, (find_walk_kernel<T, D, Query, Filter>), grid, 128, nodes, leaf_nodes, q, nq, tol, flags, perm, d_counts, pb, filt, fn);
    }
    {
      ScopedPhase ph(ctx, "find.scan");
      AXB_TRY(exclusive_scan(h, d_counts, nq, d_offsets, d_total));
    }
    unsigned int hcur[4] = {0, 0, 0, 0};
    AXB_CUDA_TRY(cudaMemcpyAsync(&htotal, d_total, sizeof(long long), cudaMemcpyDeviceToHost, ctx.stream));
    AXB_CUDA_TRY(cudaMemcpyAsync(hcur, h->f_cursor.p, sizeof(hcur), cudaMemcpyDeviceToHost, ctx.stream));
    AXB_TRY(ctx.sync());
    if(htotal > 2147483647LL)
      return fail(AXB_ERR_OVERFLOW, "candidate total " + std::to_string(htotal) + " overflows int32 offsets: split the query batch");
    h->pair_hint[kind] = htotal;
    AXB_CUDA_TRY(axb_malloc_async((void**)&d_cand, sizeof(int32_t) * (size_t)std::max<long long>(htotal, 1), ctx.stream));
    if(firsts) AXB_CUDA_TRY(axb_malloc_async((void**)&d_first, sizeof(int32_t) * (size_t)std::max<long long>(htotal, 1), ctx.stream));
    ScopedPhase ph(ctx, "find.fill");
    if(hcur[2] == 0u)
    {
      const unsigned nchunks = std::min(hcur[1], max_chunks);
      if(nchunks)
        AXB_LAUNCH(ctx, scatter_pairs_kernel, blocks_for((long long)nchunks * kPairChunk, 256), 256, pb.pairs, pb.unused, nchunks, d_offsets,
                   d_cand, d_first);
    }
    else
    {
      // the pair buffer was too small for this call: second traversal, as the reference does
      AXB_LAUNCH(ctx, (fill_kernel<T, D, Query, Filter>), blocks_for(nq, 256), 256, nodes, leaf_nodes, q, nq, tol, flags,
                 (const int32_t*)nullptr, d_offsets, d_cand, d_first, filt);
    }
  }
  ctx.phase_end(tot);
  if(out_memspace == AXB_MEM_HOST)
  {
    int32_t* hc = (int32_t*)malloc(sizeof(int32_t) * (size_t)std::max<long long>(htotal, 1));
    if(!hc) return fail(AXB_ERR_BAD_ARG, "host allocation of the candidate array failed");
    AXB_CUDA_TRY(cudaMemcpyAsync(hc, d_cand, sizeof(int32_t) * (size_t)htotal, cudaMemcpyDeviceToHost, ctx.stream));
    if(firsts)
    {
      int32_t* hf = (int32_t*)malloc(sizeof(int32_t) * (size_t)std::max<long long>(htotal, 1));
      if(!hf) return fail(AXB_ERR_BAD_ARG, "host allocation of the pair array failed");
      AXB_CUDA_TRY(cudaMemcpyAsync(hf, d_first, sizeof(int32_t) * (size_t)htotal, cudaMemcpyDeviceToHost, ctx.stream));
      AXB_CUDA_TRY(cudaFreeAsync(d_first, ctx.stream));
      *firsts = hf;
    }
    AXB_CUDA_TRY(cudaMemcpyAsync(counts, d_counts, sizeof(int32_t) * (size_t)nq, cudaMemcpyDeviceToHost, ctx.stream));
    AXB_CUDA_TRY(cudaMemcpyAsync(offsets, d_offsets, sizeof(int32_t) * (size_t)nq, cudaMemcpyDeviceToHost, ctx.stream));
    AXB_CUDA_TRY(cudaFreeAsync(d_cand, ctx.stream));
    AXB_TRY(ctx.sync());
    *candidates = hc;
  }
  else
  {
    *candidates = d_cand;
    if(firsts) *firsts = d_first;
  }
  *total = htotal;
  return ctx.finish_call();
}

bool valid_bvh(const axb_bvh* b) { return b != nullptr; }

// (FloatType, NDIMS) -> template instantiation.  CALL is a macro taking (T, D).
#define AXB_DISPATCH(h, CALL)                                             \
  ((h)->fp_bytes == 8 ? ((h)->ndims == 2 ? CALL(double, 2) : CALL(double, 3)) \
                      : ((h)->ndims == 2 ? CALL(float, 2) : CALL(float, 3)))

}  // namespace

extern "C" {

const char* axb_version(void) { return AXB_VERSION_STRING; }

int axb_trim_pool(int device, uint64_t keep_bytes)
{
  if(device < 0 || device >= kMaxDevices) return fail(AXB_ERR_BAD_ARG, "device ordinal out of range");
  cudaMemPool_t pool = nullptr;
  {
    std::lock_guard<std::mutex> lock(g_pool_mutex);
    pool = g_pool[device];
  }
  if(!pool) return AXB_OK;  // nothing was ever allocated on this device
  AXB_CUDA_TRY(cudaMemPoolTrimTo(pool, (size_t)keep_bytes));
  return AXB_OK;
}
const char* axb_last_error(void) { return g_last_error.c_str(); }

int axb_device_count(void)
{
  int c = 0;
  if(cudaGetDeviceCount(&c) != cudaSuccess)
  {
    cudaGetLastError();
    return 0;
  }
  return c;
}

const char* axb_status_string(int s)
{
  switch(s)
  {
  case AXB_OK: return "AXB_OK";
  case AXB_ERR_BAD_ARG: return "AXB_ERR_BAD_ARG";
  case AXB_ERR_CUDA: return "AXB_ERR_CUDA";
  case AXB_ERR_OVERFLOW: return "AXB_ERR_OVERFLOW";
  case AXB_ERR_NOT_BUILT: return "AXB_ERR_NOT_BUILT";
  case AXB_ERR_NO_DEVICE: return "AXB_ERR_NO_DEVICE";
  case AXB_ERR_UNSUPPORTED: return "AXB_ERR_UNSUPPORTED";
  default: return "AXB_ERR_UNKNOWN";
  }
}

int axb_bvh_create(axb_bvh** out, int ndims, int fp_bytes, int device)
{
  if(!out) return fail(AXB_ERR_BAD_ARG, "null output handle");
  *out = nullptr;
  if(ndims != 2 && ndims != 3) return fail(AXB_ERR_BAD_ARG, "The BVH class may be used only in 2D or 3D.");
  if(fp_bytes != 8 && fp_bytes != 4) return fail(AXB_ERR_BAD_ARG, "fp_bytes must be 8 (double) or 4 (float)");
  axb_bvh* h = new axb_bvh();
  h->ndims = ndims;
  h->fp_bytes = fp_bytes;
  if(fp_bytes == 4) h->tol = FLT_EPSILON;  // DEFAULT_TOLERANCE = floating_point_limits<FloatType>::epsilon()
  if(const char* e = getenv("AXB_BUILD_LEGACY")) h->legacy_build = atoi(e) != 0;
  if(const char* e = getenv("AXB_AGGLO_BLOCK")) h->agglo_block = atoi(e);
  if(const char* e = getenv("AXB_FIND_COMPACT")) h->find_compact = atoi(e) != 0;
  int s = h->ctx.init(device);
  if(s != AXB_OK)
  {
    delete h;
    return s;
  }
  *out = h;
  return AXB_OK;
}

int axb_bvh_destroy(axb_bvh* h)
{
  if(!h) return AXB_OK;
  cudaSetDevice(h->ctx.device);
  h->release_all();
  cudaStreamSynchronize(h->ctx.stream);
  h->ctx.destroy();
  delete h;
  return AXB_OK;
}

int axb_bvh_set_stream(axb_bvh* h, void* s)
{
  if(!valid_bvh(h)) return fail(AXB_ERR_BAD_ARG, "null handle");
  AXB_TRY(h->ctx.sync());
  if(h->ctx.own_stream && h->ctx.stream) cudaStreamDestroy(h->ctx.stream);
  h->ctx.stream = (cudaStream_t)s;
  h->ctx.own_stream = false;
  return AXB_OK;
}
int axb_bvh_set_async(axb_bvh* h, int e)
{
  if(!valid_bvh(h)) return fail(AXB_ERR_BAD_ARG, "null handle");
  h->ctx.async = e != 0;
  return AXB_OK;
}
int axb_bvh_synchronize(axb_bvh* h)
{
  if(!valid_bvh(h)) return fail(AXB_ERR_BAD_ARG, "null handle");
  return h->ctx.sync();
}
int axb_bvh_set_scale_factor(axb_bvh* h, double s)
{
  if(!valid_bvh(h)) return fail(AXB_ERR_BAD_ARG, "null handle");
  h->scale = s;
  return AXB_OK;
}
int axb_bvh_get_scale_factor(const axb_bvh* h, double* s)
{
  if(!valid_bvh(h) || !s) return fail(AXB_ERR_BAD_ARG, "null argument");
  *s = h->scale;
  return AXB_OK;
}
int axb_bvh_set_tolerance(axb_bvh* h, double t)
{
  if(!valid_bvh(h)) return fail(AXB_ERR_BAD_ARG, "null handle");
  h->tol = t;
  return AXB_OK;
}
int axb_bvh_get_tolerance(const axb_bvh* h, double* t)
{
  if(!valid_bvh(h) || !t) return fail(AXB_ERR_BAD_ARG, "null argument");
  *t = h->tol;
  return AXB_OK;
}

int axb_bvh_initialize(axb_bvh* h, const axb_array_desc* boxes, int32_t n)
{
  if(!valid_bvh(h)) return fail(AXB_ERR_BAD_ARG, "null handle");
  if(n < 0) return fail(AXB_ERR_BAD_ARG, "negative box count");
  if(n > (1 << 30)) return fail(AXB_ERR_BAD_ARG, "more than 2^30 boxes: int32 node ids would overflow");
  if(n > 0 && !boxes) return fail(AXB_ERR_BAD_ARG, "null boxes");
  axb_array_desc empty;
  memset(&empty, 0, sizeof(empty));
  empty.ncomp = 2 * h->ndims;
  empty.stride_bytes = h->fp_bytes;
  empty.memspace = AXB_MEM_DEVICE;
  const axb_array_desc* d = (n == 0 && !boxes) ? &empty : boxes;
#define AXB_CALL(T, D) build_impl<T, D>(h, d, n)
  return AXB_DISPATCH(h, AXB_CALL);
#undef AXB_CALL
}

int axb_bvh_is_initialized(const axb_bvh* h) { return (h && h->built) ? 1 : 0; }

int axb_bvh_get_bounds(const axb_bvh* h, double* lo, double* hi)
{
  if(!valid_bvh(h) || !lo || !hi) return fail(AXB_ERR_BAD_ARG, "null argument");
  for(int d = 0; d < h->ndims; ++d)
  {
    lo[d] = h->built ? h->bounds_lo[d] : DBL_MAX;   // invalid box when unbuilt, spin/BVH.hpp:303-307
    hi[d] = h->built ? h->bounds_hi[d] : -DBL_MAX;
  }
  return AXB_OK;
}

int axb_bvh_get_traverser(axb_bvh* h, axb_traverser* out)
{
  if(!valid_bvh(h) || !out) return fail(AXB_ERR_BAD_ARG, "null argument");
  if(!h->built) return fail(AXB_ERR_NOT_BUILT, "getTraverser() before initialize()");
  AXB_TRY(h->ctx.bind());
#define AXB_CALL(T, D) ensure_ref_view<T, D>(h)
  AXB_TRY(AXB_DISPATCH(h, AXB_CALL));
#undef AXB_CALL
  out->inner_nodes = h->ref_inner_nodes.p;
  out->inner_node_children = h->ref_children.as<int32_t>();
  out->leaf_nodes = h->leaf_nodes.as<int32_t>();
  out->num_leaves = h->n;
  out->ndims = h->ndims;
  out->fp_bytes = h->fp_bytes;
  return AXB_OK;
}

int axb_bvh_find_points(axb_bvh* h, const axb_array_desc* pts, int32_t nq, int32_t* offsets, int32_t* counts, int out_memspace,
                        int32_t** candidates, int64_t* total)
{
  if(!valid_bvh(h)) return fail(AXB_ERR_BAD_ARG, "null handle");
#define AXB_CALL(T, D) find_impl<T, D, PointQuery<T, D>>(h, 0, pts, 0, nq, offsets, counts, out_memspace, candidates, total)
  return AXB_DISPATCH(h, AXB_CALL);
#undef AXB_CALL
}

int axb_bvh_find_boxes(axb_bvh* h, const axb_array_desc* boxes, int32_t nq, int32_t* offsets, int32_t* counts, int out_memspace,
                       int32_t** candidates, int64_t* total)
{
  if(!valid_bvh(h)) return fail(AXB_ERR_BAD_ARG, "null handle");
#define AXB_CALL(T, D) find_impl<T, D, BoxQuery<T, D>>(h, 1, boxes, 0, nq, offsets, counts, out_memspace, candidates, total)
  return AXB_DISPATCH(h, AXB_CALL);
#undef AXB_CALL
}

int axb_bvh_find_rays(axb_bvh* h, const axb_array_desc* rays, int rays_normalized, int32_t nq, int32_t* offsets, int32_t* counts,
                      int out_memspace, int32_t** candidates, int64_t* total)
{
  if(!valid_bvh(h)) return fail(AXB_ERR_BAD_ARG, "null handle");
  const int f = rays_normalized ? 1 : 0;
#define AXB_CALL(T, D) find_impl<T, D, RayQuery<T, D>>(h, 2, rays, f, nq, offsets, counts, out_memspace, candidates, total)
  return AXB_DISPATCH(h, AXB_CALL);
#undef AXB_CALL
}

int axb_bvh_free_candidates(axb_bvh* h, int32_t* candidates, int memspace)
{
  if(!candidates) return AXB_OK;
  memspace = resolve_memspace(memspace, candidates);
  if(memspace == AXB_MEM_HOST)
  {
    free(candidates);
    return AXB_OK;
  }
  if(!valid_bvh(h)) return fail(AXB_ERR_BAD_ARG, "null handle");
  AXB_TRY(h->ctx.bind());
  AXB_CUDA_TRY(cudaFreeAsync(candidates, h->ctx.stream));
  return AXB_OK;
}

int axb_bvh_num_leaves(const axb_bvh* h, int32_t* n)
{
  if(!valid_bvh(h) || !n) return fail(AXB_ERR_BAD_ARG, "null argument");
  if(!h->built) return fail(AXB_ERR_NOT_BUILT, "BVH not initialized");
  *n = h->n;
  return AXB_OK;
}

int axb_bvh_copy_arrays(axb_bvh* h, uint32_t* mcodes, int32_t* leaf_nodes, void* inner_nodes, int32_t* inner_children)
{
  if(!valid_bvh(h)) return fail(AXB_ERR_BAD_ARG, "null handle");
  if(!h->built) return fail(AXB_ERR_NOT_BUILT, "BVH not initialized");
  Ctx& ctx = h->ctx;
  AXB_TRY(ctx.bind());
  const int n = h->n, inner = n - 1, D = h->ndims;
  if(mcodes)
  {
    DevBuf tmp;
    AXB_TRY(tmp.reserve(sizeof(uint32_t) * (size_t)n, ctx.stream));
    AXB_LAUNCH(ctx, extract_mcodes_kernel, blocks_for(n, 256), 256, h->sorted_keys, n, tmp.as<uint32_t>());
    AXB_CUDA_TRY(cudaMemcpyAsync(mcodes, tmp.p, sizeof(uint32_t) * (size_t)n, cudaMemcpyDeviceToHost, ctx.stream));
    AXB_TRY(ctx.sync());
    tmp.release(ctx.stream);
  }
  if(leaf_nodes)
    AXB_CUDA_TRY(cudaMemcpyAsync(leaf_nodes, h->leaf_nodes.p, sizeof(int32_t) * (size_t)n, cudaMemcpyDeviceToHost, ctx.stream));
  if(inner_nodes || inner_children)
  {
#define AXB_CALL(T, D) ensure_ref_view<T, D>(h)
    AXB_TRY(AXB_DISPATCH(h, AXB_CALL));
#undef AXB_CALL
    if(inner_nodes)
      AXB_CUDA_TRY(cudaMemcpyAsync(inner_nodes, h->ref_inner_nodes.p, (size_t)h->fp_bytes * 2 * D * 2 * (size_t)inner, cudaMemcpyDeviceToHost,
                                   ctx.stream));
    if(inner_children)
      AXB_CUDA_TRY(cudaMemcpyAsync(inner_children, h->ref_children.p, sizeof(int32_t) * 2 * (size_t)inner, cudaMemcpyDeviceToHost,
                                   ctx.stream));
  }
  return ctx.sync();
}

}  // extern "C"

// writeVtkFile (spin/BVH.hpp:405, policy/LinearBVH.hpp:404-458, internal/linear_bvh/bvh_vtkio.hpp): a debugging aid, host
// code over the reference-layout arrays.  Same traversal (root box, then both child boxes of every inner node, left
// subtree first), same stream formatting, so the file is the reference's byte for byte.
template <typename T>
static void vtk_box(int D, const T* b, int32_t& numPoints, int32_t& numBins, std::ostringstream& nodes, std::ostringstream& cells)
{
  const T* lo = b;
  const T* hi = b + D;
  if(D == 2)
  {
    nodes << lo[0] << " " << lo[1] << " 0.0\n";
    nodes << hi[0] << " " << lo[1] << " 0.0\n";
    nodes << hi[0] << " " << hi[1] << " 0.0\n";
    nodes << lo[0] << " " << hi[1] << " 0.0\n";
  }
  else
  {
    for(int k = 0; k < 2; ++k)
    {
      const T z = k == 0 ? lo[2] : hi[2];
      nodes << lo[0] << " " << lo[1] << " " << z << std::endl;
      nodes << hi[0] << " " << lo[1] << " " << z << std::endl;
      nodes << hi[0] << " " << hi[1] << " " << z << std::endl;
      nodes << lo[0] << " " << hi[1] << " " << z << std::endl;
    }
  }
  const int32_t nn = D == 2 ? 4 : 8;
  cells << nn << " ";
  for(int32_t i = 0; i < nn; ++i) cells << (numPoints + i) << " ";
  cells << "\n";
  numBins += 1;
  numPoints += nn;
}

template <typename T>
static int write_vtk_impl(axb_bvh* h, const char* file_name)
{
  const int D = h->ndims, inner = h->n - 1;
  std::vector<T> boxes((size_t)2 * inner * 2 * D);
  std::vector<int32_t> children((size_t)2 * inner);
  AXB_TRY(axb_bvh_copy_arrays(h, nullptr, nullptr, boxes.data(), children.data()));
  double lo[3], hi[3];
  AXB_TRY(axb_bvh_get_bounds(h, lo, hi));
  T root[6];
  for(int d = 0; d < D; ++d)
  {
    root[d] = (T)lo[d];
    root[D + d] = (T)hi[d];
  }
  std::ofstream ofs(file_name);
  if(!ofs) return fail(AXB_ERR_BAD_ARG, std::string("cannot open ") + file_name);
  std::ostringstream nodes, cells, levels;
  ofs << "# vtk DataFile Version 3.0\n";
  ofs << " BVHTree \n";
  ofs << "ASCII\n";
  ofs << "DATASET UNSTRUCTURED_GRID\n";
  int32_t numPoints = 0, numBins = 0;
  vtk_box<T>(D, root, numPoints, numBins, nodes, cells);
  levels << "0\n";
  // write_recursive (:155-207) with an explicit stack: (first entry of the node's pair in the flat arrays, level)
  std::vector<std::pair<int32_t, int32_t>> todo;
  todo.emplace_back(0, 1);
  while(!todo.empty())
  {
    const int32_t cur = todo.back().first, level = todo.back().second;
    todo.pop_back();
    vtk_box<T>(D, boxes.data() + (size_t)cur * 2 * D, numPoints, numBins, nodes, cells);
    levels << level << std::endl;
    vtk_box<T>(D, boxes.data() + (size_t)(cur + 1) * 2 * D, numPoints, numBins, nodes, cells);
    levels << level << std::endl;
    const int32_t l = children[cur], r = children[cur + 1];
    if(r > -1) todo.emplace_back(r, level + 1);  // popped after the whole left subtree
    if(l > -1) todo.emplace_back(l, level + 1);
  }
  ofs << "POINTS " << numPoints << " double\n";
  ofs << nodes.str() << std::endl;
  const int32_t nn = D == 2 ? 4 : 8;
  ofs << "CELLS " << numBins << " " << numBins * (nn + 1) << std::endl;
  ofs << cells.str() << std::endl;
  ofs << "CELL_TYPES " << numBins << std::endl;
  const int32_t cellType = D == 2 ? 9 : 12;
  for(int32_t i = 0; i < numBins; ++i) ofs << cellType << std::endl;
  ofs << "CELL_DATA " << numBins << std::endl;
  ofs << "SCALARS level int\n";
  ofs << "LOOKUP_TABLE default\n";
  ofs << levels.str() << std::endl;
  ofs << std::endl;
  ofs.close();
  return AXB_OK;
}

extern "C" {

int axb_bvh_write_vtk_file(axb_bvh* h, const char* file_name)
{
  if(!valid_bvh(h) || !file_name) return fail(AXB_ERR_BAD_ARG, "null argument");
  if(!h->built) return fail(AXB_ERR_NOT_BUILT, "BVH not initialized");
  return h->fp_bytes == 4 ? write_vtk_impl<float>(h, file_name) : write_vtk_impl<double>(h, file_name);
}

int axb_bvh_set_find_strategy(axb_bvh* h, int strategy)
{
  if(!valid_bvh(h)) return fail(AXB_ERR_BAD_ARG, "null handle");
  if(strategy < 0 || strategy > 2) return fail(AXB_ERR_BAD_ARG, "find strategy must be 0, 1 or 2");
  h->find_strategy = strategy;
  return AXB_OK;
}

int axb_bvh_set_profiling(axb_bvh* h, int e)
{
  if(!valid_bvh(h)) return fail(AXB_ERR_BAD_ARG, "null handle");
  h->ctx.set_profiling(e != 0);
  return AXB_OK;
}
int axb_bvh_get_phase_ms(const axb_bvh* hc, const char* name, double* ms)
{
  axb_bvh* h = const_cast<axb_bvh*>(hc);
  if(!valid_bvh(h) || !name || !ms) return fail(AXB_ERR_BAD_ARG, "null argument");
  h->ctx.resolve();
  auto it = h->ctx.acc.find(name);
  if(it == h->ctx.acc.end() || it->second.calls == 0)
    return fail(AXB_ERR_BAD_ARG, std::string("no timing recorded for phase ") + name);
  *ms = it->second.sum / (double)it->second.calls;  // mean over the calls since profiling was enabled
  return AXB_OK;
}
int axb_bvh_launch_count(const axb_bvh* h, int64_t* n)
{
  if(!valid_bvh(h) || !n) return fail(AXB_ERR_BAD_ARG, "null argument");
  *n = h->ctx.launches;
  return AXB_OK;
}

}  // extern "C"

//==========================================================================================
// quest::SignedDistance
//==========================================================================================
struct axb_sd
{
  axb_bvh* bvh = nullptr;  // owns; shares nothing else
  int nv = 3;              // vertices per cell
  int ncells = 0;
  int nnodes = 0;
  int mode = 1;  // 1 = OBB-accelerated, Morton-ordered queries (default); 0 = reference visiting order
  bool count_work = false;
  SdParams prm;
  DevBuf x, y, z, conn, offsets, soup, cell_boxes, obounds, work;
  DevBuf sdnodes, sdcens, sdup, sdnodes64;
  // per-query scratch, two sets: a host-to-host query is cut into chunks that alternate between two streams so
  // that the upload of chunk i+1 and the download of chunk i-1 overlap the kernel of chunk i
  struct QBufs
  {
    DevBuf q_stage, out_phi, out_cp, out_n, qkeys_a, qkeys_b, qscratch, qperm, qbounds, cursor, cand, cand_n, seed, solo, solo_scratch, hint, ext;
    void release(cudaStream_t st)
    {
      for(DevBuf* b : {&q_stage, &out_phi, &out_cp, &out_n, &qkeys_a, &qkeys_b, &qscratch, &qperm, &qbounds, &cursor, &cand, &cand_n, &seed, &solo,
                       &solo_scratch, &hint, &ext})
        b->release(st);
    }
  } qb[2];
  cudaStream_t pipe_stream[2] = {nullptr, nullptr};
  cudaEvent_t pipe_event[3] = {nullptr, nullptr, nullptr};
  int fast_blocks_per_sm = 0;  // occupancy of the persistent query kernel (queried once)
  int two_blocks_per_sm = 0;   // same for sd_two_phase_kernel
  int kernel = 2;              // mode 1 kernel: 2 = sd_two_phase_kernel (default), 1 = sd_fast_kernel (AXB_SD_KERNEL=fast)
  int64_t last_leaf_tests = 0, last_inner_visits = 0;
  std::map<std::string, double> setmesh_ms;  // device time of the phases of setMesh ("setmesh.*", "build.*")
  double max_diam = 0.0;                     // largest triangle diameter of the mesh (slack of the partitioned-surface bound)
  Ctx& ctx() { return bvh->ctx; }
};

extern "C" {

int axb_sd_create(axb_sd** out, int device, const double* x, const double* y, const double* z, int32_t nnodes, const int32_t* conn,
                  const int32_t* cell_node_offsets, int32_t ncells, int32_t nodes_per_cell, int mesh_memspace, int is_watertight,
                  int compute_sign)
{
  if(!out) return fail(AXB_ERR_BAD_ARG, "null output handle");
  *out = nullptr;
  const bool mixed = cell_node_offsets != nullptr;
  if(!mixed && nodes_per_cell != 3 && nodes_per_cell != 4) return fail(AXB_ERR_BAD_ARG, "surface cells must be triangles (3) or quads (4)");
  if(nnodes < 0 || ncells < 0) return fail(AXB_ERR_BAD_ARG, "negative mesh size");
  if((nnodes > 0 && (!x || !y || !z)) || (ncells > 0 && !conn)) return fail(AXB_ERR_BAD_ARG, "null mesh array");
  mesh_memspace = resolve_memspace(mesh_memspace, x);
  if(mesh_memspace != AXB_MEM_HOST && mesh_memspace != AXB_MEM_DEVICE) return fail(AXB_ERR_BAD_ARG, "unknown memspace");
  axb_sd* s = new axb_sd();
  int st = axb_bvh_create(&s->bvh, 3, 8, device);
  if(st != AXB_OK)
  {
    delete s;
    return st;
  }
  Ctx& ctx = s->ctx();
  s->nv = mixed ? 4 : nodes_per_cell;  // a mixed mesh uses the 4-vertex leaf records, NaN marks a missing 4th vertex
  s->ncells = ncells;
  s->nnodes = nnodes;
  s->prm.watertight = is_watertight != 0;
  s->prm.compute_sign = compute_sign != 0;
  cudaEvent_t tot0 = nullptr, tot1 = nullptr;
  ctx.set_profiling(true);  // setMesh always times its phases (a dozen events): "setmesh.*" of axb_sd_get_phase_ms
  auto body = [&]() -> int {
    ctx.begin_call();
    // (the total is timed with its own event pair: the BVH build inside resolves and drops the context's pending phases)
    AXB_CUDA_TRY(cudaEventCreate(&tot0));
    AXB_CUDA_TRY(cudaEventCreate(&tot1));
    AXB_CUDA_TRY(cudaEventRecord(tot0, ctx.stream));
    const int up = ctx.phase_begin("setmesh.upload");
    const cudaMemcpyKind kind = mesh_memspace == AXB_MEM_HOST ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice;
    size_t conn_len = (size_t)ncells * (mixed ? 0 : nodes_per_cell);
    if(mixed)
    {
      // the connectivity length is the last offset
      int32_t last = 0;
      if(mesh_memspace == AXB_MEM_HOST)
        last = cell_node_offsets[ncells];
      else
        AXB_CUDA_TRY(cudaMemcpy(&last, cell_node_offsets + ncells, sizeof(int32_t), cudaMemcpyDeviceToHost));
      if(last < 0) return fail(AXB_ERR_BAD_ARG, "negative cell_node_offsets");
      conn_len = (size_t)last;
      AXB_TRY(s->offsets.reserve(sizeof(int32_t) * ((size_t)ncells + 1), ctx.stream));
      AXB_CUDA_TRY(cudaMemcpyAsync(s->offsets.p, cell_node_offsets, sizeof(int32_t) * ((size_t)ncells + 1), kind, ctx.stream));
    }
    const size_t nb = sizeof(double) * (size_t)nnodes, cb = sizeof(int32_t) * conn_len;
    AXB_TRY(s->x.reserve(nb, ctx.stream));
    AXB_TRY(s->y.reserve(nb, ctx.stream));
    AXB_TRY(s->z.reserve(nb, ctx.stream));
    AXB_TRY(s->conn.reserve(cb, ctx.stream));
    if(nnodes)
    {
      AXB_CUDA_TRY(cudaMemcpyAsync(s->x.p, x, nb, kind, ctx.stream));
      AXB_CUDA_TRY(cudaMemcpyAsync(s->y.p, y, nb, kind, ctx.stream));
      AXB_CUDA_TRY(cudaMemcpyAsync(s->z.p, z, nb, kind, ctx.stream));
    }
    if(cb) AXB_CUDA_TRY(cudaMemcpyAsync(s->conn.p, conn, cb, kind, ctx.stream));
    ctx.phase_end(up);
    const int cbx = ctx.phase_begin("setmesh.cell_boxes");
    // mesh node bounds (m_boxDomain)
    AXB_TRY(s->obounds.reserve(sizeof(unsigned long long) * 6, ctx.stream));
    unsigned long long init[6];
    for(int d = 0; d < 3; ++d)
    {
      init[d] = f64_to_ordered(DBL_MAX);
      init[3 + d] = f64_to_ordered(-DBL_MAX);
    }
    AXB_CUDA_TRY(cudaMemcpyAsync(s->obounds.p, init, sizeof(init), cudaMemcpyHostToDevice, ctx.stream));
    if(nnodes)
      AXB_LAUNCH(ctx, node_bounds_kernel, capped_grid(nnodes, 256), 256, s->x.as<double>(), s->y.as<double>(), s->z.as<double>(), nnodes,
                 s->obounds.as<unsigned long long>());
    unsigned long long hb[6];
    AXB_CUDA_TRY(cudaMemcpyAsync(hb, s->obounds.p, sizeof(hb), cudaMemcpyDeviceToHost, ctx.stream));
    // per-cell AABBs
    AXB_TRY(s->cell_boxes.reserve(sizeof(Box<double, 3>) * (size_t)std::max(ncells, 1), ctx.stream));
    DevBuf badflag;
    if(ncells && mixed)
    {
      AXB_TRY(badflag.reserve(sizeof(int), ctx.stream));
      AXB_CUDA_TRY(cudaMemsetAsync(badflag.p, 0, sizeof(int), ctx.stream));
      AXB_LAUNCH(ctx, cell_boxes_mixed_kernel, blocks_for(ncells, 256), 256, s->x.as<double>(), s->y.as<double>(), s->z.as<double>(),
                 s->conn.as<int32_t>(), s->offsets.as<int32_t>(), ncells, s->cell_boxes.as<Box<double, 3>>(), badflag.as<int>());
      int hbad = 0;
      AXB_CUDA_TRY(cudaMemcpyAsync(&hbad, badflag.p, sizeof(int), cudaMemcpyDeviceToHost, ctx.stream));
      AXB_TRY(ctx.sync());
      badflag.release(ctx.stream);
      if(hbad) return fail(AXB_ERR_BAD_ARG, "mixed surface mesh has a cell that is neither a triangle nor a quad");
    }
    else if(ncells)
    {
      if(s->nv == 3)
        AXB_LAUNCH(ctx, cell_boxes_kernel<3>, blocks_for(ncells, 256), 256, s->x.as<double>(), s->y.as<double>(), s->z.as<double>(),
                   s->conn.as<int32_t>(), ncells, s->cell_boxes.as<Box<double, 3>>());
      else
        AXB_LAUNCH(ctx, cell_boxes_kernel<4>, blocks_for(ncells, 256), 256, s->x.as<double>(), s->y.as<double>(), s->z.as<double>(),
                   s->conn.as<int32_t>(), ncells, s->cell_boxes.as<Box<double, 3>>());
    }
    ctx.phase_end(cbx);
    AXB_TRY(ctx.sync());
    for(int d = 0; d < 3; ++d)
    {
      s->prm.dom_lo[d] = ordered_to_f64(hb[d]);
      s->prm.dom_hi[d] = ordered_to_f64(hb[3 + d]);
    }
    // BVH over the cell boxes with the default scale factor (quest/SignedDistance.hpp:499-500)
    axb_array_desc bd;
    memset(&bd, 0, sizeof(bd));
    for(int c = 0; c < 6; ++c) bd.comp[c] = s->cell_boxes.as<char>() + 8 * c;
    bd.stride_bytes = 48;
    bd.ncomp = 6;
    bd.memspace = AXB_MEM_DEVICE;
    AXB_TRY(axb_bvh_initialize(s->bvh, &bd, ncells));  // (its own phases: "build.*")
    // leaf geometry in sorted-leaf order
    const int nl = s->bvh->n;
    const int gs = ctx.phase_begin("setmesh.gather_soup");
    AXB_TRY(s->soup.reserve(sizeof(double) * kLeafDoubles * (size_t)nl, ctx.stream));
    if(mixed)
      AXB_LAUNCH(ctx, gather_soup_mixed_kernel, blocks_for(nl, 256), 256, s->x.as<double>(), s->y.as<double>(), s->z.as<double>(),
                 s->conn.as<int32_t>(), s->offsets.as<int32_t>(), s->bvh->leaf_nodes.as<int32_t>(), nl, ncells, s->soup.as<double>());
    else if(s->nv == 3)
      AXB_LAUNCH(ctx, gather_soup_kernel<3>, blocks_for(nl, 256), 256, s->x.as<double>(), s->y.as<double>(), s->z.as<double>(),
                 s->conn.as<int32_t>(), s->bvh->leaf_nodes.as<int32_t>(), nl, ncells, s->soup.as<double>());
    else
      AXB_LAUNCH(ctx, gather_soup_kernel<4>, blocks_for(nl, 256), 256, s->x.as<double>(), s->y.as<double>(), s->z.as<double>(),
                 s->conn.as<int32_t>(), s->bvh->leaf_nodes.as<int32_t>(), nl, ncells, s->soup.as<double>());
    s->cell_boxes.release(ctx.stream);
    {
      // largest triangle diameter (reuses the mesh-bounds scratch word 0 as an ordered-uint maximum)
      DevBuf md;
      AXB_TRY(md.reserve(sizeof(unsigned long long), ctx.stream));
      AXB_CUDA_TRY(cudaMemsetAsync(md.p, 0, sizeof(unsigned long long), ctx.stream));
      if(s->nv == 3)
        AXB_LAUNCH(ctx, sd_max_diam_kernel<3>, blocks_for(nl, 256), 256, s->soup.as<double>(), nl, md.as<unsigned long long>());
      else
        AXB_LAUNCH(ctx, sd_max_diam_kernel<4>, blocks_for(nl, 256), 256, s->soup.as<double>(), nl, md.as<unsigned long long>());
      unsigned long long hmd = 0;
      AXB_CUDA_TRY(cudaMemcpyAsync(&hmd, md.p, sizeof(hmd), cudaMemcpyDeviceToHost, ctx.stream));
      AXB_TRY(ctx.sync());
      md.release(ctx.stream);
      s->max_diam = hmd ? sqrt(ordered_to_f64(hmd)) * (1.0 + 1e-12) : 0.0;
    }
    ctx.phase_end(gs);
    const int ob = ctx.phase_begin("setmesh.obb_build");
    // traversal records of the fast query mode (sd_fast.cuh): child AABBs + oriented bounds + ids,
    // one warp per tree entity
    {
      const long long entities = 2LL * nl - 1;
      AXB_TRY(s->sdnodes.reserve(sizeof(SdNode) * (size_t)(nl - 1), ctx.stream));
      AXB_TRY(s->sdcens.reserve(sizeof(SdCen) * (size_t)(nl - 1), ctx.stream));
      AXB_TRY(s->sdnodes64.reserve(sizeof(SdNode64) * (size_t)(nl - 1), ctx.stream));
      const int blocks = blocks_for(entities, 256);  // one thread per entity; larger subtrees are queued for a warp / a block
      const Node<double, 3>* bn = s->bvh->nodes.as<Node<double, 3>>();
      int obb_max = kObbMaxRange;
      if(const char* e = getenv("AXB_SD_OBB_MAX")) obb_max = atoi(e);
      // subtrees of more than kObbWarpRange leaves are queued and bounded by whole blocks
      DevBuf big;
      const size_t big_cap = (size_t)std::max(nl - 1, 1);
      AXB_TRY(big.reserve(sizeof(int32_t) * 2 * (big_cap + 1), ctx.stream));
      unsigned int* big_count = big.as<unsigned int>();
      unsigned int* mid_count = big_count + 1;
      int32_t* big_list = big.as<int32_t>() + 2;
      int32_t* mid_list = big_list + big_cap;
      AXB_CUDA_TRY(cudaMemsetAsync(big_count, 0, 2 * sizeof(unsigned int), ctx.stream));
      int sms = kNumSMsB200;
      cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, ctx.device);
      // the subtrees' area-weighted normal sums, bottom-up in one sweep (child_sum[2 * node + side][3]; freed below)
      DevBuf nsum, narr;
      AXB_TRY(nsum.reserve(sizeof(double) * 6 * (size_t)std::max(nl - 1, 1), ctx.stream));
      AXB_TRY(narr.reserve(sizeof(unsigned int) * (size_t)std::max(nl - 1, 1), ctx.stream));
      AXB_CUDA_TRY(cudaMemsetAsync(narr.p, 0, sizeof(unsigned int) * (size_t)std::max(nl - 1, 1), ctx.stream));
      if(s->nv == 3)
        AXB_LAUNCH(ctx, normal_sums_kernel<3>, blocks_for(nl, 256), 256, s->soup.as<double>(), bn, s->bvh->leaf_parent.as<int32_t>(), nl,
                   nsum.as<double>(), narr.as<unsigned int>());
      else
        AXB_LAUNCH(ctx, normal_sums_kernel<4>, blocks_for(nl, 256), 256, s->soup.as<double>(), bn, s->bvh->leaf_parent.as<int32_t>(), nl,
                   nsum.as<double>(), narr.as<unsigned int>());
      const double* child_sum = nsum.as<double>();
      if(s->nv == 3)
      {
        AXB_LAUNCH(ctx, obb_build_kernel<3>, blocks, 256, s->soup.as<double>(), bn, s->bvh->leaf_parent.as<int32_t>(),
                   s->bvh->node_range.as<int2>(), nl, s->sdnodes.as<SdNode>(), s->sdcens.as<SdCen>(), obb_max, mid_list, mid_count, big_list,
                   big_count, child_sum);
        AXB_LAUNCH(ctx, obb_build_mid_kernel<3>, 8 * sms, 256, s->soup.as<double>(), bn, s->bvh->node_range.as<int2>(),
                   s->sdnodes.as<SdNode>(), mid_list, mid_count, child_sum);
        AXB_LAUNCH(ctx, obb_build_big_kernel<3>, 2 * sms, 512, s->soup.as<double>(), bn, s->bvh->node_range.as<int2>(),
                   s->sdnodes.as<SdNode>(), big_list, big_count, child_sum);
      }
      else
      {
        AXB_LAUNCH(ctx, obb_build_kernel<4>, blocks, 256, s->soup.as<double>(), bn, s->bvh->leaf_parent.as<int32_t>(),
                   s->bvh->node_range.as<int2>(), nl, s->sdnodes.as<SdNode>(), s->sdcens.as<SdCen>(), obb_max, mid_list, mid_count, big_list,
                   big_count, child_sum);
        AXB_LAUNCH(ctx, obb_build_mid_kernel<4>, 8 * sms, 256, s->soup.as<double>(), bn, s->bvh->node_range.as<int2>(),
                   s->sdnodes.as<SdNode>(), mid_list, mid_count, child_sum);
        AXB_LAUNCH(ctx, obb_build_big_kernel<4>, 2 * sms, 512, s->soup.as<double>(), bn, s->bvh->node_range.as<int2>(),
                   s->sdnodes.as<SdNode>(), big_list, big_count, child_sum);
      }
      big.release(ctx.stream);
      nsum.release(ctx.stream);
      narr.release(ctx.stream);
      // the same bounds as 64-byte records for the order-free search (sd_two.cuh)
      if(nl > 1) AXB_LAUNCH(ctx, sd64_pack_kernel, blocks_for(nl - 1, 256), 256, s->sdnodes.as<SdNode>(), bn, nl - 1, s->sdnodes64.as<SdNode64>());
      ctx.phase_end(ob);
      for(int k = 0; k < 2; ++k) AXB_TRY(s->qb[k].cursor.reserve(sizeof(unsigned int) * 4, ctx.stream));
      int bps = 0;
      if(s->nv == 3)
      {
        AXB_CUDA_TRY(cudaFuncSetAttribute(sd_fast_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSdFastSmem));
        AXB_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, sd_fast_kernel<3>, 128, kSdFastSmem));
      }
      else
      {
        AXB_CUDA_TRY(cudaFuncSetAttribute(sd_fast_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSdFastSmem));
        AXB_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, sd_fast_kernel<4>, 128, kSdFastSmem));
      }
      s->fast_blocks_per_sm = std::max(1, bps);
      // the two-phase kernel: parent / range of every inner node for phase 2's climb
      AXB_TRY(s->sdup.reserve(sizeof(SdUp) * (size_t)std::max(nl - 1, 1), ctx.stream));
      AXB_LAUNCH(ctx, sd_up_kernel, blocks_for(nl - 1, 256), 256, bn, s->bvh->node_range.as<int2>(), nl - 1, s->sdup.as<SdUp>());
      bps = 0;
      if(s->nv == 3)
      {
        AXB_CUDA_TRY(cudaFuncSetAttribute(sd_min_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSd2SmemMin));
        AXB_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, sd_min_kernel<3>, kSd2Threads, kSd2SmemMin));
      }
      else
      {
        AXB_CUDA_TRY(cudaFuncSetAttribute(sd_min_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSd2SmemMin));
        AXB_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, sd_min_kernel<4>, kSd2Threads, kSd2SmemMin));
      }
      s->two_blocks_per_sm = std::max(1, bps);
      s->kernel = 2;
      if(const char* e = getenv("AXB_SD_KERNEL"))
        if(!strcmp(e, "fast")) s->kernel = 1;
    }
    AXB_CUDA_TRY(cudaEventRecord(tot1, ctx.stream));
    return ctx.sync();
  };
  st = body();
  if(st == AXB_OK)
  {
    ctx.resolve();
    float tms = 0.f;
    if(tot0 && tot1 && cudaEventElapsedTime(&tms, tot0, tot1) == cudaSuccess) s->setmesh_ms["setmesh.total"] = tms;
    cudaGetLastError();
    for(const auto& kv : ctx.acc)
      if(kv.second.calls) s->setmesh_ms[kv.first] = kv.second.sum / (double)kv.second.calls;
  }
  ctx.set_profiling(false);
  if(tot0) cudaEventDestroy(tot0);
  if(tot1) cudaEventDestroy(tot1);
  if(st != AXB_OK)
  {
    axb_sd_destroy(s);
    return st;
  }
  *out = s;
  return AXB_OK;
}

int axb_sd_destroy(axb_sd* s)
{
  if(!s) return AXB_OK;
  if(s->bvh)
  {
    cudaSetDevice(s->ctx().device);
    cudaStream_t st = s->ctx().stream;
    for(DevBuf* b : {&s->x, &s->y, &s->z, &s->conn, &s->offsets, &s->soup, &s->cell_boxes, &s->obounds, &s->work, &s->sdnodes, &s->sdcens, &s->sdup, &s->sdnodes64})
      b->release(st);
    for(int k = 0; k < 2; ++k)
    {
      s->qb[k].release(st);
      if(s->pipe_stream[k])
      {
        cudaStreamSynchronize(s->pipe_stream[k]);
        cudaStreamDestroy(s->pipe_stream[k]);
      }
    }
    for(cudaEvent_t& e : s->pipe_event)
      if(e) cudaEventDestroy(e);
    axb_bvh_destroy(s->bvh);
  }
  delete s;
  return AXB_OK;
}

int axb_sd_set_stream(axb_sd* s, void* st) { return s ? axb_bvh_set_stream(s->bvh, st) : fail(AXB_ERR_BAD_ARG, "null handle"); }
int axb_sd_set_async(axb_sd* s, int e) { return s ? axb_bvh_set_async(s->bvh, e) : fail(AXB_ERR_BAD_ARG, "null handle"); }
int axb_sd_synchronize(axb_sd* s) { return s ? axb_bvh_synchronize(s->bvh) : fail(AXB_ERR_BAD_ARG, "null handle"); }
int axb_sd_set_profiling(axb_sd* s, int e)
{
  if(!s) return fail(AXB_ERR_BAD_ARG, "null handle");
  s->count_work = (e >= 2);  // level 2 also counts leaf tests / inner visits (adds a sync per query call)
  return axb_bvh_set_profiling(s->bvh, e);
}
int axb_sd_get_phase_ms(const axb_sd* s, const char* name, double* ms)
{
  if(!s || !name || !ms) return fail(AXB_ERR_BAD_ARG, "null argument");
  if(!strncmp(name, "setmesh.", 8) || !strncmp(name, "setmesh_build.", 14))
  {
    // the phases of setMesh, always recorded: "setmesh.total|upload|cell_boxes|gather_soup|obb_build", and the BVH build inside it
    // as "setmesh_build.total|bounds|morton|sort|agglo"
    const std::string key = !strncmp(name, "setmesh_build.", 14) ? std::string("build.") + (name + 14) : std::string(name);
    auto it = s->setmesh_ms.find(key);
    if(it == s->setmesh_ms.end()) return fail(AXB_ERR_BAD_ARG, std::string("no timing recorded for phase ") + name);
    *ms = it->second;
    return AXB_OK;
  }
  return axb_bvh_get_phase_ms(s->bvh, name, ms);
}
int axb_sd_launch_count(const axb_sd* s, int64_t* n) { return s ? axb_bvh_launch_count(s->bvh, n) : fail(AXB_ERR_BAD_ARG, "null handle"); }

int axb_sd_get_bvh(axb_sd* s, axb_bvh** b)
{
  if(!s || !b) return fail(AXB_ERR_BAD_ARG, "null argument");
  *b = s->bvh;
  return AXB_OK;
}

int axb_sd_get_mesh_bounds(const axb_sd* s, double* lo, double* hi)
{
  if(!s || !lo || !hi) return fail(AXB_ERR_BAD_ARG, "null argument");
  for(int d = 0; d < 3; ++d)
  {
    lo[d] = s->prm.dom_lo[d];
    hi[d] = s->prm.dom_hi[d];
  }
  return AXB_OK;
}

int axb_sd_set_mode(axb_sd* s, int mode)
{
  if(!s) return fail(AXB_ERR_BAD_ARG, "null handle");
  if(mode != 0 && mode != 1) return fail(AXB_ERR_BAD_ARG, "mode must be 0 (reference visiting order) or 1 (fast)");
  s->mode = mode;
  return AXB_OK;
}

int axb_sd_get_work_counters(const axb_sd* s, int64_t* leaf_tests, int64_t* inner_visits)
{
  if(!s || !leaf_tests || !inner_visits) return fail(AXB_ERR_BAD_ARG, "null argument");
  *leaf_tests = s->last_leaf_tests;
  *inner_visits = s->last_inner_visits;
  return AXB_OK;
}

static bool solo_on_env() { return getenv("AXB_SD_NO_SOLO") == nullptr; }

// One contiguous range of queries, enqueued on ctx.stream with the scratch set B: stage (host inputs), Morton
// order, the query kernel, and -- for host outputs -- the copies back.  Does not synchronise.
// What a partitioned-surface query exchanges between the sample pass and the search proper (sd_query_range calls it with
// the per-query distances to this part's sample points; it returns them MIN-reduced over the ranks, and the slack)
struct SdBoundExchange
{
  axb_comm* comm;
  double slack;                   // largest triangle diameter over ALL parts
  const double* given = nullptr;  // no exchange: the caller's own per-query bounds (distances, device memory)
};
static int sd_query_range(axb_sd* s, axb_sd::QBufs& B, const axb_array_desc* qpts, int32_t npts, double* phi, double* cps, double* nrms,
                          int out_memspace, unsigned long long* d_work, SdBoundExchange* ex = nullptr)
{
  Ctx& ctx = s->ctx();
  Desc<3> q;
  AXB_TRY(stage_desc<3>(ctx, qpts, npts, sizeof(double), B.q_stage, q));
  double *d_phi = phi, *d_cp = cps, *d_n = nrms;
  if(out_memspace == AXB_MEM_HOST)
  {
    AXB_TRY(B.out_phi.reserve(sizeof(double) * (size_t)npts, ctx.stream));
    d_phi = B.out_phi.as<double>();
    if(cps)
    {
      AXB_TRY(B.out_cp.reserve(sizeof(double) * 3 * (size_t)npts, ctx.stream));
      d_cp = B.out_cp.as<double>();
    }
    if(nrms)
    {
      AXB_TRY(B.out_n.reserve(sizeof(double) * 3 * (size_t)npts, ctx.stream));
      d_n = B.out_n.as<double>();
    }
  }
  const Node<double, 3>* nodes = s->bvh->nodes.as<Node<double, 3>>();
  if(s->mode == 1)
  {
    const int32_t* perm = nullptr;
    if(npts >= 4096)
    {
      // Morton order of the query points: neighbouring threads walk the same nodes
      ScopedPhase ph(ctx, "query.sortq");
      const size_t kb = sizeof(unsigned long long) * (size_t)npts;
      AXB_TRY(B.qkeys_a.reserve(kb, ctx.stream));
      AXB_TRY(B.qkeys_b.reserve(kb, ctx.stream));
      AXB_TRY(B.qperm.reserve(sizeof(int32_t) * (size_t)npts, ctx.stream));
      AXB_TRY(B.qbounds.reserve(sizeof(unsigned long long) * 6, ctx.stream));
      const size_t scratch = rsort::scratch_bytes(npts);
      AXB_TRY(B.qscratch.reserve(scratch, ctx.stream));
      AXB_CUDA_TRY(cudaMemsetAsync(B.qscratch.p, 0, scratch, ctx.stream));
      uint32_t* ghist = B.qscratch.as<uint32_t>();
      uint32_t* tile_counters = ghist + rsort::MAX_PASSES * rsort::RADIX;
      uint32_t* lookback = tile_counters + 64;
      // (a kernel, not a copy from pageable host memory: that would block the host until the stream gets there and
      // serialise the chunk pipeline)
      AXB_LAUNCH(ctx, init_query_bounds_kernel, 1, 32, B.qbounds.as<unsigned long long>());
      AXB_LAUNCH(ctx, query_bounds_kernel, capped_grid(npts, 256), 256, q, npts, B.qbounds.as<unsigned long long>());
      AXB_LAUNCH(ctx, query_keys_kernel, capped_grid(npts, 256), 256, q, npts, B.qbounds.as<unsigned long long>(),
                 B.qkeys_a.as<unsigned long long>(), ghist);
      unsigned long long* sorted = nullptr;
      AXB_TRY(sort_keys_generic(ctx, B.qkeys_a.as<unsigned long long>(), B.qkeys_b.as<unsigned long long>(), npts, ghist,
                                tile_counters, lookback, &sorted, 1));
      AXB_LAUNCH(ctx, keys_to_perm_kernel, blocks_for(npts, 256), 256, sorted, npts, B.qperm.as<int32_t>());
      perm = B.qperm.as<int32_t>();
    }
    ScopedPhase ph(ctx, "query.kernel");
    // Opt-in (AXB_SD_L2_PERSIST=<fraction of L2 to set aside, e.g. 0.5>): an L2 persisting access-policy window over the
    // compact node records for the duration of the query kernels.  Measured on C2 (profiles/r2zc_*): no gain -- the
    // search is not bound by any memory level and its L2 hit rate is 94 % without it -- so it is off by default.
    bool l2_window = false;
    if(const char* e = getenv("AXB_SD_L2_PERSIST"))
    {
      const double frac = atof(e);
      int l2 = 0, maxwin = 0;
      cudaDeviceGetAttribute(&l2, cudaDevAttrL2CacheSize, ctx.device);
      cudaDeviceGetAttribute(&maxwin, cudaDevAttrMaxAccessPolicyWindowSize, ctx.device);
      if(frac > 0.0 && l2 > 0 && maxwin > 0 && s->kernel == 2)
      {
        const size_t carve = (size_t)(frac * (double)l2);
        cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, carve);
        cudaStreamAttrValue av;
        memset(&av, 0, sizeof(av));
        const size_t bytes = std::min<size_t>(s->sdnodes64.cap, (size_t)maxwin);
        av.accessPolicyWindow.base_ptr = s->sdnodes64.p;
        av.accessPolicyWindow.num_bytes = bytes;
        av.accessPolicyWindow.hitRatio = (float)std::min(1.0, (double)carve / (double)std::max<size_t>(bytes, 1));
        av.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
        av.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
        l2_window = cudaStreamSetAttribute(ctx.stream, cudaStreamAttributeAccessPolicyWindow, &av) == cudaSuccess;
        cudaGetLastError();
      }
    }
    struct WindowOff
    {
      cudaStream_t st;
      bool on;
      ~WindowOff()
      {
        if(!on) return;
        cudaStreamAttrValue av;
        memset(&av, 0, sizeof(av));
        av.accessPolicyWindow.num_bytes = 0;
        cudaStreamSetAttribute(st, cudaStreamAttributeAccessPolicyWindow, &av);
        cudaCtxResetPersistingL2Cache();
        cudaGetLastError();
      }
    } window_off {ctx.stream, l2_window};
    // persistent warps: one resident wave, queries pulled from a device-side cursor
    AXB_CUDA_TRY(cudaMemsetAsync(B.cursor.p, 0, sizeof(unsigned int) * 4, ctx.stream));
    int sms = kNumSMsB200;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, ctx.device);
    const int grid = (int)std::min<long long>(blocks_for(npts, 128), (long long)sms * (s->kernel == 2 ? s->two_blocks_per_sm : s->fast_blocks_per_sm));
    // queries per cursor grab: a run of Morton neighbours, so a lane's consecutive queries are close and
    // the previous closest point is a useful first bound; small inputs keep every warp busy instead
    unsigned chunk = (unsigned)kQueryChunk;
    if(const char* e = getenv("AXB_SD_CHUNK")) chunk = (unsigned)std::max(32, atoi(e));
    if(s->kernel == 2)
    {
      // phase 1 (exact minimum + in-window leaves per query slot), phase 2 (state machine in the reference's order)
      AXB_TRY(B.cand.reserve(sizeof(int32_t) * kCandCap * (size_t)npts, ctx.stream));
      AXB_TRY(B.cand_n.reserve((size_t)npts, ctx.stream));
      AXB_TRY(B.seed.reserve(sizeof(double) * (size_t)npts, ctx.stream));
      const int grid2 = blocks_for(npts, kSd2Threads);
      // the sample pass: one query in 2^hint_shift (Morton order) searched first, its closest point is the first bound of
      // its neighbours in the search proper
      int hint_shift = 5;
      if(const char* e = getenv("AXB_SD_HINT_SHIFT")) hint_shift = atoi(e);
      double* hint_tab = nullptr;
      int hint_n = 0, hint_grid = 0;
      if(perm && hint_shift >= 3 && hint_shift <= 16 && npts >= (64 << hint_shift))
      {
        hint_n = (int)(((long long)npts + (1ll << hint_shift) - 1) >> hint_shift);
        hint_grid = (int)std::min<long long>(blocks_for(hint_n, 128), (long long)sms * s->two_blocks_per_sm);
        AXB_TRY(B.hint.reserve(sizeof(double) * 3 * (size_t)hint_n, ctx.stream));
        hint_tab = B.hint.as<double>();
      }
      // partitioned surface: after the sample pass the ranks agree on a bound per query (MIN over the ranks of the distance
      // to each part's sample point); a part far from the query then prunes it at the root
      const double* ext_bound = nullptr;
      double ext_slack = 0.0;
      auto exchange_bounds = [&]() -> int {
        if(!ex) return AXB_OK;
        if(ex->given)
        {
          ext_bound = ex->given;
          ext_slack = ex->slack;
          return AXB_OK;
        }
        if(!hint_tab) return AXB_OK;
        AXB_TRY(B.ext.reserve(sizeof(double) * (size_t)npts, ctx.stream));
        AXB_LAUNCH(ctx, sd_ext_bound_kernel, blocks_for(npts, 256), 256, q, perm, npts, (const double*)hint_tab, hint_shift, B.ext.as<double>());
        ScopedPhase pe(ctx, "query.bound_exchange");
        AXB_NCCL_TRY(ex->comm->api, ex->comm->api->AllReduce(B.ext.p, B.ext.p, (size_t)npts, ncclDouble, ncclMin, ex->comm->comm, ctx.stream));
        ex->comm->bytes += 8ll * npts;
        ex->comm->calls += 1;
        ext_bound = B.ext.as<double>();
        ext_slack = ex->slack;
        return AXB_OK;
      };
      unsigned heavy_visits = solo_on_env() ? 384u : 0xffffffffu;
      if(const char* e = getenv("AXB_SD_HEAVY")) heavy_visits = (unsigned)std::max(1, atoi(e));
      // heavy queries (list overflow, sign too close to call) are listed by the resolve kernel and finished one WARP each
      const unsigned solo_cap = (unsigned)std::min<long long>(npts, 1 << 20);
      const int solo_grid = std::min(2 * sms, blocks_for(npts, 32));
      unsigned int* solo_ctr = nullptr;
      int32_t* solo_list = nullptr;
      const bool solo_on = solo_on_env();  // (off: a test hook, the serial walk in the resolve lane)
      if(solo_on)
      {
        AXB_TRY(B.solo.reserve(16 + sizeof(int32_t) * (size_t)solo_cap, ctx.stream));
        AXB_TRY(B.solo_scratch.reserve(sizeof(unsigned long long) * kSoloCap * (size_t)solo_grid * (kSoloThreads / 32), ctx.stream));
        AXB_CUDA_TRY(cudaMemsetAsync(B.solo.p, 0, 16, ctx.stream));
        solo_ctr = B.solo.as<unsigned int>();
        solo_list = reinterpret_cast<int32_t*>(B.solo.as<char>() + 16);
      }
      if(s->nv == 3)
      {
        {
          ScopedPhase p1(ctx, "query.min");
          if(hint_tab)
          {
            AXB_LAUNCH_SMEM(ctx, sd_min_kernel<3>, hint_grid, kSd2Threads, kSd2SmemMin, s->sdnodes64.as<SdNode64>(), s->soup.as<double>(), q, hint_n, perm,
                            (int32_t*)nullptr, (uint8_t*)nullptr, (double*)nullptr, d_work, B.cursor.as<unsigned int>() + 1, 32u,
                            s->prm.compute_sign ? kTieWindow : 0.0, (const double*)nullptr, hint_shift, hint_tab, npts, heavy_visits, (const double*)nullptr, 0.0);
            AXB_TRY(exchange_bounds());
          }
          else if(ex && ex->given)
            AXB_TRY(exchange_bounds());
          AXB_LAUNCH_SMEM(ctx, sd_min_kernel<3>, grid, kSd2Threads, kSd2SmemMin, s->sdnodes64.as<SdNode64>(), s->soup.as<double>(), q, npts, perm,
                          B.cand.as<int32_t>(), B.cand_n.as<uint8_t>(), B.seed.as<double>(), d_work, B.cursor.as<unsigned int>(), chunk,
                          s->prm.compute_sign ? kTieWindow : 0.0, (const double*)hint_tab, hint_shift, (double*)nullptr, npts, heavy_visits, ext_bound, ext_slack);
        }
        ScopedPhase p2(ctx, "query.resolve");
        AXB_LAUNCH(ctx, sd_resolve_kernel<3>, grid2, kSd2Threads, s->sdnodes.as<SdNode>(), s->sdcens.as<SdCen>(), s->sdup.as<SdUp>(),
                   s->bvh->leaf_parent.as<int32_t>(), s->soup.as<double>(), s->prm, q, npts, perm, B.cand.as<int32_t>(), B.cand_n.as<uint8_t>(),
                   B.seed.as<double>(), d_phi, d_cp, d_n, d_work, solo_ctr, solo_cap, solo_list);
        if(solo_on)
          AXB_LAUNCH(ctx, sd_solo_kernel<3>, solo_grid, kSoloThreads, s->sdnodes.as<SdNode>(), s->sdcens.as<SdCen>(), s->soup.as<double>(), s->prm, q,
                     perm, B.seed.as<double>(), solo_ctr, solo_cap, solo_list, B.solo_scratch.as<unsigned long long>(), d_phi, d_cp, d_n, d_work);
      }
      else
      {
        {
          ScopedPhase p1(ctx, "query.min");
          if(hint_tab)
          {
            AXB_LAUNCH_SMEM(ctx, sd_min_kernel<4>, hint_grid, kSd2Threads, kSd2SmemMin, s->sdnodes64.as<SdNode64>(), s->soup.as<double>(), q, hint_n, perm,
                            (int32_t*)nullptr, (uint8_t*)nullptr, (double*)nullptr, d_work, B.cursor.as<unsigned int>() + 1, 32u,
                            s->prm.compute_sign ? kTieWindow : 0.0, (const double*)nullptr, hint_shift, hint_tab, npts, heavy_visits, (const double*)nullptr, 0.0);
            AXB_TRY(exchange_bounds());
          }
          else if(ex && ex->given)
            AXB_TRY(exchange_bounds());
          AXB_LAUNCH_SMEM(ctx, sd_min_kernel<4>, grid, kSd2Threads, kSd2SmemMin, s->sdnodes64.as<SdNode64>(), s->soup.as<double>(), q, npts, perm,
                          B.cand.as<int32_t>(), B.cand_n.as<uint8_t>(), B.seed.as<double>(), d_work, B.cursor.as<unsigned int>(), chunk,
                          s->prm.compute_sign ? kTieWindow : 0.0, (const double*)hint_tab, hint_shift, (double*)nullptr, npts, heavy_visits, ext_bound, ext_slack);
        }
        ScopedPhase p2(ctx, "query.resolve");
        AXB_LAUNCH(ctx, sd_resolve_kernel<4>, grid2, kSd2Threads, s->sdnodes.as<SdNode>(), s->sdcens.as<SdCen>(), s->sdup.as<SdUp>(),
                   s->bvh->leaf_parent.as<int32_t>(), s->soup.as<double>(), s->prm, q, npts, perm, B.cand.as<int32_t>(), B.cand_n.as<uint8_t>(),
                   B.seed.as<double>(), d_phi, d_cp, d_n, d_work, solo_ctr, solo_cap, solo_list);
        if(solo_on)
          AXB_LAUNCH(ctx, sd_solo_kernel<4>, solo_grid, kSoloThreads, s->sdnodes.as<SdNode>(), s->sdcens.as<SdCen>(), s->soup.as<double>(), s->prm, q,
                     perm, B.seed.as<double>(), solo_ctr, solo_cap, solo_list, B.solo_scratch.as<unsigned long long>(), d_phi, d_cp, d_n, d_work);
      }
    }
    else if(s->nv == 3)
      AXB_LAUNCH_SMEM(ctx, sd_fast_kernel<3>, grid, 128, kSdFastSmem, s->sdnodes.as<SdNode>(), s->sdcens.as<SdCen>(), s->soup.as<double>(), s->prm, q,
                      npts, perm, d_phi, d_cp, d_n, d_work, B.cursor.as<unsigned int>(), chunk);
    else
      AXB_LAUNCH_SMEM(ctx, sd_fast_kernel<4>, grid, 128, kSdFastSmem, s->sdnodes.as<SdNode>(), s->sdcens.as<SdCen>(), s->soup.as<double>(), s->prm, q,
                      npts, perm, d_phi, d_cp, d_n, d_work, B.cursor.as<unsigned int>(), chunk);
  }
  else
  {
    ScopedPhase ph(ctx, "query.kernel");
    if(s->nv == 3)
      AXB_LAUNCH(ctx, sd_reference_order_kernel<3>, blocks_for(npts, 128), 128, nodes, s->soup.as<double>(), s->prm, q, npts,
                 (const int32_t*)nullptr, d_phi, d_cp, d_n, d_work);
    else
      AXB_LAUNCH(ctx, sd_reference_order_kernel<4>, blocks_for(npts, 128), 128, nodes, s->soup.as<double>(), s->prm, q, npts,
                 (const int32_t*)nullptr, d_phi, d_cp, d_n, d_work);
  }
  if(out_memspace == AXB_MEM_HOST)
  {
    AXB_CUDA_TRY(cudaMemcpyAsync(phi, d_phi, sizeof(double) * (size_t)npts, cudaMemcpyDeviceToHost, ctx.stream));
    if(cps) AXB_CUDA_TRY(cudaMemcpyAsync(cps, d_cp, sizeof(double) * 3 * (size_t)npts, cudaMemcpyDeviceToHost, ctx.stream));
    if(nrms) AXB_CUDA_TRY(cudaMemcpyAsync(nrms, d_n, sizeof(double) * 3 * (size_t)npts, cudaMemcpyDeviceToHost, ctx.stream));
  }
  return AXB_OK;
}

int axb_sd_compute_distances(axb_sd* s, const axb_array_desc* qpts, int32_t npts, double* phi, double* cps, double* nrms, int out_memspace)
{
  if(!s) return fail(AXB_ERR_BAD_ARG, "null handle");
  if(npts < 0) return fail(AXB_ERR_BAD_ARG, "negative point count");
  if(npts > 0 && !phi) return fail(AXB_ERR_BAD_ARG, "outSgnDist != nullptr");
  if(npts > 0 && !qpts) return fail(AXB_ERR_BAD_ARG, "null query descriptor");
  out_memspace = resolve_memspace(out_memspace, phi);
  if(out_memspace != AXB_MEM_HOST && out_memspace != AXB_MEM_DEVICE) return fail(AXB_ERR_BAD_ARG, "unknown output memspace");
  if(!s->bvh->built) return fail(AXB_ERR_NOT_BUILT, "SignedDistance query before setMesh()");
  Ctx& ctx = s->ctx();
  AXB_TRY(ctx.bind());
  ctx.begin_call();
  if(npts == 0) return AXB_OK;
  const int tot = ctx.phase_begin("query.total");
  unsigned long long* d_work = nullptr;
  if(s->count_work)
  {
    AXB_TRY(s->work.reserve(sizeof(unsigned long long) * 80, ctx.stream));
    AXB_CUDA_TRY(cudaMemsetAsync(s->work.p, 0, sizeof(unsigned long long) * 80, ctx.stream));
    d_work = s->work.as<unsigned long long>();
  }
  // With host inputs AND host outputs, cut the call into chunks (AXB_SD_PIPE_CHUNK = points per chunk, 0 = off) that
  // alternate between two streams, so the PCIe traffic of neighbouring chunks hides behind the kernel (pinned host
  // buffers).  The result does not depend on the chunking: queries are independent.
  // Default: three chunks once the call is large (>= 4 M points).  Measured on C2 with pinned buffers (profiles/r2zd_*):
  // unchunked 69.4 ms, 2 chunks 66.9, 3 chunks 66.1, 4 chunks 67.1, 8 chunks 69.0 (58.6 ms of it kernels): since the
  // sample pass and the cooperative heavy-query kernel, a 5.6 M-point launch costs little more per point than a 16.8 M one.
  long long pipe_chunk = npts >= (4 << 20) ? ((long long)npts + 2) / 3 : 0;
  if(const char* e = getenv("AXB_SD_PIPE_CHUNK")) pipe_chunk = atoll(e);
  const bool host_in = resolve_memspace(qpts->memspace, qpts->comp[0]) == AXB_MEM_HOST;
  if(host_in && out_memspace == AXB_MEM_HOST && !ctx.async && pipe_chunk > 0 && (long long)npts >= 2 * pipe_chunk)
  {
    for(int k = 0; k < 2; ++k)
      if(!s->pipe_stream[k]) AXB_CUDA_TRY(cudaStreamCreateWithFlags(&s->pipe_stream[k], cudaStreamNonBlocking));
    for(cudaEvent_t& e : s->pipe_event)
      if(!e) AXB_CUDA_TRY(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    cudaStream_t main_stream = ctx.stream;
    AXB_CUDA_TRY(cudaEventRecord(s->pipe_event[2], main_stream));
    for(int k = 0; k < 2; ++k) AXB_CUDA_TRY(cudaStreamWaitEvent(s->pipe_stream[k], s->pipe_event[2], 0));
    int st = AXB_OK;
    int k = 0;
    for(long long first = 0; first < npts && st == AXB_OK; first += pipe_chunk, k ^= 1)
    {
      const int32_t n = (int32_t)std::min<long long>(pipe_chunk, npts - first);
      axb_array_desc sub = *qpts;
      for(int c = 0; c < sub.ncomp && c < 6; ++c) sub.comp[c] = (const char*)qpts->comp[c] + first * qpts->stride_bytes;
      ctx.stream = s->pipe_stream[k];  // everything sd_query_range enqueues goes to this chunk's stream
      st = sd_query_range(s, s->qb[k], &sub, n, phi + first, cps ? cps + 3 * first : nullptr, nrms ? nrms + 3 * first : nullptr,
                          AXB_MEM_HOST, d_work);
    }
    ctx.stream = main_stream;
    for(int j = 0; j < 2; ++j)
    {
      AXB_CUDA_TRY(cudaEventRecord(s->pipe_event[j], s->pipe_stream[j]));
      AXB_CUDA_TRY(cudaStreamWaitEvent(main_stream, s->pipe_event[j], 0));
    }
    AXB_TRY(st);
  }
  else
  {
    AXB_TRY(sd_query_range(s, s->qb[0], qpts, npts, phi, cps, nrms, out_memspace, d_work));
  }
  ctx.phase_end(tot);
  if(out_memspace == AXB_MEM_HOST) AXB_TRY(ctx.sync());
  if(d_work)
  {
    unsigned long long hw[2];
    AXB_CUDA_TRY(cudaMemcpyAsync(hw, d_work, sizeof(hw), cudaMemcpyDeviceToHost, ctx.stream));
    AXB_TRY(ctx.sync());
    s->last_leaf_tests = (int64_t)hw[0];
    s->last_inner_visits = (int64_t)hw[1];
#ifdef AXB_SD_DEBUG_MISS
    {
      unsigned long long dbg[80];
      AXB_CUDA_TRY(cudaMemcpy(dbg, d_work, sizeof(dbg), cudaMemcpyDeviceToHost));
      fprintf(stderr, "[sd debug] queries that found nothing: %llu\n", dbg[2]);
      for(unsigned long long k = 0; k < std::min<unsigned long long>(dbg[2], 4); ++k)
      {
        const unsigned long long* w = dbg + 4 + 18 * k;
        auto D = [](unsigned long long b) { double v; memcpy(&v, &b, 8); return v; };
        unsigned ml = (unsigned)w[8];
        float mlf;
        memcpy(&mlf, &ml, 4);
        fprintf(stderr, "   slot %llu lane %llu: visits %llu, leaves pushed %llu returned %llu, h_own %.17g h_warp %.17g h_tab %.17g thr0 %.17g thr %.17g min leaf bound %.9g\n",
                w[0], w[9], w[1], w[2] >> 32, w[2] & 0xffffffffull, D(w[3]), D(w[4]), D(w[5]), D(w[6]), D(w[7]), (double)mlf);
        fprintf(stderr, "   MISS q %.17g %.17g %.17g hint %.17g %.17g %.17g\n", D(w[10]), D(w[11]), D(w[12]), D(w[13]), D(w[14]), D(w[15]));
      }
    }
#endif
  }
  return ctx.finish_call();
}

}  // extern "C"

//==========================================================================================
// quest::findTriMeshIntersectionsBVH (quest/MeshTester.hpp:67-104): broad phase + exact narrow phase
//==========================================================================================
struct axb_meshtester
{
  axb_bvh* bvh = nullptr;  // owns
  int ncells = 0;
  int nnodes = 0;
  DevBuf tris, boxes, degflag, off, cnt;
  Ctx& ctx() { return bvh->ctx; }
};

extern "C" {

int axb_meshtester_create(axb_meshtester** out, int device, const double* x, const double* y, const double* z, int32_t nnodes,
                          const int32_t* conn, int32_t ncells, int mesh_memspace)
{
  if(!out) return fail(AXB_ERR_BAD_ARG, "null output handle");
  *out = nullptr;
  if(nnodes < 0 || ncells < 0) return fail(AXB_ERR_BAD_ARG, "negative mesh size");
  if((nnodes > 0 && (!x || !y || !z)) || (ncells > 0 && !conn)) return fail(AXB_ERR_BAD_ARG, "null mesh array");
  mesh_memspace = resolve_memspace(mesh_memspace, x);
  if(mesh_memspace != AXB_MEM_HOST && mesh_memspace != AXB_MEM_DEVICE) return fail(AXB_ERR_BAD_ARG, "unknown memspace");
  axb_meshtester* m = new axb_meshtester();
  int st = axb_bvh_create(&m->bvh, 3, 8, device);
  if(st != AXB_OK)
  {
    delete m;
    return st;
  }
  m->ncells = ncells;
  m->nnodes = nnodes;
  Ctx& ctx = m->ctx();
  auto body = [&]() -> int {
    AXB_TRY(ctx.bind());
    const cudaMemcpyKind kind = mesh_memspace == AXB_MEM_HOST ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice;
    const size_t nb = sizeof(double) * (size_t)nnodes, cb = sizeof(int32_t) * 3 * (size_t)ncells;
    DevBuf dx, dy, dz, dc;
    const double *px = x, *py = y, *pz = z;
    const int32_t* pc = conn;
    if(mesh_memspace == AXB_MEM_HOST)
    {
      AXB_TRY(dx.reserve(nb, ctx.stream));
      AXB_TRY(dy.reserve(nb, ctx.stream));
      AXB_TRY(dz.reserve(nb, ctx.stream));
      AXB_TRY(dc.reserve(cb, ctx.stream));
      if(nnodes)
      {
        AXB_CUDA_TRY(cudaMemcpyAsync(dx.p, x, nb, kind, ctx.stream));
        AXB_CUDA_TRY(cudaMemcpyAsync(dy.p, y, nb, kind, ctx.stream));
        AXB_CUDA_TRY(cudaMemcpyAsync(dz.p, z, nb, kind, ctx.stream));
      }
      if(cb) AXB_CUDA_TRY(cudaMemcpyAsync(dc.p, conn, cb, kind, ctx.stream));
      px = dx.as<double>();
      py = dy.as<double>();
      pz = dz.as<double>();
      pc = dc.as<int32_t>();
    }
    const size_t nc = (size_t)std::max(ncells, 1);
    AXB_TRY(m->tris.reserve(sizeof(double) * tt::kTriRecDoubles * nc, ctx.stream));
    AXB_TRY(m->boxes.reserve(sizeof(Box<double, 3>) * nc, ctx.stream));
    AXB_TRY(m->degflag.reserve(sizeof(int32_t) * nc, ctx.stream));
    if(ncells)
      AXB_LAUNCH(ctx, tri_prepare_kernel, blocks_for(ncells, 256), 256, px, py, pz, pc, ncells, m->tris.as<double>(),
                 m->boxes.as<Box<double, 3>>(), m->degflag.as<int32_t>());
    for(DevBuf* b : {&dx, &dy, &dz, &dc}) b->release(ctx.stream);
    // spin::BVH over the triangle AABBs, default scale factor (MeshTester_detail.hpp:325-327)
    axb_array_desc bd;
    memset(&bd, 0, sizeof(bd));
    for(int c = 0; c < 6; ++c) bd.comp[c] = m->boxes.as<char>() + 8 * c;
    bd.stride_bytes = 48;
    bd.ncomp = 6;
    bd.memspace = AXB_MEM_DEVICE;
    return axb_bvh_initialize(m->bvh, &bd, ncells);
  };
  st = body();
  if(st != AXB_OK)
  {
    axb_meshtester_destroy(m);
    return st;
  }
  *out = m;
  return AXB_OK;
}

int axb_meshtester_destroy(axb_meshtester* m)
{
  if(!m) return AXB_OK;
  if(m->bvh)
  {
    cudaSetDevice(m->ctx().device);
    cudaStream_t st = m->ctx().stream;
    for(DevBuf* b : {&m->tris, &m->boxes, &m->degflag, &m->off, &m->cnt}) b->release(st);
    axb_bvh_destroy(m->bvh);
  }
  delete m;
  return AXB_OK;
}

int axb_meshtester_get_bvh(axb_meshtester* m, axb_bvh** bvh)
{
  if(!m || !bvh) return fail(AXB_ERR_BAD_ARG, "null argument");
  *bvh = m->bvh;
  return AXB_OK;
}

// CandidateFinderBase::findTriMeshIntersections (MeshTester_detail.hpp:201-307) with the BVH candidate finder
// (:313-340): findBoundingBoxes(own AABBs), keep i < candidate, primal::intersect(tri_i, tri_c, false, threshold).
// One BVH walk does all three; pairs come back in the SEQ_EXEC order (i ascending, then candidate DFS order).
int axb_meshtester_find_intersections(axb_meshtester* m, double intersection_threshold, int out_memspace, int32_t** first, int32_t** second,
                                      int64_t* npairs)
{
  if(!m || !first || !second || !npairs) return fail(AXB_ERR_BAD_ARG, "null argument");
  if(out_memspace != AXB_MEM_HOST && out_memspace != AXB_MEM_DEVICE) return fail(AXB_ERR_BAD_ARG, "out_memspace must be HOST or DEVICE");
  *first = *second = nullptr;
  *npairs = 0;
  if(m->ncells == 0) return AXB_OK;
  Ctx& ctx = m->ctx();
  AXB_TRY(ctx.bind());
  AXB_TRY(m->off.reserve(sizeof(int32_t) * (size_t)m->ncells, ctx.stream));
  AXB_TRY(m->cnt.reserve(sizeof(int32_t) * (size_t)m->ncells, ctx.stream));
  axb_array_desc bd;
  memset(&bd, 0, sizeof(bd));
  for(int c = 0; c < 6; ++c) bd.comp[c] = m->boxes.as<char>() + 8 * c;
  bd.stride_bytes = 48;
  bd.ncomp = 6;
  bd.memspace = AXB_MEM_DEVICE;
  TriTriFilter filt;
  filt.query_tris = m->tris.as<double>();
  filt.tree_tris = m->tris.as<double>();
  filt.eps = intersection_threshold;
  filt.upper_only = 1;
  filt.include_boundary = 0;
  int32_t *d_second = nullptr, *d_first = nullptr;
  int64_t total = 0;
  AXB_TRY((find_impl<double, 3, BoxQuery<double, 3>, TriTriFilter>(m->bvh, 3, &bd, 0, m->ncells, m->off.as<int32_t>(), m->cnt.as<int32_t>(),
                                                                   AXB_MEM_DEVICE, &d_second, &total, filt, &d_first)));
  if(out_memspace == AXB_MEM_HOST)
  {
    int32_t* hf = (int32_t*)malloc(sizeof(int32_t) * (size_t)std::max<int64_t>(total, 1));
    int32_t* hs = (int32_t*)malloc(sizeof(int32_t) * (size_t)std::max<int64_t>(total, 1));
    if(!hf || !hs) return fail(AXB_ERR_BAD_ARG, "host allocation of the pair arrays failed");
    AXB_CUDA_TRY(cudaMemcpyAsync(hf, d_first, sizeof(int32_t) * (size_t)total, cudaMemcpyDeviceToHost, ctx.stream));
    AXB_CUDA_TRY(cudaMemcpyAsync(hs, d_second, sizeof(int32_t) * (size_t)total, cudaMemcpyDeviceToHost, ctx.stream));
    AXB_CUDA_TRY(cudaFreeAsync(d_first, ctx.stream));
    AXB_CUDA_TRY(cudaFreeAsync(d_second, ctx.stream));
    AXB_TRY(ctx.sync());
    *first = hf;
    *second = hs;
  }
  else
  {
    *first = d_first;
    *second = d_second;
  }
  *npairs = total;
  return AXB_OK;
}

// indices of the degenerate triangles, ascending (MeshTester_detail.hpp:297-305)
int axb_meshtester_get_degenerate(axb_meshtester* m, int out_memspace, int32_t** indices, int64_t* n)
{
  if(!m || !indices || !n) return fail(AXB_ERR_BAD_ARG, "null argument");
  if(out_memspace != AXB_MEM_HOST && out_memspace != AXB_MEM_DEVICE) return fail(AXB_ERR_BAD_ARG, "out_memspace must be HOST or DEVICE");
  *indices = nullptr;
  *n = 0;
  if(m->ncells == 0) return AXB_OK;
  Ctx& ctx = m->ctx();
  AXB_TRY(ctx.bind());
  AXB_TRY(m->off.reserve(sizeof(int32_t) * (size_t)m->ncells, ctx.stream));
  AXB_TRY(m->bvh->q_total.reserve(sizeof(long long), ctx.stream));
  long long* d_total = m->bvh->q_total.as<long long>();
  AXB_TRY(exclusive_scan(m->bvh, m->degflag.as<int32_t>(), m->ncells, m->off.as<int32_t>(), d_total));
  long long htotal = 0;
  AXB_CUDA_TRY(cudaMemcpyAsync(&htotal, d_total, sizeof(long long), cudaMemcpyDeviceToHost, ctx.stream));
  AXB_TRY(ctx.sync());
  int32_t* d_idx = nullptr;
  AXB_CUDA_TRY(axb_malloc_async((void**)&d_idx, sizeof(int32_t) * (size_t)std::max<long long>(htotal, 1), ctx.stream));
  AXB_LAUNCH(ctx, compact_flagged_kernel, blocks_for(m->ncells, 256), 256, m->degflag.as<int32_t>(), m->off.as<int32_t>(), m->ncells, d_idx);
  if(out_memspace == AXB_MEM_HOST)
  {
    int32_t* h = (int32_t*)malloc(sizeof(int32_t) * (size_t)std::max<long long>(htotal, 1));
    if(!h) return fail(AXB_ERR_BAD_ARG, "host allocation failed");
    AXB_CUDA_TRY(cudaMemcpyAsync(h, d_idx, sizeof(int32_t) * (size_t)htotal, cudaMemcpyDeviceToHost, ctx.stream));
    AXB_CUDA_TRY(cudaFreeAsync(d_idx, ctx.stream));
    AXB_TRY(ctx.sync());
    *indices = h;
  }
  else
  {
    AXB_TRY(ctx.sync());
    *indices = d_idx;
  }
  *n = htotal;
  return AXB_OK;
}

int axb_meshtester_free(axb_meshtester* m, int32_t* p, int memspace)
{
  if(!p) return AXB_OK;
  if(!m) return fail(AXB_ERR_BAD_ARG, "null handle");
  return axb_bvh_free_candidates(m->bvh, p, memspace);
}

// primal::intersect(Triangle3, Triangle3, includeBoundary, EPS) on n explicit pairs (9 doubles per triangle);
// tris1 / tris2 / out live in `memspace` (host buffers are staged).  Synchronous.
int axb_tri_tri_intersect(int device, const double* tris1, const double* tris2, int64_t n, int memspace, int include_boundary, double eps,
                          uint8_t* out)
{
  if(n < 0 || (n > 0 && (!tris1 || !tris2 || !out))) return fail(AXB_ERR_BAD_ARG, "null or negative argument");
  memspace = resolve_memspace(memspace, tris1);
  if(memspace != AXB_MEM_HOST && memspace != AXB_MEM_DEVICE) return fail(AXB_ERR_BAD_ARG, "unknown memspace");
  Ctx ctx;
  AXB_TRY(ctx.init(device));
  auto body = [&]() -> int {
    if(n == 0) return AXB_OK;
    DevBuf a, b, o;
    const double *pa = tris1, *pb = tris2;
    uint8_t* po = out;
    const size_t tb = sizeof(double) * 9 * (size_t)n;
    if(memspace == AXB_MEM_HOST)
    {
      AXB_TRY(a.reserve(tb, ctx.stream));
      AXB_TRY(b.reserve(tb, ctx.stream));
      AXB_TRY(o.reserve((size_t)n, ctx.stream));
      AXB_CUDA_TRY(cudaMemcpyAsync(a.p, tris1, tb, cudaMemcpyHostToDevice, ctx.stream));
      AXB_CUDA_TRY(cudaMemcpyAsync(b.p, tris2, tb, cudaMemcpyHostToDevice, ctx.stream));
      pa = a.as<double>();
      pb = b.as<double>();
      po = o.as<uint8_t>();
    }
    AXB_LAUNCH(ctx, tri_tri_pairs_kernel, blocks_for(n, 256), 256, pa, pb, (long long)n, include_boundary, eps, po);
    if(memspace == AXB_MEM_HOST) AXB_CUDA_TRY(cudaMemcpyAsync(out, po, (size_t)n, cudaMemcpyDeviceToHost, ctx.stream));
    for(DevBuf* d : {&a, &b, &o}) d->release(ctx.stream);
    return ctx.sync();
  };
  const int st = body();
  ctx.destroy();
  return st;
}

}  // extern "C"

// ---- leaf arithmetic on n independent items (leafmath.cuh) ----
namespace
{
struct LeafIO
{
  const void* in[2] = {nullptr, nullptr};
  size_t in_bytes[2] = {0, 0};
  void* out[2] = {nullptr, nullptr};
  size_t out_bytes[2] = {0, 0};
};
// stage host inputs, run `launch(in0, in1, out0, out1)`, copy host outputs back
template <class F>
int leaf_op(int device, int memspace, int64_t n, LeafIO io, F&& launch)
{
  if(n < 0) return fail(AXB_ERR_BAD_ARG, "negative item count");
  for(int k = 0; k < 2; ++k)
    if(n > 0 && ((io.in_bytes[k] && !io.in[k]) || (io.out_bytes[k] && !io.out[k]))) return fail(AXB_ERR_BAD_ARG, "null argument");
  memspace = resolve_memspace(memspace, io.in[0]);
  if(memspace != AXB_MEM_HOST && memspace != AXB_MEM_DEVICE) return fail(AXB_ERR_BAD_ARG, "unknown memspace");
  Ctx ctx;
  AXB_TRY(ctx.init(device));
  auto body = [&]() -> int {
    if(n == 0) return AXB_OK;
    DevBuf bi[2], bo[2];
    const void* pi[2] = {io.in[0], io.in[1]};
    void* po[2] = {io.out[0], io.out[1]};
    if(memspace == AXB_MEM_HOST)
    {
      for(int k = 0; k < 2; ++k)
      {
        if(io.in_bytes[k])
        {
          AXB_TRY(bi[k].reserve(io.in_bytes[k], ctx.stream));
          AXB_CUDA_TRY(cudaMemcpyAsync(bi[k].p, io.in[k], io.in_bytes[k], cudaMemcpyHostToDevice, ctx.stream));
          pi[k] = bi[k].p;
        }
        if(io.out_bytes[k])
        {
          AXB_TRY(bo[k].reserve(io.out_bytes[k], ctx.stream));
          po[k] = bo[k].p;
        }
      }
    }
    AXB_TRY(launch(ctx, pi[0], pi[1], po[0], po[1]));
    if(memspace == AXB_MEM_HOST)
      for(int k = 0; k < 2; ++k)
        if(io.out_bytes[k]) AXB_CUDA_TRY(cudaMemcpyAsync(io.out[k], po[k], io.out_bytes[k], cudaMemcpyDeviceToHost, ctx.stream));
    for(int k = 0; k < 2; ++k)
    {
      bi[k].release(ctx.stream);
      bo[k].release(ctx.stream);
    }
    return ctx.sync();
  };
  const int st = body();
  ctx.destroy();
  return st;
}
}  // namespace

extern "C" {

int axb_closest_point_tri(int device, const double* pts, const double* tris, int64_t n, int memspace, double eps, double* cp, int32_t* loc)
{
  LeafIO io;
  io.in[0] = pts, io.in_bytes[0] = sizeof(double) * 3 * (size_t)std::max<int64_t>(n, 0);
  io.in[1] = tris, io.in_bytes[1] = sizeof(double) * 9 * (size_t)std::max<int64_t>(n, 0);
  io.out[0] = cp, io.out_bytes[0] = sizeof(double) * 3 * (size_t)std::max<int64_t>(n, 0);
  io.out[1] = loc, io.out_bytes[1] = sizeof(int32_t) * (size_t)std::max<int64_t>(n, 0);
  return leaf_op(device, memspace, n, io, [&](Ctx& ctx, const void* a, const void* b, void* o0, void* o1) -> int {
    AXB_LAUNCH(ctx, closest_point_tri_kernel, blocks_for(n, 256), 256, (const double*)a, (const double*)b, (long long)n, eps, (double*)o0,
               (int32_t*)o1);
    return AXB_OK;
  });
}

int axb_squared_distance_point_box(int device, const double* pts, const double* boxes, int64_t n, int memspace, double* out)
{
  LeafIO io;
  io.in[0] = pts, io.in_bytes[0] = sizeof(double) * 3 * (size_t)std::max<int64_t>(n, 0);
  io.in[1] = boxes, io.in_bytes[1] = sizeof(double) * 6 * (size_t)std::max<int64_t>(n, 0);
  io.out[0] = out, io.out_bytes[0] = sizeof(double) * (size_t)std::max<int64_t>(n, 0);
  return leaf_op(device, memspace, n, io, [&](Ctx& ctx, const void* a, const void* b, void* o0, void*) -> int {
    AXB_LAUNCH(ctx, sqdist_point_box_kernel, blocks_for(n, 256), 256, (const double*)a, (const double*)b, (long long)n, (double*)o0);
    return AXB_OK;
  });
}

int axb_intersect_ray_box(int device, const double* rays, const double* boxes, int64_t n, int memspace, int rays_normalized, double tol,
                          uint8_t* out)
{
  LeafIO io;
  io.in[0] = rays, io.in_bytes[0] = sizeof(double) * 6 * (size_t)std::max<int64_t>(n, 0);
  io.in[1] = boxes, io.in_bytes[1] = sizeof(double) * 6 * (size_t)std::max<int64_t>(n, 0);
  io.out[0] = out, io.out_bytes[0] = (size_t)std::max<int64_t>(n, 0);
  return leaf_op(device, memspace, n, io, [&](Ctx& ctx, const void* a, const void* b, void* o0, void*) -> int {
    AXB_LAUNCH(ctx, ray_box_kernel, blocks_for(n, 256), 256, (const double*)a, (const double*)b, (long long)n, rays_normalized, tol,
               (uint8_t*)o0);
    return AXB_OK;
  });
}

int axb_box_scale(int device, const double* boxes_in, int64_t n, int memspace, double scale_factor, double* boxes_out)
{
  LeafIO io;
  io.in[0] = boxes_in, io.in_bytes[0] = sizeof(double) * 6 * (size_t)std::max<int64_t>(n, 0);
  io.out[0] = boxes_out, io.out_bytes[0] = sizeof(double) * 6 * (size_t)std::max<int64_t>(n, 0);
  const double half_scale = static_cast<double>(scale_factor * 0.5);
  return leaf_op(device, memspace, n, io, [&](Ctx& ctx, const void* a, const void*, void* o0, void*) -> int {
    AXB_LAUNCH(ctx, box_scale_kernel, blocks_for(n, 256), 256, (const double*)a, (long long)n, half_scale, (double*)o0);
    return AXB_OK;
  });
}

}  // extern "C"

//==========================================================================================
// quest::DistributedClosestPoint, per-rank step (quest/detail/DistributedClosestPointImpl.hpp:883-1079)
//==========================================================================================
struct axb_dcp
{
  axb_bvh* bvh = nullptr;  // owns; built over the object points' zero-size boxes
  int ndims = 3;
  int npts = 0;
  bool tree_built = false;
  double sq_thresh = DBL_MAX;  // m_sqDistanceThreshold default (:252)
  DevBuf pts, dom, boxes;
  DevBuf q_stage, st_idx, st_dom, st_rank, st_coords, st_dist, keys_a, keys_b, scratch, perm;
  int mode = 1;  // 1 = nearest-first search with explicit tie-break (default), 0 = the reference's traversal order
  Ctx& ctx() { return bvh->ctx; }
};

static void dcpx_release(axb_dcp* h, cudaStream_t st);

extern "C" {

int axb_dcp_set_mode(axb_dcp* h, int mode)
{
  if(!h) return fail(AXB_ERR_BAD_ARG, "null handle");
  if(mode != 0 && mode != 1) return fail(AXB_ERR_BAD_ARG, "mode must be 0 or 1");
  h->mode = mode;
  return AXB_OK;
}

int axb_dcp_create(axb_dcp** out, int ndims, int device)
{
  if(!out) return fail(AXB_ERR_BAD_ARG, "null output handle");
  *out = nullptr;
  if(ndims != 2 && ndims != 3) return fail(AXB_ERR_BAD_ARG, "DistributedClosestPoint is 2-D or 3-D");
  axb_dcp* h = new axb_dcp();
  h->ndims = ndims;
  const int st = axb_bvh_create(&h->bvh, ndims, 8, device);
  if(st != AXB_OK)
  {
    delete h;
    return st;
  }
  *out = h;
  return AXB_OK;
}

int axb_dcp_destroy(axb_dcp* h)
{
  if(!h) return AXB_OK;
  if(h->bvh)
  {
    cudaSetDevice(h->ctx().device);
    cudaStream_t st = h->ctx().stream;
    for(DevBuf* b : {&h->pts, &h->dom, &h->boxes, &h->q_stage, &h->st_idx, &h->st_dom, &h->st_rank, &h->st_coords, &h->st_dist, &h->keys_a,
                     &h->keys_b, &h->scratch, &h->perm})
      b->release(st);
    dcpx_release(h, st);
    axb_bvh_destroy(h->bvh);
  }
  delete h;
  return AXB_OK;
}

// importObjectPoints (:590-649): interleaved coordinates of all local domains, one domain id per point
int axb_dcp_set_object_points(axb_dcp* h, const double* coords, const int32_t* domain_ids, int32_t npts, int memspace)
{
  if(!h) return fail(AXB_ERR_BAD_ARG, "null handle");
  if(npts < 0 || (npts > 0 && (!coords || !domain_ids))) return fail(AXB_ERR_BAD_ARG, "null or negative object point arrays");
  memspace = resolve_memspace(memspace, coords);
  if(memspace != AXB_MEM_HOST && memspace != AXB_MEM_DEVICE) return fail(AXB_ERR_BAD_ARG, "unknown memspace");
  Ctx& ctx = h->ctx();
  AXB_TRY(ctx.bind());
  const cudaMemcpyKind kind = memspace == AXB_MEM_HOST ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice;
  h->npts = npts;
  h->tree_built = false;
  AXB_TRY(h->pts.reserve(sizeof(double) * h->ndims * (size_t)std::max(npts, 1), ctx.stream));
  AXB_TRY(h->dom.reserve(sizeof(int32_t) * (size_t)std::max(npts, 1), ctx.stream));
  if(npts)
  {
    AXB_CUDA_TRY(cudaMemcpyAsync(h->pts.p, coords, sizeof(double) * h->ndims * (size_t)npts, kind, ctx.stream));
    AXB_CUDA_TRY(cudaMemcpyAsync(h->dom.p, domain_ids, sizeof(int32_t) * (size_t)npts, kind, ctx.stream));
  }
  return ctx.sync();
}

// generateBVHTreeImpl (:883-903)
int axb_dcp_generate_bvh_tree(axb_dcp* h)
{
  if(!h) return fail(AXB_ERR_BAD_ARG, "null handle");
  Ctx& ctx = h->ctx();
  AXB_TRY(ctx.bind());
  h->tree_built = true;
  if(h->npts == 0) return AXB_OK;  // no object points on this rank: queries only get initialised
  const int n = h->npts, D = h->ndims;
  AXB_TRY(h->boxes.reserve(sizeof(double) * 2 * D * (size_t)n, ctx.stream));
  if(D == 3)
    AXB_LAUNCH(ctx, dcp_point_boxes_kernel<3>, blocks_for(n, 256), 256, h->pts.as<double>(), n, h->boxes.as<Box<double, 3>>());
  else
    AXB_LAUNCH(ctx, dcp_point_boxes_kernel<2>, blocks_for(n, 256), 256, h->pts.as<double>(), n, h->boxes.as<Box<double, 2>>());
  axb_array_desc bd;
  memset(&bd, 0, sizeof(bd));
  for(int c = 0; c < 2 * D; ++c) bd.comp[c] = h->boxes.as<char>() + 8 * c;
  bd.stride_bytes = 16 * D;
  bd.ncomp = 2 * D;
  bd.memspace = AXB_MEM_DEVICE;
  AXB_TRY(axb_bvh_initialize(h->bvh, &bd, n));
  h->boxes.release(ctx.stream);
  return AXB_OK;
}

int axb_dcp_set_squared_distance_threshold(axb_dcp* h, double sq_threshold)
{
  if(!h) return fail(AXB_ERR_BAD_ARG, "null handle");
  if(sq_threshold < 0.0) return fail(AXB_ERR_BAD_ARG, "Squared distance-threshold must be non-negative.");
  h->sq_thresh = sq_threshold;
  return AXB_OK;
}

int axb_dcp_get_bvh(axb_dcp* h, axb_bvh** bvh)
{
  if(!h || !bvh) return fail(AXB_ERR_BAD_ARG, "null argument");
  *bvh = h->bvh;
  return AXB_OK;
}

// computeLocalClosestPoints (:905-1079) for one block of query points.  The five state arrays are the xferDom fields
// cp_index / cp_domain_index / cp_rank / cp_coords (interleaved) / debug/cp_distance (may be NULL); is_first != 0
// initialises them (-1, signalling NaN) before the search, otherwise they carry what earlier ranks of the ring found.
static int dcp_compute(axb_dcp* h, int rank, const double* query_coords, int32_t nq, int is_first, const double* bound_sq, int32_t* cp_index,
                       int32_t* cp_domain_index, int32_t* cp_rank, double* cp_coords, double* cp_distance, int memspace);

int axb_dcp_compute_local_closest_points(axb_dcp* h, int rank, const double* query_coords, int32_t nq, int is_first, int32_t* cp_index,
                                         int32_t* cp_domain_index, int32_t* cp_rank, double* cp_coords, double* cp_distance, int memspace)
{
  return dcp_compute(h, rank, query_coords, nq, is_first, nullptr, cp_index, cp_domain_index, cp_rank, cp_coords, cp_distance, memspace);
}

// The first-visit search (is_first) restricted to points within a per-query squared-distance bound that some other rank
// has already achieved: the state arrays are initialised, and an entry is filled only if this rank holds a point with
// squared distance <= bound_sq[i] (ties included: who wins a tie is decided by ring order afterwards).  Device memory
// only; used by the collective replacement of the ring (axom_b200/distributed_closest_point.py).
int axb_dcp_compute_bounded_closest_points(axb_dcp* h, int rank, const double* query_coords, int32_t nq, const double* bound_sq,
                                           int32_t* cp_index, int32_t* cp_domain_index, int32_t* cp_rank, double* cp_coords,
                                           double* cp_distance)
{
  if(nq > 0 && !bound_sq) return fail(AXB_ERR_BAD_ARG, "null bound array");
  if(h && h->mode != 1) return fail(AXB_ERR_UNSUPPORTED, "bounded search needs the nearest-first mode (axb_dcp_set_mode 1)");
  return dcp_compute(h, rank, query_coords, nq, 1, bound_sq, cp_index, cp_domain_index, cp_rank, cp_coords, cp_distance, AXB_MEM_DEVICE);
}

static int dcp_compute(axb_dcp* h, int rank, const double* query_coords, int32_t nq, int is_first, const double* bound_sq, int32_t* cp_index,
                       int32_t* cp_domain_index, int32_t* cp_rank, double* cp_coords, double* cp_distance, int memspace)
{
  if(!h) return fail(AXB_ERR_BAD_ARG, "null handle");
  if(nq < 0) return fail(AXB_ERR_BAD_ARG, "negative query count");
  if(nq > 0 && (!query_coords || !cp_index || !cp_domain_index || !cp_rank || !cp_coords)) return fail(AXB_ERR_BAD_ARG, "null query / state array");
  if(!h->tree_built) return fail(AXB_ERR_NOT_BUILT, "BVH tree must be initialized before calling 'computeClosestPoints");
  memspace = resolve_memspace(memspace, query_coords);
  if(memspace != AXB_MEM_HOST && memspace != AXB_MEM_DEVICE) return fail(AXB_ERR_BAD_ARG, "unknown memspace");
  if(nq == 0) return AXB_OK;
  Ctx& ctx = h->ctx();
  AXB_TRY(ctx.bind());
  ctx.begin_call();
  const int D = h->ndims;
  const size_t cb = sizeof(double) * D * (size_t)nq, ib = sizeof(int32_t) * (size_t)nq, db = sizeof(double) * (size_t)nq;
  const double* d_q = query_coords;
  int32_t *d_idx = cp_index, *d_dom = cp_domain_index, *d_rank = cp_rank;
  double *d_coords = cp_coords, *d_dist = cp_distance;
  if(memspace == AXB_MEM_HOST)
  {
    AXB_TRY(h->q_stage.reserve(cb, ctx.stream));
    AXB_TRY(h->st_idx.reserve(ib, ctx.stream));
    AXB_TRY(h->st_dom.reserve(ib, ctx.stream));
    AXB_TRY(h->st_rank.reserve(ib, ctx.stream));
    AXB_TRY(h->st_coords.reserve(cb, ctx.stream));
    AXB_CUDA_TRY(cudaMemcpyAsync(h->q_stage.p, query_coords, cb, cudaMemcpyHostToDevice, ctx.stream));
    d_q = h->q_stage.as<double>();
    d_idx = h->st_idx.as<int32_t>();
    d_dom = h->st_dom.as<int32_t>();
    d_rank = h->st_rank.as<int32_t>();
    d_coords = h->st_coords.as<double>();
    if(cp_distance)
    {
      AXB_TRY(h->st_dist.reserve(db, ctx.stream));
      d_dist = h->st_dist.as<double>();
    }
    if(!is_first)
    {
      AXB_CUDA_TRY(cudaMemcpyAsync(d_idx, cp_index, ib, cudaMemcpyHostToDevice, ctx.stream));
      AXB_CUDA_TRY(cudaMemcpyAsync(d_dom, cp_domain_index, ib, cudaMemcpyHostToDevice, ctx.stream));
      AXB_CUDA_TRY(cudaMemcpyAsync(d_rank, cp_rank, ib, cudaMemcpyHostToDevice, ctx.stream));
      AXB_CUDA_TRY(cudaMemcpyAsync(d_coords, cp_coords, cb, cudaMemcpyHostToDevice, ctx.stream));
      if(cp_distance) AXB_CUDA_TRY(cudaMemcpyAsync(d_dist, cp_distance, db, cudaMemcpyHostToDevice, ctx.stream));
    }
  }
  const bool has = h->npts > 0;
  const int32_t* perm = nullptr;
  if(h->mode == 1 && has && nq >= 4096)
  {
    // Morton order of the queries (processing order only): neighbouring threads walk the same nodes
    ScopedPhase ph(ctx, "dcp.sortq");
    const size_t kb = sizeof(unsigned long long) * (size_t)nq;
    AXB_TRY(h->keys_a.reserve(kb, ctx.stream));
    AXB_TRY(h->keys_b.reserve(kb, ctx.stream));
    AXB_TRY(h->perm.reserve(sizeof(int32_t) * (size_t)nq, ctx.stream));
    const size_t scratch = rsort::scratch_bytes(nq);
    AXB_TRY(h->scratch.reserve(scratch, ctx.stream));
    AXB_CUDA_TRY(cudaMemsetAsync(h->scratch.p, 0, scratch, ctx.stream));
    uint32_t* ghist = h->scratch.as<uint32_t>();
    uint32_t* tile_counters = ghist + rsort::MAX_PASSES * rsort::RADIX;
    uint32_t* lookback = tile_counters + 64;
    if(D == 3)
      AXB_LAUNCH(ctx, (dcp_query_keys_kernel<3, BuildState<double, 3>>), capped_grid(nq, 256), 256, d_q, nq,
                 h->bvh->state.as<BuildState<double, 3>>(), h->keys_a.as<unsigned long long>(), ghist);
    else
      AXB_LAUNCH(ctx, (dcp_query_keys_kernel<2, BuildState<double, 2>>), capped_grid(nq, 256), 256, d_q, nq,
                 h->bvh->state.as<BuildState<double, 2>>(), h->keys_a.as<unsigned long long>(), ghist);
    unsigned long long* sorted = nullptr;
    AXB_TRY(sort_keys_generic(ctx, h->keys_a.as<unsigned long long>(), h->keys_b.as<unsigned long long>(), nq, ghist, tile_counters, lookback,
                              &sorted, 1));
    AXB_LAUNCH(ctx, keys_to_perm_kernel, blocks_for(nq, 256), 256, sorted, nq, h->perm.as<int32_t>());
    perm = h->perm.as<int32_t>();
  }
  {
    ScopedPhase ph(ctx, "dcp.kernel");
#define AXB_DCP_LAUNCH(KERNEL, DD)                                                                                                 \
  AXB_LAUNCH(ctx, KERNEL<DD>, blocks_for(nq, 128), 128, has ? h->bvh->nodes.as<Node<double, DD>>() : nullptr,                      \
             h->bvh->leaf_nodes.as<int32_t>(), h->pts.as<double>(), h->dom.as<int32_t>(), rank, h->sq_thresh, d_q, nq, perm, is_first, \
             d_idx, d_dom, d_rank, d_coords, d_dist)
    if(h->mode == 1 && bound_sq)
    {
#define AXB_DCP_LAUNCH_B(DD)                                                                                                       \
  AXB_LAUNCH(ctx, dcp_nearest_kernel<DD>, blocks_for(nq, 128), 128, has ? h->bvh->nodes.as<Node<double, DD>>() : nullptr,          \
             h->bvh->leaf_nodes.as<int32_t>(), h->pts.as<double>(), h->dom.as<int32_t>(), rank, h->sq_thresh, d_q, nq, perm, is_first, \
             d_idx, d_dom, d_rank, d_coords, d_dist, bound_sq)
      if(D == 3)
        AXB_DCP_LAUNCH_B(3);
      else
        AXB_DCP_LAUNCH_B(2);
#undef AXB_DCP_LAUNCH_B
    }
    else if(h->mode == 1)
    {
      if(D == 3)
        AXB_DCP_LAUNCH(dcp_nearest_kernel, 3);
      else
        AXB_DCP_LAUNCH(dcp_nearest_kernel, 2);
    }
    else
    {
      if(D == 3)
        AXB_DCP_LAUNCH(dcp_local_kernel, 3);
      else
        AXB_DCP_LAUNCH(dcp_local_kernel, 2);
    }
#undef AXB_DCP_LAUNCH
  }
  if(memspace == AXB_MEM_HOST)
  {
    AXB_CUDA_TRY(cudaMemcpyAsync(cp_index, d_idx, ib, cudaMemcpyDeviceToHost, ctx.stream));
    AXB_CUDA_TRY(cudaMemcpyAsync(cp_domain_index, d_dom, ib, cudaMemcpyDeviceToHost, ctx.stream));
    AXB_CUDA_TRY(cudaMemcpyAsync(cp_rank, d_rank, ib, cudaMemcpyDeviceToHost, ctx.stream));
    AXB_CUDA_TRY(cudaMemcpyAsync(cp_coords, d_coords, cb, cudaMemcpyDeviceToHost, ctx.stream));
    if(cp_distance) AXB_CUDA_TRY(cudaMemcpyAsync(cp_distance, d_dist, db, cudaMemcpyDeviceToHost, ctx.stream));
  }
  return ctx.finish_call();
}

}  // extern "C"

//==========================================================================================
// The exchange steps of the distributed cases (comm.cuh): NCCL communicator, partitioned-surface MIN,
// DistributedClosestPoint::computeClosestPoints (quest/detail/DistributedClosestPointImpl.hpp:687-693, :737-851)
//==========================================================================================
struct DcpxScratch
{
  DevBuf counts, offs, boxes, Q, mine, nearest, slot, list, listn, q_pack, b_pack, st_idx, st_dom, st_rank, st_coords, st_dist, sq, smin, bound,
    pos, win, cnt, send_off, send, recv, o_idx, o_dom, o_rank, o_coords, o_dist, q_own;
  void release(cudaStream_t st)
  {
    for(DevBuf* b : {&counts, &offs, &boxes, &Q, &mine, &nearest, &slot, &list, &listn, &q_pack, &b_pack, &st_idx, &st_dom, &st_rank, &st_coords,
                     &st_dist, &sq, &smin, &bound, &pos, &win, &cnt, &send_off, &send, &recv, &o_idx, &o_dom, &o_rank, &o_coords, &o_dist, &q_own})
      b->release(st);
  }
};
static std::map<axb_dcp*, DcpxScratch> g_dcpx;  // per-handle scratch of the collective path (released by axb_dcp_destroy)
static std::mutex g_dcpx_mutex;

static void dcpx_release(axb_dcp* h, cudaStream_t st)
{
  std::lock_guard<std::mutex> lock(g_dcpx_mutex);
  auto it = g_dcpx.find(h);
  if(it == g_dcpx.end()) return;
  it->second.release(st);
  g_dcpx.erase(it);
}

extern "C" {

int axb_comm_get_unique_id(uint8_t* id_bytes)
{
  if(!id_bytes) return fail(AXB_ERR_BAD_ARG, "null id buffer");
  NcclApi* api = nullptr;
  AXB_TRY(nccl_api(&api));
  ncclUniqueId id;
  static_assert(sizeof(ncclUniqueId) == AXB_COMM_ID_BYTES, "ncclUniqueId is 128 bytes");
  AXB_NCCL_TRY(api, api->GetUniqueId(&id));
  memcpy(id_bytes, &id, sizeof(id));
  return AXB_OK;
}

int axb_comm_create(axb_comm** out, int nranks, int rank, const uint8_t* id_bytes, int device)
{
  if(!out) return fail(AXB_ERR_BAD_ARG, "null output handle");
  *out = nullptr;
  if(nranks < 1 || rank < 0 || rank >= nranks || !id_bytes) return fail(AXB_ERR_BAD_ARG, "bad communicator arguments");
  int count = 0;
  if(cudaGetDeviceCount(&count) != cudaSuccess || count <= 0)
  {
    cudaGetLastError();
    return fail(AXB_ERR_NO_DEVICE, "no CUDA device available (libaxb200 has no CPU fallback)");
  }
  if(device < 0 || device >= count) return fail(AXB_ERR_BAD_ARG, "device ordinal out of range");
  NcclApi* api = nullptr;
  AXB_TRY(nccl_api(&api));
  AXB_CUDA_TRY(cudaSetDevice(device));
  ncclUniqueId id;
  memcpy(&id, id_bytes, sizeof(id));
  ncclComm_t c = nullptr;
  AXB_NCCL_TRY(api, api->CommInitRank(&c, nranks, id, rank));
  axb_comm* h = new axb_comm();
  h->api = api;
  h->comm = c;
  h->nranks = nranks;
  h->rank = rank;
  h->device = device;
  *out = h;
  return AXB_OK;
}

int axb_comm_destroy(axb_comm* c)
{
  if(!c) return AXB_OK;
  if(c->comm)
  {
    cudaSetDevice(c->device);
    c->api->CommDestroy(c->comm);
  }
  delete c;
  return AXB_OK;
}

int axb_comm_get_rank(const axb_comm* c, int* rank, int* nranks)
{
  if(!c) return fail(AXB_ERR_BAD_ARG, "null communicator");
  if(rank) *rank = c->rank;
  if(nranks) *nranks = c->nranks;
  return AXB_OK;
}

int axb_comm_get_traffic(const axb_comm* c, int64_t* bytes, int64_t* collectives)
{
  if(!c) return fail(AXB_ERR_BAD_ARG, "null communicator");
  if(bytes) *bytes = c->bytes;
  if(collectives) *collectives = c->calls;
  return AXB_OK;
}

const char* axb_comm_library(void)
{
  NcclApi* api = nullptr;
  if(nccl_api(&api) != AXB_OK) return "";
  static thread_local std::string s;
  int v = 0;
  api->GetVersion(&v);
  s = api->where + " version " + std::to_string(v);
  return s.c_str();
}

// elementwise MIN / MAX / SUM of a device array of doubles over the ranks, in place, on `cuda_stream`
int axb_comm_allreduce_f64(axb_comm* c, double* device_buf, int64_t n, int op, void* cuda_stream)
{
  if(!c) return fail(AXB_ERR_BAD_ARG, "null communicator");
  if(n < 0 || (n > 0 && !device_buf)) return fail(AXB_ERR_BAD_ARG, "null or negative buffer");
  if(op < 0 || op > 2) return fail(AXB_ERR_BAD_ARG, "op must be 0 (min), 1 (max) or 2 (sum)");
  AXB_CUDA_TRY(cudaSetDevice(c->device));
  const ncclRedOp_t ops[3] = {ncclMin, ncclMax, ncclSum};
  AXB_NCCL_TRY(c->api, c->api->AllReduce(device_buf, device_buf, (size_t)n, ncclDouble, ops[op], c->comm, (cudaStream_t)cuda_stream));
  c->bytes += 8 * n;
  c->calls += 1;
  return AXB_OK;
}

// C5: every rank holds ONE PART of the surface in `s` (built with compute_sign = 0) and the SAME query points; the result on
// every rank is the distance to the whole surface = the elementwise MIN of the ranks' partial distances.  The query kernel
// writes into the buffer the reduction then runs on, in place, on the same stream: no copy, no host round trip between.
int axb_sd_compute_distances_minreduce(axb_sd* s, axb_comm* c, const axb_array_desc* qpts, int32_t npts, double* dist, int out_memspace)
{
  if(!s || !c) return fail(AXB_ERR_BAD_ARG, "null handle");
  if(npts < 0) return fail(AXB_ERR_BAD_ARG, "negative point count");
  if(npts > 0 && (!dist || !qpts)) return fail(AXB_ERR_BAD_ARG, "null query descriptor or output");
  if(s->prm.compute_sign) return fail(AXB_ERR_BAD_ARG, "the MIN over surface parts is defined for unsigned distances: create the handle with compute_sign = 0");
  if(c->device != s->ctx().device) return fail(AXB_ERR_BAD_ARG, "communicator and SignedDistance handle live on different devices");
  out_memspace = resolve_memspace(out_memspace, dist);
  if(out_memspace != AXB_MEM_HOST && out_memspace != AXB_MEM_DEVICE) return fail(AXB_ERR_BAD_ARG, "unknown output memspace");
  if(!s->bvh->built) return fail(AXB_ERR_NOT_BUILT, "SignedDistance query before setMesh()");
  Ctx& ctx = s->ctx();
  AXB_TRY(ctx.bind());
  ctx.begin_call();
  if(npts == 0) return AXB_OK;  // (every rank passes the same count, so every rank returns here)
  const int tot = ctx.phase_begin("query.total");
  axb_sd::QBufs& B = s->qb[0];
  double* d_out = dist;
  if(out_memspace == AXB_MEM_HOST)
  {
    AXB_TRY(B.out_phi.reserve(sizeof(double) * (size_t)npts, ctx.stream));
    d_out = B.out_phi.as<double>();
  }
  // the slack of the exchanged bound: the largest triangle diameter over ALL parts (8 bytes, MAX)
  SdBoundExchange ex {c, s->max_diam};
  const bool bounded = c->nranks > 1 && getenv("AXB_SD_NO_BOUND_EXCHANGE") == nullptr;
  if(bounded)
  {
    AXB_TRY(B.ext.reserve(sizeof(double) * (size_t)std::max(npts, 1), ctx.stream));
    AXB_CUDA_TRY(cudaMemcpyAsync(B.ext.p, &s->max_diam, sizeof(double), cudaMemcpyHostToDevice, ctx.stream));
    AXB_NCCL_TRY(c->api, c->api->AllReduce(B.ext.p, B.ext.p, 1, ncclDouble, ncclMax, c->comm, ctx.stream));
    AXB_CUDA_TRY(cudaMemcpyAsync(&ex.slack, B.ext.p, sizeof(double), cudaMemcpyDeviceToHost, ctx.stream));
    AXB_TRY(ctx.sync());
    c->bytes += 8;
    c->calls += 1;
  }
  AXB_TRY(sd_query_range(s, B, qpts, npts, d_out, nullptr, nullptr, AXB_MEM_DEVICE, nullptr, bounded ? &ex : nullptr));
  {
    ScopedPhase ph(ctx, "query.minreduce");
    AXB_NCCL_TRY(c->api, c->api->AllReduce(d_out, d_out, (size_t)npts, ncclDouble, ncclMin, c->comm, ctx.stream));
    c->bytes += 8ll * npts;
    c->calls += 1;
  }
  if(out_memspace == AXB_MEM_HOST) AXB_CUDA_TRY(cudaMemcpyAsync(dist, d_out, sizeof(double) * (size_t)npts, cudaMemcpyDeviceToHost, ctx.stream));
  ctx.phase_end(tot);
  if(out_memspace == AXB_MEM_HOST) AXB_TRY(ctx.sync());
  return ctx.finish_call();
}

}  // extern "C"

__global__ void __launch_bounds__(256) min_update_kernel(double* __restrict__ inout, const double* __restrict__ v, int n)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if(i < n) inout[i] = fmin(inout[i], v[i]);
}

extern "C" {

// The same partitioned-surface query on ONE GPU: the parts are evaluated in turn and `dist` is in/out -- on entry the
// distance to the parts evaluated so far (DBL_MAX: none), on exit min(entry, distance to this handle's part).  The entry
// value bounds the search (a part farther than it cannot change the minimum: values computed by the same arithmetic are
// compared, so no slack is needed), which prunes far parts at the root.
int axb_sd_update_min_distances(axb_sd* s, const axb_array_desc* qpts, int32_t npts, double* dist, int memspace)
{
  if(!s) return fail(AXB_ERR_BAD_ARG, "null handle");
  if(npts < 0) return fail(AXB_ERR_BAD_ARG, "negative point count");
  if(npts > 0 && (!dist || !qpts)) return fail(AXB_ERR_BAD_ARG, "null query descriptor or distance array");
  if(s->prm.compute_sign) return fail(AXB_ERR_BAD_ARG, "a running minimum over surface parts is defined for unsigned distances: create the handle with compute_sign = 0");
  memspace = resolve_memspace(memspace, dist);
  if(memspace != AXB_MEM_HOST && memspace != AXB_MEM_DEVICE) return fail(AXB_ERR_BAD_ARG, "unknown memspace");
  if(!s->bvh->built) return fail(AXB_ERR_NOT_BUILT, "SignedDistance query before setMesh()");
  Ctx& ctx = s->ctx();
  AXB_TRY(ctx.bind());
  ctx.begin_call();
  if(npts == 0) return AXB_OK;
  const int tot = ctx.phase_begin("query.total");
  axb_sd::QBufs& B = s->qb[0];
  AXB_TRY(B.out_phi.reserve(sizeof(double) * (size_t)npts, ctx.stream));
  double* d_io = dist;
  if(memspace == AXB_MEM_HOST)
  {
    AXB_TRY(B.ext.reserve(sizeof(double) * (size_t)npts, ctx.stream));
    AXB_CUDA_TRY(cudaMemcpyAsync(B.ext.p, dist, sizeof(double) * (size_t)npts, cudaMemcpyHostToDevice, ctx.stream));
    d_io = B.ext.as<double>();
  }
  SdBoundExchange ex {nullptr, 0.0, d_io};
  AXB_TRY(sd_query_range(s, B, qpts, npts, B.out_phi.as<double>(), nullptr, nullptr, AXB_MEM_DEVICE, nullptr, &ex));
  AXB_LAUNCH(ctx, min_update_kernel, blocks_for(npts, 256), 256, d_io, B.out_phi.as<double>(), npts);
  if(memspace == AXB_MEM_HOST) AXB_CUDA_TRY(cudaMemcpyAsync(dist, d_io, sizeof(double) * (size_t)npts, cudaMemcpyDeviceToHost, ctx.stream));
  ctx.phase_end(tot);
  if(memspace == AXB_MEM_HOST) AXB_TRY(ctx.sync());
  return ctx.finish_call();
}

}  // extern "C"

// computeClosestPoints (:737-851) as collectives; see comm.cuh for the protocol
template <int D>
static int dcpx_compute(axb_dcp* h, axb_comm* c, DcpxScratch& S, const double* d_q, int32_t nq, int32_t* o_idx, int32_t* o_dom, int32_t* o_rank,
                        double* o_coords, double* o_dist)
{
  Ctx& ctx = h->ctx();
  NcclApi* api = c->api;
  const int N = c->nranks, rank = c->rank;
  cudaStream_t st = ctx.stream;
  const bool was_async = ctx.async;
  // ---- 1: counts, query blocks, object boxes ----
  std::vector<long long> counts(N), offs(N + 1, 0);
  {
    ScopedPhase ph(ctx, "dcpx.counts");
    AXB_TRY(S.counts.reserve(sizeof(long long) * (size_t)(N + 1), st));
    const long long mine = nq;
    AXB_CUDA_TRY(cudaMemcpyAsync(S.counts.as<long long>() + N, &mine, sizeof(long long), cudaMemcpyHostToDevice, st));
    AXB_NCCL_TRY(api, api->AllGather(S.counts.as<long long>() + N, S.counts.p, 1, ncclInt64, c->comm, st));
    AXB_CUDA_TRY(cudaMemcpyAsync(counts.data(), S.counts.p, sizeof(long long) * (size_t)N, cudaMemcpyDeviceToHost, st));
    AXB_TRY(ctx.sync());
    c->bytes += 8;
    c->calls += 1;
  }
  for(int r = 0; r < N; ++r) offs[r + 1] = offs[r] + counts[r];
  const long long ntot = offs[N];
  if(counts[rank] != nq) return fail(AXB_ERR_CUDA, "the gathered query count of this rank differs from the one passed in");
  if(ntot > 2147483647LL) return fail(AXB_ERR_OVERFLOW, "more than 2^31-1 query points in one computeClosestPoints: split the call");
  if(ntot == 0) return AXB_OK;
  const size_t nt = (size_t)ntot;
  {
    ScopedPhase ph(ctx, "dcpx.gather");
    AXB_TRY(S.offs.reserve(sizeof(long long) * (size_t)(N + 1), st));
    AXB_CUDA_TRY(cudaMemcpyAsync(S.offs.p, offs.data(), sizeof(long long) * (size_t)(N + 1), cudaMemcpyHostToDevice, st));
    AXB_TRY(S.Q.reserve(sizeof(double) * D * nt, st));
    AXB_NCCL_TRY(api, api->GroupStart());
    for(int r = 0; r < N; ++r)
      if(counts[r] > 0)
        AXB_NCCL_TRY(api, api->Broadcast(r == rank ? (const void*)d_q : (const void*)(S.Q.as<double>() + (size_t)offs[r] * D),
                                         S.Q.as<double>() + (size_t)offs[r] * D, (size_t)counts[r] * D, ncclDouble, r, c->comm, st));
    AXB_NCCL_TRY(api, api->GroupEnd());
    c->bytes += 8ll * D * nq;
    c->calls += 1;
    // gatherBVHRoots (:671-693): every rank's object bounding box (invalid = no object points)
    double hb[6] = {DBL_MAX, DBL_MAX, DBL_MAX, -DBL_MAX, -DBL_MAX, -DBL_MAX};
    if(h->npts > 0)
    {
      double lo[3] = {0, 0, 0}, hi[3] = {0, 0, 0};
      AXB_TRY(axb_bvh_get_bounds(h->bvh, lo, hi));
      for(int d = 0; d < 3; ++d)
      {
        hb[d] = d < D ? lo[d] : 0.0;
        hb[3 + d] = d < D ? hi[d] : 0.0;
      }
    }
    AXB_TRY(S.boxes.reserve(sizeof(double) * 6 * (size_t)(N + 1), st));
    AXB_CUDA_TRY(cudaMemcpyAsync(S.boxes.as<double>() + 6 * (size_t)N, hb, sizeof(hb), cudaMemcpyHostToDevice, st));
    AXB_NCCL_TRY(api, api->AllGather(S.boxes.as<double>() + 6 * (size_t)N, S.boxes.p, 6, ncclDouble, c->comm, st));
    AXB_TRY(ctx.sync());  // hb and offs leave scope / are reused
    c->bytes += 48;
    c->calls += 1;
  }
  DcpxRanks R;
  R.offs = S.offs.as<long long>();
  R.boxes = S.boxes.as<double>();
  R.nranks = N;
  R.rank = rank;
  // ---- 2: first search (the rank whose box centre is nearest), bound = MIN over ranks ----
  AXB_TRY(S.mine.reserve(sizeof(double) * nt, st));
  AXB_TRY(S.nearest.reserve(sizeof(int32_t) * nt, st));
  AXB_TRY(S.slot.reserve(sizeof(int32_t) * nt, st));
  AXB_TRY(S.list.reserve(sizeof(int32_t) * nt, st));
  AXB_TRY(S.listn.reserve(sizeof(unsigned int) * 2, st));
  AXB_TRY(S.q_pack.reserve(sizeof(double) * D * nt, st));
  AXB_TRY(S.b_pack.reserve(sizeof(double) * nt, st));
  AXB_TRY(S.st_idx.reserve(sizeof(int32_t) * nt, st));
  AXB_TRY(S.st_dom.reserve(sizeof(int32_t) * nt, st));
  AXB_TRY(S.st_rank.reserve(sizeof(int32_t) * nt, st));
  AXB_TRY(S.st_coords.reserve(sizeof(double) * D * nt, st));
  AXB_TRY(S.st_dist.reserve(sizeof(double) * nt, st));
  AXB_TRY(S.bound.reserve(sizeof(double) * nt, st));
  AXB_TRY(S.sq.reserve(sizeof(double) * nt, st));
  AXB_TRY(S.smin.reserve(sizeof(double) * nt, st));
  AXB_TRY(S.pos.reserve(nt, st));
  AXB_TRY(S.win.reserve(nt, st));
  unsigned int hn[2] = {0u, 0u};
  int n1 = 0, n2 = 0;
  {
    ScopedPhase ph(ctx, "dcpx.search1");
    AXB_CUDA_TRY(cudaMemsetAsync(S.listn.p, 0, sizeof(unsigned int) * 2, st));
    AXB_LAUNCH(ctx, dcpx_classify_kernel<D>, blocks_for(ntot, 256), 256, S.Q.as<double>(), ntot, R, h->sq_thresh, S.mine.as<double>(),
               S.nearest.as<int32_t>(), S.slot.as<int32_t>(), S.list.as<int32_t>(), S.listn.as<unsigned int>());
    AXB_CUDA_TRY(cudaMemcpyAsync(hn, S.listn.p, sizeof(unsigned int), cudaMemcpyDeviceToHost, st));
    AXB_TRY(ctx.sync());
    n1 = (int)hn[0];
    if(n1 > 0)
    {
      AXB_LAUNCH(ctx, dcpx_gather_kernel<D>, blocks_for(n1, 256), 256, S.Q.as<double>(), S.list.as<int32_t>(), n1, 0, (const double*)nullptr,
                 S.q_pack.as<double>(), (double*)nullptr, S.slot.as<int32_t>());
      ctx.async = true;  // stay on the stream: the collectives below are ordered behind the search
      const int rc = axb_dcp_compute_local_closest_points(h, rank, S.q_pack.as<double>(), n1, 1, S.st_idx.as<int32_t>(), S.st_dom.as<int32_t>(),
                                                          S.st_rank.as<int32_t>(), S.st_coords.as<double>(), S.st_dist.as<double>(), AXB_MEM_DEVICE);
      ctx.async = was_async;
      AXB_TRY(rc);
    }
    AXB_LAUNCH(ctx, dcpx_sq_kernel<D>, blocks_for(ntot, 256), 256, S.Q.as<double>(), ntot, S.slot.as<int32_t>(), S.st_rank.as<int32_t>(),
               S.st_coords.as<double>(), S.bound.as<double>(), (double*)nullptr);
  }
  {
    ScopedPhase ph(ctx, "dcpx.bound_allreduce");
    AXB_NCCL_TRY(api, api->AllReduce(S.bound.p, S.bound.p, nt, ncclDouble, ncclMin, c->comm, st));
    c->bytes += 8ll * ntot;
    c->calls += 1;
  }
  // ---- 3: bounded search on the other ranks (the ring prunes with the same two tests, per block: :762-775, :859-878, :1037-1040) ----
  {
    ScopedPhase ph(ctx, "dcpx.search2");
    AXB_LAUNCH(ctx, dcpx_classify2_kernel, blocks_for(ntot, 256), 256, ntot, rank, h->sq_thresh, S.mine.as<double>(), S.nearest.as<int32_t>(),
               S.bound.as<double>(), S.list.as<int32_t>(), S.listn.as<unsigned int>() + 1);
    AXB_CUDA_TRY(cudaMemcpyAsync(hn + 1, S.listn.as<unsigned int>() + 1, sizeof(unsigned int), cudaMemcpyDeviceToHost, st));
    AXB_TRY(ctx.sync());
    n2 = (int)hn[1];
    if(n2 > 0)
    {
      AXB_LAUNCH(ctx, dcpx_gather_kernel<D>, blocks_for(n2, 256), 256, S.Q.as<double>(), S.list.as<int32_t>(), n2, n1, S.bound.as<double>(),
                 S.q_pack.as<double>(), S.b_pack.as<double>(), S.slot.as<int32_t>());
      ctx.async = true;
      const int rc = axb_dcp_compute_bounded_closest_points(h, rank, S.q_pack.as<double>(), n2, S.b_pack.as<double>(), S.st_idx.as<int32_t>() + n1,
                                                            S.st_dom.as<int32_t>() + n1, S.st_rank.as<int32_t>() + n1,
                                                            S.st_coords.as<double>() + (size_t)n1 * D, S.st_dist.as<double>() + n1);
      ctx.async = was_async;
      AXB_TRY(rc);
    }
  }
  // ---- 4: the winner of every query, exactly as the ring would pick it ----
  std::vector<unsigned int> hc(2 * (size_t)N, 0u);
  {
    ScopedPhase ph(ctx, "dcpx.combine");
    AXB_LAUNCH(ctx, dcpx_sq_kernel<D>, blocks_for(ntot, 256), 256, S.Q.as<double>(), ntot, S.slot.as<int32_t>(), S.st_rank.as<int32_t>(),
               S.st_coords.as<double>(), S.sq.as<double>(), S.smin.as<double>());
    AXB_NCCL_TRY(api, api->AllReduce(S.smin.p, S.smin.p, nt, ncclDouble, ncclMin, c->comm, st));
    AXB_LAUNCH(ctx, dcpx_pos_kernel, blocks_for(ntot, 256), 256, ntot, R, S.sq.as<double>(), S.smin.as<double>(), S.pos.as<uint8_t>(),
               S.win.as<uint8_t>());
    AXB_NCCL_TRY(api, api->AllReduce(S.win.p, S.win.p, nt, ncclUint8, ncclMin, c->comm, st));
    c->bytes += 9ll * ntot;
    c->calls += 2;
    AXB_TRY(S.cnt.reserve(sizeof(unsigned int) * 3 * (size_t)N, st));
    AXB_CUDA_TRY(cudaMemsetAsync(S.cnt.p, 0, sizeof(unsigned int) * 3 * (size_t)N, st));
    AXB_LAUNCH(ctx, dcpx_count_kernel, capped_grid(ntot, 256), 256, ntot, R, S.pos.as<uint8_t>(), S.win.as<uint8_t>(), S.cnt.as<unsigned int>(),
               S.cnt.as<unsigned int>() + N);
    AXB_CUDA_TRY(cudaMemcpyAsync(hc.data(), S.cnt.p, sizeof(unsigned int) * 2 * (size_t)N, cudaMemcpyDeviceToHost, st));
    AXB_TRY(ctx.sync());
  }
  // ---- 5: every winner sends its record to the query's home rank ----
  {
    ScopedPhase ph(ctx, "dcpx.exchange");
    std::vector<long long> soff(N + 1, 0), roff(N + 1, 0);
    for(int r = 0; r < N; ++r)
    {
      soff[r + 1] = soff[r] + hc[r];
      roff[r + 1] = roff[r] + hc[N + r];
    }
    if(roff[N] > nq) return fail(AXB_ERR_CUDA, "internal error: more winners than query points");
    AXB_TRY(S.send_off.reserve(sizeof(long long) * (size_t)(N + 1), st));
    AXB_CUDA_TRY(cudaMemcpyAsync(S.send_off.p, soff.data(), sizeof(long long) * (size_t)(N + 1), cudaMemcpyHostToDevice, st));
    AXB_TRY(S.send.reserve(sizeof(DcpxRecord) * (size_t)std::max<long long>(soff[N], 1), st));
    AXB_TRY(S.recv.reserve(sizeof(DcpxRecord) * (size_t)std::max<long long>(roff[N], 1), st));
    if(soff[N] > 0)
      AXB_LAUNCH(ctx, dcpx_pack_kernel<D>, blocks_for(ntot, 256), 256, ntot, R, S.pos.as<uint8_t>(), S.win.as<uint8_t>(), S.slot.as<int32_t>(),
                 S.st_idx.as<int32_t>(), S.st_dom.as<int32_t>(), S.st_rank.as<int32_t>(), S.st_coords.as<double>(), S.st_dist.as<double>(),
                 S.send_off.as<long long>(), S.cnt.as<unsigned int>() + 2 * N, S.send.as<DcpxRecord>());
    constexpr size_t W = sizeof(DcpxRecord) / sizeof(long long);
    AXB_NCCL_TRY(api, api->GroupStart());
    for(int r = 0; r < N; ++r)
    {
      if(r == rank) continue;
      if(hc[r] > 0) AXB_NCCL_TRY(api, api->Send(S.send.as<DcpxRecord>() + soff[r], (size_t)hc[r] * W, ncclInt64, r, c->comm, st));
      if(hc[N + r] > 0) AXB_NCCL_TRY(api, api->Recv(S.recv.as<DcpxRecord>() + roff[r], (size_t)hc[N + r] * W, ncclInt64, r, c->comm, st));
    }
    AXB_NCCL_TRY(api, api->GroupEnd());
    if(hc[rank] != hc[N + rank]) return fail(AXB_ERR_CUDA, "internal error: this rank's own send / receive counts differ");
    if(hc[rank] > 0)
      AXB_CUDA_TRY(cudaMemcpyAsync(S.recv.as<DcpxRecord>() + roff[rank], S.send.as<DcpxRecord>() + soff[rank], sizeof(DcpxRecord) * (size_t)hc[rank],
                                   cudaMemcpyDeviceToDevice, st));
    c->bytes += (long long)sizeof(DcpxRecord) * (soff[N] - hc[rank]);
    c->calls += 1;
    if(nq > 0)
    {
      AXB_LAUNCH(ctx, dcpx_init_outputs_kernel<D>, blocks_for(nq, 256), 256, nq, o_idx, o_dom, o_rank, o_coords, o_dist);
      if(roff[N] > 0)
        AXB_LAUNCH(ctx, dcpx_unpack_kernel<D>, blocks_for(roff[N], 256), 256, S.recv.as<DcpxRecord>(), roff[N], o_idx, o_dom, o_rank, o_coords, o_dist);
    }
    AXB_TRY(ctx.sync());  // soff / roff were read by asynchronous copies
  }
  return AXB_OK;
}

extern "C" {

int axb_dcp_compute_closest_points(axb_dcp* h, axb_comm* c, const double* query_coords, int32_t nq, int memspace, int32_t* cp_index,
                                   int32_t* cp_domain_index, int32_t* cp_rank, double* cp_coords, double* cp_distance)
{
  if(!h) return fail(AXB_ERR_BAD_ARG, "null handle");
  if(nq < 0) return fail(AXB_ERR_BAD_ARG, "negative query count");
  if(nq > 0 && !query_coords) return fail(AXB_ERR_BAD_ARG, "null query array");
  if(!h->tree_built) return fail(AXB_ERR_NOT_BUILT, "BVH tree must be initialized before calling 'computeClosestPoints");
  if(c && c->nranks > kDcpxMaxRanks) return fail(AXB_ERR_UNSUPPORTED, "more than 254 ranks");
  if(c && c->device != h->ctx().device) return fail(AXB_ERR_BAD_ARG, "communicator and DistributedClosestPoint handle live on different devices");
  if(c && h->mode != 1) return fail(AXB_ERR_UNSUPPORTED, "the collective path needs the nearest-first mode (axb_dcp_set_mode 1)");
  memspace = resolve_memspace(memspace, nq > 0 ? (const void*)query_coords : (const void*)cp_index);
  if(memspace != AXB_MEM_HOST && memspace != AXB_MEM_DEVICE) return fail(AXB_ERR_BAD_ARG, "unknown memspace");
  Ctx& ctx = h->ctx();
  AXB_TRY(ctx.bind());
  const int D = h->ndims;
  DcpxScratch* S = nullptr;
  {
    std::lock_guard<std::mutex> lock(g_dcpx_mutex);
    S = &g_dcpx[h];
  }
  cudaStream_t st = ctx.stream;
  const size_t n = (size_t)std::max(nq, 1);
  // outputs the caller did not ask for still exist on the device (the records carry all five fields)
  const bool host = memspace == AXB_MEM_HOST;
  int32_t *d_idx = cp_index, *d_dom = cp_domain_index, *d_rank = cp_rank;
  double *d_coords = cp_coords, *d_dist = cp_distance;
  if(host || !d_idx) { AXB_TRY(S->o_idx.reserve(sizeof(int32_t) * n, st)); d_idx = S->o_idx.as<int32_t>(); }
  if(host || !d_dom) { AXB_TRY(S->o_dom.reserve(sizeof(int32_t) * n, st)); d_dom = S->o_dom.as<int32_t>(); }
  if(host || !d_rank) { AXB_TRY(S->o_rank.reserve(sizeof(int32_t) * n, st)); d_rank = S->o_rank.as<int32_t>(); }
  if(host || !d_coords) { AXB_TRY(S->o_coords.reserve(sizeof(double) * D * n, st)); d_coords = S->o_coords.as<double>(); }
  if(host || !d_dist) { AXB_TRY(S->o_dist.reserve(sizeof(double) * n, st)); d_dist = S->o_dist.as<double>(); }
  const double* d_q = query_coords;
  if(host && nq > 0)
  {
    AXB_TRY(S->q_own.reserve(sizeof(double) * D * n, st));
    AXB_CUDA_TRY(cudaMemcpyAsync(S->q_own.p, query_coords, sizeof(double) * D * (size_t)nq, cudaMemcpyHostToDevice, st));
    d_q = S->q_own.as<double>();
  }
  const int tot = ctx.phase_begin("dcpx.total");
  if(!c)
  {
    // one rank, no exchange: the ring of one
    const bool was_async = ctx.async;
    ctx.async = true;
    const int rc = nq > 0 ? axb_dcp_compute_local_closest_points(h, 0, d_q, nq, 1, d_idx, d_dom, d_rank, d_coords, d_dist, AXB_MEM_DEVICE) : AXB_OK;
    ctx.async = was_async;
    AXB_TRY(rc);
  }
  else if(D == 3)
    AXB_TRY(dcpx_compute<3>(h, c, *S, d_q, nq, d_idx, d_dom, d_rank, d_coords, d_dist));
  else
    AXB_TRY(dcpx_compute<2>(h, c, *S, d_q, nq, d_idx, d_dom, d_rank, d_coords, d_dist));
  ctx.phase_end(tot);
  if(host && nq > 0)
  {
    if(cp_index) AXB_CUDA_TRY(cudaMemcpyAsync(cp_index, d_idx, sizeof(int32_t) * (size_t)nq, cudaMemcpyDeviceToHost, st));
    if(cp_domain_index) AXB_CUDA_TRY(cudaMemcpyAsync(cp_domain_index, d_dom, sizeof(int32_t) * (size_t)nq, cudaMemcpyDeviceToHost, st));
    if(cp_rank) AXB_CUDA_TRY(cudaMemcpyAsync(cp_rank, d_rank, sizeof(int32_t) * (size_t)nq, cudaMemcpyDeviceToHost, st));
    if(cp_coords) AXB_CUDA_TRY(cudaMemcpyAsync(cp_coords, d_coords, sizeof(double) * D * (size_t)nq, cudaMemcpyDeviceToHost, st));
    if(cp_distance) AXB_CUDA_TRY(cudaMemcpyAsync(cp_distance, d_dist, sizeof(double) * (size_t)nq, cudaMemcpyDeviceToHost, st));
    AXB_TRY(ctx.sync());
  }
  return ctx.finish_call();
}

}  // extern "C"

//==========================================================================================
// quest::MarchingCubes
//==========================================================================================

struct axb_mc
{
  Ctx ctx;
  int ndims = 3;
  int mask_val = 1;
  struct Domain
  {
    axb_mc_domain d;        // pointers resolved to device memory
    int slowest[3] = {0, 1, 2};
    long long case_stride[3] = {0, 0, 0};
    long long num_cells = 0;
    int num_tiles = 0;
    DevBuf staged[5];       // host inputs: coords x/y/z, fcn, mask
    DevBuf case_ids, tile_offsets, active_tiles, lut;
    uint8_t h_lut[256];     // compact corner bits -> case id for this domain's direction order (mark_rows_kernel)
  };
  std::vector<Domain> doms;
  long long* h_totals = nullptr;  // pinned + device-mapped: (facets, active tiles) per domain, written by the scan kernel
  size_t h_totals_cap = 0;
  long long facet_count = 0;
  DevBuf node_ids, node_coords, parent_ids, domain_ids;
  DevBuf ticket;            // the count kernel's block ticket (zero between launches)
  void release_domains()
  {
    for(auto& dm : doms)
    {
      for(auto& b : dm.staged) b.release(ctx.stream);
      dm.case_ids.release(ctx.stream);
      dm.tile_offsets.release(ctx.stream);
      dm.lut.release(ctx.stream);
      dm.active_tiles.release(ctx.stream);
    }
    doms.clear();
  }
};

namespace
{
// the span of elements a ghost-free strided view touches: 1 + sum (shape[d] - 1) * stride[d]
// (extra = 1: nodal arrays, cell_shape + 1 entries per direction; extra = 0: cell-centred arrays)
long long view_span(const int64_t* cell_shape, int extra, const int64_t* strides, int nd)
{
  long long span = 1;
  for(int d = 0; d < nd; ++d) span += (long long)(cell_shape[d] + extra - 1) * strides[d];
  return span;
}

template <int DIM>
mc::DomainView<DIM> make_view(const axb_mc::Domain& dm)
{
  mc::DomainView<DIM> v;
  for(int d = 0; d < DIM; ++d)
  {
    v.case_div[d] = mc::make_fastdiv((uint32_t)dm.case_stride[d]);
    v.slowest[d] = dm.slowest[d];
    v.fcn_stride[d] = dm.d.fcn_strides[d];
    v.coords[d] = dm.d.coords[d];
    v.coords_stride[d] = dm.d.coords_strides[d];
    v.mask_stride[d] = dm.d.mask_strides[d];
  }
  v.fcn = dm.d.fcn;
  v.mask = dm.d.mask;
  v.num_cells = (uint32_t)dm.num_cells;
  v.fast_extent = (uint32_t)dm.d.cell_shape[dm.slowest[DIM - 1]];
  return v;
}

template <int DIM>
mc::RowView<DIM> make_row_view(const axb_mc::Domain& dm)
{
  mc::RowView<DIM> v;
  const int F = dm.slowest[DIM - 1], M = dm.slowest[DIM - 2], S = dm.slowest[0];
  const int dir[3] = {F, M, S};
  for(int a = 0; a < 3; ++a)
  {
    const bool on = a < DIM;
    v.fs[a] = on ? dm.d.fcn_strides[dir[a]] : 0;
    v.ms[a] = on ? dm.d.mask_strides[dir[a]] : 0;
  }
  v.fcn = dm.d.fcn;
  v.mask = dm.d.mask;
  v.nf = (uint32_t)dm.d.cell_shape[F];
  v.nm = (uint32_t)dm.d.cell_shape[M];
  v.ns = DIM == 3 ? (uint32_t)dm.d.cell_shape[S] : 1u;
  v.cs_m = (uint32_t)dm.case_stride[M];
  v.cs_s = DIM == 3 ? (uint32_t)dm.case_stride[S] : 0u;
  const uint32_t chunks = (v.nf + mc::kUnitCells - 1) / mc::kUnitCells, mgroups = (v.nm + mc::kRows - 1) / mc::kRows;
  v.chunks = mc::make_fastdiv(chunks);
  v.mgroups = mc::make_fastdiv(mgroups);
  const uint32_t sgroups = DIM == 3 ? (v.ns + mc::kPlanes - 1) / mc::kPlanes : 1u;
  v.num_units = chunks * mgroups * sgroups;
  v.lut = dm.lut.as<uint8_t>();
  return v;
}

// compact corner bits of the row kernel -> the reference's case id: bit p of the code is the node at offsets
// (df, dm[, ds]) along (F, M[, S]) with p = df * CB + dm * NS + ds
template <int DIM>
void build_row_lut(axb_mc::Domain& dm)
{
  constexpr int NS = DIM == 3 ? 2 : 1, CB = DIM == 3 ? 4 : 2;
  const int F = dm.slowest[DIM - 1], M = dm.slowest[DIM - 2], S = dm.slowest[0];
  for(int code = 0; code < 256; ++code)
  {
    int case_id = 0;
    for(int p = 0; p < 2 * CB; ++p)
    {
      if(!((code >> p) & 1)) continue;
      int o[DIM];
      o[F] = p / CB;
      o[M] = (p % CB) / NS;
      if(DIM == 3) o[S] = p % NS;
      case_id |= 1 << mc::corner_id<DIM>(o);
    }
    dm.h_lut[code] = (uint8_t)case_id;
  }
}

// grow a device array to new_bytes keeping its first keep_bytes (Array::resize keeps the earlier contour)
int grow_preserve(Ctx& ctx, DevBuf& b, size_t keep_bytes, size_t new_bytes)
{
  if(new_bytes <= b.cap && b.p) return AXB_OK;
  DevBuf nb;
  AXB_TRY(nb.reserve(std::max(new_bytes, b.cap + b.cap / 2), ctx.stream));
  if(keep_bytes && b.p) AXB_CUDA_TRY(cudaMemcpyAsync(nb.p, b.p, keep_bytes, cudaMemcpyDeviceToDevice, ctx.stream));
  b.release(ctx.stream);
  b = nb;
  return AXB_OK;
}

template <int DIM>
int mc_compute(axb_mc* h, double contour_val)
{
  Ctx& ctx = h->ctx;
  const int nd = (int)h->doms.size();
  if(nd == 0) return ctx.finish_call();
  // (facets, active tiles) per domain land straight in pinned, device-mapped host memory: the scan kernel writes them
  // over PCIe and the stream synchronisation below makes them visible -- no copy to enqueue
  if(h->h_totals_cap < (size_t)nd)
  {
    if(h->h_totals) cudaFreeHost(h->h_totals);
    h->h_totals = nullptr;
    AXB_CUDA_TRY(cudaHostAlloc(&h->h_totals, sizeof(long long) * 2 * nd, cudaHostAllocMapped));
    h->h_totals_cap = nd;
  }
  long long* d_totals = nullptr;
  AXB_CUDA_TRY(cudaHostGetDevicePointer(&d_totals, h->h_totals, 0));
  // tuning / A-B switch: AXB_MC_MARK_PLAIN=1 selects the mark kernel without the lane-sharing of corner bits
  const char* env = getenv("AXB_MC_MARK_PLAIN");
  const bool plain_mark = env && env[0] == '1';
  // markCrossings + scanCrossings for every domain (MarchingCubes.cpp:112-121)
  for(int k = 0; k < nd; ++k)
  {
    axb_mc::Domain& dm = h->doms[k];
    long long* tot = d_totals + 2 * k;
    if(dm.num_cells == 0)
    {
      h->h_totals[2 * k] = h->h_totals[2 * k + 1] = 0;
      continue;
    }
    const mc::DomainView<DIM> v = make_view<DIM>(dm);
    {
      ScopedPhase ph(ctx, "mc.mark");
      if(plain_mark)
        AXB_LAUNCH(ctx, mc::mark_plain_kernel<DIM>, dm.num_tiles, mc::kTileThreads, v, contour_val, h->mask_val, dm.case_ids.as<uint8_t>());
      else
      {
        const mc::RowView<DIM> rv = make_row_view<DIM>(dm);
        const int warps_per_block = mc::kTileThreads / 32;
        const int grid = (int)std::min<long long>(((long long)rv.num_units + warps_per_block - 1) / warps_per_block, (long long)kNumSMsB200 * 8);
        if(rv.mask)
          AXB_LAUNCH(ctx, (mc::mark_rows_kernel<DIM, true>), grid, mc::kTileThreads, rv, contour_val, h->mask_val, dm.case_ids.as<uint8_t>());
        else
          AXB_LAUNCH(ctx, (mc::mark_rows_kernel<DIM, false>), grid, mc::kTileThreads, rv, contour_val, h->mask_val, dm.case_ids.as<uint8_t>());
      }
    }
    {
      // tile counts + (last block) their scan, totals and the active-tile list: one launch
      ScopedPhase ph(ctx, "mc.count");
      const int count_grid = (int)std::min<long long>(((long long)dm.num_tiles + 31) / 32, (long long)kNumSMsB200 * 2);
      AXB_LAUNCH(ctx, mc::count_scan_kernel<DIM>, count_grid, mc::kScanThreads, dm.case_ids.as<uint8_t>(), (uint32_t)dm.num_cells,
                 (uint32_t)dm.num_tiles, dm.tile_offsets.as<int32_t>(), tot, dm.active_tiles.as<int32_t>(), h->ticket.as<unsigned int>());
    }
  }
  AXB_TRY(ctx.sync());  // the facet count sizes the output, as m_facetCount does in the reference
  std::vector<long long> first(nd);
  long long count = h->facet_count;
  for(int k = 0; k < nd; ++k)
  {
    first[k] = count;  // m_facetIndexOffsets[d]
    count += h->h_totals[2 * k];
  }
  if(count * DIM > (long long)INT32_MAX)
    return fail(AXB_ERR_OVERFLOW, "contour node count does not fit the reference's 32-bit IndexType");
  // allocateOutputBuffers (:235-248)
  const size_t old = (size_t)h->facet_count, now = (size_t)count;
  AXB_TRY(grow_preserve(ctx, h->node_ids, old * DIM * sizeof(int32_t), now * DIM * sizeof(int32_t)));
  AXB_TRY(grow_preserve(ctx, h->node_coords, old * DIM * DIM * sizeof(double), now * DIM * DIM * sizeof(double)));
  AXB_TRY(grow_preserve(ctx, h->parent_ids, old * sizeof(int32_t), now * sizeof(int32_t)));
  AXB_TRY(grow_preserve(ctx, h->domain_ids, old * sizeof(int32_t), now * sizeof(int32_t)));
  // computeFacets per domain (:133-136)
  for(int k = 0; k < nd; ++k)
  {
    axb_mc::Domain& dm = h->doms[k];
    if(dm.num_cells == 0 || h->h_totals[2 * k] == 0) continue;
    const mc::DomainView<DIM> v = make_view<DIM>(dm);
    ScopedPhase ph(ctx, "mc.emit");
    AXB_LAUNCH(ctx, mc::emit_kernel<DIM>, (int)h->h_totals[2 * k + 1], mc::kTileThreads, v, contour_val, dm.case_ids.as<uint8_t>(),
               dm.tile_offsets.as<int32_t>(), dm.active_tiles.as<int32_t>(), (int32_t)first[k], (int32_t)dm.d.domain_id, h->node_ids.as<int32_t>(),
               h->node_coords.as<double>(), h->parent_ids.as<int32_t>(), h->domain_ids.as<int32_t>());
  }
  h->facet_count = count;
  return ctx.finish_call();
}
}  // namespace

extern "C" {

int axb_mc_create(axb_mc** out, int ndims, int device)
{
  if(!out) return fail(AXB_ERR_BAD_ARG, "null output handle");
  *out = nullptr;
  if(ndims != 2 && ndims != 3) return fail(AXB_ERR_BAD_ARG, "ndims must be 2 or 3");
  axb_mc* h = new axb_mc();
  h->ndims = ndims;
  int st = h->ctx.init(device);
  if(st != AXB_OK)
  {
    delete h;
    return st;
  }
  *out = h;
  return AXB_OK;
}

int axb_mc_destroy(axb_mc* h)
{
  if(!h) return AXB_OK;
  cudaSetDevice(h->ctx.device);
  h->release_domains();
  for(DevBuf* b : {&h->node_ids, &h->node_coords, &h->parent_ids, &h->domain_ids, &h->ticket}) b->release(h->ctx.stream);
  if(h->h_totals) cudaFreeHost(h->h_totals);
  cudaStreamSynchronize(h->ctx.stream);
  h->ctx.destroy();
  delete h;
  return AXB_OK;
}

int axb_mc_set_stream(axb_mc* h, void* s)
{
  if(!h) return fail(AXB_ERR_BAD_ARG, "null handle");
  AXB_TRY(h->ctx.sync());
  if(h->ctx.own_stream && h->ctx.stream) cudaStreamDestroy(h->ctx.stream);
  h->ctx.stream = (cudaStream_t)s;
  h->ctx.own_stream = false;
  return AXB_OK;
}

int axb_mc_set_mask_value(axb_mc* h, int mask_val)
{
  if(!h) return fail(AXB_ERR_BAD_ARG, "null handle");
  h->mask_val = mask_val;
  return AXB_OK;
}

int axb_mc_set_mesh(axb_mc* h, const axb_mc_domain* domains, int32_t num_domains, int memspace)
{
  if(!h) return fail(AXB_ERR_BAD_ARG, "null handle");
  if(num_domains < 0 || (num_domains > 0 && !domains)) return fail(AXB_ERR_BAD_ARG, "bad domain list");
  Ctx& ctx = h->ctx;
  AXB_TRY(ctx.bind());
  const int D = h->ndims;
  // validate everything before touching the handle's state
  for(int k = 0; k < num_domains; ++k)
  {
    const axb_mc_domain& in = domains[k];
    long long cells = 1;
    for(int d = 0; d < D; ++d)
    {
      if(in.cell_shape[d] < 0) return fail(AXB_ERR_BAD_ARG, "negative cell shape");
      cells *= in.cell_shape[d];
      if(cells > (long long)INT32_MAX) return fail(AXB_ERR_OVERFLOW, "cell count does not fit the reference's 32-bit IndexType");
    }
    if(cells == 0) continue;
    if(!in.fcn) return fail(AXB_ERR_BAD_ARG, "null function field");
    for(int d = 0; d < D; ++d)
    {
      if(!in.coords[d]) return fail(AXB_ERR_BAD_ARG, "null coordinate array");
      if(in.fcn_strides[d] <= 0 || in.coords_strides[d] <= 0 || (in.mask && in.mask_strides[d] <= 0))
        return fail(AXB_ERR_BAD_ARG, "strides must be positive");
      for(int e = 0; e < d; ++e)
        if(in.fcn_strides[d] == in.fcn_strides[e])
          return fail(AXB_ERR_BAD_ARG, "non-unique function strides: impossible to compute index ordering (MDMapping)");
    }
  }
  h->release_domains();
  h->doms.resize(num_domains);
  auto stage = [&]() -> int {
  AXB_TRY(h->ticket.reserve(sizeof(unsigned int), ctx.stream));
  AXB_CUDA_TRY(cudaMemsetAsync(h->ticket.p, 0, sizeof(unsigned int), ctx.stream));
  for(int k = 0; k < num_domains; ++k)
  {
    axb_mc::Domain& dm = h->doms[k];
    dm.d = domains[k];
    long long cells = 1;
    for(int d = 0; d < D; ++d) cells *= dm.d.cell_shape[d];
    dm.num_cells = cells;
    if(cells == 0) continue;
    // MDMapping::initializeStrides(strides, ROW) (core/MDMapping.hpp:255-275): directions by decreasing stride
    for(int d = 0; d < D; ++d) dm.slowest[d] = d;
    for(int s = 0; s < D; ++s)
      for(int d = s; d < D; ++d)
        if(dm.d.fcn_strides[dm.slowest[s]] < dm.d.fcn_strides[dm.slowest[d]]) std::swap(dm.slowest[s], dm.slowest[d]);
    // MDMapping::initializeShape(bShape, slowestDirs) (:182-198): compact strides of the case-id array in that order
    dm.case_stride[dm.slowest[D - 1]] = 1;
    for(int s = D - 2; s >= 0; --s)
      dm.case_stride[dm.slowest[s]] = dm.case_stride[dm.slowest[s + 1]] * dm.d.cell_shape[dm.slowest[s + 1]];
    dm.num_tiles = (int)((cells + mc::kTileCells - 1) / mc::kTileCells);
    const int ms = resolve_memspace(memspace, dm.d.fcn);
    if(ms != AXB_MEM_HOST && ms != AXB_MEM_DEVICE) return fail(AXB_ERR_BAD_ARG, "unknown memspace");
    if(ms == AXB_MEM_HOST)
    {
      const long long nspan = view_span(dm.d.cell_shape, 1, dm.d.coords_strides, D);
      for(int d = 0; d < D; ++d)
      {
        AXB_TRY(dm.staged[d].reserve(sizeof(double) * nspan, ctx.stream));
        AXB_CUDA_TRY(cudaMemcpyAsync(dm.staged[d].p, dm.d.coords[d], sizeof(double) * nspan, cudaMemcpyHostToDevice, ctx.stream));
        dm.d.coords[d] = dm.staged[d].as<double>();
      }
      const long long fspan = view_span(dm.d.cell_shape, 1, dm.d.fcn_strides, D);
      AXB_TRY(dm.staged[3].reserve(sizeof(double) * fspan, ctx.stream));
      AXB_CUDA_TRY(cudaMemcpyAsync(dm.staged[3].p, dm.d.fcn, sizeof(double) * fspan, cudaMemcpyHostToDevice, ctx.stream));
      dm.d.fcn = dm.staged[3].as<double>();
      if(dm.d.mask)
      {
        const long long mspan = view_span(dm.d.cell_shape, 0, dm.d.mask_strides, D);
        AXB_TRY(dm.staged[4].reserve(sizeof(int32_t) * mspan, ctx.stream));
        AXB_CUDA_TRY(cudaMemcpyAsync(dm.staged[4].p, dm.d.mask, sizeof(int32_t) * mspan, cudaMemcpyHostToDevice, ctx.stream));
        dm.d.mask = dm.staged[4].as<int32_t>();
      }
    }
    if(D == 2)
      build_row_lut<2>(dm);
    else
      build_row_lut<3>(dm);
    AXB_TRY(dm.lut.reserve(256, ctx.stream));
    AXB_CUDA_TRY(cudaMemcpyAsync(dm.lut.p, dm.h_lut, 256, cudaMemcpyHostToDevice, ctx.stream));
    AXB_TRY(dm.case_ids.reserve((size_t)dm.num_tiles * mc::kTileCells, ctx.stream));
    AXB_TRY(dm.tile_offsets.reserve(sizeof(int32_t) * ((size_t)dm.num_tiles + 1), ctx.stream));
    AXB_TRY(dm.active_tiles.reserve(sizeof(int32_t) * (size_t)dm.num_tiles, ctx.stream));
  }
  return ctx.sync();  // host inputs may be released by the caller on return
  };
  const int st = stage();
  if(st != AXB_OK) h->release_domains();  // never leave a half-staged mesh behind
  return st;
}

int axb_mc_compute_isocontour(axb_mc* h, double contour_val)
{
  if(!h) return fail(AXB_ERR_BAD_ARG, "null handle");
  AXB_TRY(h->ctx.bind());
  return h->ndims == 2 ? mc_compute<2>(h, contour_val) : mc_compute<3>(h, contour_val);
}

int axb_mc_get_contour_cell_count(const axb_mc* h, int64_t* n)
{
  if(!h || !n) return fail(AXB_ERR_BAD_ARG, "null argument");
  *n = h->facet_count;
  return AXB_OK;
}

int axb_mc_get_contour_node_count(const axb_mc* h, int64_t* n)
{
  if(!h || !n) return fail(AXB_ERR_BAD_ARG, "null argument");
  *n = h->doms.empty() ? 0 : h->facet_count * h->ndims;  // MarchingCubes.cpp:149-154
  return AXB_OK;
}

int axb_mc_get_contour_views(axb_mc* h, const int32_t** ids, const double** coords, const int32_t** parents, const int32_t** domains)
{
  if(!h) return fail(AXB_ERR_BAD_ARG, "null handle");
  const bool any = h->facet_count > 0;
  if(ids) *ids = any ? h->node_ids.as<int32_t>() : nullptr;
  if(coords) *coords = any ? h->node_coords.as<double>() : nullptr;
  if(parents) *parents = any ? h->parent_ids.as<int32_t>() : nullptr;
  if(domains) *domains = any ? h->domain_ids.as<int32_t>() : nullptr;
  return AXB_OK;
}

int axb_mc_copy_contour(axb_mc* h, int memspace, int32_t* ids, double* coords, int32_t* parents, int32_t* domains)
{
  if(!h) return fail(AXB_ERR_BAD_ARG, "null handle");
  if(memspace != AXB_MEM_HOST && memspace != AXB_MEM_DEVICE) return fail(AXB_ERR_BAD_ARG, "memspace must be HOST or DEVICE");
  Ctx& ctx = h->ctx;
  AXB_TRY(ctx.bind());
  const size_t n = (size_t)h->facet_count, D = (size_t)h->ndims;
  const cudaMemcpyKind kind = memspace == AXB_MEM_HOST ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice;
  if(n)
  {
    if(ids) AXB_CUDA_TRY(cudaMemcpyAsync(ids, h->node_ids.p, n * D * sizeof(int32_t), kind, ctx.stream));
    if(coords) AXB_CUDA_TRY(cudaMemcpyAsync(coords, h->node_coords.p, n * D * D * sizeof(double), kind, ctx.stream));
    if(parents) AXB_CUDA_TRY(cudaMemcpyAsync(parents, h->parent_ids.p, n * sizeof(int32_t), kind, ctx.stream));
    if(domains) AXB_CUDA_TRY(cudaMemcpyAsync(domains, h->domain_ids.p, n * sizeof(int32_t), kind, ctx.stream));
  }
  return ctx.sync();
}

int axb_mc_clear_output(axb_mc* h)
{
  if(!h) return fail(AXB_ERR_BAD_ARG, "null handle");
  h->facet_count = 0;  // buffers are kept for the next contour, like Array::clear()
  return AXB_OK;
}

int axb_mc_set_profiling(axb_mc* h, int enabled)
{
  if(!h) return fail(AXB_ERR_BAD_ARG, "null handle");
  h->ctx.set_profiling(enabled != 0);
  return AXB_OK;
}

int axb_mc_get_phase_ms(const axb_mc* hc, const char* name, double* ms)
{
  axb_mc* h = const_cast<axb_mc*>(hc);
  if(!h || !name || !ms) return fail(AXB_ERR_BAD_ARG, "null argument");
  h->ctx.resolve();
  auto it = h->ctx.acc.find(name);
  if(it == h->ctx.acc.end() || it->second.calls == 0) return fail(AXB_ERR_BAD_ARG, std::string("no timing recorded for phase ") + name);
  *ms = it->second.sum / (double)it->second.calls;
  return AXB_OK;
}

int axb_mc_launch_count(const axb_mc* h, int64_t* n)
{
  if(!h || !n) return fail(AXB_ERR_BAD_ARG, "null argument");
  *n = h->ctx.launches;
  return AXB_OK;
}

}  // extern "C"
