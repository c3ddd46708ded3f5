// comm.cuh -- the exchange steps of the distributed cases, inside the library: one NCCL communicator per rank
// (axb_comm), the partitioned-surface MIN (C5) and quest::DistributedClosestPoint::computeClosestPoints as a whole.
//
// Reference path replaced (quest/detail/DistributedClosestPointImpl.hpp):
//   :687-693   gatherBVHRoots         MPI_Allgather of every rank's object bounding box
//   :737-851   computeClosestPoints   the ring: a query block visits owner, owner+1, ... as Conduit messages over
//                                     MPI_Isend / MPI_Irecv; ranks whose box is farther than the threshold are skipped
//   :905-1079  computeLocalClosestPoints (dcp.cuh) keeps an entry only if this rank holds a STRICTLY nearer point
// so among equidistant points the first rank in ring order from the owner wins.  On one NVSwitch box the ring is
// replaced by collectives on the handle's stream (every rank reaches every peer at full bandwidth; a ring of N hops
// would serialise N searches):
//   1  all-gather of the query blocks and of the object bounding boxes
//   2  every query is searched, unbounded, by the rank whose box CENTRE is nearest; MIN all-reduce of the squared
//      distances found = an upper bound for everybody                                        8 B / query
//   3  the other ranks search a query only if their box is within that bound, and only for points at least that near
//   4  MIN all-reduce of the squared distances (recomputed with the reference's expression, equal values are
//      bit-equal), MIN all-reduce of the ring position of the ranks that attain it            8 + 1 B / query
//   5  every query's winner is now known to all: ONE all-to-all (grouped ncclSend / ncclRecv) carries the winner's
//      48-byte record to the query's home rank -- each payload crosses NVLink once            48 B / query
// NCCL is bound at run time (dlopen): libaxb200.so has no link-time dependency on it, and in a torch process the
// copy torch already loaded is the one used.
#pragma once
#include <dlfcn.h>
#include <nccl.h>

namespace axb
{
//------------------------------------------------------------------------------------------
// NCCL, resolved at first use
//------------------------------------------------------------------------------------------
struct NcclApi
{
  void* lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  ncclResult_t (*GetVersion)(int*) = nullptr;
  std::string where;
};

static NcclApi g_nccl;
static std::mutex g_nccl_mutex;

static int nccl_api(NcclApi** out)
{
  std::lock_guard<std::mutex> lock(g_nccl_mutex);
  if(!g_nccl.lib)
  {
    void* lib = nullptr;
    std::string tried;
    if(const char* env = getenv("AXB_NCCL_LIB"))
    {
      lib = dlopen(env, RTLD_NOW | RTLD_GLOBAL);
      tried += std::string(env) + " ";
      if(lib) g_nccl.where = env;
    }
    // a copy that is already mapped (torch's, or the host code's own) comes first
    for(const char* name : {"libnccl.so.2", "libnccl.so"})
    {
      if(lib) break;
      lib = dlopen(name, RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);
      if(lib) g_nccl.where = std::string(name) + " (already loaded)";
    }
    for(const char* name : {"libnccl.so.2", "libnccl.so"})
    {
      if(lib) break;
      lib = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
      tried += std::string(name) + " ";
      if(lib) g_nccl.where = name;
    }
    if(!lib) return fail(AXB_ERR_UNSUPPORTED, "NCCL not found (tried " + tried + "; set AXB_NCCL_LIB to the path of libnccl.so.2)");
    bool ok = true;
    auto sym = [&](const char* n) {
      void* p = dlsym(lib, n);
      ok = ok && p != nullptr;
      return p;
    };
    g_nccl.GetUniqueId = (decltype(g_nccl.GetUniqueId))sym("ncclGetUniqueId");
    g_nccl.CommInitRank = (decltype(g_nccl.CommInitRank))sym("ncclCommInitRank");
    g_nccl.CommDestroy = (decltype(g_nccl.CommDestroy))sym("ncclCommDestroy");
    g_nccl.GetErrorString = (decltype(g_nccl.GetErrorString))sym("ncclGetErrorString");
    g_nccl.AllReduce = (decltype(g_nccl.AllReduce))sym("ncclAllReduce");
    g_nccl.AllGather = (decltype(g_nccl.AllGather))sym("ncclAllGather");
    g_nccl.Broadcast = (decltype(g_nccl.Broadcast))sym("ncclBroadcast");
    g_nccl.Send = (decltype(g_nccl.Send))sym("ncclSend");
    g_nccl.Recv = (decltype(g_nccl.Recv))sym("ncclRecv");
    g_nccl.GroupStart = (decltype(g_nccl.GroupStart))sym("ncclGroupStart");
    g_nccl.GroupEnd = (decltype(g_nccl.GroupEnd))sym("ncclGroupEnd");
    g_nccl.GetVersion = (decltype(g_nccl.GetVersion))sym("ncclGetVersion");
    if(!ok) return fail(AXB_ERR_UNSUPPORTED, "the NCCL library found (" + g_nccl.where + ") lacks a required symbol");
    g_nccl.lib = lib;
  }
  *out = &g_nccl;
  return AXB_OK;
}

#define AXB_NCCL_TRY(api, expr)                                                                                              \
  do                                                                                                                         \
  {                                                                                                                          \
    ncclResult_t _r = (expr);                                                                                                \
    if(_r != ncclSuccess)                                                                                                    \
      return fail(AXB_ERR_CUDA, std::string(#expr) + ": " + (api)->GetErrorString(_r) + " (" + __FILE__ + ":" + std::to_string(__LINE__) + ")"); \
  } while(0)

}  // namespace axb

struct axb_comm
{
  axb::NcclApi* api = nullptr;
  ncclComm_t comm = nullptr;
  int nranks = 1, rank = 0, device = 0;
  int64_t bytes = 0;  // payload bytes this rank has put into collectives (send side), cumulative
  int64_t calls = 0;
};

namespace axb
{
//------------------------------------------------------------------------------------------
// kernels of the collective form of computeClosestPoints
//------------------------------------------------------------------------------------------
constexpr int kDcpxMaxRanks = 254;  // ring positions travel as one byte

struct alignas(16) DcpxRecord  // what the winner sends to the query's home rank
{
  int32_t q;  // index of the query in its home rank's block
  int32_t cp_index, cp_domain_index, cp_rank;
  double cp_distance;
  double cp_coords[3];
};
static_assert(sizeof(DcpxRecord) == 48, "DcpxRecord is 48 bytes");

struct DcpxRanks  // per-call constants, by value
{
  const long long* offs;  // [nranks + 1] first query of every home rank in the gathered order
  const double* boxes;    // [nranks][2 * 3] object bounding boxes lo, hi (unused components 0)
  int nranks, rank;
};

__device__ __forceinline__ int dcpx_home(const long long* __restrict__ offs, int nranks, long long g)
{
  int lo = 0, hi = nranks;  // offs[lo] <= g < offs[hi]
  while(hi - lo > 1)
  {
    const int mid = (lo + hi) >> 1;
    if(__ldg(offs + mid) <= g)
      lo = mid;
    else
      hi = mid;
  }
  return lo;
}

// squared_distance(qpt, cp) as the reference's checkMinDist computes it (sum of separately rounded squares, in order)
template <int D>
__device__ __forceinline__ double dcpx_sq(const double* __restrict__ cp, const double* __restrict__ q)
{
  double s = 0.0;
#pragma unroll
  for(int d = 0; d < D; ++d)
  {
    const double v = cp[d] - q[d];
    s += v * v;
  }
  return s;
}

// warp-aggregated append of the flagged items' indices
__device__ __forceinline__ void dcpx_append(bool flag, int32_t value, int32_t* __restrict__ list, unsigned int* __restrict__ count)
{
  const unsigned m = __ballot_sync(0xffffffffu, flag);
  if(m == 0u) return;
  const unsigned lane = threadIdx.x & 31u;
  unsigned base = 0;
  if(lane == (unsigned)(__ffs(m) - 1)) base = atomicAdd(count, (unsigned)__popc(m));
  base = __shfl_sync(0xffffffffu, base, __ffs(m) - 1);
  if(flag) list[base + __popc(m & ((1u << lane) - 1u))] = value;
}

// step 2 classification: distance from every gathered query to THIS rank's object box (`mine`, only ever used to
// prune), the rank whose box centre is nearest (`nearest`: who searches the query first -- any rule every rank
// evaluates identically is correct, this one balances the ranks where boxes overlap), and the first-search list
template <int D>
__global__ void __launch_bounds__(256) dcpx_classify_kernel(const double* __restrict__ Q, long long ntot, DcpxRanks R, double sq_thresh,
                                                            double* __restrict__ mine, int32_t* __restrict__ nearest, int32_t* __restrict__ slot,
                                                            int32_t* __restrict__ list1, unsigned int* __restrict__ count1)
{
  const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  bool first = false;
  if(g < ntot)
  {
    double q[D];
#pragma unroll
    for(int d = 0; d < D; ++d) q[d] = Q[g * D + d];
    const double inf = __longlong_as_double(0x7ff0000000000000ll);
    double best = inf;
    int who = 0;
    double my = inf;
    for(int r = 0; r < R.nranks; ++r)
    {
      const double* b = R.boxes + 6 * r;
      bool valid = true;
#pragma unroll
      for(int d = 0; d < D; ++d) valid = valid && !(b[d] > b[3 + d]);
      if(!valid) continue;
      double c2 = 0.0, g2 = 0.0;
#pragma unroll
      for(int d = 0; d < D; ++d)
      {
        const double dc = q[d] - 0.5 * (b[d] + b[3 + d]);
        c2 += dc * dc;
        const double gap = fmax(fmax(b[d] - q[d], q[d] - b[3 + d]), 0.0);
        g2 += gap * gap;
      }
      if(c2 < best)
      {
        best = c2;
        who = r;
      }
      if(r == R.rank) my = g2;
    }
    mine[g] = my;
    nearest[g] = who;
    slot[g] = -1;
    first = (who == R.rank) && (my <= sq_thresh);
  }
  dcpx_append(first, (int32_t)g, list1, count1);
}

// the second-search list: queries some OTHER rank searched first, whose bound (and the threshold) reaches this rank's box
__global__ void __launch_bounds__(256) dcpx_classify2_kernel(long long ntot, int rank, double sq_thresh, const double* __restrict__ mine,
                                                             const int32_t* __restrict__ nearest, const double* __restrict__ bound,
                                                             int32_t* __restrict__ list2, unsigned int* __restrict__ count2)
{
  const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  bool second = false;
  if(g < ntot) second = (nearest[g] != rank) && (mine[g] <= fmin(bound[g], sq_thresh));
  dcpx_append(second, (int32_t)g, list2, count2);
}

// queries (and their bounds) of a list, packed for the search kernels; slot[g] = where the search leaves its result
template <int D>
__global__ void __launch_bounds__(256) dcpx_gather_kernel(const double* __restrict__ Q, const int32_t* __restrict__ list, int n, int slot0,
                                                          const double* __restrict__ bound, double* __restrict__ q_out, double* __restrict__ bound_out,
                                                          int32_t* __restrict__ slot)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if(i >= n) return;
  const long long g = list[i];
#pragma unroll
  for(int d = 0; d < D; ++d) q_out[(size_t)i * D + d] = Q[g * D + d];
  if(bound_out) bound_out[i] = bound[g];
  slot[g] = slot0 + i;
}

// squared distance of what this rank holds for every gathered query (+inf: nothing)
template <int D>
__global__ void __launch_bounds__(256) dcpx_sq_kernel(const double* __restrict__ Q, long long ntot, const int32_t* __restrict__ slot,
                                                      const int32_t* __restrict__ st_rank, const double* __restrict__ st_coords,
                                                      double* __restrict__ sq_a, double* __restrict__ sq_b)
{
  const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if(g >= ntot) return;
  const int s = slot[g];
  double v = __longlong_as_double(0x7ff0000000000000ll);
  if(s >= 0 && st_rank[s] >= 0) v = dcpx_sq<D>(st_coords + (size_t)s * D, Q + g * D);
  sq_a[g] = v;
  if(sq_b) sq_b[g] = v;
}

// ring position (rank - home) mod N of this rank where it attains the global minimum, else N
__global__ void __launch_bounds__(256) dcpx_pos_kernel(long long ntot, DcpxRanks R, const double* __restrict__ sq, const double* __restrict__ smin,
                                                       uint8_t* __restrict__ pos_a, uint8_t* __restrict__ pos_b)
{
  const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if(g >= ntot) return;
  const double v = sq[g];
  uint8_t p = (uint8_t)R.nranks;
  if(v < __longlong_as_double(0x7ff0000000000000ll) && v == smin[g])
  {
    const int home = dcpx_home(R.offs, R.nranks, g);
    p = (uint8_t)((R.rank - home + R.nranks) % R.nranks);
  }
  pos_a[g] = p;
  pos_b[g] = p;
}

// how many records this rank sends to every home rank, and how many it receives from every winner
__global__ void __launch_bounds__(256) dcpx_count_kernel(long long ntot, DcpxRanks R, const uint8_t* __restrict__ pos, const uint8_t* __restrict__ win,
                                                         unsigned int* __restrict__ send_count, unsigned int* __restrict__ recv_count)
{
  __shared__ unsigned int s_send[kDcpxMaxRanks + 2], s_recv[kDcpxMaxRanks + 2];
  for(int i = threadIdx.x; i < R.nranks; i += blockDim.x) s_send[i] = s_recv[i] = 0u;
  __syncthreads();
  for(long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x; g < ntot; g += (long long)gridDim.x * blockDim.x)
  {
    const int w = win[g];
    if(w >= R.nranks) continue;  // nobody holds a point for this query
    const int home = dcpx_home(R.offs, R.nranks, g);
    if(pos[g] == w) atomicAdd(&s_send[home], 1u);
    if(home == R.rank) atomicAdd(&s_recv[(home + w) % R.nranks], 1u);
  }
  __syncthreads();
  for(int i = threadIdx.x; i < R.nranks; i += blockDim.x)
  {
    if(s_send[i]) atomicAdd(send_count + i, s_send[i]);
    if(s_recv[i]) atomicAdd(recv_count + i, s_recv[i]);
  }
}

// the records this rank won, grouped by home rank (order inside a group is irrelevant: the query index travels along)
template <int D>
__global__ void __launch_bounds__(256) dcpx_pack_kernel(long long ntot, DcpxRanks R, const uint8_t* __restrict__ pos, const uint8_t* __restrict__ win,
                                                        const int32_t* __restrict__ slot, const int32_t* __restrict__ st_idx,
                                                        const int32_t* __restrict__ st_dom, const int32_t* __restrict__ st_rank,
                                                        const double* __restrict__ st_coords, const double* __restrict__ st_dist,
                                                        const long long* __restrict__ send_off, unsigned int* __restrict__ fill,
                                                        DcpxRecord* __restrict__ out)
{
  const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if(g >= ntot) return;
  const int w = win[g];
  if(w >= R.nranks || pos[g] != w) return;
  const int home = dcpx_home(R.offs, R.nranks, g);
  const int s = slot[g];
  const long long at = send_off[home] + (long long)atomicAdd(fill + home, 1u);
  DcpxRecord r;
  r.q = (int32_t)(g - R.offs[home]);
  r.cp_index = st_idx[s];
  r.cp_domain_index = st_dom[s];
  r.cp_rank = st_rank[s];
  r.cp_distance = st_dist[s];
#pragma unroll
  for(int d = 0; d < 3; ++d) r.cp_coords[d] = d < D ? st_coords[(size_t)s * D + d] : 0.0;
  out[at] = r;
}

template <int D>
__global__ void __launch_bounds__(256) dcpx_init_outputs_kernel(int n, int32_t* __restrict__ cp_index, int32_t* __restrict__ cp_dom,
                                                                int32_t* __restrict__ cp_rank, double* __restrict__ cp_coords,
                                                                double* __restrict__ cp_dist)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if(i >= n) return;
  const double snan = __longlong_as_double(0x7ff4000000000000ll);  // std::numeric_limits<double>::signaling_NaN (:1010-1016)
  cp_index[i] = -1;
  cp_dom[i] = -1;
  cp_rank[i] = -1;
#pragma unroll
  for(int d = 0; d < D; ++d) cp_coords[(size_t)i * D + d] = snan;
  cp_dist[i] = snan;
}

template <int D>
__global__ void __launch_bounds__(256) dcpx_unpack_kernel(const DcpxRecord* __restrict__ rec, long long n, int32_t* __restrict__ cp_index,
                                                          int32_t* __restrict__ cp_dom, int32_t* __restrict__ cp_rank, double* __restrict__ cp_coords,
                                                          double* __restrict__ cp_dist)
{
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if(i >= n) return;
  const DcpxRecord r = rec[i];
  cp_index[r.q] = r.cp_index;
  cp_dom[r.q] = r.cp_domain_index;
  cp_rank[r.q] = r.cp_rank;
#pragma unroll
  for(int d = 0; d < D; ++d) cp_coords[(size_t)r.q * D + d] = r.cp_coords[d];
  cp_dist[r.q] = r.cp_distance;
}

}  // namespace axb
