// mc.cuh -- quest::MarchingCubes on the device (the consumer of the 256^3 distance field).
//
// Reference: quest/detail/MarchingCubesImpl.hpp (markCrossings :157-362, scanCrossings :364-573, computeFacets
// :575-808) driven by quest/MarchingCubes.cpp:107-147.  The reference runs, per domain, a case-id pass, a flag pass, a
// scan, a compaction pass, a second scan and the facet pass, keeping 2 + 4 + 4 + 4 bytes per CELL of scratch and
// 2 + 4 + 4 + 4 bytes per crossing.  Here:
//
//   mark_count_kernel   one read of the nodal function (8 B / node, the compulsory traffic): case id per cell
//                       (1 byte, kept) and the facet count of each 1024-cell tile
//   scan_tiles_kernel   exclusive scan of the tile counts (one block; 16 K tiles for 255^3 cells) + the domain total
//   emit_kernel         tiles without facets exit on two loads; the others rescan their 1024 case bytes in shared
//                       memory and write the facets of their crossing cells straight into the output arrays
//
// Output order is the reference's: facets sorted by the parent cell's flat index in the case-id array (whose stride
// order follows the function's, MarchingCubesImpl.hpp:164-169), then by the case table's facet order; facet f owns
// nodes DIM*f .. DIM*f + DIM-1 (:601-616).  Both data-parallel variants of the reference (hybridParallel /
// fullParallel) produce exactly this order, so there is one device path.
//
// Arithmetic: linear_interp (:755-807) is restated operation by operation; the library is built with -fmad=false so
// p1 + w * (p2 - p1) rounds twice as in the reference's x86-64 build, and double division is IEEE on the device.
#pragma once
#include "common.cuh"
#include "mc_tables.h"

namespace axb
{
namespace mc
{
constexpr int kTileThreads = 256;
constexpr int kCellsPerThread = 4;
constexpr int kTileCells = kTileThreads * kCellsPerThread;  // 1024 cells of the flat case-id order per block

__constant__ uint64_t c_cases3d[256] = {AXB_MC_CASES3D_WORDS};
__constant__ uint16_t c_cases2d[16] = {AXB_MC_CASES2D_WORDS};

// exact n / d for 0 <= n < 2^31 by multiply-high (d >= 1): the flat cell index is a 32-bit IndexType in the reference
struct FastDiv
{
  uint32_t d, mul, shr;
};
inline FastDiv make_fastdiv(uint32_t d)
{
  FastDiv f;
  f.d = d;
  f.mul = 0;
  f.shr = 0;
  if(d > 1)
  {
    uint32_t lg = 0;
    while((1ull << lg) < d) ++lg;  // ceil(log2 d)
    const uint32_t p = 31 + lg;
    f.mul = (uint32_t)(((1ull << p) + d - 1) / d);
    f.shr = p - 32;
  }
  return f;
}
__host__ __device__ __forceinline__ uint32_t fastdiv(uint32_t n, const FastDiv& f)
{
#ifdef __CUDA_ARCH__
  return f.d == 1 ? n : (__umulhi(n, f.mul) >> f.shr);
#else
  return f.d == 1 ? n : (uint32_t)((((uint64_t)n * f.mul) >> 32) >> f.shr);
#endif
}

// one domain: the views MarchingCubesImpl::setDomain / setFunctionField hold (:100-145), ghost-free, strides in elements
template <int DIM>
struct DomainView
{
  FastDiv case_div[DIM];     // case-id strides (MDMapping::initializeShape(bShape, slowestDirs), core/MDMapping.hpp:182-198)
  int slowest[DIM];          // slowestDirs of the function's strides
  const double* fcn;
  long long fcn_stride[DIM];
  const double* coords[DIM];
  long long coords_stride[DIM];
  const int32_t* mask;       // nullptr: no mask
  long long mask_stride[DIM];
  uint32_t num_cells;
};

// MDMapping::toMultiIndex (core/MDMapping.hpp:361-371)
template <int DIM>
__device__ __forceinline__ void to_multi_index(const DomainView<DIM>& v, uint32_t flat, uint32_t idx[DIM])
{
#pragma unroll
  for(int s = 0; s < DIM; ++s)
  {
    const int dir = v.slowest[s];
    const uint32_t q = fastdiv(flat, v.case_div[dir]);
    idx[dir] = q;
    flat -= q * v.case_div[dir].d;
  }
}

// corner c of a cell as a node offset from the cell's (i,j[,k]) node: MarchingCubesImpl.hpp:329-333 (2-D), :349-357 (3-D)
template <int DIM>
__device__ __forceinline__ void corner_offset(int c, int o[DIM])
{
  if(DIM == 2)
  {
    o[0] = (c == 1 || c == 2);
    o[1] = (c >= 2);
  }
  else
  {
    o[0] = ((c & 3) < 2);          // corners 0,1,4,5 sit at i+1
    o[1] = ((c & 3) == 1 || (c & 3) == 2);
    o[DIM - 1] = (c >> 2);
  }
}

template <int DIM>
__device__ __forceinline__ int used_entries(int case_id)
{
  // entries are packed from nibble 0 and terminated by 0xF nibbles
  if(DIM == 2)
  {
    const uint32_t w = c_cases2d[case_id];
    const uint32_t m = w & (w >> 1) & (w >> 2) & (w >> 3) & 0x1111u;
    return m ? (__ffs(m) - 1) >> 2 : 4;
  }
  const uint64_t w = c_cases3d[case_id];
  const uint64_t m = w & (w >> 1) & (w >> 2) & (w >> 3) & 0x1111111111111111ull;
  return (__ffsll((long long)m) - 1) >> 2;  // nibble 15 is never used, so m != 0
}

//------------------------------------------------------------------------------------------
// pass 1: computeCaseId (:306-361) for every cell + per-tile facet count (num_contour_cells, :815-843)
//------------------------------------------------------------------------------------------
template <int DIM>
__global__ void __launch_bounds__(kTileThreads) mark_count_kernel(DomainView<DIM> v, double contour_val, int mask_val,
                                                                  uint8_t* __restrict__ case_ids, int32_t* __restrict__ tile_facets)
{
  constexpr int NCASE = DIM == 2 ? 16 : 256;
  constexpr int NCORNER = DIM == 2 ? 4 : 8;
  __shared__ uint8_t s_nfacets[NCASE];
  __shared__ int s_warp[kTileThreads / 32];
  if(threadIdx.x < NCASE) s_nfacets[threadIdx.x] = (uint8_t)(used_entries<DIM>(threadIdx.x) / DIM);
  __syncthreads();

  const uint32_t base = blockIdx.x * (uint32_t)kTileCells;
  int nf = 0;
#pragma unroll
  for(int r = 0; r < kCellsPerThread; ++r)
  {
    const uint32_t n = base + r * kTileThreads + threadIdx.x;  // consecutive lanes -> consecutive cells of the fastest direction
    int case_id = 0;                                           // m_caseIdsFlat.fill(0) (:162)
    if(n < v.num_cells)
    {
      uint32_t idx[DIM];
      to_multi_index<DIM>(v, n, idx);
      bool use_zone = true;
      if(v.mask)
      {
        long long mo = 0;
#pragma unroll
        for(int d = 0; d < DIM; ++d) mo += (long long)idx[d] * v.mask_stride[d];
        use_zone = (__ldg(v.mask + mo) == mask_val);
      }
      if(use_zone)
      {
        long long fo = 0;
#pragma unroll
        for(int d = 0; d < DIM; ++d) fo += (long long)idx[d] * v.fcn_stride[d];
#pragma unroll
        for(int c = 0; c < NCORNER; ++c)
        {
          int o[DIM];
          corner_offset<DIM>(c, o);
          long long off = fo;
#pragma unroll
          for(int d = 0; d < DIM; ++d) off += o[d] ? v.fcn_stride[d] : 0;
          if(__ldg(v.fcn + off) >= contour_val) case_id |= (1 << c);  // computeCrossingCase (:307-319)
        }
      }
    }
    // the tail of the last tile is written too (case 0), so pass 3 can read whole tiles
    case_ids[n] = (uint8_t)case_id;
    nf += s_nfacets[case_id];
  }
#pragma unroll
  for(int o = 16; o > 0; o >>= 1) nf += __shfl_xor_sync(0xffffffffu, nf, o);
  if((threadIdx.x & 31) == 0) s_warp[threadIdx.x >> 5] = nf;
  __syncthreads();
  if(threadIdx.x == 0)
  {
    int t = 0;
#pragma unroll
    for(int w = 0; w < kTileThreads / 32; ++w) t += s_warp[w];
    tile_facets[blockIdx.x] = t;
  }
}

//------------------------------------------------------------------------------------------
// pass 2: exclusive scan of the tile counts, in place, + total (the two inclusive scans of :413-483 collapse
// into this one because crossing ids are never an output)
//------------------------------------------------------------------------------------------
constexpr int kScanThreads = 1024;
__global__ void __launch_bounds__(kScanThreads) scan_tiles_kernel(int32_t* __restrict__ tile_facets, int num_tiles, long long* __restrict__ total)
{
  __shared__ long long s_part[kScanThreads];
  const int per = (num_tiles + kScanThreads - 1) / kScanThreads;
  const int lo = min(num_tiles, (int)threadIdx.x * per), hi = min(num_tiles, lo + per);
  long long sum = 0;
  for(int i = lo; i < hi; ++i) sum += tile_facets[i];
  s_part[threadIdx.x] = sum;
  __syncthreads();
  // Hillis-Steele inclusive scan over the 1024 partial sums
  for(int o = 1; o < kScanThreads; o <<= 1)
  {
    const long long add = (int)threadIdx.x >= o ? s_part[threadIdx.x - o] : 0;
    __syncthreads();
    s_part[threadIdx.x] += add;
    __syncthreads();
  }
  long long run = s_part[threadIdx.x] - sum;
  for(int i = lo; i < hi; ++i)
  {
    const int c = tile_facets[i];
    // offsets are 32-bit like the reference's IndexType; the host rejects totals that do not fit before pass 3 runs
    tile_facets[i] = (int32_t)run;
    run += c;
  }
  if(threadIdx.x == kScanThreads - 1)
  {
    tile_facets[num_tiles] = (int32_t)s_part[threadIdx.x];
    *total = s_part[threadIdx.x];
  }
}

//------------------------------------------------------------------------------------------
// pass 3: computeFacets (:575-620) -- get_corner_coords_and_values (:644-703) + linear_interp (:706-807)
//------------------------------------------------------------------------------------------
template <int DIM>
__device__ __forceinline__ void linear_interp(int edge, const double (*cc)[DIM], const double* cv, double contour_val, double* out)
{
  int n1, n2;
  if(DIM == 2)
  {
    n1 = edge;
    n2 = (edge == 3) ? 0 : edge + 1;
  }
  else
  {
    // hex_edge_table (:766-770): base 0-1 1-2 2-3 3-0, top 4-5 5-6 6-7 7-4, vertical 0-4 1-5 2-6 3-7
    if(edge < 8)
    {
      n1 = edge;
      n2 = (edge & 4) | ((edge + 1) & 3);
    }
    else
    {
      n1 = edge - 8;
      n2 = edge - 4;
    }
  }
  const double f1 = cv[n1], f2 = cv[n2];
  const double* p1 = cc[n1];
  const double* p2 = cc[n2];
  // isNearlyEqual(a, b) = abs(a - b) <= 1e-8 (core/utilities/Utilities.hpp:317-321)
  if(fabs(contour_val - f1) <= 1.0e-8 || fabs(f1 - f2) <= 1.0e-8)
  {
#pragma unroll
    for(int d = 0; d < DIM; ++d) out[d] = p1[d];
    return;
  }
  if(fabs(contour_val - f2) <= 1.0e-8)
  {
#pragma unroll
    for(int d = 0; d < DIM; ++d) out[d] = p2[d];
    return;
  }
  const double df = f2 - f1 + 1.0e-50;  // PRIMAL_TINY (primal/constants.hpp)
  const double w = (contour_val - f1) / df;
#pragma unroll
  for(int d = 0; d < DIM; ++d) out[d] = p1[d] + w * (p2[d] - p1[d]);
}

template <int DIM>
__global__ void __launch_bounds__(kTileThreads) emit_kernel(DomainView<DIM> v, double contour_val, const uint8_t* __restrict__ case_ids,
                                                            const int32_t* __restrict__ tile_offsets, int32_t facet_index_offset,
                                                            int32_t domain_id, int32_t* __restrict__ facet_node_ids,
                                                            double* __restrict__ facet_node_coords, int32_t* __restrict__ facet_parent_ids,
                                                            int32_t* __restrict__ facet_domain_ids)
{
  constexpr int NCORNER = DIM == 2 ? 4 : 8;
  const int32_t tile_first = tile_offsets[blockIdx.x];
  if(tile_offsets[blockIdx.x + 1] == tile_first) return;  // no crossing in this tile (the common case)

  __shared__ int s_warp[kTileThreads / 32];
  const uint32_t cell0 = blockIdx.x * (uint32_t)kTileCells + threadIdx.x * kCellsPerThread;  // 4 consecutive cells per thread
  const uchar4 cs4 = *reinterpret_cast<const uchar4*>(case_ids + cell0);
  const int cs[4] = {cs4.x, cs4.y, cs4.z, cs4.w};
  int cnt[4], mine = 0;
#pragma unroll
  for(int r = 0; r < 4; ++r)
  {
    cnt[r] = used_entries<DIM>(cs[r]) / DIM;
    mine += cnt[r];
  }
  // block-wide exclusive scan of `mine` in thread order = flat cell order
  int incl = mine;
#pragma unroll
  for(int o = 1; o < 32; o <<= 1)
  {
    const int u = __shfl_up_sync(0xffffffffu, incl, o);
    if((threadIdx.x & 31) >= o) incl += u;
  }
  if((threadIdx.x & 31) == 31) s_warp[threadIdx.x >> 5] = incl;
  __syncthreads();
  int before = 0;
  for(int w = 0; w < (int)(threadIdx.x >> 5); ++w) before += s_warp[w];
  int facet = facet_index_offset + tile_first + before + incl - mine;  // firstFacetId (:599)

#pragma unroll 1
  for(int r = 0; r < 4; ++r)
  {
    if(cnt[r] == 0) continue;
    const uint32_t parent = cell0 + r;
    uint32_t idx[DIM];
    to_multi_index<DIM>(v, parent, idx);
    long long fo = 0, co = 0;
#pragma unroll
    for(int d = 0; d < DIM; ++d)
    {
      fo += (long long)idx[d] * v.fcn_stride[d];
      co += (long long)idx[d] * v.coords_stride[d];
    }
    double cc[NCORNER][DIM], cv[NCORNER];
#pragma unroll
    for(int c = 0; c < NCORNER; ++c)
    {
      int o[DIM];
      corner_offset<DIM>(c, o);
      long long f = fo, x = co;
#pragma unroll
      for(int d = 0; d < DIM; ++d)
      {
        f += o[d] ? v.fcn_stride[d] : 0;
        x += o[d] ? v.coords_stride[d] : 0;
      }
      cv[c] = __ldg(v.fcn + f);
#pragma unroll
      for(int d = 0; d < DIM; ++d) cc[c][d] = __ldg(v.coords[d] + x);
    }
    const uint64_t word = DIM == 2 ? (uint64_t)c_cases2d[cs[r]] : c_cases3d[cs[r]];
    for(int f = 0; f < cnt[r]; ++f, ++facet)
    {
      facet_parent_ids[facet] = (int32_t)parent;
      facet_domain_ids[facet] = domain_id;  // m_facetDomainIds.fill (MarchingCubes.cpp:140-146)
#pragma unroll
      for(int d = 0; d < DIM; ++d)
      {
        const int corner = facet * DIM + d;
        facet_node_ids[corner] = corner;
        const int edge = (int)((word >> (4 * (f * DIM + d))) & 0xF);  // cases_table(caseId, fId * DIM + d)
        double p[DIM];
        linear_interp<DIM>(edge, cc, cv, contour_val, p);
#pragma unroll
        for(int k = 0; k < DIM; ++k) facet_node_coords[(long long)corner * DIM + k] = p[k];
      }
    }
  }
}

}  // namespace mc
}  // namespace axb
