// mc.cuh -- quest::MarchingCubes on the device (the consumer of the 256^3 distance field).
//
// Reference: quest/detail/MarchingCubesImpl.hpp (markCrossings :157-362, scanCrossings :364-573, computeFacets
// :575-808) driven by quest/MarchingCubes.cpp:107-147.  The reference runs, per domain, a case-id pass, a flag pass, a
// scan, a compaction pass, a second scan and the facet pass, keeping 2 + 4 + 4 + 4 bytes per CELL of scratch and
// 2 + 4 + 4 + 4 bytes per crossing.  Here:
//
//   mark_rows_kernel    one read of the nodal function (8 B / node, the compulsory traffic): case id per cell (1 byte, kept)
//   count_scan_kernel   facet count of each 1024-cell tile from the case bytes; the last block to finish then runs the
//                       exclusive scan of the tile counts (16 K tiles for 255^3 cells), the domain total and the ordered
//                       list of tiles that hold facets
//   emit_kernel         one block per listed tile: the facets are listed in shared memory and their corners spread over
//                       all threads (one output node per thread)
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

// in global memory, read through the read-only path: the index differs per thread (a divergent index serialises on the
// constant bank, L1 serves it in one pass per cache line)
__device__ const uint64_t g_cases3d[256] = {AXB_MC_CASES3D_WORDS};
__device__ const uint16_t g_cases2d[16] = {AXB_MC_CASES2D_WORDS};

template <int DIM>
__device__ __forceinline__ uint64_t case_word(int case_id)
{
  return DIM == 2 ? (uint64_t)__ldg(&g_cases2d[case_id]) : __ldg(&g_cases3d[case_id]);
}

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
  uint32_t fast_extent;      // cells along the fastest direction slowest[DIM-1]: consecutive flat ids are neighbours along it
};

// MDMapping::toMultiIndex (core/MDMapping.hpp:361-371)
template <int DIM>
__device__ __forceinline__ void to_multi_index(const DomainView<DIM>& v, uint32_t flat, uint32_t idx[DIM])
{
#pragma unroll
  for(int s = 0; s < DIM; ++s)
  {
    const int dir = v.slowest[s];
    const FastDiv f = v.case_div[dir];  // kernel-parameter (constant bank) indexing
    const uint32_t q = fastdiv(flat, f);
#pragma unroll
    for(int d = 0; d < DIM; ++d) idx[d] = (d == dir) ? q : idx[d];
    flat -= q * f.d;
  }
}

// corner c of a cell as a node offset from the cell's (i,j[,k]) node: MarchingCubesImpl.hpp:329-333 (2-D), :349-357 (3-D)
template <int DIM>
__host__ __device__ __forceinline__ void corner_offset(int c, int o[DIM])
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

// inverse of corner_offset
template <int DIM>
__host__ __device__ __forceinline__ int corner_id(const int o[DIM])
{
  if(DIM == 2) return o[1] ? (o[0] ? 2 : 3) : (o[0] ? 1 : 0);
  return (o[0] ? (o[1] ? 1 : 0) : (o[1] ? 2 : 3)) + 4 * o[DIM - 1];
}

template <int DIM>
__device__ __forceinline__ int used_entries(int case_id)
{
  // entries are packed from nibble 0 and terminated by 0xF nibbles
  if(DIM == 2)
  {
    const uint32_t w = (uint32_t)case_word<2>(case_id);
    const uint32_t m = w & (w >> 1) & (w >> 2) & (w >> 3) & 0x1111u;
    return m ? (__ffs(m) - 1) >> 2 : 4;
  }
  const uint64_t w = case_word<3>(case_id);
  const uint64_t m = w & (w >> 1) & (w >> 2) & (w >> 3) & 0x1111111111111111ull;
  return (__ffsll((long long)m) - 1) >> 2;  // nibble 15 is never used, so m != 0
}

//------------------------------------------------------------------------------------------
// pass 1: computeCaseId (:306-361) for every cell + per-tile facet count (num_contour_cells, :815-843)
//------------------------------------------------------------------------------------------
// The plain form: one cell per thread-iteration, every lane loads all its corners.  Kept as the A/B reference of the row
// kernel below (AXB_MC_MARK_PLAIN=1); ~100 instructions per cell, issue-bound.
template <int DIM>
__global__ void __launch_bounds__(kTileThreads) mark_plain_kernel(DomainView<DIM> v, double contour_val, int mask_val,
                                                                  uint8_t* __restrict__ case_ids)
{
  constexpr int NCORNER = DIM == 2 ? 4 : 8;
  const uint32_t base = blockIdx.x * (uint32_t)kTileCells;
#pragma unroll
  for(int r = 0; r < kCellsPerThread; ++r)
  {
    const uint32_t n = base + r * kTileThreads + threadIdx.x;
    int case_id = 0;  // m_caseIdsFlat.fill(0) (:162)
    if(n < v.num_cells)
    {
      uint32_t idx[DIM] = {};
      to_multi_index<DIM>(v, n, idx);
      bool use_zone = true;
      if(v.mask)
      {
        long long mo = 0;
#pragma unroll
        for(int d = 0; d < DIM; ++d) mo += (long long)idx[d] * v.mask_stride[d];
        use_zone = (__ldg(v.mask + mo) == mask_val);  // :325 / :345
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
    case_ids[n] = (uint8_t)case_id;  // the tail of the last tile is written too (case 0): later passes read whole tiles
  }
}

// The row kernel (default).  Directions are named by speed in the case-id / function layout: F fastest, M middle, S slowest
// (2-D: F, M).  A warp takes a unit of 31 cells along F x kRows cells along M x kPlanes cells along S: lane l evaluates
// (value >= contour_val) ONCE for the (kPlanes + 1) * (kRows + 1) nodes of its node column f = 31 * chunk + l (27 loads for
// 16 cells instead of 128), packs the bits into one word and hands them to lane l - 1 with a single shuffle; lane 31 only supplies
// bits.  The 8 bits of a cell (4 own, 4 from the right-hand neighbour) index a 256-byte table, built on the host from the
// direction permutation, that yields the reference's case id (corner numbering of :349-357).  Warps walk the units
// grid-stride, so the table is staged in shared memory once per block.
constexpr int kRows = 8;
#ifndef AXB_MC_PLANES
  #define AXB_MC_PLANES 2
#endif
constexpr int kPlanes = AXB_MC_PLANES;  // cell planes per unit along S (3-D): kPlanes + 1 node planes are read once for kPlanes cell planes
constexpr int kUnitCells = 31;

template <int DIM>
struct RowView
{
  const double* fcn;
  const int32_t* mask;     // nullptr: no mask
  long long fs[3];         // function strides along F, M, S
  long long ms[3];         // mask strides along F, M, S
  uint32_t cs_m, cs_s;     // case-id strides along M, S (F has stride 1)
  uint32_t nf, nm, ns;     // cells along F, M, S (ns = 1 in 2-D)
  FastDiv chunks, mgroups; // units per row / row groups per S
  uint32_t num_units;
  const uint8_t* lut;      // compact corner bits -> case id
};

// one unit; FULL: all kRows rows and all kPlanes cell planes of the unit exist (no bounds checks), MASK: a cell mask is present
template <int DIM, bool MASK, bool FULL>
__device__ __forceinline__ void mark_rows_unit(const RowView<DIM>& v, const uint8_t* s_lut, double contour_val, int mask_val, uint32_t lane,
                                               uint32_t f, uint32_t m0, uint32_t s, uint8_t* __restrict__ case_ids)
{
  constexpr int SP = DIM == 3 ? kPlanes : 1;   // cell planes of the unit along S
  constexpr int NSN = DIM == 3 ? SP + 1 : 1;   // node planes of the unit along S
  constexpr int CB = DIM == 3 ? 4 : 2;         // corner bits a cell takes from one node column
  // (value >= contour_val) for nodes (f, m0 + mm, s + ss): bit mm * NSN + ss.  All (kRows + 1) * NSN loads are issued before
  // the first comparison, so a warp keeps several KB in flight per unit (the kernel is latency-bound otherwise).
  uint32_t bits = 0;
  if(f <= v.nf)
  {
    const double* p = v.fcn + (long long)f * v.fs[0] + (long long)m0 * v.fs[1] + (DIM == 3 ? (long long)s * v.fs[2] : 0);
    double a[kRows + 1][NSN];
#pragma unroll
    for(int mm = 0; mm <= kRows; ++mm)
    {
      const bool row = FULL || m0 + mm <= v.nm;  // a missing node row / plane is never used by an existing cell
#pragma unroll
      for(int ss = 0; ss < NSN; ++ss) a[mm][ss] = (row && (FULL || ss == 0 || s + ss <= v.ns)) ? __ldg(p + ss * v.fs[2]) : 0.0;
      p += v.fs[1];
    }
#pragma unroll
    for(int mm = 0; mm <= kRows; ++mm)
#pragma unroll
      for(int ss = 0; ss < NSN; ++ss)
        if(a[mm][ss] >= contour_val) bits |= 1u << (mm * NSN + ss);  // computeCrossingCase (:307-319)
  }
  const uint32_t right = __shfl_down_sync(0xffffffffu, bits, 1);
  if(lane < kUnitCells && f < v.nf)
  {
#pragma unroll
    for(int cs = 0; cs < SP; ++cs)
    {
      if(FULL || s + cs < v.ns)
      {
        uint8_t* out = case_ids + (f + m0 * v.cs_m + (DIM == 3 ? (s + cs) * v.cs_s : 0));
        const int32_t* mp =
          MASK ? v.mask + (long long)f * v.ms[0] + (long long)m0 * v.ms[1] + (DIM == 3 ? (long long)(s + cs) * v.ms[2] : 0) : nullptr;
#pragma unroll
        for(int r = 0; r < kRows; ++r)
        {
          if(FULL || m0 + r < v.nm)
          {
            // the cell's nodes in this column: (r, cs), (r, cs + 1), (r + 1, cs), (r + 1, cs + 1) -> 4 compact bits (2 in 2-D)
            const uint32_t o = bits >> (r * NSN + cs), q = right >> (r * NSN + cs);
            const uint32_t own = DIM == 3 ? ((o & 3u) | ((o >> (NSN - 2)) & 0xCu)) : (o & 3u);
            const uint32_t rgt = DIM == 3 ? ((q & 3u) | ((q >> (NSN - 2)) & 0xCu)) : (q & 3u);
            int case_id = s_lut[own | (rgt << CB)];
            if(MASK && __ldg(mp) != mask_val) case_id = 0;  // :325 / :345 (m_caseIdsFlat.fill(0), :162)
            *out = (uint8_t)case_id;
          }
          out += v.cs_m;
          if(MASK) mp += v.ms[1];
        }
      }
    }
  }
}

template <int DIM, bool MASK>
__global__ void __launch_bounds__(kTileThreads) mark_rows_kernel(RowView<DIM> v, double contour_val, int mask_val,
                                                                 uint8_t* __restrict__ case_ids)
{
  constexpr uint32_t SP = DIM == 3 ? kPlanes : 1;
  __shared__ uint8_t s_lut[256];
  if(threadIdx.x < 64) reinterpret_cast<uint32_t*>(s_lut)[threadIdx.x] = __ldg(reinterpret_cast<const uint32_t*>(v.lut) + threadIdx.x);
  __syncthreads();
  const uint32_t lane = threadIdx.x & 31;
  const uint32_t warps = (gridDim.x * blockDim.x) >> 5;
  for(uint32_t u = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; u < v.num_units; u += warps)
  {
    const uint32_t t = fastdiv(u, v.chunks);
    const uint32_t chunk = u - t * v.chunks.d;
    const uint32_t sg = fastdiv(t, v.mgroups);
    const uint32_t m0 = (t - sg * v.mgroups.d) * kRows;
    const uint32_t s = sg * SP;                    // first cell plane of the unit
    const uint32_t f = chunk * kUnitCells + lane;  // node column
    if(m0 + kRows <= v.nm && s + SP <= v.ns)       // warp-uniform
      mark_rows_unit<DIM, MASK, true>(v, s_lut, contour_val, mask_val, lane, f, m0, s, case_ids);
    else
      mark_rows_unit<DIM, MASK, false>(v, s_lut, contour_val, mask_val, lane, f, m0, s, case_ids);
  }
}

//------------------------------------------------------------------------------------------
// pass 2: tile counts + their exclusive scan, in place, + total (the two inclusive scans of :413-483 collapse into this
// one because crossing ids are never an output)
//------------------------------------------------------------------------------------------
// One block walks the tile counts in chunks of 16384 (four int4 per thread, issued together), carrying the running
// totals: 255^3 cells = 16193 tiles = one chunk.  It also lists the tiles that hold facets, in order, so that pass 3
// launches one block per ACTIVE tile only.  totals[0] = facets of the domain, totals[1] = active tiles.
constexpr int kScanThreads = 1024;
constexpr int kScanPerThread = 16;
__device__ __forceinline__ void scan_tiles_block(int32_t* tile_facets, int num_tiles, long long* __restrict__ totals,
                                                 int32_t* __restrict__ active_tiles)
{
  __shared__ unsigned long long s_warp[kScanThreads / 32];
  __shared__ unsigned long long s_total;
  long long carry = 0;  // facets before this chunk
  int carry_active = 0;
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for(int base = 0; base < num_tiles; base += kScanThreads * kScanPerThread)
  {
    const int i0 = base + (int)threadIdx.x * kScanPerThread;
    int c[kScanPerThread];
    if(i0 + kScanPerThread <= num_tiles)
    {
#pragma unroll
      for(int k = 0; k < kScanPerThread / 4; ++k)
      {
        const int4 q = __ldcg(reinterpret_cast<const int4*>(tile_facets + i0 + 4 * k));  // written by other blocks: L2, not L1
        c[4 * k + 0] = q.x;
        c[4 * k + 1] = q.y;
        c[4 * k + 2] = q.z;
        c[4 * k + 3] = q.w;
      }
    }
    else
    {
#pragma unroll
      for(int k = 0; k < kScanPerThread; ++k) c[k] = (i0 + k < num_tiles) ? __ldcg(tile_facets + i0 + k) : 0;
    }
    int sum = 0, act = 0;
#pragma unroll
    for(int k = 0; k < kScanPerThread; ++k)
    {
      sum += c[k];
      act += (c[k] > 0);
    }
    // one scan for both: a chunk holds < 2^27 facets (16384 tiles x 5120) and <= 16384 active tiles
    const unsigned long long mine = ((unsigned long long)act << 40) | (unsigned long long)sum;
    unsigned long long incl = mine;
#pragma unroll
    for(int o = 1; o < 32; o <<= 1)
    {
      const unsigned long long u = __shfl_up_sync(0xffffffffu, incl, o);
      if((int)lane >= o) incl += u;
    }
    if(lane == 31) s_warp[warp] = incl;
    __syncthreads();
    if(warp == 0)
    {
      unsigned long long w = s_warp[lane];
#pragma unroll
      for(int o = 1; o < 32; o <<= 1)
      {
        const unsigned long long u = __shfl_up_sync(0xffffffffu, w, o);
        if((int)lane >= o) w += u;
      }
      s_warp[lane] = w;  // inclusive totals of warps 0 .. lane
      if(lane == 31) s_total = w;
    }
    __syncthreads();
    const unsigned long long excl = (warp ? s_warp[warp - 1] : 0ull) + incl - mine;
    long long run = carry + (long long)(excl & ((1ull << 40) - 1));
    int slot = carry_active + (int)(excl >> 40);
#pragma unroll
    for(int k = 0; k < kScanPerThread; ++k)
      if(i0 + k < num_tiles)
      {
        // offsets are 32-bit like the reference's IndexType; the host rejects totals that do not fit before pass 3 runs
        tile_facets[i0 + k] = (int32_t)run;
        run += c[k];
        if(c[k] > 0) active_tiles[slot++] = i0 + k;
      }
    const unsigned long long tot = s_total;
    carry += (long long)(tot & ((1ull << 40) - 1));
    carry_active += (int)(tot >> 40);
    __syncthreads();  // s_warp / s_total are rewritten by the next chunk
  }
  if(threadIdx.x == 0)
  {
    tile_facets[num_tiles] = (int32_t)carry;
    totals[0] = carry;
    totals[1] = carry_active;
  }
}

// pass 2 (one launch): per-tile facet count from the case ids (num_contour_cells, :815-843; one warp per tile, warps walk
// the tiles grid-stride; the tail of the last tile is zeroed), then the LAST block to finish -- a ticket taken after a
// __threadfence(), the classic single-pass reduction hand-over -- runs the exclusive scan of the tile counts below.
template <int DIM>
__global__ void __launch_bounds__(kScanThreads) count_scan_kernel(uint8_t* __restrict__ case_ids, uint32_t num_cells, uint32_t num_tiles,
                                                                  int32_t* tile_facets, long long* __restrict__ totals,
                                                                  int32_t* __restrict__ active_tiles, unsigned int* __restrict__ ticket)
{
  constexpr int NCASE = DIM == 2 ? 16 : 256;
  __shared__ uint8_t s_nfacets[NCASE];
  __shared__ bool s_last;
  if(threadIdx.x < NCASE) s_nfacets[threadIdx.x] = (uint8_t)(used_entries<DIM>(threadIdx.x) / DIM);
  __syncthreads();
  const uint32_t lane = threadIdx.x & 31;
  const uint32_t warps = (gridDim.x * blockDim.x) >> 5;
  for(uint32_t tile = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; tile < num_tiles; tile += warps)
  {
    const uint32_t tile0 = tile * (uint32_t)kTileCells;
    int nf = 0;
#pragma unroll
    for(int k = 0; k < kTileCells / 128; ++k)
    {
      const uint32_t cell0 = tile0 + (k * 32 + lane) * 4;  // coalesced 128 B per warp load
      uchar4 c = *reinterpret_cast<const uchar4*>(case_ids + cell0);
      if(cell0 + 3 >= num_cells)  // ragged tail: cells past the end count (and are later read) as case 0
      {
        if(cell0 + 0 >= num_cells) c.x = 0;
        if(cell0 + 1 >= num_cells) c.y = 0;
        if(cell0 + 2 >= num_cells) c.z = 0;
        c.w = 0;
        *reinterpret_cast<uchar4*>(case_ids + cell0) = c;
      }
      // away from the contour the four cells are all "outside" (0) or all "inside" (2^corners - 1): no facets, no look-up
      const uint32_t w = (uint32_t)c.x | ((uint32_t)c.y << 8) | ((uint32_t)c.z << 16) | ((uint32_t)c.w << 24);
      constexpr uint32_t kAllInside = DIM == 2 ? 0x0F0F0F0Fu : 0xFFFFFFFFu;
      if(w != 0u && w != kAllInside) nf += s_nfacets[c.x] + s_nfacets[c.y] + s_nfacets[c.z] + s_nfacets[c.w];
    }
#pragma unroll
    for(int o = 16; o > 0; o >>= 1) nf += __shfl_xor_sync(0xffffffffu, nf, o);
    if(lane == 0) tile_facets[tile] = nf;
  }
  // hand-over: make this block's counts visible, take a ticket; the block that draws the last one scans
  __threadfence();
  __syncthreads();
  if(threadIdx.x == 0)
  {
    const unsigned int t = atomicAdd(ticket, 1u);
    s_last = (t == gridDim.x - 1);
    if(s_last) *ticket = 0;  // ready for the next launch
  }
  __syncthreads();
  if(!s_last) return;
  __threadfence();
  scan_tiles_block(tile_facets, (int)num_tiles, totals, active_tiles);
}

//------------------------------------------------------------------------------------------
// pass 3: computeFacets (:575-620) -- get_corner_coords_and_values (:644-703) + linear_interp (:706-807)
//------------------------------------------------------------------------------------------
// the two corners of cell edge `edge`: 2-D (:718-719), 3-D hex_edge_table (:766-770): base 0-1 1-2 2-3 3-0, top 4-5 5-6 6-7
// 7-4, vertical 0-4 1-5 2-6 3-7
template <int DIM>
__device__ __forceinline__ void edge_corners(int edge, int& n1, int& n2)
{
  if(DIM == 2)
  {
    n1 = edge;
    n2 = (edge == 3) ? 0 : edge + 1;
  }
  else if(edge < 8)
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

// linear_interp (:706-807) on the edge's two end nodes
template <int DIM>
__device__ __forceinline__ void linear_interp(double f1, double f2, const double p1[DIM], const double p2[DIM], double contour_val,
                                              double out[DIM])
{
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

// One block per ACTIVE 1024-cell tile (the list pass 2 made; tiles without facets are never launched).  A block (a) scans its case bytes and lists
// its facets as (cell in tile, facet in cell) in shared memory, in output order, then (b) spread the facet CORNERS over
// all threads: a thread owns one output node, reads only the two end nodes of its edge (2 values + 2 * DIM coordinates
// instead of the cell's 8 * (1 + DIM)), interpolates and writes DIM consecutive doubles -- consecutive threads write
// consecutive nodes.  The reference walks crossing cells instead (:590-619); the arithmetic per node is the same.
template <int DIM>
__global__ void __launch_bounds__(kTileThreads) emit_kernel(DomainView<DIM> v, double contour_val, const uint8_t* __restrict__ case_ids,
                                                            const int32_t* __restrict__ tile_offsets, const int32_t* __restrict__ active_tiles,
                                                            int32_t facet_index_offset, int32_t domain_id, int32_t* __restrict__ facet_node_ids,
                                                            double* __restrict__ facet_node_coords, int32_t* __restrict__ facet_parent_ids,
                                                            int32_t* __restrict__ facet_domain_ids)
{
  constexpr int MAXF = DIM == 2 ? 2 : 5;  // facets per cell
  const uint32_t tile = (uint32_t)active_tiles[blockIdx.x];
  const int32_t tile_first = tile_offsets[tile];
  const int tile_total = tile_offsets[tile + 1] - tile_first;
  if(tile_total == 0) return;  // cannot happen for a listed tile

  __shared__ int s_warp[kTileThreads / 32];
  __shared__ uint8_t s_case[kTileCells];
  __shared__ uint16_t s_cell[kTileCells * MAXF];  // facet (in tile order) -> cell within the tile
  __shared__ uint8_t s_fid[kTileCells * MAXF];    //                      -> facet within the cell
  const uint32_t tile0 = tile * (uint32_t)kTileCells;
  const uint32_t local0 = threadIdx.x * kCellsPerThread;  // 4 consecutive cells per thread
  const uchar4 cs4 = *reinterpret_cast<const uchar4*>(case_ids + tile0 + local0);
  *reinterpret_cast<uchar4*>(s_case + local0) = cs4;
  const int cs[4] = {cs4.x, cs4.y, cs4.z, cs4.w};
  int cnt[4], mine = 0;
#pragma unroll
  for(int r = 0; r < 4; ++r)
  {
    cnt[r] = (cs[r] == 0 || cs[r] == (DIM == 2 ? 15 : 255)) ? 0 : used_entries<DIM>(cs[r]) / DIM;
    mine += cnt[r];
  }
  // block-wide exclusive scan of `mine` in thread order = flat cell order
  int incl = mine;
#pragma unroll
  for(int o = 1; o < 32; o <<= 1)
  {
    const int u = __shfl_up_sync(0xffffffffu, incl, o);
    if((int)(threadIdx.x & 31) >= o) incl += u;
  }
  if((threadIdx.x & 31) == 31) s_warp[threadIdx.x >> 5] = incl;
  __syncthreads();
  int pos = incl - mine;
  for(int w = 0; w < (int)(threadIdx.x >> 5); ++w) pos += s_warp[w];
#pragma unroll
  for(int r = 0; r < 4; ++r)
    for(int f = 0; f < cnt[r]; ++f, ++pos)
    {
      s_cell[pos] = (uint16_t)(local0 + r);
      s_fid[pos] = (uint8_t)f;
    }
  __syncthreads();

  const int first = facet_index_offset + tile_first;  // firstFacetId of the tile's first crossing (:599)
  for(int q = threadIdx.x; q < tile_total * DIM; q += kTileThreads)
  {
    const int fl = q / DIM, d = q - fl * DIM;  // facet within the tile, corner within the facet
    const uint32_t local = s_cell[fl];
    const int f = s_fid[fl];
    const uint32_t parent = tile0 + local;
    const int facet = first + fl;
    const int corner = facet * DIM + d;  // newCornerId (:610)
    if(d == 0)
    {
      facet_parent_ids[facet] = (int32_t)parent;
      facet_domain_ids[facet] = domain_id;  // m_facetDomainIds.fill (MarchingCubes.cpp:140-146)
    }
    facet_node_ids[corner] = corner;
    const int edge = (int)((case_word<DIM>(s_case[local]) >> (4 * (f * DIM + d))) & 0xF);  // cases_table(caseId, fId * DIM + d)
    int n1, n2;
    edge_corners<DIM>(edge, n1, n2);
    uint32_t idx[DIM] = {};
    to_multi_index<DIM>(v, parent, idx);
    long long fo = 0, co = 0;
#pragma unroll
    for(int k = 0; k < DIM; ++k)
    {
      fo += (long long)idx[k] * v.fcn_stride[k];
      co += (long long)idx[k] * v.coords_stride[k];
    }
    int o1[DIM], o2[DIM];
    corner_offset<DIM>(n1, o1);
    corner_offset<DIM>(n2, o2);
    long long f1o = fo, f2o = fo, c1o = co, c2o = co;
#pragma unroll
    for(int k = 0; k < DIM; ++k)
    {
      f1o += o1[k] ? v.fcn_stride[k] : 0;
      f2o += o2[k] ? v.fcn_stride[k] : 0;
      c1o += o1[k] ? v.coords_stride[k] : 0;
      c2o += o2[k] ? v.coords_stride[k] : 0;
    }
    const double f1 = __ldg(v.fcn + f1o), f2 = __ldg(v.fcn + f2o);
    double p1[DIM], p2[DIM], out[DIM];
#pragma unroll
    for(int k = 0; k < DIM; ++k)
    {
      p1[k] = __ldg(v.coords[k] + c1o);
      p2[k] = __ldg(v.coords[k] + c2o);
    }
    linear_interp<DIM>(f1, f2, p1, p2, contour_val, out);
#pragma unroll
    for(int k = 0; k < DIM; ++k) facet_node_coords[(long long)corner * DIM + k] = out[k];
  }
}

}  // namespace mc
}  // namespace axb
