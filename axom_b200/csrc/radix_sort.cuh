// radix_sort.cuh -- hand-written device LSD radix sort (Onesweep: one global histogram
// pass + one chained-scan scatter pass per 8-bit digit, decoupled look-back between tiles).
//
// Replaces RAJA::stable_sort_pairs -> cub::DeviceRadixSort::SortPairs on the reference's
// GPU path (spin/internal/linear_bvh/build_radix_tree.hpp:225-238).  Keys are 64-bit
// (code32 << 32) | original_index: the low word is already ascending on input and every
// pass is stable, so sorting only the digits of the high word reproduces the reference's
// *stable* (code, index) order bit for bit -- and the sorted low words are the permutation
// (RadixTree::m_leafs).
//
// HBM traffic per key: 8 B (histogram read) + passes * 16 B.  3-D: 30 code bits -> 4 passes
// -> 72 B/key.
#pragma once
#include "common.cuh"

namespace axb
{
namespace rsort
{
constexpr int RADIX_BITS = 8;
constexpr int RADIX = 1 << RADIX_BITS;
constexpr int BLOCK = 256;
constexpr int WARPS = BLOCK / 32;
constexpr int ITEMS = 16;  // keys per thread
constexpr int TILE = BLOCK * ITEMS;
constexpr int MAX_PASSES = 4;

constexpr uint32_t FLAG_NONE = 0u;
constexpr uint32_t FLAG_AGG = 1u << 30;
constexpr uint32_t FLAG_PREFIX = 2u << 30;
constexpr uint32_t FLAG_MASK = 3u << 30;
constexpr uint32_t VALUE_MASK = ~FLAG_MASK;

inline int num_tiles(long long n) { return (int)((n + TILE - 1) / TILE); }

// bytes of scratch (lookback words + tile counters + histograms) for n keys
inline size_t scratch_bytes(long long n)
{
  const size_t tiles = (size_t)num_tiles(n);
  return sizeof(uint32_t) * (MAX_PASSES * RADIX            // global histograms
                             + MAX_PASSES                  // dynamic tile counters
                             + MAX_PASSES * tiles * RADIX  // look-back status words
                             + 64);
}

// Global digit histograms of all passes in one read of the keys.  Fused into the Morton
// kernel for the build (see build.cu); this standalone version serves other key sources.
__global__ void __launch_bounds__(BLOCK) histogram_kernel(const unsigned long long* __restrict__ keys, long long n, int first_bit,
                                                           int passes, uint32_t* __restrict__ ghist)
{
  __shared__ uint32_t sh[MAX_PASSES * RADIX];
  for(int i = threadIdx.x; i < MAX_PASSES * RADIX; i += BLOCK) sh[i] = 0;
  __syncthreads();
  for(long long i = (long long)blockIdx.x * BLOCK + threadIdx.x; i < n; i += (long long)gridDim.x * BLOCK)
  {
    const unsigned long long k = keys[i];
    for(int p = 0; p < passes; ++p) atomicAdd(&sh[p * RADIX + (unsigned)((k >> (first_bit + p * RADIX_BITS)) & (RADIX - 1))], 1u);
  }
  __syncthreads();
  for(int i = threadIdx.x; i < passes * RADIX; i += BLOCK)
    if(sh[i]) atomicAdd(&ghist[i], sh[i]);
}

// One digit pass.  grid = num_tiles(n) blocks; tiles are claimed through a global counter so
// that a block only ever waits on tiles that are already running (forward progress).
__global__ void __launch_bounds__(BLOCK) onesweep_kernel(const unsigned long long* __restrict__ in, unsigned long long* __restrict__ out,
                                                          long long n, int shift, const uint32_t* __restrict__ ghist, /* RADIX counts */
                                                          volatile uint32_t* lookback, /* tiles*RADIX, zeroed */
                                                          uint32_t* tile_counter)
{
  __shared__ unsigned long long skeys[TILE];  // 32 KB staging for coalesced scatter
  __shared__ uint32_t whist[WARPS][RADIX];    // per-warp digit counts -> per-warp digit offsets
  __shared__ uint32_t bin_start[RADIX];       // first staged slot of each digit in this tile
  __shared__ long long gbase[RADIX];          // global position of staged slot 0 of each digit, minus bin_start
  __shared__ uint32_t scan_tmp[RADIX];
  __shared__ uint32_t s_tile;

  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int lane = tid & 31;

  if(tid == 0) s_tile = atomicAdd(tile_counter, 1u);
#pragma unroll
  for(int w = 0; w < WARPS; ++w) whist[w][tid] = 0;
  __syncthreads();
  const uint32_t tile = s_tile;
  const long long tile_base = (long long)tile * TILE;

  // ---- load: warp-striped, so (item, lane) lexicographic order == memory order ----
  unsigned long long key[ITEMS];
  const long long wbase = tile_base + (long long)warp * (32 * ITEMS);
#pragma unroll
  for(int i = 0; i < ITEMS; ++i)
  {
    const long long idx = wbase + i * 32 + lane;
    key[i] = (idx < n) ? __ldcs(in + idx) : ~0ull;  // sentinel sorts last inside the tile
  }

  // ---- stable intra-warp ranking with match.any ----
  uint32_t rank[ITEMS];
#pragma unroll
  for(int i = 0; i < ITEMS; ++i)
  {
    const uint32_t digit = (uint32_t)(key[i] >> shift) & (RADIX - 1);
    const unsigned peers = __match_any_sync(0xffffffffu, digit);
    const uint32_t pre = whist[warp][digit];
    __syncwarp();
    const int leader = 31 - __clz(peers);
    if(lane == leader) whist[warp][digit] = pre + __popc(peers);
    __syncwarp();
    rank[i] = pre + __popc(peers & ((1u << lane) - 1u));
  }
  __syncthreads();

  // ---- per-digit: exclusive scan over warps, tile total (thread == digit) ----
  uint32_t tile_count = 0;
#pragma unroll
  for(int w = 0; w < WARPS; ++w)
  {
    const uint32_t c = whist[w][tid];
    whist[w][tid] = tile_count;
    tile_count += c;
  }
  // publish this tile's aggregate as early as possible
  lookback[(size_t)tile * RADIX + tid] = (tile == 0 ? FLAG_PREFIX : FLAG_AGG) | tile_count;

  // ---- block exclusive scans over digits: staged slot of each digit, global digit start ----
  // (two 256-wide Hillis-Steele scans in shared memory)
  uint32_t gcount = ghist[tid];
  scan_tmp[tid] = tile_count;
  __syncthreads();
  uint32_t incl = tile_count;
  for(int off = 1; off < RADIX; off <<= 1)
  {
    const uint32_t v = (tid >= off) ? scan_tmp[tid - off] : 0u;
    __syncthreads();
    incl += v;
    scan_tmp[tid] = incl;
    __syncthreads();
  }
  bin_start[tid] = incl - tile_count;
  __syncthreads();
  scan_tmp[tid] = gcount;
  __syncthreads();
  uint32_t gincl = gcount;
  for(int off = 1; off < RADIX; off <<= 1)
  {
    const uint32_t v = (tid >= off) ? scan_tmp[tid - off] : 0u;
    __syncthreads();
    gincl += v;
    scan_tmp[tid] = gincl;
    __syncthreads();
  }
  const uint32_t gstart = gincl - gcount;

  // ---- decoupled look-back for digit `tid` ----
  uint32_t excl = 0;
  if(tile > 0)
  {
    long long t = (long long)tile - 1;
    while(true)
    {
      uint32_t s;
      do
      {
        s = lookback[(size_t)t * RADIX + tid];
      } while((s & FLAG_MASK) == FLAG_NONE);
      excl += s & VALUE_MASK;
      if((s & FLAG_MASK) == FLAG_PREFIX) break;
      --t;
    }
    lookback[(size_t)tile * RADIX + tid] = FLAG_PREFIX | (excl + tile_count);
  }
  gbase[tid] = (long long)gstart + (long long)excl - (long long)bin_start[tid];
  __syncthreads();

  // ---- scatter into the staging buffer by (digit, warp, rank) ----
#pragma unroll
  for(int i = 0; i < ITEMS; ++i)
  {
    const uint32_t digit = (uint32_t)(key[i] >> shift) & (RADIX - 1);
    skeys[bin_start[digit] + whist[warp][digit] + rank[i]] = key[i];
  }
  __syncthreads();

  // ---- coalesced write-out: consecutive threads -> consecutive slots -> runs per digit ----
  const long long remaining = n - tile_base;
  const int valid = remaining < TILE ? (int)remaining : TILE;
#pragma unroll
  for(int i = 0; i < ITEMS; ++i)
  {
    const int slot = i * BLOCK + tid;
    if(slot < valid)
    {
      const unsigned long long k = skeys[slot];
      const uint32_t digit = (uint32_t)(k >> shift) & (RADIX - 1);
      out[gbase[digit] + slot] = k;
    }
  }
}

}  // namespace rsort
}  // namespace axb
