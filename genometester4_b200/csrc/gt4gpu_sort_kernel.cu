// gt4gpu_sort_kernel.cu -- the list-building back end: raw canonical words -> sorted (word, count) list.
//
// Replaces the CPU path  wordtable_sort (src/word-table.c, hybridInPlaceRadixSort256 in
// src/utils.c:127-198)  ->  merge_tables_to_file (src/glistmaker.c:1080-1144), i.e. "sort the words
// of a table, then count the run length of every distinct word".
//
//   radix_hist_kernel        one read of the keys: 256-bin histograms of every 8-bit digit that will
//                            be sorted (only ceil(2k/8) digits: words are < 4^k)
//   radix_bins_kernel        exclusive scan of each histogram -> first output slot of every bin
//   radix_onesweep_kernel    one least-significant-digit pass: a tile of 512 x 16 keys is ranked inside
//                            the CTA (eight warp votes per key + per-warp digit counters, stable), the
//                            per-digit tile counts are published and then chained across tiles by a
//                            decoupled look-back (256 chains in parallel, one per thread) while the keys
//                            are regrouped by digit in shared memory; they leave in runs of consecutive
//                            addresses.  8 B read + 8 B written per key and pass; no separate "upsweep"
//                            pass over the data.  <true>: every key drags a 32-bit payload along (the
//                            key's position: the sorting permutation, used by the lookup batches).
//   rle_heads_kernel         sorted keys -> distinct words + index of each run's first element
//                            (warp votes, block scan, decoupled look-back)
//   rle_counts_kernel        count = distance to the next run's first element
//
// Integer work, HBM-bound; no tensor cores.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include "gt4gpu_device.cuh"
#include "gt4gpu_internal.h"

namespace gt4gpu {

namespace {

constexpr int RADIX = 256;
constexpr int SORT_NT = 512;
constexpr int SORT_WARPS = SORT_NT / 32;
#ifndef GT4_SORT_ITEMS
#define GT4_SORT_ITEMS 16
#endif
constexpr int SORT_ITEMS = GT4_SORT_ITEMS;
constexpr int SORT_TILE = SORT_NT * SORT_ITEMS;     // 8192 keys = 64 KiB of shared memory
#ifndef GT4_LB_BATCH
#define GT4_LB_BATCH 4
#endif
constexpr int LB_BATCH = GT4_LB_BATCH;
static_assert (SORT_ITEMS % 2 == 0 && SORT_TILE < 65536, "ranks are packed two per register");

// look-back descriptor: status (2 bits) | pass tag (6 bits) | value (56 bits).  The tag makes the
// descriptors of an earlier pass read as "not ready", so one memset serves all passes of a sort.
constexpr uint64_t ST_PARTIAL = 1ull << 62;
constexpr uint64_t ST_INCLUSIVE = 2ull << 62;
constexpr int TAG_SHIFT = 56;
constexpr uint64_t VALUE_MASK = (1ull << TAG_SHIFT) - 1;

using namespace dev;

// ------------------------------------------------------------------------------------------
// histograms of all digits in one pass over the keys
// ------------------------------------------------------------------------------------------
constexpr int HIST_COPIES = 4;     // privatised per group of warps: fewer same-address shared atomics

__global__ void __launch_bounds__ (512)
radix_hist_kernel (const uint64_t *__restrict__ keys, uint64_t n, int n_pass, unsigned long long *__restrict__ hist, unsigned long long *__restrict__ or_all)
{
  uint64_t seen = 0;       // OR of every key: the caller checks that no bit above the sorted digits is set
  __shared__ uint32_t s_hist[HIST_COPIES][SORT_MAX_PASSES][RADIX];
  for (int i = threadIdx.x; i < HIST_COPIES * SORT_MAX_PASSES * RADIX; i += blockDim.x) (&s_hist[0][0][0])[i] = 0;
  __syncthreads ();
  uint32_t (*mine)[RADIX] = s_hist[(threadIdx.x >> 5) & (HIST_COPIES - 1)];
  // every CTA takes a contiguous chunk; counts per CTA stay far below 2^32
  const uint64_t per_cta = (n + gridDim.x - 1) / gridDim.x;
  const uint64_t lo = per_cta * blockIdx.x;
  const uint64_t hi = lo + per_cta < n ? lo + per_cta : n;
  constexpr int U = 4;         // independent loads in flight per thread
  uint64_t i = lo + threadIdx.x;
  for (; i + (U - 1) * (uint64_t) blockDim.x < hi; i += U * (uint64_t) blockDim.x) {
    uint64_t key[U];
#pragma unroll
    for (int u = 0; u < U; u++) key[u] = keys[i + u * (uint64_t) blockDim.x];
#pragma unroll
    for (int u = 0; u < U; u++) {
      seen |= key[u];
      for (int p = 0; p < n_pass; p++) atomicAdd (&mine[p][(key[u] >> (8 * p)) & 255u], 1u);
    }
  }
  for (; i < hi; i += blockDim.x) {
    const uint64_t key = keys[i];
    seen |= key;
    for (int p = 0; p < n_pass; p++) atomicAdd (&mine[p][(key >> (8 * p)) & 255u], 1u);
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) seen |= __shfl_xor_sync (0xffffffffu, seen, off);
  if ((threadIdx.x & 31) == 0 && seen) atomicOr (or_all, (unsigned long long) seen);
  __syncthreads ();
  for (int i = threadIdx.x; i < n_pass * RADIX; i += blockDim.x) {
    unsigned long long sum = 0;
#pragma unroll
    for (int c = 0; c < HIST_COPIES; c++) sum += (&s_hist[c][0][0])[i];
    if (sum) atomicAdd (&hist[i], sum);
  }
}

// hist[p][d] -> bins[p][d] = number of keys whose digit p is smaller than d
__global__ void __launch_bounds__ (RADIX)
radix_bins_kernel (const unsigned long long *__restrict__ hist, unsigned long long *__restrict__ bins)
{
  __shared__ unsigned long long s_warp[RADIX / 32];
  const int d = threadIdx.x, lane = d & 31, warp = d >> 5;
  const unsigned long long own = hist[blockIdx.x * RADIX + d];
  unsigned long long incl = own;
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    const unsigned long long t = __shfl_up_sync (0xffffffffu, incl, off);
    if (lane >= off) incl += t;
  }
  if (lane == 31) s_warp[warp] = incl;
  __syncthreads ();
  unsigned long long before = 0;
  for (int w = 0; w < warp; w++) before += s_warp[w];
  bins[blockIdx.x * RADIX + d] = before + incl - own;
}

// ------------------------------------------------------------------------------------------
// one least-significant-digit pass
// ------------------------------------------------------------------------------------------
struct SweepArgs {
  const uint64_t *in;
  uint64_t *out;
  uint64_t n;
  uint64_t n_tiles;
  int shift;                            // bit position of the digit
  uint64_t tag;                         // pass number + 1, pre-shifted to TAG_SHIFT
  const unsigned long long *bins;       // [RADIX] of this pass
  uint64_t *desc;                       // [n_tiles][RADIX]
  uint32_t *ticket;                     // one per pass, zeroed
  const uint32_t *vin;                  // key-value passes: payload of every key (NULL in the first pass: the key's index)
  uint32_t *vout;
  int debug;                            // experiments (GT4GPU_DEBUG): bit 0 = skip the look-back (WRONG output)
};

#ifndef GT4_SORT_RANK
#define GT4_SORT_RANK 0     // 0: eight votes per word, 1: shared-memory atomicOr match, 2: match.any
#endif
#ifndef GT4_SORT_MIN_CTAS
#define GT4_SORT_MIN_CTAS 2
#endif
template <bool VALUES>
__global__ void __launch_bounds__ (SORT_NT, GT4_SORT_MIN_CTAS)
radix_onesweep_kernel (const SweepArgs a)
{
  extern __shared__ __align__ (16) unsigned char smem_raw[];
  uint64_t *s_keys = reinterpret_cast<uint64_t *> (smem_raw);                 // SORT_TILE keys
  __shared__ uint32_t s_whist[SORT_WARPS][RADIX];    // per warp: digit counts, then exclusive prefix over the warps
#if GT4_SORT_RANK == 1
  __shared__ uint32_t s_match[SORT_WARPS][RADIX];    // per warp and digit: the lanes that hold it in the current row
#endif
  __shared__ uint32_t s_dbase[RADIX];                // first slot of every digit inside the regrouped tile
  __shared__ uint64_t s_gbase[RADIX];                // global slot of that first slot, minus s_dbase
  __shared__ uint32_t s_scan[RADIX / 32];
  __shared__ uint32_t s_tile;
  __shared__ uint8_t s_digit[VALUES ? SORT_TILE : 1];   // key-value passes: digit of every regrouped slot

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) s_tile = atomicAdd (a.ticket, 1u);   // tiles start in order: a predecessor is always running
  for (int i = tid; i < SORT_WARPS * RADIX; i += SORT_NT) {
    (&s_whist[0][0])[i] = 0;
#if GT4_SORT_RANK == 1
    (&s_match[0][0])[i] = 0;
#endif
  }
  __syncthreads ();
  const uint64_t tile = s_tile;
  const uint64_t base = tile * SORT_TILE;
  const int n_valid = (a.n - base < (uint64_t) SORT_TILE) ? (int) (a.n - base) : SORT_TILE;

  // warp w owns the contiguous slice [w * 32 * ITEMS, (w + 1) * 32 * ITEMS); item j of lane l is element j * 32 + l
  // of it, so loads are coalesced and (item, lane) order is the input order
  uint64_t key[SORT_ITEMS];
  uint32_t rank2[SORT_ITEMS / 2];      // two 16-bit ranks (within warp and digit) per register
#pragma unroll
  for (int j = 0; j < SORT_ITEMS; j++) {
    const int idx = warp * 32 * SORT_ITEMS + j * 32 + lane;
    key[j] = idx < n_valid ? a.in[base + idx] : ~0ull;     // padding sorts behind every real key of the tile
  }
  uint32_t val[VALUES ? SORT_ITEMS : 1];
  if (VALUES) {
#pragma unroll
    for (int j = 0; j < SORT_ITEMS; j++) {
      const int idx = warp * 32 * SORT_ITEMS + j * 32 + lane;
      val[j] = idx < n_valid ? (a.vin ? a.vin[base + idx] : (uint32_t) (base + idx)) : 0u;
    }
  }
  const uint32_t lt_mask = (1u << lane) - 1u;
#pragma unroll
  for (int j = 0; j < SORT_ITEMS; j++) {
    const uint32_t d = (uint32_t) (key[j] >> a.shift) & 255u;
#if GT4_SORT_RANK == 2
    const uint32_t peers = __match_any_sync (0xffffffffu, d);
#elif GT4_SORT_RANK == 1
    // lanes holding the same digit meet in one shared-memory word: every lane ORs its bit in, then reads the word
    atomicOr (&s_match[warp][d], 1u << lane);
    __syncwarp ();
    const uint32_t peers = s_match[warp][d];
    __syncwarp ();
#else
    // lanes holding the same digit, from one ballot per digit bit (match.any is far slower than 8 votes here)
    uint32_t peers = 0xffffffffu;
#pragma unroll
    for (int b = 0; b < 8; b++) {
#ifdef GT4_SORT_VOTE_SELECT
      const bool bit = (d >> b) & 1u;
      const uint32_t vote = __ballot_sync (0xffffffffu, bit);
      peers &= bit ? vote : ~vote;
#else
      // all-ones when the bit is set (signed 1-bit field extract), so peers &= ~(vote ^ ones) is a single LOP3
      int ones;
      asm ("bfe.s32 %0, %1, %2, 1;" : "=r"(ones) : "r"(d), "r"(b));
      const uint32_t vote = __ballot_sync (0xffffffffu, ones != 0);
      peers &= ~(vote ^ (uint32_t) ones);
#endif
    }
#endif
    const int leader = __ffs (peers) - 1;
    uint32_t before = 0;
    if (lane == leader) {
      before = s_whist[warp][d];
      s_whist[warp][d] = before + __popc (peers);
#if GT4_SORT_RANK == 1
      s_match[warp][d] = 0;
#endif
    }
    before = __shfl_sync (0xffffffffu, before, leader);
    const uint32_t r = before + __popc (peers & lt_mask);
    if (j & 1) rank2[j / 2] |= r << 16; else rank2[j / 2] = r;
    __syncwarp ();
  }
  __syncthreads ();

  // thread d < 256: digit d's counts over the warps -> exclusive prefix per warp, tile total
  uint32_t tile_cnt = 0;
  if (tid < RADIX) {
#pragma unroll
    for (int w = 0; w < SORT_WARPS; w++) {
      const uint32_t c = s_whist[w][tid];
      s_whist[w][tid] = tile_cnt;
      tile_cnt += c;
    }
    uint32_t incl = tile_cnt;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      const uint32_t t = __shfl_up_sync (0xffffffffu, incl, off);
      if (lane >= off) incl += t;
    }
    if (lane == 31) s_scan[warp] = incl;
    s_dbase[tid] = incl - tile_cnt;                  // completed below with the totals of the lower warps
  }
  __syncthreads ();
  uint64_t pub = 0;
  uint32_t dbase = 0;
  if (tid < RADIX) {
    uint32_t before = 0;
    for (int w = 0; w < warp; w++) before += s_scan[w];
    dbase = s_dbase[tid] + before;
    s_dbase[tid] = dbase;
    // publish this tile's count of digit `tid` before anything else: successors only need the partial value
    pub = tile_cnt - ((tid == RADIX - 1) ? (uint32_t) (SORT_TILE - n_valid) : 0u);   // padding is not data
    st_relaxed (a.desc + tile * RADIX + tid, ((tile == 0 || (a.debug & 1)) ? ST_INCLUSIVE : ST_PARTIAL) | a.tag | pub);
  }
  __syncthreads ();

  // regroup by digit in shared memory (stable); the predecessors' descriptors arrive meanwhile
#pragma unroll
  for (int j = 0; j < SORT_ITEMS; j++) {
    const uint32_t d = (uint32_t) (key[j] >> a.shift) & 255u;
    const uint32_t r = (j & 1) ? rank2[j / 2] >> 16 : rank2[j / 2] & 0xffffu;
    s_keys[s_dbase[d] + s_whist[warp][d] + r] = key[j];
  }

  if (tid < RADIX) {
    // chain this digit's count over the tiles
    uint64_t excl = 0;
    if (tile != 0 && !(a.debug & 1)) {
      // walk back over the predecessors LB_BATCH descriptors at a time: the loads of a batch are independent, so a
      // long walk costs one L2 round trip per batch instead of one per tile
      int64_t t = (int64_t) tile - 1;
      bool done = false;
      while (!done) {
        uint64_t v[LB_BATCH];
#pragma unroll
        for (int i = 0; i < LB_BATCH; i++) v[i] = (t - i >= 0) ? ld_relaxed (a.desc + (uint64_t) (t - i) * RADIX + tid) : 0;
#pragma unroll
        for (int i = 0; i < LB_BATCH; i++) {
          if (done || t - i < 0) continue;
          const uint64_t *p = a.desc + (uint64_t) (t - i) * RADIX + tid;
          while ((v[i] >> 62) == 0 || (v[i] & (0x3full << TAG_SHIFT)) != a.tag) v[i] = ld_relaxed (p);
          excl += v[i] & VALUE_MASK;
          if ((v[i] >> 62) == 2) done = true;
        }
        t -= LB_BATCH;
      }
      st_relaxed (a.desc + tile * RADIX + tid, ST_INCLUSIVE | a.tag | (excl + pub));
    }
    s_gbase[tid] = a.bins[tid] + excl - dbase;
  }
  __syncthreads ();
  // leave in runs of consecutive addresses
#pragma unroll
  for (int i = 0; i < SORT_ITEMS; i++) {
    const int idx = tid + i * SORT_NT;
    if (idx < n_valid && !(a.debug & 2)) {
      const uint64_t k = s_keys[idx];
      const uint32_t d = (uint32_t) (k >> a.shift) & 255u;
      a.out[s_gbase[d] + idx] = k;
      if (VALUES) s_digit[idx] = (uint8_t) d;
    }
  }
  if (VALUES) {
    // the payloads take the same route through the (now free) staging buffer
    uint32_t *s_vals = reinterpret_cast<uint32_t *> (smem_raw);
    __syncthreads ();
#pragma unroll
    for (int j = 0; j < SORT_ITEMS; j++) {
      const uint32_t d = (uint32_t) (key[j] >> a.shift) & 255u;
      const uint32_t r = (j & 1) ? rank2[j / 2] >> 16 : rank2[j / 2] & 0xffffu;
      s_vals[s_dbase[d] + s_whist[warp][d] + r] = val[j];
    }
    __syncthreads ();
#pragma unroll
    for (int i = 0; i < SORT_ITEMS; i++) {
      const int idx = tid + i * SORT_NT;
      if (idx < n_valid) a.vout[s_gbase[s_digit[idx]] + idx] = s_vals[idx];
    }
  }
}

// ------------------------------------------------------------------------------------------
// run-length encoding of the sorted keys
// ------------------------------------------------------------------------------------------
#ifndef GT4_RLE_NT
#define GT4_RLE_NT 128
#endif
#ifndef GT4_RLE_ITEMS
#define GT4_RLE_ITEMS 16
#endif
#ifndef GT4_RLE_MIN_CTAS
#define GT4_RLE_MIN_CTAS 6
#endif
constexpr int RLE_NT = GT4_RLE_NT;
constexpr int RLE_WARPS = RLE_NT / 32;
constexpr int RLE_ITEMS = GT4_RLE_ITEMS;
constexpr int RLE_TILE = RLE_NT * RLE_ITEMS;

#ifndef GT4_RLE_LB_W
#define GT4_RLE_LB_W 4      // rows of 32 descriptors per look-back hop (gt4gpu_device.cuh)
#endif
constexpr int RLE_LB_W = GT4_RLE_LB_W;

__global__ void __launch_bounds__ (RLE_NT, GT4_RLE_MIN_CTAS)
rle_heads_kernel (const uint64_t *__restrict__ keys, uint64_t n, uint64_t *__restrict__ words, uint64_t *__restrict__ first,
                  uint64_t *desc, uint32_t *ticket, unsigned long long *n_unique, int debug)
{
  __shared__ uint32_t s_wcnt[RLE_WARPS];
  __shared__ uint64_t s_base;
  __shared__ uint32_t s_tile;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) s_tile = atomicAdd (ticket, 1u);
  __syncthreads ();
  const uint64_t tile = s_tile;
  const uint64_t base = tile * RLE_TILE + (uint64_t) warp * 32 * RLE_ITEMS;

  // warp-striped like the sort: element j * 32 + lane of the warp's slice
  uint64_t key[RLE_ITEMS];
  uint32_t head[RLE_ITEMS];     // ballot of "first element of a run" per item row
  uint32_t below = 0;           // heads of this warp in earlier rows
  uint64_t carry = 0;           // the element just before the row
  if (base > 0 && base <= n) carry = keys[base - 1];
#pragma unroll
  for (int j = 0; j < RLE_ITEMS; j++) {
    const uint64_t i = base + j * 32 + lane;
    key[j] = i < n ? keys[i] : 0;
    uint64_t prev = __shfl_up_sync (0xffffffffu, key[j], 1);
    if (lane == 0) prev = carry;
    const bool is_head = i < n && (i == 0 || key[j] != prev);
    head[j] = __ballot_sync (0xffffffffu, is_head);
    carry = __shfl_sync (0xffffffffu, key[j], 31);
  }
  uint32_t warp_cnt = 0;
#pragma unroll
  for (int j = 0; j < RLE_ITEMS; j++) warp_cnt += __popc (head[j]);
  if (lane == 0) s_wcnt[warp] = warp_cnt;
  __syncthreads ();
  if (warp == 0) {
    uint32_t v = lane < RLE_WARPS ? s_wcnt[lane] : 0;
    uint32_t incl = v;
#pragma unroll
    for (int off = 1; off < RLE_WARPS; off <<= 1) {
      const uint32_t t = __shfl_up_sync (0xffffffffu, incl, off);
      if (lane >= off) incl += t;
    }
    const uint32_t tile_cnt = __shfl_sync (0xffffffffu, incl, RLE_WARPS - 1);
    if (lane < RLE_WARPS) s_wcnt[lane] = incl - v;
    const uint64_t excl = (debug & 16) ? tile * RLE_TILE : lookback_exclusive<RLE_LB_W> (desc, tile, tile_cnt, lane);
    if (lane == 0) {
      s_base = excl;
      if ((tile + 1) * RLE_TILE >= n) *n_unique = excl + tile_cnt;     // the last tile knows the total
    }
  }
  __syncthreads ();
  const uint64_t out0 = s_base + s_wcnt[warp];
  const uint32_t lt_mask = (1u << lane) - 1u;
#pragma unroll
  for (int j = 0; j < RLE_ITEMS; j++) {
    if ((head[j] >> lane) & 1u) {
      const uint64_t o = out0 + below + __popc (head[j] & lt_mask);
      words[o] = key[j];
      first[o] = base + j * 32 + lane;
    }
    below += __popc (head[j]);
  }
}

__global__ void __launch_bounds__ (256)
rle_counts_kernel (const uint64_t *__restrict__ words_in, const uint64_t *__restrict__ first, uint64_t n_unique, uint64_t n,
                   uint64_t *__restrict__ words_out, uint32_t *__restrict__ counts)
{
  const uint64_t u = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
  if (u >= n_unique) return;
  const uint64_t next = (u + 1 < n_unique) ? first[u + 1] : n;
  words_out[u] = words_in[u];
  counts[u] = (uint32_t) (next - first[u]);        // unsigned int freq of merge_tables_to_file (:1108)
}

}  // namespace

// ------------------------------------------------------------------------------------------
// launchers
// ------------------------------------------------------------------------------------------
size_t sort_scratch_bytes (uint64_t n)
{
  const uint64_t n_tiles = (n + SORT_TILE - 1) / SORT_TILE;
  return SORT_SCRATCH_HEAD + (size_t) n_tiles * RADIX * sizeof (uint64_t);
}

// Sorts n keys ascending on their low 8 * n_pass bits (n_pass >= 1).  `keys` and `alt` are work buffers of n entries; the
// keys are read from `input`, which is either `keys` itself or an array that is only read.  The result lands in one of
// the two buffers (returned through *sorted).  scratch: sort_scratch_bytes (n), any content.
// With vals / valt (n u32 each) every key drags a payload along: vals is filled by the first pass with the key's
// original index, *sorted_vals is the permutation that sorts the input.
static cudaError_t radix_sort_impl (const uint64_t *input, uint64_t *keys, uint64_t *alt, uint32_t *vals, uint32_t *valt, uint64_t n, int n_pass,
                                    unsigned char *scratch, int sm_count, uint64_t **sorted, uint32_t **sorted_vals, cudaStream_t st)
{
  *sorted = keys;
  if (sorted_vals) *sorted_vals = vals;
  if (n == 0) return cudaSuccess;
  if (n_pass > SORT_MAX_PASSES || n_pass < 1) return cudaErrorInvalidValue;
  const uint64_t n_tiles = (n + SORT_TILE - 1) / SORT_TILE;
  if (n_tiles > 0x7fffffffull) return cudaErrorInvalidConfiguration;
  // scratch head: hist [8][256] u64 | bins [8][256] u64 | tickets [8] u32
  unsigned long long *hist = reinterpret_cast<unsigned long long *> (scratch);
  unsigned long long *bins = hist + SORT_MAX_PASSES * RADIX;
  uint32_t *tickets = reinterpret_cast<uint32_t *> (bins + SORT_MAX_PASSES * RADIX);
  uint64_t *desc = reinterpret_cast<uint64_t *> (scratch + SORT_SCRATCH_HEAD);
  cudaError_t e = cudaMemsetAsync (scratch, 0, sort_scratch_bytes (n), st);
  if (e != cudaSuccess) return e;

  uint64_t hist_grid = (uint64_t) sm_count * 2;
  if (hist_grid > (n + 511) / 512) hist_grid = (n + 511) / 512;
  radix_hist_kernel<<<(unsigned) hist_grid, 512, 0, st>>> (input, n, n_pass, hist, reinterpret_cast<unsigned long long *> (tickets) + SORT_OR_SLOT);
  radix_bins_kernel<<<n_pass, RADIX, 0, st>>> (hist, bins);

  static bool configured = false;   // benign race: the attribute is idempotent
  const size_t smem = (size_t) SORT_TILE * sizeof (uint64_t);
  if (!configured) {
    e = cudaFuncSetAttribute (radix_onesweep_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute (radix_onesweep_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  // the first pass may read the caller's own array (input != keys): nothing is copied, nothing of the caller's is written
  const uint64_t *src = input;
  uint64_t *dst = (input == keys) ? alt : keys;
  uint32_t *vsrc = nullptr, *vdst = vals;      // the first pass writes the identity permutation's image into vals
  for (int p = 0; p < n_pass; p++) {
    SweepArgs a;
    a.in = src; a.out = dst; a.n = n; a.n_tiles = n_tiles;
    a.shift = 8 * p;
    a.tag = (uint64_t) (p + 1) << TAG_SHIFT;
    a.bins = bins + p * RADIX;
    a.desc = desc;
    a.ticket = tickets + p;
    a.vin = vsrc; a.vout = vdst;
    a.debug = debug_flags ();
    if (vals) radix_onesweep_kernel<true><<<(unsigned) n_tiles, SORT_NT, smem, st>>> (a);
    else radix_onesweep_kernel<false><<<(unsigned) n_tiles, SORT_NT, smem, st>>> (a);
    src = dst;
    dst = (dst == keys) ? alt : keys;
    if (vals) {
      vsrc = vdst;
      vdst = (vdst == vals) ? valt : vals;
    }
  }
  *sorted = const_cast<uint64_t *> (src);       // one of keys / alt: n_pass >= 1
  if (sorted_vals) *sorted_vals = vsrc;
  return cudaGetLastError ();
}

cudaError_t launch_radix_sort (const uint64_t *input, uint64_t *keys, uint64_t *alt, uint64_t n, int n_pass, unsigned char *scratch,
                               int sm_count, uint64_t **sorted, cudaStream_t st)
{
  return radix_sort_impl (input, keys, alt, nullptr, nullptr, n, n_pass, scratch, sm_count, sorted, nullptr, st);
}

cudaError_t launch_radix_sort_pairs (uint64_t *keys, uint64_t *alt, uint32_t *vals, uint32_t *valt, uint64_t n, int n_pass,
                                     unsigned char *scratch, int sm_count, uint64_t **sorted, uint32_t **sorted_vals, cudaStream_t st)
{
  if (n > 0xffffffffull) return cudaErrorInvalidValue;     // payloads are 32-bit indices
  return radix_sort_impl (keys, keys, alt, vals, valt, n, n_pass, scratch, sm_count, sorted, sorted_vals, st);
}

size_t rle_scratch_bytes (uint64_t n)
{
  const uint64_t n_tiles = (n + RLE_TILE - 1) / RLE_TILE;
  return 256 + (size_t) n_tiles * sizeof (uint64_t);
}

// sorted keys -> words_tmp[u], first[u] for every run u; *d_n_unique (device, u64) receives the number of runs
cudaError_t launch_rle_heads (const uint64_t *sorted, uint64_t n, uint64_t *words_tmp, uint64_t *first, unsigned char *scratch,
                              int sm_count, unsigned long long **d_n_unique, cudaStream_t st)
{
  *d_n_unique = reinterpret_cast<unsigned long long *> (scratch);
  cudaError_t e = cudaMemsetAsync (scratch, 0, rle_scratch_bytes (n), st);
  if (e != cudaSuccess || n == 0) return e;
  const uint64_t n_tiles = (n + RLE_TILE - 1) / RLE_TILE;
  if (n_tiles > 0x7fffffffull) return cudaErrorInvalidConfiguration;
  uint32_t *ticket = reinterpret_cast<uint32_t *> (scratch + 8);
  uint64_t *desc = reinterpret_cast<uint64_t *> (scratch + 256);
  (void) sm_count;     // one CTA per tile: a persistent loop over tickets measured slower (the next CTA's loads overlap the stores)
  rle_heads_kernel<<<(unsigned) n_tiles, RLE_NT, 0, st>>> (sorted, n, words_tmp, first, desc, ticket, *d_n_unique, debug_flags ());
  return cudaGetLastError ();
}

cudaError_t launch_rle_counts (const uint64_t *words_tmp, const uint64_t *first, uint64_t n_unique, uint64_t n,
                               uint64_t *words, uint32_t *counts, cudaStream_t st)
{
  if (n_unique == 0) return cudaSuccess;
  rle_counts_kernel<<<(unsigned) ((n_unique + 255) / 256), 256, 0, st>>> (words_tmp, first, n_unique, n, words, counts);
  return cudaGetLastError ();
}

}  // namespace gt4gpu
